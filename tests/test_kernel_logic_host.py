"""Kernel SOURCE vs oracle, on the CPU: clv_device.cuh compiled for the host by tests/hostemu (test infrastructure)
must reproduce the shared-math oracle bit for bit, stage by stage.  The `-m gpu` tests repeat this on the device."""
import numpy as np
import pytest

from clive2_b200 import struct_types as T
from oracle.oracle import Oracle, OracleRenderer
from tests.common import (AWKWARD_CAMERAS, StageBuffers, arrays, assert_bits_equal, assert_paths_equal, awkward_camera_arrays,
                          check_ordered_against_reference, fuzz_rays,
                          fuzz_soup, many_lights_arrays, oracle_stages, random_material_arrays, random_rays,
                          rel_rmse, seeds_for, zoo_arrays)
from tests.hostemu import MODE_ORDERED, MODE_REFERENCE, HostEmu

SCENES = [("empty", 40, 40), ("bench-spheres", 48, 32)]


@pytest.fixture(scope="module")
def O():
    return Oracle(shared_math=True)


@pytest.fixture(scope="module")
def E():
    return HostEmu()


def emu_stages(E, acc, a, B, mode):
    b = B.batch
    E.generate_light_rays(b, a.light_triangles, a.light_surface_areas, a.light_triangle_indices, a.materials, B.rand,
                          B.light_rays, a.light_counts)
    E.generate_camera_rays(b, a.camera, B.rand, B.indices, B.camera_rays)
    E.generate_paths(b, B.light_rays, acc, a.materials, B.rand, B.out_light, B.light_paths, B.debug, mode)
    E.generate_paths(b, B.camera_rays, acc, a.materials, B.rand, B.out_camera, B.camera_paths, B.debug, mode)
    E.connect_paths(b, B.camera_paths, B.light_paths, acc, a.materials, a.camera, B.aggs, B.out_samples, B.pix, B.pth,
                    B.ry, B.w, B.sh, None, mode)
    E.adaptive_finalize_samples(b, B.aggs, a.camera, B.finalized, B.sample_counts, B.bins_identity, B.sample_weights)
    return B


@pytest.mark.parametrize("preset,w,h", SCENES)
@pytest.mark.parametrize("mode", [MODE_REFERENCE, MODE_ORDERED])
def test_stage_buffers_bit_exact(O, E, preset, w, h, mode):
    a = arrays(preset, w, h)
    seeds = seeds_for(w * h)
    R = oracle_stages(O, a, StageBuffers(w * h, seeds), upto="finalize")
    acc = E.accel(a.boxes, a.triangles)
    try:
        S = emu_stages(E, acc, a, StageBuffers(w * h, seeds), mode)
    finally:
        E.accel_destroy(acc)
    assert_bits_equal(R.rand, S.rand, "RNG state")
    assert_bits_equal(R.light_rays, S.light_rays, "light rays")
    assert_bits_equal(R.camera_rays, S.camera_rays, "camera rays")
    assert_paths_equal(R.light_paths, S.light_paths, "light paths")
    assert_paths_equal(R.camera_paths, S.camera_paths, "camera paths")
    assert_bits_equal(R.out_camera, S.out_camera, "unidirectional estimate")
    for name in ("pix", "pth", "ry", "w", "sh"):
        assert_bits_equal(getattr(R, name), getattr(S, name), "light entries " + name)
    assert_bits_equal(R.aggs, S.aggs, "weight aggregators")
    assert_bits_equal(R.out_samples, S.out_samples, "out_samples")
    assert_bits_equal(R.finalized, S.finalized, "finalized samples")
    assert_bits_equal(R.sample_weights, S.sample_weights, "sample weights")
    assert np.array_equal(R.sample_counts, S.sample_counts)


@pytest.mark.parametrize("preset,w,h", SCENES)
def test_traversal_modes_agree_with_oracle(O, E, preset, w, h):
    a = arrays(preset, w, h)
    rays = random_rays(60000, [-9, -1.9, -9], [9, 9, 9], seed=3)
    tri, t, u, v = O.traverse(rays, a.boxes, a.triangles)
    acc = E.accel(a.boxes, a.triangles)
    try:
        for mode in (MODE_REFERENCE, MODE_ORDERED):
            tri2, t2, u2, v2 = E.traverse(rays, acc, mode)
            assert np.array_equal(tri, tri2), (mode, int((tri != tri2).sum()))
            hit = tri >= 0
            assert np.array_equal(t[hit].view(np.uint32), t2[hit].view(np.uint32))
            assert np.array_equal(u[hit].view(np.uint32), u2[hit].view(np.uint32))
            assert np.array_equal(v[hit].view(np.uint32), v2[hit].view(np.uint32))
        assert (tri >= 0).mean() > 0.99  # closed box: (almost) every ray hits something
    finally:
        E.accel_destroy(acc)


def test_ordered_traversal_does_less_work(E):
    a = arrays("bench-spheres", 48, 32)
    rays = random_rays(20000, [-9, -1.9, -9], [9, 9, 9], seed=4)
    acc = E.accel(a.boxes, a.triangles)
    try:
        E.counters()
        E.traverse(rays, acc, MODE_REFERENCE)
        ref = E.counters()
        E.traverse(rays, acc, MODE_ORDERED)
        ordd = E.counters()
    finally:
        E.accel_destroy(acc)
    assert ordd["node_fetches"] < ref["node_fetches"] and ordd["tri_fetches"] < ref["tri_fetches"]


def test_visibility_matches_oracle(O, E):
    a = arrays("bench-spheres", 48, 32)
    rays = random_rays(20000, [-9, -1.9, -9], [9, 9, 9], seed=5)
    tri, t, _, _ = O.traverse(rays, a.boxes, a.triangles)
    ok = tri >= 0
    # vertex pairs: a = ray origin (off-surface, triangle -1), b = hit point on triangle tri (visible by construction)
    # and b' = hit point of ANOTHER ray (mostly occluded or visible by chance)
    A = rays[ok].copy()
    B = np.zeros(len(A), T.Ray)
    B["origin"][:, :3] = A["origin"][:, :3] + A["direction"][:, :3] * t[ok][:, None]
    B["triangle"] = tri[ok]
    B2 = np.roll(B, 7)
    acc = E.accel(a.boxes, a.triangles)
    try:
        for b in (B, B2):
            ref = O.visibility(A, b, a.boxes, a.triangles)
            for mode in (MODE_REFERENCE, MODE_ORDERED):
                assert np.array_equal(ref, E.visibility(A, b, acc, mode)), mode
        assert O.visibility(A, B, a.boxes, a.triangles).mean() > 0.9
        assert 0.02 < O.visibility(A, B2, a.boxes, a.triangles).mean() < 0.98
    finally:
        E.accel_destroy(acc)


def test_single_leaf_scene_and_long_leaf(O, E):
    """Edge cases of the BVH encoding: the root is a leaf; a leaf longer than the 15 triangles a leaf ref encodes."""
    rng = np.random.default_rng(11)
    for n_tris in (3, 40):
        tris = np.zeros(n_tris, T.Triangle)
        base = rng.uniform(-1, 1, (n_tris, 3)).astype(np.float32)
        tris["v0"][:, :3] = base
        tris["v1"][:, :3] = base + rng.normal(0, 0.5, (n_tris, 3)).astype(np.float32)
        tris["v2"][:, :3] = base + rng.normal(0, 0.5, (n_tris, 3)).astype(np.float32)
        boxes = np.zeros(1, T.Box)
        allv = np.concatenate([tris["v0"][:, :3], tris["v1"][:, :3], tris["v2"][:, :3]])
        boxes["min"][0, :3] = allv.min(0)
        boxes["max"][0, :3] = allv.max(0)
        boxes["left"], boxes["right"] = 0, n_tris
        rays = random_rays(5000, [-3, -3, -3], [3, 3, 3], seed=n_tris)
        tri, t, u, v = O.traverse(rays, boxes, tris)
        acc = E.accel(boxes, tris)
        try:
            for mode in (MODE_REFERENCE, MODE_ORDERED):
                tri2, t2, _, _ = E.traverse(rays, acc, mode)
                assert np.array_equal(tri, tri2)
                assert np.array_equal(t[tri >= 0].view(np.uint32), t2[tri >= 0].view(np.uint32))
        finally:
            E.accel_destroy(acc)
        assert 0 < (tri >= 0).sum() < len(rays)


def test_bad_bvh_arrays_are_rejected(E):
    boxes = np.zeros(2, T.Box)
    boxes["left"][0], boxes["right"][0] = 5, 0  # child index out of range
    with pytest.raises(ValueError):
        E.accel(boxes, np.zeros(1, T.Triangle))


def _chain_bvh(levels):
    """A degenerate tree: every inner node has a one-triangle leaf on the left and the rest of the chain on the right."""
    n_boxes, n_tris = 2 * levels + 1, levels + 1
    boxes, tris = np.zeros(n_boxes, T.Box), np.zeros(n_tris, T.Triangle)
    for k in range(n_tris):
        tris["v0"][k, :3] = (k, 0, 0)
        tris["v1"][k, :3] = (k + 1, 0, 0)
        tris["v2"][k, :3] = (k, 1, 0)
    for k in range(levels):  # inner node 2k: children 2k+1 (leaf, triangle k) and 2k+2 (rest)
        boxes["left"][2 * k], boxes["right"][2 * k] = 2 * k + 1, 0
        boxes["min"][2 * k, :3], boxes["max"][2 * k, :3] = (k, 0, 0), (n_tris + 1, 1, 0)
        boxes["left"][2 * k + 1], boxes["right"][2 * k + 1] = k, k + 1
        boxes["min"][2 * k + 1, :3], boxes["max"][2 * k + 1, :3] = (k, 0, 0), (k + 1, 1, 0)
    boxes["left"][2 * levels], boxes["right"][2 * levels] = levels, levels + 1
    boxes["min"][2 * levels, :3], boxes["max"][2 * levels, :3] = (levels, 0, 0), (levels + 1, 1, 0)
    return boxes, tris


def test_trees_deeper_than_the_traversal_stacks_are_refused_not_truncated(O, E):
    """The reference's 64-entry stack silently aborts a search that overflows it (trace.metal:149).  Here a structure whose
    tree could overflow the kernels' stacks (one deferred sibling per level) is refused when it is built; a tree at the
    limit traverses exactly."""
    boxes, tris = _chain_bvh(20)  # 20 binary levels: the 4-wide collapse has up to 3 deferred siblings per level
    acc = E.accel(boxes, tris)
    try:
        rays = random_rays(2000, [0, 0, -3], [21, 1, 3], seed=4)
        tri, t, _, _ = O.traverse(rays, boxes, tris)
        for mode in (MODE_REFERENCE, MODE_ORDERED):
            tri2, t2, _, _ = E.traverse(rays, acc, mode)
            assert np.array_equal(tri, tri2)
            assert np.array_equal(t[tri >= 0].view(np.uint32), t2[tri >= 0].view(np.uint32))
        assert (tri >= 0).sum() > 100
    finally:
        E.accel_destroy(acc)
    with pytest.raises(ValueError):
        E.accel(*_chain_bvh(64))


def test_fused_pass_matches_oracle_renderer(E):
    """clv_renderer_run_sample's kernel sequence vs the reference's eight stages (oracle): everything that does not go
    through the atomically splatted light image is bit exact; the image agrees to float-summation-order tolerance."""
    w = h = 40
    a = arrays("empty", w, h)
    seeds = seeds_for(w * h)
    ref = OracleRenderer(a, seeds, shared_math=True)
    b = w * h
    s = seeds.copy()
    lp, cp = np.zeros(b, T.Path), np.zeros(b, T.Path)
    aggs = np.zeros(b, T.WeightAggregator)
    light_image = np.zeros((b, 4), np.float32)
    accum = np.zeros((b, 8), np.float32)
    acc = E.accel(a.boxes, a.triangles)
    try:
        for _ in range(3):
            ref.run_sample()
            E.fused_pass(a, acc, s, lp, cp, aggs, light_image, accum)
    finally:
        E.accel_destroy(acc)
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), s, "RNG state after 3 passes")
    assert_bits_equal(ref.unidirectional_image_buffer.reshape(b, 3), accum[:, 4:7].copy(), "unidirectional image")
    assert np.array_equal(ref.summed_sample_counts.reshape(b), accum[:, 7].astype(np.int32))
    assert (light_image == 0).all()  # cleared for the next pass
    np.testing.assert_allclose(accum[:, 3], ref.summed_sample_weights.reshape(b), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(accum[:, 0:3], ref.summed_image.reshape(b, 3), rtol=2e-4, atol=1e-6)
    est_ref = np.nan_to_num(ref.summed_image / ref.summed_sample_weights).reshape(b, 3)
    est = np.nan_to_num(accum[:, 0:3] / accum[:, 3:4])
    assert rel_rmse(est, est_ref) < 1e-4


def test_sharded_fused_pass_sums_to_full_render(E):
    """Two shards (sample ids split in 8-pixel stripes) keep the global RNG streams; their accumulators add up to the
    single-renderer result (float summation order aside)."""
    w = h = 32
    a = arrays("empty", w, h)
    b = w * h
    seeds = seeds_for(b)
    acc = E.accel(a.boxes, a.triangles)

    def render(ids):
        n = b if ids is None else len(ids)
        s = (seeds if ids is None else seeds[ids]).copy()
        lp, cp = np.zeros(n, T.Path), np.zeros(n, T.Path)
        aggs = np.zeros(b, T.WeightAggregator)
        li = np.zeros((b, 4), np.float32)
        accum = np.zeros((b, 8), np.float32)
        for _ in range(2):
            E.fused_pass(a, acc, s, lp, cp, aggs, li, accum, sample_ids=ids)
        return accum, s

    try:
        full, s_full = render(None)
        ids = np.arange(b, dtype=np.uint32)
        stripe = (ids // 8) % 2
        parts = [render(np.ascontiguousarray(ids[stripe == r])) for r in (0, 1)]
    finally:
        E.accel_destroy(acc)
    total = parts[0][0] + parts[1][0]
    assert np.array_equal(total[:, 7], full[:, 7])                      # counts
    assert_bits_equal(total[:, 4:7].copy(), full[:, 4:7].copy(), "unidirectional")  # one owner per pixel
    np.testing.assert_allclose(total[:, 0:4], full[:, 0:4], rtol=2e-4, atol=1e-6)
    for r in (0, 1):
        assert_bits_equal(parts[r][1], s_full[stripe == r], "shard RNG streams")


@pytest.mark.parametrize("mode", [MODE_REFERENCE, MODE_ORDERED])
def test_every_material_branch_bit_exact(O, E, mode):
    """Mirror, coated, smooth and rough dielectrics, alpha > 0 on a diffuse surface: all stage buffers against the oracle."""
    w, h = 40, 30
    a = zoo_arrays(w, h)
    seeds = seeds_for(w * h, seed=77)
    R = oracle_stages(O, a, StageBuffers(w * h, seeds), upto="finalize")
    acc = E.accel(a.boxes, a.triangles)
    try:
        S = emu_stages(E, acc, a, StageBuffers(w * h, seeds), mode)
    finally:
        E.accel_destroy(acc)
    used = np.unique(np.concatenate([R.camera_paths["rays"]["material"][R.camera_paths["length"] > 1, 1],
                                     R.light_paths["rays"]["material"][R.light_paths["length"] > 1, 1]]))
    assert {0, 1, 2, 3, 5}.issubset(set(used.tolist())), used  # the walk really visited every material kind
    assert_bits_equal(R.rand, S.rand, "RNG state")
    assert_paths_equal(R.light_paths, S.light_paths, "light paths")
    assert_paths_equal(R.camera_paths, S.camera_paths, "camera paths")
    for name in ("pix", "pth", "ry", "w", "sh"):
        assert_bits_equal(getattr(R, name), getattr(S, name), "light entries " + name)
    assert_bits_equal(R.aggs, S.aggs, "weight aggregators")
    assert_bits_equal(R.finalized, S.finalized, "finalized samples")
    assert_bits_equal(R.sample_weights, S.sample_weights, "sample weights")
    assert (R.camera_paths["length"] < 6).any()  # glass/mirror paths also terminate early


def test_adversarial_rays_through_vertices_and_edges(O, E):
    """rays aimed EXACTLY at mesh vertices / points on edges, where a box's slab distance and the triangle's
    own t disagree in the last bits.  The REFERENCE_ORDER flavour must still equal the reference on every ray; the
    ORDERED flavour may return an adjacent triangle (same point, t within a few ulps) in a small fraction of these
    measure-zero rays — this test puts the number on it (0.5 % here; 0 of ~10^6 random rays in the other tests)."""
    a = arrays("bench-spheres", 48, 32)
    tr = a.triangles
    rng = np.random.default_rng(0)
    n = 200000
    k = rng.integers(0, len(tr), n)
    v0, v1, v2 = tr["v0"][k, :3], tr["v1"][k, :3], tr["v2"][k, :3]
    kind = rng.integers(0, 3, n)[:, None]
    w = rng.random((n, 1)).astype(np.float32)
    target = np.where(kind == 0, v0, np.where(kind == 1, v0 * (1 - w) + v1 * w, v1 * (1 - w) + v2 * w)).astype(np.float32)
    origin = rng.uniform([-9, -1.9, -9], [9, 9, 9], (n, 3)).astype(np.float32)
    d = target - origin
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.zeros(n, T.Ray)
    rays["origin"][:, :3], rays["direction"][:, :3] = origin, d
    with np.errstate(divide="ignore"):
        rays["inv_direction"][:, :3] = np.float32(1) / d
    tri, t, _, _ = O.traverse(rays, a.boxes, a.triangles)
    acc = E.accel(a.boxes, a.triangles)
    try:
        tri_r, t_r, _, _ = E.traverse(rays, acc, MODE_REFERENCE)
        tri_o, t_o, _, _ = E.traverse(rays, acc, MODE_ORDERED)
    finally:
        E.accel_destroy(acc)
    assert np.array_equal(tri, tri_r) and np.array_equal(t.view(np.uint32), t_r.view(np.uint32))
    mm = tri != tri_o
    assert mm.mean() < 0.01, float(mm.mean())
    both = mm & (tri >= 0) & (tri_o >= 0)
    # the same surface point: within a few ulps for vertex/edge ties; up to ~1e-4 relative for grazing hits, where the
    # reference's Moeller-Trumbore test (no parallel-ray guard, f = 1/a with a -> 0) is itself that inaccurate
    rel = np.abs(t[both] - t_o[both]) / t[both]
    assert np.median(rel) < 1e-6 and rel.max() < 1e-3, (float(np.median(rel)), float(rel.max()))
    print(f"ordered-vs-reference disagreement on vertex/edge rays: {int(mm.sum())} of {n}")


@pytest.mark.parametrize("preset,w,h", [("bench-spheres", 40, 30)])
def test_soa_vertex_records_give_the_same_bits_as_path_records(E, preset, w, h):
    """The fused renderer's compact SoA vertex layout vs the reference's Path records: identical RNG streams,
    aggregators and accumulators apart from the atomic summation order of the light image."""
    a = arrays(preset, w, h)
    b = w * h
    seeds = seeds_for(b, seed=9)
    acc = E.accel(a.boxes, a.triangles)
    outs = []
    try:
        for use_soa in (False, True):
            s = seeds.copy()
            lp, cp = np.zeros(b, T.Path), np.zeros(b, T.Path)
            aggs = np.zeros(b, T.WeightAggregator)
            li = np.zeros((b, 4), np.float32)
            accum = np.zeros((b, 8), np.float32)
            soa = np.zeros((2 * 6 * 4 * b, 4), np.float32) if use_soa else None
            soa_len = np.zeros(2 * b, np.int32) if use_soa else None
            for _ in range(3):
                E.fused_pass(a, acc, s, lp, cp, aggs, li, accum, soa=soa, soa_len=soa_len)
            outs.append((s, aggs, accum, lp, soa_len))
    finally:
        E.accel_destroy(acc)
    (s0, g0, a0, lp0, _), (s1, g1, a1, _, len1) = outs
    assert_bits_equal(s0, s1, "RNG state")
    assert_bits_equal(g0, g1, "aggregators")
    assert_bits_equal(a0[:, 4:8].copy(), a1[:, 4:8].copy(), "unidirectional + counts")
    np.testing.assert_allclose(a1[:, 0:4], a0[:, 0:4], rtol=2e-4, atol=1e-6)
    assert np.array_equal(len1[:b], lp0["length"])  # SoA lengths: [light | camera]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzzed_soups_and_degenerate_rays(O, E, seed):
    """Awkward geometry (exact duplicates, coplanar overlaps, zero-area triangles, flat boxes) under awkward rays
    (axis-parallel: inv_direction = inf and 0 * inf = NaN in the slab test; origins on vertices and in triangle
    planes).  REFERENCE_ORDER must equal the oracle on every ray, bits of t/u/v included; ORDERED may differ only in
    the documented class (same t to within rounding), and never on an exact tie between duplicates."""
    boxes, tris = fuzz_soup(seed)
    rays = fuzz_rays(tris, 60000, seed + 100)
    tri, t, u, v = O.traverse(rays, boxes, tris)
    acc = E.accel(boxes, tris)
    try:
        tri_r, t_r, u_r, v_r = E.traverse(rays, acc, MODE_REFERENCE)
        tri_o, t_o, _, _ = E.traverse(rays, acc, MODE_ORDERED)
    finally:
        E.accel_destroy(acc)
    assert (tri >= 0).mean() > 0.5
    assert np.array_equal(tri, tri_r)
    hit = tri >= 0
    for x, y in ((t, t_r), (u, u_r), (v, v_r)):
        assert np.array_equal(x[hit].view(np.uint32), y[hit].view(np.uint32))
    check_ordered_against_reference(rays, tris, tri, t, tri_o, t_o)
    # the visibility predicate on the same soup: a -> the reference's own hit point (mostly visible) and a -> someone
    # else's hit point (mostly blocked)
    ok = (tri >= 0) & np.isfinite(t)
    A = rays[ok].copy()
    B = np.zeros(len(A), T.Ray)
    B["origin"][:, :3] = A["origin"][:, :3] + A["direction"][:, :3] * t[ok][:, None]
    B["triangle"] = tri[ok]
    acc = E.accel(boxes, tris)
    try:
        for b in (B, np.roll(B, 7)):
            ref = O.visibility(A, b, boxes, tris)
            assert np.array_equal(ref, E.visibility(A, b, acc, MODE_REFERENCE))
            assert (ref != E.visibility(A, b, acc, MODE_ORDERED)).mean() < 1e-3  # the artefact class again
    finally:
        E.accel_destroy(acc)


@pytest.mark.parametrize("which", range(len(AWKWARD_CAMERAS)))
def test_awkward_cameras_stage_buffers_bit_exact(O, E, which):
    """Cameras outside the enclosure, facing a wall, next to a wall: empty / one-vertex camera paths, connections from
    nothing, projections beyond the last pixel.  Every stage buffer still equals the oracle, in both traversal flavours."""
    w, h = 32, 24
    a = awkward_camera_arrays(which, w, h)
    seeds = seeds_for(w * h, seed=5)
    R = oracle_stages(O, a, StageBuffers(w * h, seeds), upto="finalize")
    acc = E.accel(a.boxes, a.triangles)
    try:
        for mode in (MODE_REFERENCE, MODE_ORDERED):
            S = emu_stages(E, acc, a, StageBuffers(w * h, seeds), mode)
            assert_bits_equal(R.rand, S.rand, "RNG state")
            assert_paths_equal(R.light_paths, S.light_paths, "light paths")
            assert_paths_equal(R.camera_paths, S.camera_paths, "camera paths")
            for name in ("pix", "pth", "ry", "w", "sh"):
                assert_bits_equal(getattr(R, name), getattr(S, name), "light entries " + name)
            assert_bits_equal(R.aggs, S.aggs, "weight aggregators")
            assert_bits_equal(R.finalized, S.finalized, "finalized samples")
            assert_bits_equal(R.sample_weights, S.sample_weights, "sample weights")
    finally:
        E.accel_destroy(acc)
    if which == 0:
        assert (R.camera_paths["length"] == 0).mean() > 0.5       # the case this camera is here for
    if which == 2:
        assert R.pix.max() >= w * h                               # a projection beyond the last pixel (kept, never gathered)


@pytest.mark.parametrize("trial", range(6))
def test_random_material_tables_bit_exact(O, E, trial):
    """Random material tables from extreme values (tests/common.py random_material_arrays) through every stage."""
    w, h = 40, 30
    a = random_material_arrays(trial, w, h)
    seeds = seeds_for(w * h, seed=trial)
    R = oracle_stages(O, a, StageBuffers(w * h, seeds), upto="finalize")
    acc = E.accel(a.boxes, a.triangles)
    try:
        for mode in (MODE_REFERENCE, MODE_ORDERED):
            S = emu_stages(E, acc, a, StageBuffers(w * h, seeds), mode)
            assert_bits_equal(R.rand, S.rand, "RNG state")
            assert_paths_equal(R.light_paths, S.light_paths, "light paths")
            assert_paths_equal(R.camera_paths, S.camera_paths, "camera paths")
            for name in ("pix", "pth", "ry", "w", "sh"):
                assert_bits_equal(getattr(R, name), getattr(S, name), "light entries " + name)
            assert_bits_equal(R.aggs, S.aggs, "weight aggregators")
            assert_bits_equal(R.finalized, S.finalized, "finalized samples")
    finally:
        E.accel_destroy(acc)


@pytest.mark.parametrize("w,h", [(1, 1), (2, 1), (1, 3), (3, 2), (33, 1), (1, 65)])
def test_tiny_frames_bit_exact(O, E, w, h):
    """One pixel, one row, one column: every border case of the 3x3 finalize neighbourhood."""
    a = arrays("empty", w, h)
    seeds = seeds_for(w * h, seed=3)
    R = oracle_stages(O, a, StageBuffers(w * h, seeds), upto="finalize")
    acc = E.accel(a.boxes, a.triangles)
    try:
        S = emu_stages(E, acc, a, StageBuffers(w * h, seeds), MODE_ORDERED)
    finally:
        E.accel_destroy(acc)
    assert_bits_equal(R.rand, S.rand, "RNG state")
    assert_bits_equal(R.aggs, S.aggs, "weight aggregators")
    assert_bits_equal(R.finalized, S.finalized, "finalized samples")
    assert_bits_equal(R.sample_weights, S.sample_weights, "sample weights")


@pytest.mark.parametrize("n_extra", [0, 1, 9])
def test_light_counts_bit_exact(O, E, n_extra):
    """1, 3 and 11 light triangles (the reference's scenes always have 2): light pick, light pdf, emitters on a mesh."""
    w, h = 40, 30
    a = many_lights_arrays(n_extra, w, h)
    assert int(a.light_counts) == (1 if n_extra == 0 else 2 + n_extra)
    seeds = seeds_for(w * h, seed=8 + n_extra)
    R = oracle_stages(O, a, StageBuffers(w * h, seeds), upto="finalize")
    acc = E.accel(a.boxes, a.triangles)
    try:
        for mode in (MODE_REFERENCE, MODE_ORDERED):
            S = emu_stages(E, acc, a, StageBuffers(w * h, seeds), mode)
            assert_bits_equal(R.rand, S.rand, "RNG state")
            assert_bits_equal(R.light_rays, S.light_rays, "light rays")
            assert_paths_equal(R.light_paths, S.light_paths, "light paths")
            assert_paths_equal(R.camera_paths, S.camera_paths, "camera paths")
            assert_bits_equal(R.aggs, S.aggs, "weight aggregators")
            assert_bits_equal(R.finalized, S.finalized, "finalized samples")
    finally:
        E.accel_destroy(acc)
    assert len(np.unique(R.light_rays["triangle"])) == int(a.light_counts)  # every light gets picked


def _fused_vs_oracle_renderer(E, a, seeds, passes=3, wavefront=False, suspend_after=0):
    b = a.pixel_width * a.pixel_height
    ref = OracleRenderer(a, seeds, shared_math=True)
    s = seeds.copy()
    aggs = np.zeros(b, T.WeightAggregator)
    light_image = np.zeros((b, 4), np.float32)
    accum = np.zeros((b, 8), np.float32)
    soa = np.zeros((2 * 6 * 4 * b, 4), np.float32)
    soa_len = np.zeros(2 * b, np.int32)
    acc = E.accel(a.boxes, a.triangles)
    try:
        for _ in range(passes):
            ref.run_sample()
            E.fused_pass(a, acc, s, None, None, aggs, light_image, accum, soa=soa, soa_len=soa_len, wavefront=wavefront,
                         suspend_after=suspend_after)
    finally:
        E.accel_destroy(acc)
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), s, "RNG state")
    assert_bits_equal(ref.buffer("aggregators", T.WeightAggregator), aggs, "aggregators")
    assert_bits_equal(ref.unidirectional_image_buffer.reshape(b, 3), accum[:, 4:7].copy(), "unidirectional image")
    assert np.array_equal(ref.summed_sample_counts.reshape(b), accum[:, 7].astype(np.int32))
    assert (light_image == 0).all()
    with np.errstate(invalid="ignore"):
        np.testing.assert_allclose(accum[:, 3], ref.summed_sample_weights.reshape(b), rtol=5e-5, atol=1e-6)
        np.testing.assert_allclose(accum[:, 0:3], ref.summed_image.reshape(b, 3), rtol=5e-4, atol=1e-6)


@pytest.mark.parametrize("wavefront", [False, True], ids=["megakernel", "wavefront"])
@pytest.mark.parametrize("which", range(len(AWKWARD_CAMERAS)))
def test_fused_pass_awkward_cameras(E, which, wavefront):
    """The fused pass (atomic splats instead of sort + bins + gather) where light-image entries fall beyond the last
    pixel or do not exist at all: dropped / absent exactly as in the reference's gather."""
    w, h = 32, 24
    _fused_vs_oracle_renderer(E, awkward_camera_arrays(which, w, h), seeds_for(w * h, seed=5), wavefront=wavefront)


@pytest.mark.parametrize("wavefront", [False, True], ids=["megakernel", "wavefront"])
@pytest.mark.parametrize("n_extra", [0, 9])
def test_fused_pass_light_counts(E, n_extra, wavefront):
    w, h = 32, 24
    _fused_vs_oracle_renderer(E, many_lights_arrays(n_extra, w, h), seeds_for(w * h, seed=6), wavefront=wavefront)


@pytest.mark.parametrize("wavefront", [1, 2], ids=["binary_nodes", "wide_nodes"])
@pytest.mark.parametrize("preset,w,h", [("empty", 40, 40), ("bench-spheres", 48, 36)])
def test_wavefront_walks_vs_oracle_renderer(E, preset, w, h, wavefront):
    """Stages 1-4 as wavefront rounds (the logic of k_walk_init / k_walk_trace / k_walk_shade: resumable single-step
    closest-hit search, per-bounce walk state in memory, live-sample queues) against the oracle renderer."""
    _fused_vs_oracle_renderer(E, arrays(preset, w, h), seeds_for(w * h, seed=12), passes=3, wavefront=wavefront)


@pytest.mark.parametrize("suspend_after", [1, 3, 9])
@pytest.mark.parametrize("wavefront", [1, 2], ids=["binary_nodes", "wide_nodes"])
def test_suspended_searches_resume_to_the_same_hit(E, wavefront, suspend_after):
    """k_walk_trace suspends the searches that outlast their round (continuation records: scalars + live stack, resumed by
    the next round while the sample waits).  Here every search longer than 1 / 3 / 9 steps is suspended, up to
    CLV_WALK_MAX_LAG times per sample: images, records and RNG rows still equal the oracle renderer's bit for bit."""
    _fused_vs_oracle_renderer(E, arrays("bench-spheres", 48, 36), seeds_for(48 * 36, seed=13), passes=2, wavefront=wavefront,
                              suspend_after=suspend_after)


@pytest.mark.parametrize("max_members", [8, 15], ids=["leaves_le_8", "leaves_le_15"])
@pytest.mark.parametrize("wide", [False, True], ids=["binary_nodes", "wide_nodes"])
def test_cooperative_leaf_steps_in_lockstep_on_the_host(E, monkeypatch, wide, max_members):
    """hit_leaf_step_warp / occ_leaf_step_warp — the warp-cooperative leaf steps of k_walk_trace / k_connect_rays (pairs
    dealt to all 32 lanes by ballots and shuffles) — run here as the very same source on 32 host threads in lockstep and
    are compared, after every leaf, with the per-lane loops they replace: same best hit / verdict, same stack, same next
    node, bit for bit; the searches end at closest_hit's hits.  Leaves of the reference's size (<= 8) and of the largest
    size a leaf reference encodes (15)."""
    from clive2_b200 import bvh
    from clive2_b200 import scene
    monkeypatch.setattr(bvh, "MAX_MEMBERS", max_members)
    a = scene.preset_arrays("bench-spheres", 48, 32, bvh_builder="numpy")  # not tests.common.arrays: that one caches
    sizes = (a.boxes["right"] - a.boxes["left"])[a.boxes["right"] != 0]
    assert sizes.max() == max_members
    rays = random_rays(32 * 40 - 5, [-9, -1.9, -9], [9, 9, 9], seed=21)  # a ragged last group
    acc = E.accel(a.boxes, a.triangles)
    try:
        r = E.coop_hit_leaves(rays, acc, wide=wide)
        if r is None:
            pytest.skip("no team of 32 OpenMP threads")
        assert r["compared"] > len(rays) and r["mismatches"] == 0 and r["wrong_hits"] == 0, r
        assert r["max_pairs"] > 64, r  # several rounds of 32 pairs in one step
        if not wide:
            # occlusion queries bounded at random distances, some skipping a triangle, some with a rank bound that matters
            rng = np.random.default_rng(5)
            tri, t, _, _ = E.traverse(rays, acc)
            t_bound = np.where(rng.random(len(rays)) < 0.5, t, rng.uniform(0.5, 12.0, len(rays))).astype(np.float32)
            rank_bound = np.where(rng.random(len(rays)) < 0.5, 0x7FFFFFFF, rng.integers(0, len(a.triangles), len(rays))).astype(np.int32)
            skip = np.where(rng.random(len(rays)) < 0.5, tri, -1).astype(np.int32)
            q = E.coop_occ_leaves(rays, t_bound, rank_bound, skip, acc)
            assert q["compared"] > len(rays) and q["mismatches"] == 0 and q["wrong_verdicts"] == 0, q
    finally:
        E.accel_destroy(acc)


@pytest.mark.parametrize("use_soa", [False, True], ids=["path_records", "soa_records"])
def test_wavefront_walks_write_the_same_records_as_the_megakernel(E, use_soa):
    """Same Path records / SoA vertex records, lengths, RNG state, aggregators and accumulators from both drivers of the
    walk (GGX + transmission scene, 3 passes), byte for byte where a record is defined."""
    w, h = 40, 30
    a = arrays("bench-spheres", w, h)
    b = w * h
    seeds = seeds_for(b, seed=21)
    acc = E.accel(a.boxes, a.triangles)
    outs = []
    try:
        for wavefront in (False, True):
            s = seeds.copy()
            lp, cp = np.zeros(b, T.Path), np.zeros(b, T.Path)
            aggs = np.zeros(b, T.WeightAggregator)
            li = np.zeros((b, 4), np.float32)
            accum = np.zeros((b, 8), np.float32)
            soa = np.zeros((2 * 6 * 4 * b, 4), np.float32) if use_soa else None
            soa_len = np.zeros(2 * b, np.int32) if use_soa else None
            for _ in range(3):
                E.fused_pass(a, acc, s, lp, cp, aggs, li, accum, soa=soa, soa_len=soa_len, wavefront=wavefront)
            outs.append((s, aggs, accum, lp, cp, soa, soa_len))
    finally:
        E.accel_destroy(acc)
    m, wv = outs
    assert_bits_equal(m[0], wv[0], "RNG state")
    assert_bits_equal(m[1], wv[1], "aggregators")
    assert_bits_equal(m[2][:, 4:8].copy(), wv[2][:, 4:8].copy(), "unidirectional + counts")
    np.testing.assert_allclose(wv[2][:, 0:4], m[2][:, 0:4], rtol=2e-4, atol=1e-6)
    if use_soa:
        assert np.array_equal(m[6], wv[6])
        # a vertex slot is defined while its index < length (+ slot 0 of an empty path); compare those
        parts = m[5].reshape(2, 6, 4, b, 4), wv[5].reshape(2, 6, 4, b, 4)
        for sub in range(2):
            for k in range(6):
                live = (m[6][sub * b:(sub + 1) * b] > k) | ((k == 0) & (m[6][sub * b:(sub + 1) * b] == 0))
                assert np.array_equal(parts[0][sub, k][:, live].view(np.uint32), parts[1][sub, k][:, live].view(np.uint32)), (sub, k)
    else:
        assert_paths_equal(m[3], wv[3], "light paths")
        assert_paths_equal(m[4], wv[4], "camera paths")


def test_random_number_exactly_one(O, E):
    """xorshift states >= 0xffffff80 convert to exactly 1.0: the light pick `(int)(r * count)`
    then equals `count` (out of bounds in the reference; clamped here and in the oracle), a barycentric draw of 1.0
    lands on a vertex, a roll of 1.0 never passes `<= fresnel`.  Seeds are chosen so that the FIRST draw of either stream
    of some samples is 1.0 (0x5e6cfce7, 0x4a4d939f, 0x55935ce6 are the xorshift32 pre-images of 0xffffffff, 0xffffff80,
    0xffffffc1)."""
    w, h = 24, 16
    a = arrays("bench-spheres", w, h)
    seeds = seeds_for(w * h, seed=12)
    pre = np.array([0x5E6CFCE7, 0x4A4D939F, 0x55935CE6], np.uint32)
    seeds[0:3, 0] = pre          # light pick r == 1.0
    seeds[3:6, 1] = pre          # second stream: barycentric / direction draws of 1.0
    seeds[6:9, 0] = pre
    seeds[6:9, 1] = pre[::-1]
    R = oracle_stages(O, a, StageBuffers(w * h, seeds), upto="finalize")
    assert (R.light_rays["triangle"][:3] == a.light_triangle_indices[-1]).all()  # clamped to the last light
    acc = E.accel(a.boxes, a.triangles)
    try:
        for mode in (MODE_REFERENCE, MODE_ORDERED):
            S = emu_stages(E, acc, a, StageBuffers(w * h, seeds), mode)
            assert_bits_equal(R.rand, S.rand, "RNG state")
            assert_bits_equal(R.light_rays, S.light_rays, "light rays")
            assert_bits_equal(R.camera_rays, S.camera_rays, "camera rays")
            assert_paths_equal(R.light_paths, S.light_paths, "light paths")
            assert_paths_equal(R.camera_paths, S.camera_paths, "camera paths")
            assert_bits_equal(R.aggs, S.aggs, "weight aggregators")
            assert_bits_equal(R.finalized, S.finalized, "finalized samples")
    finally:
        E.accel_destroy(acc)


@pytest.mark.parametrize("trial", [0, 2, 3])
def test_fused_pass_scrubs_nan_radiance_like_the_reference(E, trial):
    """Material tables that produce NaN / inf radiance for some samples: the reference zeroes them
    per pass with np.nan_to_num before accumulating (renderer.py:269-271); the fused finalize does it on the device."""
    w, h = 32, 24
    a = random_material_arrays(trial, w, h)
    _fused_vs_oracle_renderer(E, a, seeds_for(w * h, seed=trial))


def test_per_ray_work_counts(E):
    """The per-ray node / triangle counts the lane model (tools/lane_model.py) is built on add up to the global counters."""
    a = arrays("bench-spheres", 32, 24)
    rays = random_rays(5000, [-9, -1.9, -9], [9, 9, 9], seed=4)
    acc = E.accel(a.boxes, a.triangles)
    try:
        E.counters(reset=True)
        E.traverse(rays, acc, MODE_ORDERED)
        total = E.counters(reset=True)
        nodes, tris = E.traverse_counts(rays, acc)
    finally:
        E.accel_destroy(acc)
    assert nodes.sum() == total["node_fetches"] and tris.sum() == total["tri_fetches"]
    assert nodes.max() > 4 * nodes.mean() or tris.max() > 3 * tris.mean()  # the heavy tail the model is about
