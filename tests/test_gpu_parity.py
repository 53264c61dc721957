"""Parity tests proper: the sm_100a kernels, called through the C ABI (clive2_b200.runtime -> libclive_b200.so),
against the CPU oracle on identical seeded inputs.  Bit-exact wherever the reference's own float summation order is
kept (every stage buffer, RNG streams, hit records, the staged renderer's accumulators); stated tolerances where the
fused path replaces the sorted light-image gather by atomic splats."""
import numpy as np
import pytest

from clive2_b200 import runtime, scene, struct_types as T
from clive2_b200.renderer import Renderer
from oracle.oracle import Oracle, OracleRenderer
from tests.common import (StageBuffers, arrays, assert_bits_equal, assert_paths_equal, awkward_camera_arrays,
                          check_ordered_against_reference, fuzz_rays,
                          fuzz_soup, many_lights_arrays, oracle_stages, random_material_arrays, random_rays,
                          rel_rmse, seeds_for, zoo_arrays)

pytestmark = pytest.mark.gpu

SCENES = [("empty", 64, 64), ("bench-spheres", 64, 48)]
MODES = [runtime.TRAVERSE_REFERENCE_ORDER, runtime.TRAVERSE_ORDERED]


@pytest.fixture(scope="module")
def dev():
    d = runtime.Device(0)
    yield d
    d.release()


@pytest.fixture(scope="module")
def O():
    return Oracle(shared_math=True)


def test_library_is_the_cuda_one(dev):
    assert b"sm_100a" in dev.lib.clv_version()
    assert runtime.device_count() >= 1


@pytest.mark.parametrize("preset,w,h", SCENES)
@pytest.mark.parametrize("mode", MODES)
def test_closest_hit_bit_exact(dev, O, preset, w, h, mode):
    a = arrays(preset, w, h)
    acc = runtime.Accel(dev, a.boxes, a.triangles)
    # camera rays of the configuration + incoherent random rays
    B = oracle_stages(O, a, StageBuffers(w * h, seeds_for(w * h)), upto="rays")
    for rays in (B.camera_rays, B.light_rays, random_rays(200000, [-9, -1.9, -9], [9, 9, 9], seed=2)):
        tri, t, u, v = O.traverse(rays, a.boxes, a.triangles)
        tri2, t2, u2, v2 = dev.traverse(acc, rays, mode)
        assert np.array_equal(tri, tri2), int((tri != tri2).sum())
        hit = tri >= 0
        for x, y in ((t, t2), (u, u2), (v, v2)):
            assert np.array_equal(x[hit].view(np.uint32), y[hit].view(np.uint32))
    acc.release()


@pytest.mark.parametrize("mode", MODES)
def test_visibility_bit_exact(dev, O, mode):
    a = arrays("bench-spheres", 64, 48)
    acc = runtime.Accel(dev, a.boxes, a.triangles)
    rays = random_rays(100000, [-9, -1.9, -9], [9, 9, 9], seed=5)
    tri, t, _, _ = O.traverse(rays, a.boxes, a.triangles)
    ok = tri >= 0
    A = rays[ok].copy()
    Bv = np.zeros(len(A), T.Ray)
    Bv["origin"][:, :3] = A["origin"][:, :3] + A["direction"][:, :3] * t[ok][:, None]
    Bv["triangle"] = tri[ok]
    for b in (Bv, np.roll(Bv, 7)):
        assert np.array_equal(O.visibility(A, b, a.boxes, a.triangles), dev.visibility(acc, A, b, mode))
    acc.release()


def test_single_leaf_and_long_leaf_bvh(dev, O):
    rng = np.random.default_rng(11)
    for n_tris in (3, 40):
        tris = np.zeros(n_tris, T.Triangle)
        base = rng.uniform(-1, 1, (n_tris, 3)).astype(np.float32)
        tris["v0"][:, :3] = base
        tris["v1"][:, :3] = base + rng.normal(0, 0.5, (n_tris, 3)).astype(np.float32)
        tris["v2"][:, :3] = base + rng.normal(0, 0.5, (n_tris, 3)).astype(np.float32)
        boxes = np.zeros(1, T.Box)
        allv = np.concatenate([tris["v0"][:, :3], tris["v1"][:, :3], tris["v2"][:, :3]])
        boxes["min"][0, :3], boxes["max"][0, :3] = allv.min(0), allv.max(0)
        boxes["left"], boxes["right"] = 0, n_tris
        rays = random_rays(5000, [-3, -3, -3], [3, 3, 3], seed=n_tris)
        tri, t, _, _ = O.traverse(rays, boxes, tris)
        acc = runtime.Accel(dev, boxes, tris)
        for mode in MODES:
            tri2, t2, _, _ = dev.traverse(acc, rays, mode)
            assert np.array_equal(tri, tri2)
            assert np.array_equal(t[tri >= 0].view(np.uint32), t2[tri >= 0].view(np.uint32))
        acc.release()


def test_bad_inputs_are_errors(dev):
    boxes = np.zeros(2, T.Box)
    boxes["left"][0], boxes["right"][0] = 5, 0
    with pytest.raises(runtime.CliveError):
        runtime.Accel(dev, boxes, np.zeros(1, T.Triangle))
    with pytest.raises(runtime.CliveError):
        dev.kernel().function("no_such_kernel")
    # zero-thread dispatches are no-ops
    buf = dev.buffer(64)
    dev.kernel().function("reset_light_indices")(0, buf, buf, buf, buf, buf)
    dev.synchronize()


@pytest.mark.parametrize("preset,w,h", SCENES)
@pytest.mark.parametrize("mode", MODES)
def test_stage_kernels_bit_exact(dev, O, preset, w, h, mode):
    """The eight kernels through `device.kernel().function(name)(n, *buffers)` with the reference's argument orders."""
    a = arrays(preset, w, h)
    b = w * h
    seeds = seeds_for(b)
    R = oracle_stages(O, a, StageBuffers(b, seeds), upto="finalize")
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, fused=False, seeds=seeds, traverse_mode=mode)
    r.make_light_rays()
    r.make_camera_rays()
    assert_bits_equal(R.light_rays, r.light_ray_buffer.numpy(T.Ray), "light rays")
    assert_bits_equal(R.camera_rays, r.camera_ray_buffer.numpy(T.Ray), "camera rays")
    r.trace_light_rays()
    r.trace_camera_rays()
    assert_paths_equal(R.light_paths, r.out_light_paths.numpy(T.Path), "light paths")
    assert_paths_equal(R.camera_paths, r.out_camera_paths.numpy(T.Path), "camera paths")
    assert_bits_equal(R.rand, r.rand_buffer.numpy(np.uint32).reshape(b, 2), "RNG state")
    assert_bits_equal(R.out_camera, r.out_camera_image.numpy(np.float32).reshape(b, 4), "unidirectional estimate")
    r.join_paths()
    for name, buf in (("pix", r.out_light_indices), ("pth", r.out_light_path_indices), ("ry", r.out_light_ray_indices),
                      ("w", r.out_light_weights), ("sh", r.out_light_shade)):
        assert_bits_equal(getattr(R, name), buf.numpy(getattr(R, name).dtype), "light entries " + name)
    assert_bits_equal(R.aggs, r.weight_aggregators.numpy(T.WeightAggregator), "weight aggregators")
    assert_bits_equal(R.out_samples, r.out_samples.numpy(np.float32).reshape(b, 4), "out_samples")
    r.finalize_samples()
    assert_bits_equal(R.finalized, r.finalized_samples.numpy(np.float32).reshape(b, 4), "finalized samples")
    assert_bits_equal(R.sample_weights, r.sample_weights.numpy(np.float32), "sample weights")
    assert np.array_equal(R.sample_counts, r.sample_counts.numpy(np.uint32))
    r.release()
    sc.release()


def test_staged_renderer_equals_oracle_renderer_bit_for_bit(dev):
    """Full reference pipeline (bitonic sort, host bins, gather, numpy accumulation) over several passes: the sort
    network is deterministic, so even the light image keeps the reference's summation order."""
    w = h = 32
    a = arrays("empty", w, h)
    seeds = seeds_for(w * h)
    ref = OracleRenderer(a, seeds, shared_math=True)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, fused=False, seeds=seeds)
    for _ in range(2):
        ref.run_sample()
        r.run_sample()
    assert_bits_equal(ref.summed_image.reshape(-1, 3), r.summed_image.reshape(-1, 3), "summed_image")
    assert_bits_equal(ref.summed_sample_weights.reshape(-1, 1), r.summed_sample_weights.reshape(-1, 1), "weights")
    assert_bits_equal(ref.unidirectional_image_buffer.reshape(-1, 3), r.unidirectional_image_buffer.reshape(-1, 3), "uni")
    assert np.array_equal(ref.summed_sample_counts, r.summed_sample_counts)
    assert r.image.shape == (h, w, 3) and r.image.dtype == np.uint8
    r.release()
    sc.release()


@pytest.mark.parametrize("preset,w,h", SCENES)
def test_fused_renderer_vs_oracle(dev, preset, w, h):
    a = arrays(preset, w, h)
    b = w * h
    seeds = seeds_for(b)
    ref = OracleRenderer(a, seeds, shared_math=True)
    ref.oracle.counters_reset()
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds, path_records=True)
    dev.counters(reset=True)
    dev.counters_enable(True)
    for _ in range(3):
        ref.run_sample()
        r.run_sample()
    cnt = dev.counters()
    dev.counters_enable(False)
    # same random walks and same connection decisions => same number of reference-equivalent rays
    assert cnt["rays_ref"] == ref.oracle.counters()["rays"]
    assert 0 < cnt["rays_traced"] <= cnt["rays_ref"]
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b).reshape(b, 2), "RNG")
    assert_paths_equal(ref.buffer("camera_paths", T.Path), r._fr.read(runtime.BUF_CAMERA_PATHS, T.Path, b), "camera paths")
    assert_paths_equal(ref.buffer("light_paths", T.Path), r._fr.read(runtime.BUF_LIGHT_PATHS, T.Path, b), "light paths")
    assert_bits_equal(ref.buffer("aggregators", T.WeightAggregator), r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b), "aggregators")
    assert_bits_equal(ref.unidirectional_image_buffer.reshape(b, 3), np.ascontiguousarray(r.unidirectional_image_buffer).reshape(b, 3), "uni")
    assert np.array_equal(ref.summed_sample_counts, r.summed_sample_counts)
    # light image: atomic splats instead of sorted gather => float summation order differs
    np.testing.assert_allclose(r.summed_sample_weights, ref.summed_sample_weights, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(r.summed_image, ref.summed_image, rtol=5e-4, atol=1e-6)
    est_ref = np.nan_to_num(ref.summed_image / ref.summed_sample_weights)
    assert rel_rmse(r.estimate(), est_ref) < 1e-4
    r.release()
    sc.release()


def test_fused_reference_order_mode_gives_same_paths(dev):
    w = h = 48
    a = arrays("bench-spheres", w, h)
    seeds = seeds_for(w * h)
    sc = scene.scene_from_arrays(a, dev)
    outs = []
    for mode in MODES:
        r = Renderer(sc, seeds=seeds, traverse_mode=mode, path_records=True)
        r.run_sample()
        r.run_sample()
        outs.append((r._fr.read(runtime.BUF_CAMERA_PATHS, T.Path, w * h), r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * w * h)))
        r.release()
    assert_paths_equal(outs[0][0], outs[1][0], "camera paths, reference-order vs ordered traversal")
    assert np.array_equal(outs[0][1], outs[1][1])
    sc.release()


def test_sharded_renderers_sum_to_full_render(dev):
    w = h = 48
    a = arrays("empty", w, h)
    b = w * h
    seeds = seeds_for(b)
    sc = scene.scene_from_arrays(a, dev)
    full = Renderer(sc, seeds=seeds)
    ids = np.arange(b, dtype=np.uint32)
    stripe = (ids // 16) % 2
    parts = [Renderer(sc, seeds=seeds, sample_ids=ids[stripe == k]) for k in (0, 1)]
    for _ in range(2):
        full.run_sample()
        for p in parts:
            p.run_sample()
    acc_full = full._accum()
    total = parts[0]._accum() + parts[1]._accum()
    assert np.array_equal(total[..., 7], acc_full[..., 7])
    assert_bits_equal(np.ascontiguousarray(total[..., 4:7]).reshape(b, 3), np.ascontiguousarray(acc_full[..., 4:7]).reshape(b, 3), "uni")
    np.testing.assert_allclose(total[..., 0:4], acc_full[..., 0:4], rtol=5e-4, atol=1e-6)
    for r in parts + [full]:
        r.release()
    sc.release()


def test_end_to_end_call_with_host_buffers(dev):
    w = h = 48
    a = arrays("empty", w, h)
    b = w * h
    seeds = seeds_for(b)
    sc = scene.scene_from_arrays(a, dev)
    r1 = Renderer(sc, seeds=seeds)
    r2 = Renderer(sc, seeds=seeds_for(b, seed=1))  # will be overwritten by the uploaded seeds
    h_seeds = runtime.PinnedArray((b, 2), np.uint32)
    h_acc = runtime.PinnedArray((h, w, 8), np.float32)
    h_seeds.array[:] = seeds
    r1.run_sample()
    r2._fr.run_sample_e2e(h_seeds.ptr, h_acc.ptr)
    a1 = r1._accum()
    assert_bits_equal(np.ascontiguousarray(a1[..., 4:8]).reshape(b, 4), np.ascontiguousarray(h_acc.array[..., 4:8]).reshape(b, 4), "uni+count")
    np.testing.assert_allclose(h_acc.array[..., 0:4], a1[..., 0:4], rtol=5e-4, atol=1e-6)
    # the same call with the pass cut into slices (forced: this frame is smaller than one wave)
    before = dev.pass_slices
    dev.pass_slices = 3
    try:
        r3 = Renderer(sc, seeds=seeds_for(b, seed=2))
        h_img = runtime.PinnedArray((h, w, 3), np.uint8)
        r3._fr.run_sample_preview(h_seeds.ptr, 4.0, h_img.ptr)
        a3 = r3._accum()
        assert_bits_equal(np.ascontiguousarray(a1[..., 4:8]).reshape(b, 4), np.ascontiguousarray(a3[..., 4:8]).reshape(b, 4), "uni+count, 3 slices")
        assert np.array_equal(r1._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b), r3._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b))
        assert np.abs(h_img.array.astype(int) - r1.image.astype(int)).max() <= 2
        r3.release()
        h_img.release()
    finally:
        dev.pass_slices = before
    for x in (r1, r2):
        x.release()
    h_seeds.release()
    h_acc.release()
    sc.release()


def test_images_match_libm_transcription_within_noise(dev):
    """Equal-spp relative RMSE against the canonical (glibc) C transcription, calibrated by the transcription's own
    seed-to-seed noise, and convergence towards a higher-spp reference."""
    w = h = 40
    a = arrays("empty", w, h)
    b = w * h

    def oracle_est(seed, spp):
        r = OracleRenderer(a, seeds_for(b, seed=seed), shared_math=False)
        for _ in range(spp):
            r.run_sample()
        return np.nan_to_num(r.summed_image / r.summed_sample_weights)

    ref_hi = oracle_est(101, 192)
    noise = rel_rmse(oracle_est(102, 32), ref_hi)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds_for(b, seed=103))
    for _ in range(32):
        r.run_sample()
    e32 = rel_rmse(r.estimate(), ref_hi)
    for _ in range(224):
        r.run_sample()
    e256 = rel_rmse(r.estimate(), ref_hi)
    # tolerance: 1.5x the transcription's own 32-spp noise against the same reference
    assert e32 < 1.5 * noise, (e32, noise)
    assert e256 < e32
    assert abs(r.estimate().mean() - ref_hi.mean()) / ref_hi.mean() < 0.05
    r.release()
    sc.release()


def test_full_size_properties(dev):
    """BASELINE config 2 size (512x512 glossy spheres): properties that need no oracle."""
    a = arrays("bench-spheres", 512, 512)
    b = 512 * 512
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seed=7)
    n0 = dev.launch_count
    for _ in range(4):
        r.run_sample()
    dev.synchronize()
    # walks + three connection kernels per slice of the samples, one finalize kernel per pass
    per_pass, rem = divmod(dev.launch_count - n0, 4)
    assert dev.walk_mode == runtime.WALK_AUTO  # cache-resident scene: the megakernel walks
    assert rem == 0 and (per_pass - 1) % (len(runtime.FusedRenderer.KERNELS) - 1) == 0 and 5 <= per_pass <= 33
    assert (r.summed_sample_counts == 4).all()
    img = r.summed_image
    assert np.isfinite(img).all() and (img >= 0).all()
    assert (r.summed_sample_weights > 0).mean() > 0.9
    assert r.image.dtype == np.uint8
    # linearity of accumulation: a second renderer with the same seeds stepped 2+2 passes gives identical streams
    seeds_after = r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b)
    r2 = Renderer(sc, seed=7)
    for _ in range(4):
        r2.run_sample()
    assert np.array_equal(seeds_after, r2._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b))
    assert np.array_equal(r.summed_sample_counts, r2.summed_sample_counts)
    np.testing.assert_allclose(r2.summed_image, img, rtol=5e-4, atol=1e-6)
    for x in (r, r2):
        x.release()
    sc.release()


def test_movie_frame_renders_and_differs_between_frames(dev, tmp_path):
    """BASELINE config 5 in miniature: two orbit frames of the glossy-spheres scene through the headless driver."""
    from clive2_b200.movie import render_frame, write_image
    imgs = [render_frame(dev, "bench-spheres", 64, 36, 4, frame_idx=f, total_frames=120, seed=3) for f in (0, 30)]
    for img in imgs:
        assert img.shape == (36, 64, 3) and img.dtype == np.uint8 and img.max() > 0
    assert (imgs[0] != imgs[1]).mean() > 0.2  # the camera moved a quarter turn
    again = render_frame(dev, "bench-spheres", 64, 36, 4, frame_idx=0, total_frames=120, seed=3)
    assert np.abs(again.astype(int) - imgs[0].astype(int)).max() <= 2  # same seeds => same image (atomic order aside)
    assert write_image(str(tmp_path / "f.png"), imgs[0])


def test_device_tone_map_matches_host_tone_map(dev):
    w, h = 96, 64
    a = arrays("bench-spheres", w, h)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seed=5)
    for _ in range(6):
        r.run_sample()
    host = r.image.astype(int)
    devimg = r.image_device().astype(int)
    assert devimg.shape == (h, w, 3)
    assert np.abs(devimg - host).max() <= 1 and (devimg != host).mean() < 0.01
    # preview call = pass + tone map + download
    pin = runtime.PinnedArray((h, w, 3), np.uint8)
    r.run_sample_preview(None, pin)
    assert np.abs(pin.array.astype(int) - r.image.astype(int)).max() <= 1
    pin.release()
    r.release()
    sc.release()


def test_accel_cache_does_not_survive_its_buffers(dev, O):
    """The allocator hands the addresses of a released scene to the next one; a BVH cached for the old buffers must
    not be served for the new scene."""
    first = None
    for preset in ("empty", "bench-spheres", "empty"):
        a = arrays(preset, 32, 32)
        sc = scene.scene_from_arrays(a, dev)
        acc = dev.accel(sc.boxes, sc.triangles)
        assert acc.n_triangles == len(a.triangles)
        rays = random_rays(20000, [-9, -1.9, -9], [9, 9, 9], seed=9)
        tri, _, _, _ = O.traverse(rays, a.boxes, a.triangles)
        tri2, _, _, _ = dev.traverse(acc, rays)
        assert np.array_equal(tri, tri2)
        first = first or acc
        sc.release()
        assert not dev._accels


def test_renderer_argument_errors(dev):
    a = arrays("empty", 16, 16)
    sc = scene.scene_from_arrays(a, dev)
    acc = dev.accel(sc.boxes, sc.triangles)
    seeds = seeds_for(256)
    with pytest.raises(runtime.CliveError):  # sample id outside the frame
        runtime.FusedRenderer(dev, acc, a, seeds, sample_ids=np.array([0, 999999], np.uint32))
    import copy
    dark = copy.copy(a)
    dark.light_triangles = a.light_triangles[:0]
    dark.light_surface_areas = a.light_surface_areas[:0]
    dark.light_triangle_indices = a.light_triangle_indices[:0]
    with pytest.raises(runtime.CliveError):  # the light-ray generator needs at least one emitter
        runtime.FusedRenderer(dev, acc, dark, seeds)
    r = Renderer(sc, seeds=seeds)
    with pytest.raises(runtime.CliveError):  # wrong readback size
        r._fr.read(runtime.BUF_SEEDS, np.uint32, 7)
    r.release()
    sc.release()


def test_tiled_sample_order_gives_the_same_render(dev):
    from clive2_b200 import sharding
    w, h = 48, 40
    a = arrays("bench-spheres", w, h)
    seeds = seeds_for(w * h)
    sc = scene.scene_from_arrays(a, dev)
    plain = Renderer(sc, seeds=seeds)
    tiled = Renderer(sc, seeds=seeds, sample_ids=sharding.tiled_order(w, h, 8, 4))
    for _ in range(2):
        plain.run_sample()
        tiled.run_sample()
    p, t = plain._accum(), tiled._accum()
    assert np.array_equal(p[..., 4:8].view(np.uint32), t[..., 4:8].view(np.uint32))
    np.testing.assert_allclose(t[..., 0:4], p[..., 0:4], rtol=5e-4, atol=1e-6)
    for x in (plain, tiled):
        x.release()
    sc.release()


def test_ragged_frame_size_partial_warps(dev):
    """37 x 25 = 925 samples: neither a multiple of the warp nor of the block size; also a shard of 77 ids."""
    w, h = 37, 25
    a = arrays("bench-spheres", w, h)
    b = w * h
    seeds = seeds_for(b)
    ref = OracleRenderer(a, seeds, shared_math=True)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds, path_records=True)
    for _ in range(2):
        ref.run_sample()
        r.run_sample()
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b).reshape(b, 2), "RNG")
    assert_paths_equal(ref.buffer("light_paths", T.Path), r._fr.read(runtime.BUF_LIGHT_PATHS, T.Path, b), "light paths")
    assert_bits_equal(ref.buffer("aggregators", T.WeightAggregator), r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b), "aggregators")
    np.testing.assert_allclose(r.summed_image, ref.summed_image, rtol=5e-4, atol=1e-6)
    np.testing.assert_allclose(r.summed_sample_weights, ref.summed_sample_weights, rtol=5e-5, atol=1e-6)
    ids = np.arange(5, b, 12, dtype=np.uint32)
    part = Renderer(sc, seeds=seeds, sample_ids=ids, path_records=True)
    part.run_sample()
    got = part._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * len(ids)).reshape(-1, 2)
    one = OracleRenderer(a, seeds, shared_math=True)
    one.run_sample()
    assert_bits_equal(one.buffer("rand", np.uint32).reshape(b, 2)[ids], got, "shard RNG")
    assert (part.summed_sample_counts.reshape(-1)[ids] == 1).all() and part.summed_sample_counts.sum() == len(ids)
    for x in (r, part):
        x.release()
    sc.release()


def test_every_material_branch_bit_exact(dev):
    """Mirror / coated / smooth + rough dielectric / rough-normal diffuse materials through the fused renderer."""
    w, h = 64, 40
    a = zoo_arrays(w, h)
    b = w * h
    seeds = seeds_for(b, seed=77)
    ref = OracleRenderer(a, seeds, shared_math=True)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds, path_records=True)
    for _ in range(3):
        ref.run_sample()
        r.run_sample()
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b).reshape(b, 2), "RNG")
    assert_paths_equal(ref.buffer("camera_paths", T.Path), r._fr.read(runtime.BUF_CAMERA_PATHS, T.Path, b), "camera paths")
    assert_paths_equal(ref.buffer("light_paths", T.Path), r._fr.read(runtime.BUF_LIGHT_PATHS, T.Path, b), "light paths")
    assert_bits_equal(ref.buffer("aggregators", T.WeightAggregator), r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b), "aggregators")
    assert_bits_equal(ref.unidirectional_image_buffer.reshape(b, 3), np.ascontiguousarray(r.unidirectional_image_buffer).reshape(b, 3), "uni")
    np.testing.assert_allclose(r.summed_sample_weights, ref.summed_sample_weights, rtol=5e-5, atol=1e-6)
    np.testing.assert_allclose(r.summed_image, ref.summed_image, rtol=5e-4, atol=1e-6)
    r.release()
    sc.release()


@pytest.mark.parametrize("preset,w,h", SCENES)
def test_default_soa_vertex_records_vs_oracle(dev, preset, w, h):
    """The default fused renderer keeps path vertices in its compact SoA records: RNG streams, aggregators and the
    unidirectional image must equal the oracle bit for bit, exactly as with the reference's Path records."""
    a = arrays(preset, w, h)
    b = w * h
    seeds = seeds_for(b, seed=31)
    ref = OracleRenderer(a, seeds, shared_math=True)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds)
    for _ in range(3):
        ref.run_sample()
        r.run_sample()
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b).reshape(b, 2), "RNG")
    assert_bits_equal(ref.buffer("aggregators", T.WeightAggregator), r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b), "aggregators")
    assert_bits_equal(ref.unidirectional_image_buffer.reshape(b, 3), np.ascontiguousarray(r.unidirectional_image_buffer).reshape(b, 3), "uni")
    assert np.array_equal(ref.summed_sample_counts, r.summed_sample_counts)
    np.testing.assert_allclose(r.summed_sample_weights, ref.summed_sample_weights, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(r.summed_image, ref.summed_image, rtol=5e-4, atol=1e-6)
    with pytest.raises(runtime.CliveError):  # Path records are not kept unless asked for
        r._fr.read(runtime.BUF_CAMERA_PATHS, T.Path, b)
    r.release()
    sc.release()


def test_light_image_splat_equals_sort_and_gather(dev):
    """clv_light_image_splat (atomics on the unsorted entries) vs the reference's bitonic sort + host bins + gather."""
    w, h = 48, 40
    a = arrays("empty", w, h)
    b = w * h
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, fused=False, seeds=seeds_for(b, seed=12))
    r.make_light_rays(); r.make_camera_rays(); r.trace_light_rays(); r.trace_camera_rays(); r.join_paths()
    r.finalize_samples()
    n = len(r.out_light_indices.numpy(np.int32))
    splat_image = dev.buffer(b * 16)
    splat_weights = dev.buffer(r.sample_weights.numpy(np.float32))  # finalize's weights, as gather adds onto them
    dev.light_image_splat(n, b, r.out_light_paths, sc.materials, r.out_light_indices, r.out_light_path_indices,
                          r.out_light_ray_indices, r.out_light_weights, r.out_light_shade, splat_image, splat_weights)
    r.gather_light_image()  # sorts the entries in place, bins on the host, gathers
    ref_img = r.out_light_image.numpy(np.float32).reshape(b, 4)
    ref_w = r.sample_weights.numpy(np.float32)
    got_img = splat_image.numpy(np.float32).reshape(b, 4)
    assert (ref_img[:, 3] == 1).all() and (got_img[:, 3] == 1).all()
    assert (ref_img[:, :3] > 0).any()
    np.testing.assert_allclose(got_img[:, :3], ref_img[:, :3], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(splat_weights.numpy(np.float32), ref_w, rtol=1e-5, atol=1e-7)
    r.release()
    sc.release()


def test_overlapped_movie_loop_equals_frame_by_frame(dev):
    """render_movie (host prep / tone map + encode on worker threads, pooled device memory reused between frames)
    gives the images render_frame gives."""
    from clive2_b200.movie import render_frame, render_movie
    frames = [0, 17, 45]
    got = {}
    render_movie(dev, "bench-spheres", 64, 36, 3, frames, 120, seed=5, sink=lambda f, img: got.__setitem__(f, img))
    assert sorted(got) == frames
    for f in frames:
        assert np.array_equal(got[f], render_frame(dev, "bench-spheres", 64, 36, 3, frame_idx=f, total_frames=120, seed=5))


def test_pass_slices_do_not_change_the_result(dev):
    """The fused pass split into 1, 3 or 5 independently launched slices of the samples (ragged last slice, a slice
    count that leaves some slices empty at this size): same RNG streams, same aggregators, same unidirectional image."""
    w, h = 50, 37  # 1850 samples: not a multiple of the block size
    a = arrays("bench-spheres", w, h)
    sc = scene.scene_from_arrays(a, dev)
    before = dev.pass_slices
    results = []
    try:
        for slices in (1, 3, 5, 8):
            dev.pass_slices = slices
            assert dev.pass_slices == slices
            r = Renderer(sc, seed=21)
            for _ in range(3):
                r.run_sample()
            results.append((r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * w * h), r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, w * h),
                            np.array(r.unidirectional_image_buffer), np.array(r.summed_image), np.array(r.summed_sample_counts)))
            r.release()
    finally:
        dev.pass_slices = before
    for other in results[1:]:
        assert np.array_equal(other[0], results[0][0])
        assert_bits_equal(results[0][1], other[1], "aggregators")
        assert np.array_equal(other[2].view(np.uint32), results[0][2].view(np.uint32))
        assert np.array_equal(other[4], results[0][4])
        np.testing.assert_allclose(other[3], results[0][3], rtol=5e-4, atol=1e-6)
    sc.release()


def test_large_scene_kernel_flavours_vs_oracle(dev):
    """A scene whose traversal records exceed the cache-resident size (1 M-triangle terrain, BASELINE config 3's
    geometry) runs the high-occupancy flavours of the walk and occlusion kernels (32 / 48 registers per thread): same
    bits as the oracle, as with the default flavours."""
    w, h = 64, 40
    a = scene.preset_arrays("bench-terrain-1m", w, h)  # native BVH builder (byte-identical to the numpy one, tests/test_native_bvh.py)
    b = w * h
    seeds = seeds_for(b, seed=77)
    sc = scene.scene_from_arrays(a, dev)
    assert dev.accel(sc.boxes, sc.triangles).stats()["large_scene_tuning"]
    ref = OracleRenderer(a, seeds, shared_math=True)
    r = Renderer(sc, seeds=seeds)
    for _ in range(2):
        ref.run_sample()
        r.run_sample()
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b).reshape(b, 2), "RNG")
    assert_bits_equal(ref.buffer("aggregators", T.WeightAggregator), r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b), "aggregators")
    assert_bits_equal(ref.unidirectional_image_buffer.reshape(b, 3), np.ascontiguousarray(r.unidirectional_image_buffer).reshape(b, 3), "uni")
    np.testing.assert_allclose(r.summed_image, ref.summed_image, rtol=5e-4, atol=1e-6)
    r.release()
    sc.release()
    small = scene.scene_from_arrays(arrays("empty", 16, 16), dev)
    assert not dev.accel(small.boxes, small.triangles).stats()["large_scene_tuning"]
    small.release()


@pytest.mark.parametrize("seed", [1, 2])
def test_fuzzed_soups_and_degenerate_rays(dev, O, seed):
    """The kernels on awkward geometry and rays (tests/common.py fuzz_soup / fuzz_rays: exact duplicates, coplanar
    overlaps, zero-area triangles, flat boxes; axis-parallel rays with inv_direction = inf, origins on vertices and in
    triangle planes): REFERENCE_ORDER equals the oracle bit for bit, ORDERED leaves it only where the reference's own
    hit is a rounding artefact of its unguarded Moeller-Trumbore test."""
    boxes, tris = fuzz_soup(seed)
    rays = fuzz_rays(tris, 200000, seed + 100)
    tri, t, u, v = O.traverse(rays, boxes, tris)
    acc = runtime.Accel(dev, boxes, tris)
    tri_r, t_r, u_r, v_r = dev.traverse(acc, rays, runtime.TRAVERSE_REFERENCE_ORDER)
    tri_o, t_o, _, _ = dev.traverse(acc, rays, runtime.TRAVERSE_ORDERED)
    assert np.array_equal(tri, tri_r)
    hit = tri >= 0
    for x, y in ((t, t_r), (u, u_r), (v, v_r)):
        assert np.array_equal(x[hit].view(np.uint32), y[hit].view(np.uint32))
    check_ordered_against_reference(rays, tris, tri, t, tri_o, t_o)
    ok = (tri >= 0) & np.isfinite(t)
    A = rays[ok].copy()
    B = np.zeros(len(A), T.Ray)
    B["origin"][:, :3] = A["origin"][:, :3] + A["direction"][:, :3] * t[ok][:, None]
    B["triangle"] = tri[ok]
    for b in (B, np.roll(B, 7)):
        ref = O.visibility(A, b, boxes, tris)
        assert np.array_equal(ref, dev.visibility(acc, A, b, runtime.TRAVERSE_REFERENCE_ORDER))
        assert (ref != dev.visibility(acc, A, b, runtime.TRAVERSE_ORDERED)).mean() < 1e-3
    acc.release()


@pytest.mark.parametrize("which", [0, 2, 3])
def test_awkward_cameras_fused_vs_oracle(dev, which):
    """Camera far outside the enclosure (empty camera paths), straight down next to the light (projections beyond the
    last pixel), centimetres from a wall: the fused renderer against the oracle renderer over three passes."""
    w, h = 48, 36
    a = awkward_camera_arrays(which, w, h)
    b = w * h
    seeds = seeds_for(b, seed=9)
    ref = OracleRenderer(a, seeds, shared_math=True)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds)
    for _ in range(3):
        ref.run_sample()
        r.run_sample()
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b).reshape(b, 2), "RNG")
    assert_bits_equal(ref.buffer("aggregators", T.WeightAggregator), r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b), "aggregators")
    assert_bits_equal(ref.unidirectional_image_buffer.reshape(b, 3), np.ascontiguousarray(r.unidirectional_image_buffer).reshape(b, 3), "uni")
    assert np.array_equal(ref.summed_sample_counts, r.summed_sample_counts)
    np.testing.assert_allclose(r.summed_sample_weights, ref.summed_sample_weights, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(r.summed_image, ref.summed_image, rtol=5e-4, atol=1e-6)
    r.release()
    sc.release()


@pytest.mark.parametrize("trial", [0, 2, 3, 5])
@pytest.mark.parametrize("mode", MODES)
def test_random_material_tables_stage_kernels(dev, O, trial, mode):
    """Random material tables from extreme values (alpha 0 ... 3, ior 0.7 ... 10, every type on every surface), some of
    which produce NaN radiance: every stage buffer equals the oracle, NaNs compared as NaNs (their payload bits differ
    between SSE and the GPU and never steer a path)."""
    w, h = 40, 30
    a = random_material_arrays(trial, w, h)
    b = w * h
    seeds = seeds_for(b, seed=trial)
    R = oracle_stages(O, a, StageBuffers(b, seeds), upto="finalize")
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, fused=False, seeds=seeds, traverse_mode=mode)
    r.make_light_rays(); r.make_camera_rays(); r.trace_light_rays(); r.trace_camera_rays()
    lp, cp = r.out_light_paths.numpy(T.Path), r.out_camera_paths.numpy(T.Path)
    for ref, got, what in ((R.light_paths, lp, "light paths"), (R.camera_paths, cp, "camera paths")):
        assert np.array_equal(ref["length"], got["length"]), what
        for k in range(6):
            m = ref["length"] > k
            if m.any():
                assert_bits_equal(ref["rays"][m, k], got["rays"][m, k], f"{what} vertex {k}", nan_equal=True)
    assert_bits_equal(R.rand, r.rand_buffer.numpy(np.uint32).reshape(b, 2), "RNG state")
    r.join_paths()
    for name, buf in (("pix", r.out_light_indices), ("pth", r.out_light_path_indices), ("ry", r.out_light_ray_indices),
                      ("w", r.out_light_weights), ("sh", r.out_light_shade)):
        assert_bits_equal(getattr(R, name), buf.numpy(getattr(R, name).dtype), "light entries " + name, nan_equal=True)
    assert_bits_equal(R.aggs, r.weight_aggregators.numpy(T.WeightAggregator), "weight aggregators", nan_equal=True)
    r.finalize_samples()
    assert_bits_equal(R.finalized, r.finalized_samples.numpy(np.float32).reshape(b, 4), "finalized samples", nan_equal=True)
    r.release()
    sc.release()


@pytest.mark.parametrize("w,h", [(1, 1), (3, 2), (33, 1), (1, 65)])
def test_tiny_frames(dev, w, h):
    """One pixel, one row, one column: the 3x3 neighbourhoods of the finalize stage at every border, warps with a
    single live lane, a pass that is smaller than one block.  Fused renderer and stage kernels against the oracle."""
    a = arrays("empty", w, h)
    b = w * h
    seeds = seeds_for(b, seed=3)
    ref = OracleRenderer(a, seeds, shared_math=True)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds)
    staged = Renderer(sc, fused=False, seeds=seeds)
    for _ in range(3):
        ref.run_sample()
        r.run_sample()
        staged.run_sample()
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b).reshape(b, 2), "RNG")
    assert_bits_equal(ref.buffer("aggregators", T.WeightAggregator), r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b), "aggregators")
    np.testing.assert_allclose(r.summed_image, ref.summed_image, rtol=5e-4, atol=1e-6)
    np.testing.assert_allclose(r.summed_sample_weights, ref.summed_sample_weights, rtol=2e-5, atol=1e-6)
    assert_bits_equal(ref.summed_image.reshape(-1, 3), staged.summed_image.reshape(-1, 3), "staged summed_image")
    assert_bits_equal(ref.summed_sample_weights.reshape(-1, 1), staged.summed_sample_weights.reshape(-1, 1), "staged weights")
    assert r.image.shape == (h, w, 3)
    for x in (r, staged):
        x.release()
    sc.release()


def test_reset_and_sharded_host_seed_upload(dev):
    """`reset(seeds)` restarts a render exactly (same streams, same accumulators), and the end-to-end call on a renderer
    that owns a shard of the sample ids takes the rows of ITS samples out of the whole-frame host seed buffer."""
    w, h = 40, 28
    a = arrays("bench-spheres", w, h)
    b = w * h
    seeds = seeds_for(b, seed=41)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds)
    for _ in range(2):
        r.run_sample()
    first = (r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b), np.array(r.unidirectional_image_buffer), np.array(r.summed_image))
    r._fr.reset(seeds)
    r._cache = None
    assert r.summed_image.sum() == 0 and r.summed_sample_counts.sum() == 0
    for _ in range(2):
        r.run_sample()
    assert np.array_equal(first[0], r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b))
    assert np.array_equal(first[1].view(np.uint32), np.array(r.unidirectional_image_buffer).view(np.uint32))
    np.testing.assert_allclose(r.summed_image, first[2], rtol=5e-4, atol=1e-6)
    # a shard (every third id) driven through the host-buffer call
    ids = np.arange(1, b, 3, dtype=np.uint32)
    part = Renderer(sc, seeds=seeds_for(b, seed=42), sample_ids=ids)  # wrong seeds on purpose: the call replaces them
    h_seeds = runtime.PinnedArray((b, 2), np.uint32)
    h_seeds.array[:] = seeds
    h_acc = runtime.PinnedArray((h, w, 8), np.float32)
    part._fr.run_sample_e2e(h_seeds.ptr, h_acc.ptr)
    full = Renderer(sc, seeds=seeds)
    full.run_sample()
    after = full._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b).reshape(b, 2)
    assert np.array_equal(after[ids], part._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * len(ids)).reshape(-1, 2))
    counts = h_acc.array[..., 7].reshape(-1)
    assert (counts[ids] == 1).all() and counts.sum() == len(ids)
    uni_full = np.ascontiguousarray(full._accum()[..., 4:7]).reshape(b, 3)
    assert np.array_equal(uni_full[ids].view(np.uint32), np.ascontiguousarray(h_acc.array[..., 4:7]).reshape(b, 3)[ids].view(np.uint32))
    for x in (r, part, full):
        x.release()
    h_seeds.release()
    h_acc.release()
    sc.release()


@pytest.mark.parametrize("n_extra", [0, 9])
def test_light_counts_fused_vs_oracle(dev, n_extra):
    """One light and eleven lights (nine of them faces of a mesh): fused renderer against the oracle renderer."""
    w, h = 48, 36
    a = many_lights_arrays(n_extra, w, h)
    b = w * h
    seeds = seeds_for(b, seed=8 + n_extra)
    ref = OracleRenderer(a, seeds, shared_math=True)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds)
    for _ in range(3):
        ref.run_sample()
        r.run_sample()
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b).reshape(b, 2), "RNG")
    assert_bits_equal(ref.buffer("aggregators", T.WeightAggregator), r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b), "aggregators")
    assert_bits_equal(ref.unidirectional_image_buffer.reshape(b, 3), np.ascontiguousarray(r.unidirectional_image_buffer).reshape(b, 3), "uni")
    np.testing.assert_allclose(r.summed_image, ref.summed_image, rtol=5e-4, atol=1e-6)
    r.release()
    sc.release()
