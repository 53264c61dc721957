"""Parity at BASELINE.json's own configurations and frame sizes (round-1 tests used 64x64 / 64x48 frames).

Every test renders the configuration's scene at the configuration's resolution with the fused CUDA renderer (through
the C ABI) and with the CPU oracle in its shared-math build on the same seeds, and compares RNG state, MIS aggregators
and the unidirectional image bit for bit and the accumulated image within the float-summation-order tolerance of the
atomically splatted light image (rtol 5e-4).  Where oracle/_ref exists (the reference's own trace.metal compiled as
C++, prebuilt in the build container and shipped with the snapshot) the checker is THAT, not the restatement.

A second group renders each configuration at bench size with the reference's visiting order (REFERENCE_ORDER: unordered
DFS, right child first, trace.metal:144-176) and with the product's ordered traversal and requires IDENTICAL path
records, RNG state and aggregators: zero rays with a different hit, not "few".
"""
import numpy as np
import pytest

from clive2_b200 import runtime, scene, struct_types as T
from clive2_b200.renderer import Renderer
from oracle import oracle as oracle_mod
from oracle.oracle import OracleRenderer
from tests.common import arrays, assert_bits_equal, assert_paths_equal, rel_rmse, seeds_for

pytestmark = pytest.mark.gpu

USE_REF = oracle_mod.have_ref()  # checker = the compiled reference when it is available


@pytest.fixture(scope="module")
def dev():
    d = runtime.Device(0)
    yield d
    d.release()


def checker(a, seeds):
    r = OracleRenderer(a, seeds, shared_math=True, reference=USE_REF)
    r.oracle.use_all_cores()
    return r


def compare_with_checker(dev, a, passes, seed=20240928, rtol=5e-4):
    w, h = a.pixel_width, a.pixel_height
    b = w * h
    seeds = seeds_for(b, seed=seed)
    ref = checker(a, seeds)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds)
    for _ in range(passes):
        ref.run_sample()
        r.run_sample()
    assert_bits_equal(ref.buffer("rand", np.uint32).reshape(b, 2), r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b).reshape(b, 2), "RNG state")
    assert_bits_equal(ref.buffer("aggregators", T.WeightAggregator), r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b),
                      "aggregators of the last pass", nan_equal=True)
    assert_bits_equal(ref.unidirectional_image_buffer.reshape(b, 3), np.ascontiguousarray(r.unidirectional_image_buffer).reshape(b, 3),
                      "unidirectional image", nan_equal=True)
    assert np.array_equal(ref.summed_sample_counts, r.summed_sample_counts)
    np.testing.assert_allclose(r.summed_sample_weights, ref.summed_sample_weights, rtol=5e-5, atol=1e-6)
    np.testing.assert_allclose(r.summed_image, ref.summed_image, rtol=rtol, atol=1e-6)
    est_ref = np.nan_to_num(ref.summed_image / ref.summed_sample_weights)
    err = rel_rmse(r.estimate(), est_ref)
    assert err < 1e-4, err
    ref.close()
    r.release()
    sc.release()
    return err


def test_c1_cornell_256x256_16spp_vs_checker(dev):
    """BASELINE.json configs[0]: built-in Cornell box, 256x256, 16 spp — the whole configuration."""
    compare_with_checker(dev, arrays("empty", 256, 256), passes=16)


def test_c2_ggx_transmission_512x512_vs_checker(dev):
    """BASELINE.json configs[1] at its frame size: GGX specular + dielectric transmission, 512x512, 4 of its 64 passes."""
    a = arrays("bench-spheres", 512, 512)
    compare_with_checker(dev, a, passes=4)


@pytest.mark.parametrize("walk_mode", [runtime.WALK_AUTO, runtime.WALK_WAVEFRONT, runtime.WALK_PIPELINE, runtime.WALK_MEGAKERNEL],
                         ids=["auto", "rounds", "pipelined", "megakernel"])
def test_c3_terrain_1m_256x256_vs_checker(dev, walk_mode):
    """BASELINE.json configs[2]'s scene (1 002 528-triangle mesh, large-scene kernel flavours) at 256x256, 2 passes, under
    every driver of the walks."""
    a = scene.preset_arrays("bench-terrain-1m", 256, 256)
    before = dev.walk_mode
    try:
        dev.walk_mode = walk_mode
        compare_with_checker(dev, a, passes=2, seed=5)
    finally:
        dev.walk_mode = before


@pytest.mark.parametrize("budget", [1, 16, 64])
@pytest.mark.parametrize("which", ["c2", "c3"])
def test_suspended_searches_vs_checker(dev, which, budget):
    """k_walk_trace suspends the searches still running `walk_tail_budget` iterations after their round's queue ran dry
    (continuation records, resumed by the next round).  Budget 1 suspends every search a warp still holds at that moment,
    up to CLV_WALK_MAX_LAG times per sample; 64 is the production value.  Same bits as the checker under every budget, on
    the binary nodes (c2) and the 4-wide nodes (c3)."""
    a = arrays("bench-spheres", 512, 512) if which == "c2" else scene.preset_arrays("bench-terrain-1m", 256, 256)
    before, before_budget = dev.walk_mode, dev.knob("walk_tail_budget", -1)
    try:
        dev.walk_mode = runtime.WALK_WAVEFRONT
        dev.knob("walk_tail_budget", budget)
        compare_with_checker(dev, a, passes=2, seed=17)
    finally:
        dev.walk_mode = before
        dev.knob("walk_tail_budget", before_budget)


def test_tile_shards_under_wavefront_walks_sum_to_the_full_frame(dev):
    """Two tile shards (16x16 tiles, the multi-GPU split) of C3's scene at 256x256, each rendered by the wavefront walks with
    searches suspended early: the shards' accumulators add up to the full-frame render (counts and the unidirectional image
    bit for bit — every sample consumed its 1-GPU RNG stream and found its 1-GPU hits — the atomically splatted sums within
    float reordering)."""
    from clive2_b200 import sharding
    w = h = 256
    a = scene.preset_arrays("bench-terrain-1m", w, h)
    b = w * h
    seeds = seeds_for(b, seed=23)
    sc = scene.scene_from_arrays(a, dev)
    before, before_budget = dev.walk_mode, dev.knob("walk_tail_budget", -1)
    try:
        dev.walk_mode = runtime.WALK_WAVEFRONT
        dev.knob("walk_tail_budget", 8)
        full = Renderer(sc, seeds=seeds)
        parts = [Renderer(sc, seeds=seeds, sample_ids=sharding.tile_ids(w, h, k, 2, 16)) for k in (0, 1)]
        for _ in range(2):
            full.run_sample()
            for p in parts:
                p.run_sample()
        acc_full = full._accum()
        total = parts[0]._accum() + parts[1]._accum()
        assert np.array_equal(total[..., 7], acc_full[..., 7])
        assert_bits_equal(np.ascontiguousarray(total[..., 4:7]).reshape(b, 3), np.ascontiguousarray(acc_full[..., 4:7]).reshape(b, 3),
                          "unidirectional image", nan_equal=True)
        np.testing.assert_allclose(total[..., 0:4], acc_full[..., 0:4], rtol=5e-4, atol=1e-6)
        for r in parts + [full]:
            r.release()
    finally:
        dev.walk_mode = before
        dev.knob("walk_tail_budget", before_budget)
        sc.release()


def test_cooperative_leaf_steps_with_leaves_of_up_to_15_triangles(dev, monkeypatch):
    """The warp-cooperative leaf steps of k_walk_trace / k_connect_rays deal (ray, triangle) pairs by a 4-bit prefix sum of
    the leaf sizes.  The reference's builder stops at 8 triangles per leaf; here C2's scene is built with leaves of up to 15
    (the largest a leaf reference can encode) and run through the large-scene kernel flavours (threshold lowered by the
    test hook), wavefront walks with searches suspended early: same bits as the checker."""
    from clive2_b200 import bvh
    monkeypatch.setattr(bvh, "MAX_MEMBERS", 15)
    monkeypatch.setenv("CLV_LARGE_SCENE_BYTES", "1")
    a = scene.preset_arrays("bench-spheres", 160, 120, bvh_builder="numpy")  # not tests.common.arrays: that one caches
    leaves = a.boxes[a.boxes["right"] != 0]
    sizes = leaves["right"] - leaves["left"]
    assert sizes.max() == 15 and (sizes > 8).sum() > 100
    acc = runtime.Accel(dev, a.boxes, a.triangles)
    st = acc.stats()
    acc.release()
    assert st["large_scene_tuning"] and st["trace_uses_wide"] and st["occlusion_uses_wide"]
    before, before_budget = dev.walk_mode, dev.knob("walk_tail_budget", -1)
    try:
        dev.walk_mode = runtime.WALK_WAVEFRONT
        dev.knob("walk_tail_budget", 4)
        compare_with_checker(dev, a, passes=3, seed=29)
    finally:
        dev.walk_mode = before
        dev.knob("walk_tail_budget", before_budget)


@pytest.fixture(scope="module")
def c4_arrays():
    """BASELINE.json configs[3]'s scene (9 999 392-triangle mesh + enclosure) at a 160x90 frame (1080p / 12): built once."""
    a = scene.preset_arrays("bench-terrain-10m", 160, 90)
    assert len(a.triangles) > 9_900_000
    return a


@pytest.mark.parametrize("walk_mode,budget", [(runtime.WALK_AUTO, None), (runtime.WALK_WAVEFRONT, 0), (runtime.WALK_WAVEFRONT, 8)],
                         ids=["auto", "rounds", "rounds_suspending"])
def test_c4_terrain_10m_160x90_vs_checker(dev, c4_arrays, walk_mode, budget):
    """The north-star configuration's geometry — i.e. the kernel flavours, the 29-level tree and the 4-wide nodes the
    headline number runs on — 2 passes against the checker; `rounds` forces the wavefront walks the bench-size frame uses
    (a 160x90 pass is below the automatic mode's sample threshold)."""
    acc = runtime.Accel(dev, c4_arrays.boxes, c4_arrays.triangles)
    st = acc.stats()
    assert st["large_scene_tuning"] and st["trace_uses_wide"] and 25 <= st["binary_depth"] <= 40 and st["wide_depth"] <= 21
    acc.release()
    before, before_budget = dev.walk_mode, dev.knob("walk_tail_budget", -1)
    try:
        dev.walk_mode = walk_mode
        if budget is not None:
            dev.knob("walk_tail_budget", budget)
        compare_with_checker(dev, c4_arrays, passes=2, seed=11)
    finally:
        dev.walk_mode = before
        dev.knob("walk_tail_budget", before_budget)


@pytest.mark.parametrize("wave_mode", [runtime.WALK_WAVEFRONT, runtime.WALK_PIPELINE], ids=["rounds", "pipelined"])
@pytest.mark.parametrize("preset,w,h,passes", [("empty", 256, 256, 4), ("bench-spheres", 512, 512, 3)])
def test_wavefront_walks_at_bench_size_vs_checker_and_megakernel(dev, preset, w, h, passes, wave_mode):
    """Stages 1-4 as wavefront rounds (k_walk_init / k_walk_trace / k_walk_shade) or as the pipelined persistent kernel
    (k_walk_pipeline), forced on the cache-resident configurations where the automatic mode keeps the megakernel: same
    bits as the checker and as the megakernel."""
    a = arrays(preset, w, h)
    b = w * h
    before = dev.walk_mode
    try:
        dev.walk_mode = wave_mode
        n0 = dev.launch_count
        compare_with_checker(dev, a, passes=passes, seed=77)
        if wave_mode == runtime.WALK_WAVEFRONT:
            assert dev.launch_count - n0 > passes * 25  # init + 12 x (trace, shade) + connection + finalize per slice
        outs = []
        for mode in (runtime.WALK_MEGAKERNEL, wave_mode):
            dev.walk_mode = mode
            sc = scene.scene_from_arrays(a, dev)
            r = Renderer(sc, seeds=seeds_for(b, seed=78), path_records=True)
            for _ in range(2):
                r.run_sample()
            outs.append(dict(seeds=r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b), cam=r._fr.read(runtime.BUF_CAMERA_PATHS, T.Path, b),
                             lig=r._fr.read(runtime.BUF_LIGHT_PATHS, T.Path, b),
                             aggs=r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b), acc=r._accum().copy()))
            r.release()
            sc.release()
    finally:
        dev.walk_mode = before
    mega, wave = outs
    assert np.array_equal(mega["seeds"], wave["seeds"])
    assert_paths_equal(mega["cam"], wave["cam"], "camera Path records, megakernel vs wavefront")
    assert_paths_equal(mega["lig"], wave["lig"], "light Path records, megakernel vs wavefront")
    assert_bits_equal(mega["aggs"], wave["aggs"], "aggregators", nan_equal=True)
    assert np.array_equal(mega["acc"][..., 4:8].view(np.uint32), wave["acc"][..., 4:8].view(np.uint32))


def both_orders(dev, a, passes, seed):
    w, h = a.pixel_width, a.pixel_height
    b = w * h
    seeds = seeds_for(b, seed=seed)
    sc = scene.scene_from_arrays(a, dev)
    out = []
    for mode in (runtime.TRAVERSE_REFERENCE_ORDER, runtime.TRAVERSE_ORDERED):
        r = Renderer(sc, seeds=seeds, traverse_mode=mode, path_records=True)
        for _ in range(passes):
            r.run_sample()
        out.append(dict(seeds=r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * b),
                        cam=r._fr.read(runtime.BUF_CAMERA_PATHS, T.Path, b), lig=r._fr.read(runtime.BUF_LIGHT_PATHS, T.Path, b),
                        aggs=r._fr.read(runtime.BUF_AGGREGATORS, T.WeightAggregator, b), acc=r._accum().copy()))
        r.release()
    sc.release()
    return out


@pytest.mark.parametrize("preset,w,h,passes,native", [
    ("empty", 256, 256, 4, False),              # C1 at bench size
    ("bench-spheres", 512, 512, 2, False),      # C2 at bench size
    ("bench-terrain-1m", 1024, 1024, 1, True),  # C3 at bench size
])
def test_ordered_traversal_equals_reference_order_at_bench_size(dev, preset, w, h, passes, native):
    """Every ray of the pass — closest hits of both walks, shadow rays, projections — decides the same way under the
    product's ordered traversal as under the reference's visiting order: identical Path records (hit triangle of every
    vertex included), RNG state and aggregators.  north_star: 'bit-exact ... ray-triangle hit indices for identical rays'."""
    a = scene.preset_arrays(preset, w, h) if native else arrays(preset, w, h)
    ref_order, ordered = both_orders(dev, a, passes, seed=31)
    assert np.array_equal(ref_order["seeds"], ordered["seeds"])
    assert_paths_equal(ref_order["cam"], ordered["cam"], "camera paths, reference order vs ordered")
    assert_paths_equal(ref_order["lig"], ordered["lig"], "light paths, reference order vs ordered")
    assert_bits_equal(ref_order["aggs"], ordered["aggs"], "aggregators (every shadow ray decided identically)", nan_equal=True)
    # accumulators: unidirectional + count exact, image up to the splat's summation order
    assert np.array_equal(ref_order["acc"][..., 4:8].view(np.uint32), ordered["acc"][..., 4:8].view(np.uint32))
    np.testing.assert_allclose(ordered["acc"][..., 0:4], ref_order["acc"][..., 0:4], rtol=5e-4, atol=1e-6)


def test_c4_scene_ordered_equals_reference_order(dev, c4_arrays):
    ref_order, ordered = both_orders(dev, c4_arrays, 2, seed=41)
    assert np.array_equal(ref_order["seeds"], ordered["seeds"])
    assert_paths_equal(ref_order["cam"], ordered["cam"], "camera paths")
    assert_paths_equal(ref_order["lig"], ordered["lig"], "light paths")
    assert_bits_equal(ref_order["aggs"], ordered["aggs"], "aggregators", nan_equal=True)


def test_c2_image_against_glibc_transcription_within_noise(dev):
    """GGX / transmission scene against the canonical glibc build of the checker (different elementary functions, so
    paths diverge after the first bounce: statistical agreement).  Tolerance: 1.5 x the transcription's own seed-to-seed
    relative RMSE at equal spp against a 128-spp transcription image, plus agreement of the image means within 3 %."""
    w, h = 96, 96
    a = arrays("bench-spheres", w, h)
    b = w * h

    def cpu_est(seed, spp):
        r = OracleRenderer(a, seeds_for(b, seed=seed), shared_math=False, reference=USE_REF)
        r.oracle.use_all_cores()
        for _ in range(spp):
            r.run_sample()
        est = np.nan_to_num(r.summed_image / r.summed_sample_weights)
        r.close()
        return est

    hi = cpu_est(201, 128)
    noise = rel_rmse(cpu_est(202, 16), hi)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds_for(b, seed=203))
    for _ in range(16):
        r.run_sample()
    e16 = rel_rmse(r.estimate(), hi)
    for _ in range(112):
        r.run_sample()
    e128 = rel_rmse(r.estimate(), hi)
    assert e16 < 1.5 * noise, (e16, noise)
    assert e128 < e16
    assert abs(r.estimate().mean() - hi.mean()) / hi.mean() < 0.03
    r.release()
    sc.release()


def test_accel_cache_follows_buffer_contents_and_renderer_outlives_scene(dev):
    """a cached acceleration structure must not survive an upload into its buffers, and releasing the scene
    before a renderer that still uses its structure must be harmless (clv_accel is reference counted)."""
    a = arrays("empty", 32, 32)
    sc = scene.scene_from_arrays(a, dev)
    acc1 = dev.accel(sc.boxes, sc.triangles)
    assert dev.accel(sc.boxes, sc.triangles) is acc1
    assert not sc.triangles.resident  # the fused path never needs the reference-layout device copy
    sc.triangles.upload(a.triangles)
    acc2 = dev.accel(sc.boxes, sc.triangles)
    assert acc2 is not acc1 and acc1.handle is None
    fresh = dev.accel(a.boxes, a.triangles)  # host arrays: never cached, caller owns
    assert fresh is not acc2 and dev.accel(a.boxes, a.triangles) is not fresh
    fresh.release()
    seeds = seeds_for(32 * 32)
    r = Renderer(sc, seeds=seeds)
    keep = Renderer(sc, seeds=seeds)
    r.run_sample()
    sc.release()      # scene (and the cache's reference) gone first ...
    r.run_sample()    # ... the renderer still traverses a live structure
    keep.run_sample()
    keep.run_sample()
    assert np.array_equal(r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * 32 * 32), keep._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * 32 * 32))
    r.release()
    keep.release()
