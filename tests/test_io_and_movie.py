""""Next" rows of SURVEY.md §8f: OBJ/PLY readers (the reference uses objloader/plyfile, load.py:76-95) and the
frame-parallel movie driver's host logic."""
import numpy as np
import pytest

from clive2_b200 import load, meshgen, scene
from clive2_b200.movie import frames_for_rank


def test_obj_roundtrip_and_negative_indices(tmp_path):
    v, f = meshgen.icosphere(1, 1.5)
    p = tmp_path / "m.obj"
    load.write_obj(str(p), v, f)
    v2, f2 = load.read_obj(str(p))
    assert v2.dtype == np.float64 and np.array_equal(f2, f)
    np.testing.assert_allclose(v2, v, rtol=1e-8)
    q = tmp_path / "neg.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf -3/1/1 -2/1/1 -1/1/1\n")
    _, fneg = load.read_obj(str(q))
    assert fneg.tolist() == [[0, 1, 2]]
    bad = tmp_path / "quad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nf 1 2 3 4\n")
    with pytest.raises(ValueError):
        load.read_obj(str(bad))


def test_ply_binary_and_ascii(tmp_path):
    v, f = meshgen.terrain(6, 5, seed=2)
    p = tmp_path / "m.ply"
    load.write_ply(str(p), v, f)
    v2, f2 = load.read_ply(str(p))
    assert v2.dtype == np.float32 and np.array_equal(v2, v.astype(np.float32)) and np.array_equal(f2, f)
    a = tmp_path / "a.ply"
    a.write_text("ply\nformat ascii 1.0\ncomment x\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    v3, f3 = load.read_ply(str(a))
    assert v3.shape == (3, 3) and f3.tolist() == [[0, 1, 2]]
    with pytest.raises(ValueError):
        load.read_ply(str(tmp_path / "m.obj")) if (tmp_path / "m.obj").exists() else (_ for _ in ()).throw(ValueError())


def test_file_scene_equals_in_memory_scene(tmp_path):
    """A mesh loaded from OBJ / PLY gives the same kernel arrays as the same mesh passed in memory (scene.py:49-66)."""
    v, f = meshgen.icosphere(2, 1.0)
    v = v + np.random.default_rng(3).normal(0, 1e-3, v.shape)
    obj, ply = tmp_path / "s.obj", tmp_path / "s.ply"
    load.write_obj(str(obj), v, f)
    load.write_ply(str(ply), v, f)
    kw = dict(pixel_width=32, pixel_height=32, cam_center=np.array([0, 1.5, 6]), cam_direction=np.array([0, 0, -1]))
    a_mem = scene.build_scene_arrays(file_specs=[{"mesh": (load.read_obj(str(obj))[0], f), "material": 5}], **kw)
    a_obj = scene.build_scene_arrays(file_specs=[{"file_path": str(obj), "material": 5}], **kw)
    assert np.array_equal(a_mem.triangles.view(np.uint8), a_obj.triangles.view(np.uint8))
    assert np.array_equal(a_mem.boxes.view(np.uint8), a_obj.boxes.view(np.uint8))
    a_ply = scene.build_scene_arrays(file_specs=[{"file_path": str(ply), "material": 5}], **kw)
    a_m32 = scene.build_scene_arrays(file_specs=[{"mesh": (v.astype(np.float32), f), "material": 5}], **kw)
    assert np.array_equal(a_ply.triangles.view(np.uint8), a_m32.triangles.view(np.uint8))


def test_frames_are_dealt_round_robin():
    world = 8
    parts = [frames_for_rank(0, 120, r, world) for r in range(world)]
    assert sorted(sum(parts, [])) == list(range(120)) and all(len(p) == 15 for p in parts)
    assert frames_for_rank(100, 120, 0, 8) == [100, 108, 116]  # --start-frame resumes (movie.py:19,22-29)


def test_orbit_camera_matches_reference_formula():
    c, d = scene.orbit_camera(30, 120)
    assert np.allclose(c, [7.5, 1.5, 0.0], atol=1e-12) and np.allclose(d, [-1.0, 0.0, 0.0], atol=1e-12)


def test_render_movie_pipeline_keeps_frame_order_and_surfaces_writer_errors(monkeypatch):
    """Host logic of the overlapped movie loop (no GPU: the two stages are stubbed)."""
    import pytest
    from clive2_b200 import movie
    prepared = []
    monkeypatch.setattr(movie, "frame_arrays", lambda preset, w, h, f, total: (prepared.append(f), {"frame": f})[1])
    monkeypatch.setattr(movie, "render_arrays",
                        lambda dev, arrays, samples, seed=None, raw=False: np.full((2, 2, 3), arrays["frame"], np.uint8))
    monkeypatch.setattr(movie, "image_from_accumulators", lambda a: a)
    for overlap in (False, True):
        got = []
        prepared.clear()
        movie.render_movie(None, "x", 2, 2, 1, [3, 11, 19], 120, sink=lambda f, img: got.append((f, int(img[0, 0, 0]))),
                           overlap=overlap)
        assert got == [(3, 3), (11, 11), (19, 19)] and prepared == [3, 11, 19]
    movie.render_movie(None, "x", 2, 2, 1, [], 120, sink=None)  # nothing to do

    def bad_sink(f, img):
        raise OSError("disk full")
    with pytest.raises(OSError):
        movie.render_movie(None, "x", 2, 2, 1, [0, 1], 120, sink=bad_sink)


def test_obj_relative_indices_interleaved_blocks_and_slashes(tmp_path):
    """Negative OBJ indices count back from the vertices defined so far; `v/vt/vn` corners, tabs, comments, other records."""
    import pytest
    from clive2_b200.load import read_obj
    p = tmp_path / "blocks.obj"
    p.write_text("# two blocks\n"
                 "v 0 0 0\nv 1 0 0\nv 0 1 0\n"
                 "vn 0 0 1\nvt 0 0\n"
                 "f -3/1/1 -2/1/1 -1/1/1\n"
                 "v\t0 0 1\nv 1 0 1\n  v 0 1 1\n"
                 "f -3//1 -2//1 -1//1\n"
                 "f 1 5 6\n")
    v, f = read_obj(str(p))
    assert v.shape == (6, 3) and f.tolist() == [[0, 1, 2], [3, 4, 5], [0, 4, 5]]
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n")
    with pytest.raises(ValueError):
        read_obj(str(bad))
    quad = tmp_path / "quad.obj"
    quad.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n")
    with pytest.raises(ValueError):
        read_obj(str(quad))


def test_passes_are_dealt_to_ranks_completely():
    from clive2_b200.movie import passes_for_rank
    for samples in (0, 1, 7, 8, 256, 1001):
        for world in (1, 2, 3, 8):
            shares = [passes_for_rank(samples, r, world) for r in range(world)]
            assert sum(shares) == samples and max(shares) - min(shares) <= 1 and shares == sorted(shares, reverse=True)
