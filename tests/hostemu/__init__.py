"""TEST INFRASTRUCTURE: the per-sample kernel bodies of clive2_b200/csrc/clv_device.cuh compiled for the host.

Lets the CPU-only test tier diff the *kernel source* against the oracle (bit for bit, shared-math build) before any
GPU time is spent.  Never imported by clive2_b200; the `-m gpu` tests run the same comparisons on the CUDA library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_build", "libhostemu.so")
SOURCES = [os.path.join(HERE, "hostemu.cpp")] + [os.path.join(REPO, "clive2_b200", "csrc", f)
                                                 for f in ("clv_device.cuh", "clv_accel_build.h", "clv_math.h")]
MODE_REFERENCE, MODE_ORDERED = 0, 1


def build(force=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in SOURCES):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    extra = os.environ.get("CLV_HOSTEMU_EXTRA", "").split()  # e.g. -DCLV_BINARY_NODE_ORDER=1: check a build option on the host
    subprocess.run([gxx, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                    "-fvisibility=hidden", "-shared", "-Wno-unknown-pragmas"] + extra + ["-o", LIB, SOURCES[0]], check=True)
    return LIB


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class HostEmu:
    def __init__(self):
        self.lib = C.CDLL(build())
        self.lib.emu_accel_create.restype = C.c_void_p

    def accel(self, boxes, triangles):
        boxes, triangles = np.ascontiguousarray(boxes), np.ascontiguousarray(triangles)
        h = self.lib.emu_accel_create(_p(boxes), C.c_int32(len(boxes)), _p(triangles), C.c_int32(len(triangles)))
        if not h:
            raise ValueError("bad BVH arrays")
        return C.c_void_p(h)

    def accel_destroy(self, h):
        self.lib.emu_accel_destroy(h)

    def counters(self, reset=True):
        out = (C.c_ulonglong * 5)()
        self.lib.emu_counters(out, C.c_int(int(reset)))
        return dict(zip(("rays_ref", "rays_traced", "node_fetches", "tri_fetches", "vertices"), (int(x) for x in out)))

    def generate_light_rays(self, n, lt, areas, lti, mats, rand, out, counts):
        cnt = np.ascontiguousarray(counts, dtype=np.int32).reshape(1)
        self.lib.emu_generate_light_rays(C.c_uint32(n), _p(lt), _p(areas), _p(lti), _p(mats), _p(rand), _p(out), _p(cnt))

    def generate_camera_rays(self, n, camera, rand, indices, out):
        self.lib.emu_generate_camera_rays(C.c_uint32(n), _p(camera), _p(rand), _p(indices), _p(out))

    def generate_paths(self, n, rays, accel, mats, rand, out, paths, debug, mode=MODE_ORDERED):
        self.lib.emu_generate_paths(C.c_uint32(n), _p(rays), accel, _p(mats), _p(rand), _p(out), _p(paths), _p(debug),
                                    C.c_int(mode))

    def connect_paths(self, n, cp, lp, accel, mats, camera, aggs, out, pix, path, ray, w, shade, splat_image=None,
                      mode=MODE_ORDERED):
        self.lib.emu_connect_paths(C.c_uint32(n), _p(cp), _p(lp), accel, _p(mats), _p(camera), _p(aggs), _p(out),
                                   _p(pix), _p(path), _p(ray), _p(w), _p(shade), _p(splat_image), C.c_int(mode))

    def adaptive_finalize_samples(self, n, aggs, camera, out, counts, bins, weights):
        self.lib.emu_adaptive_finalize_samples(C.c_uint32(n), _p(aggs), _p(camera), _p(out), _p(counts), _p(bins),
                                               _p(weights))

    def traverse(self, rays, accel, mode=MODE_ORDERED):
        n = len(rays)
        tri = np.empty(n, np.int32)
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        self.lib.emu_traverse(C.c_uint32(n), _p(rays), accel, _p(tri), _p(t), _p(u), _p(v), C.c_int(mode))
        return tri, t, u, v

    def traverse_counts(self, rays, accel):
        """Node and triangle records each ray's ORDERED closest-hit search fetches (work model for tools/lane_model.py)."""
        n = len(rays)
        nodes, tris = np.empty(n, np.uint32), np.empty(n, np.uint32)
        self.lib.emu_traverse_counts(C.c_uint32(n), _p(rays), accel, _p(nodes), _p(tris))
        return nodes, tris

    def visibility(self, a, b, accel, mode=MODE_ORDERED):
        out = np.empty(len(a), np.int32)
        self.lib.emu_visibility(C.c_uint32(len(a)), _p(a), _p(b), accel, _p(out), C.c_int(mode))
        return out

    def coop_hit_leaves(self, rays, accel, wide=False):
        """hit_leaf_step_warp (the warp-cooperative leaf step of k_walk_trace) on 32 host threads in lockstep against the
        per-lane hit_leaf_step, leaf by leaf: dict(compared, mismatches, max_pairs, wrong_hits); None without 32 threads."""
        out = (C.c_int64 * 4)()
        rc = self.lib.emu_coop_hit_leaves(C.c_uint32(len(rays)), _p(rays), accel, C.c_int(int(wide)), out)
        return None if rc else dict(zip(("compared", "mismatches", "max_pairs", "wrong_hits"), (int(x) for x in out)))

    def coop_occ_leaves(self, rays, t_bound, rank_bound, skip, accel):
        """occ_leaf_step_warp (k_connect_rays) against occ_leaf_step, the same way."""
        out = (C.c_int64 * 4)()
        t_bound, rank_bound, skip = (np.ascontiguousarray(x) for x in (t_bound, rank_bound, skip))
        rc = self.lib.emu_coop_occ_leaves(C.c_uint32(len(rays)), _p(rays), _p(t_bound), _p(rank_bound), _p(skip), accel, out)
        return None if rc else dict(zip(("compared", "mismatches", "max_pairs", "wrong_verdicts"), (int(x) for x in out)))

    def fused_pass(self, arrays, accel, seeds, light_paths, camera_paths, aggs, light_image, accum, sample_ids=None,
                   mode=MODE_ORDERED, soa=None, soa_len=None, wavefront=False, suspend_after=0):
        """One fused pass.  With `soa` (float32 (2*6*4*n_local, 4)) and `soa_len` (int32 (2*n_local,)) the renderer's compact
        SoA vertex records are used instead of the Path arrays.  `wavefront`: stages 1-4 as the rounds of k_walk_init /
        k_walk_trace / k_walk_shade instead of the per-sample megakernel (2: the trace steps over the 4-wide nodes);
        `suspend_after` > 0: a search still running after that many steps is suspended into a continuation record and
        resumed by the next round (hit_save / hit_restore), as k_walk_trace does with the searches that outlast a round."""
        n_local = len(seeds)
        self.lib.emu_fused_pass(C.c_uint32(n_local), _p(sample_ids), _p(arrays.light_triangles),
                                _p(arrays.light_surface_areas), _p(arrays.light_triangle_indices),
                                C.c_int(int(arrays.light_counts)), _p(arrays.materials), _p(arrays.camera), accel,
                                _p(seeds), _p(light_paths), _p(camera_paths), _p(aggs), _p(light_image), _p(accum),
                                C.c_int(mode), _p(soa), _p(soa_len), C.c_int(int(wavefront) | (int(suspend_after) << 8)))
