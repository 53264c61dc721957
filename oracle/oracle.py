"""ctypes binding of the CPU oracle (oracle/trace_ref.c).  TEST INFRASTRUCTURE — never imported by clive2_b200.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this module.
Arrays are numpy arrays in the `clive2_b200.struct_types` layouts and are modified in place, exactly like the device
buffers the Metal kernels write (positional argument order = SURVEY.md Appendix D).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
REF_DIR = os.path.join(HERE, "_ref")  # the reference's own trace.metal compiled as C++ (git-ignored build output)
REFERENCE_TREE = os.environ.get("CLV_REFERENCE_TREE", "/root/reference")
STAGES = ("make_light_rays", "make_camera_rays", "trace_light_rays", "trace_camera_rays", "join_paths",
          "finalize_samples", "gather_light_image", "process_images")


def build(force=False):
    """Compile both oracle variants with oracle/Makefile (gcc -O3 -fopenmp -ffp-contract=off)."""
    libs = [os.path.join(BUILD_DIR, n) for n in ("libclive_oracle.so", "libclive_oracle_sm.so")]
    srcs = [os.path.join(HERE, n) for n in ("trace_ref.c", "ref_glue.c", "ref_types.h", "Makefile")]
    newest = max(os.path.getmtime(p) for p in srcs)
    if force or not all(os.path.exists(p) and os.path.getmtime(p) >= newest for p in libs):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return libs


def ref_libs():
    return [os.path.join(REF_DIR, n) for n in ("libclive_ref.so", "libclive_ref_sm.so")]


def build_ref(force=False):
    """Compile the reference's src/trace.metal, unmodified and where it lies, as C++ (`make -C oracle ref`) into
    oracle/_ref/.  Possible only where the reference tree exists (the build container); on the GPU box the prebuilt
    files that travelled with the snapshot are used.  Returns the library paths, or [] when neither is available."""
    libs = ref_libs()
    have_tree = os.path.exists(os.path.join(REFERENCE_TREE, "src", "trace.metal"))
    if have_tree and (force or not all(os.path.exists(p) for p in libs)):
        subprocess.run(["make", "-C", HERE, "ref", f"REFERENCE={REFERENCE_TREE}"] + (["-B"] if force else []),
                       check=True, capture_output=True)
    return libs if all(os.path.exists(p) for p in libs) else []


def have_ref():
    return all(os.path.exists(p) for p in ref_libs())


def _ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle buffers must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One of four builds: the C restatement (oracle/trace_ref.c) or, with reference=True, the reference's own
    trace.metal compiled as C++ (oracle/_ref); each with glibc libm or with the shared clv_math.h functions."""

    def __init__(self, shared_math=False, reference=False):
        if reference:
            if not build_ref():
                raise FileNotFoundError("oracle/_ref is not built (needs /root/reference: `make -C oracle ref`)")
            name = os.path.join(REF_DIR, "libclive_ref_sm.so" if shared_math else "libclive_ref.so")
        else:
            build()
            name = os.path.join(BUILD_DIR, "libclive_oracle_sm.so" if shared_math else "libclive_oracle.so")
        self.lib = C.CDLL(name)
        self.shared_math = bool(shared_math)
        self.reference = bool(reference)
        L = self.lib
        L.ref_renderer_create.restype = C.c_void_p
        L.ref_renderer_buffer.restype = C.c_void_p
        L.ref_renderer_buffer.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64)]
        L.ref_renderer_stage.argtypes = [C.c_void_p, C.c_int]
        L.ref_renderer_run_sample.argtypes = [C.c_void_p]
        L.ref_renderer_destroy.argtypes = [C.c_void_p]
        L.ref_xorshift32.restype = C.c_uint32
        assert L.ref_uses_shared_math() == int(self.shared_math)

    # ---- counters -------------------------------------------------------------------------------------------
    def counters_reset(self):
        self.lib.ref_counters_reset()

    def counters(self):
        out = (C.c_uint64 * 3)()
        self.lib.ref_counters_get(out)
        return {"rays": int(out[0]), "node_visits": int(out[1]), "tri_tests": int(out[2])}

    @property
    def num_threads(self):
        return int(self.lib.ref_num_threads())

    def use_all_cores(self):
        """Run on every host core even when the launcher (torchrun) exported OMP_NUM_THREADS=1."""
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        self.lib.ref_set_num_threads(C.c_int(n))
        return self.num_threads

    # ---- the eight kernels, reference positional order ----------------------------------------------------------
    def _call(self, fn, n, *args):
        conv = []
        for a in args:
            conv.append(_ptr(a) if isinstance(a, np.ndarray) else a)
        getattr(self.lib, fn)(C.c_uint32(n), *conv)

    def generate_light_rays(self, n, light_triangles, surface_areas, light_triangle_indices, materials, rand, out, counts):
        self._call("ref_generate_light_rays", n, light_triangles, surface_areas, light_triangle_indices, materials, rand,
                   out, np.ascontiguousarray(counts, dtype=np.int32).reshape(1))

    def generate_camera_rays(self, n, camera, rand, indices, out):
        self._call("ref_generate_camera_rays", n, camera, rand, indices, out)

    def generate_paths(self, n, rays, boxes, triangles, materials, rand, out, paths, debug):
        self._call("ref_generate_paths", n, rays, boxes, triangles, materials, rand, out, paths, debug)

    def reset_light_indices(self, n, pix, path, ray, w, shade):
        self._call("ref_reset_light_indices", n, pix, path, ray, w, shade)

    def connect_paths(self, n, camera_paths, light_paths, triangles, materials, boxes, camera, aggregators, out, pix,
                      path, ray, w, shade):
        self._call("ref_connect_paths", n, camera_paths, light_paths, triangles, materials, boxes, camera, aggregators,
                   out, pix, path, ray, w, shade)

    def adaptive_finalize_samples(self, n, aggregators, camera, out, counts, bin_offsets, weights):
        self._call("ref_adaptive_finalize_samples", n, aggregators, camera, out, counts, bin_offsets, weights)

    def light_sort(self, n_threads, pix, path, ray, w, shade, stage, pass_of_stage, n, pairs_per_thread):
        self._call("ref_light_sort", n_threads, pix, path, ray, w, shade, C.c_uint32(stage), C.c_uint32(pass_of_stage),
                   C.c_uint32(n), C.c_uint32(pairs_per_thread))

    def light_image_gather(self, n, light_paths, materials, path_idx, ray_idx, bins, offset, weights, shades,
                           light_image, sum_weights):
        off = np.array([offset], dtype=np.uint32)
        self._call("ref_light_image_gather", n, light_paths, materials, path_idx, ray_idx, bins, off, weights, shades,
                   light_image, sum_weights)

    # ---- parity hooks ---------------------------------------------------------------------------------------------
    def traverse(self, rays, boxes, triangles):
        n = len(rays)
        tri = np.empty(n, np.int32)
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        self._call("ref_traverse", n, rays, boxes, triangles, tri, t, u, v)
        return tri, t, u, v

    def visibility(self, a, b, boxes, triangles):
        out = np.empty(len(a), np.int32)
        self._call("ref_visibility", len(a), a, b, boxes, triangles, out)
        return out

    def math_probe(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty_like(x)
        self._call("ref_math_probe", len(x), x, C.c_int({"sin": 0, "cos": 1, "acos": 2, "atan": 3, "exp": 4}[which]), y)
        return y

    def xorshift32(self, seed):
        f = C.c_float()
        s = self.lib.ref_xorshift32(C.c_uint32(seed), C.byref(f))
        return int(s), float(f.value)


class OracleRenderer:
    """renderer.py's per-pass dispatcher restated in C (ref_renderer_*): same buffers, same stage order."""

    def __init__(self, arrays, seeds, shared_math=False, oracle=None, reference=False):
        from clive2_b200 import struct_types as T
        self.T = T
        self.oracle = oracle or Oracle(shared_math, reference=reference)
        self.arrays = arrays  # keeps the borrowed numpy arrays alive
        self.width, self.height = arrays.pixel_width, arrays.pixel_height
        self.batch = self.width * self.height
        self.seeds = np.ascontiguousarray(seeds, dtype=np.uint32).reshape(self.batch, 2)
        a = arrays
        self._keep = [np.ascontiguousarray(x) for x in (a.boxes, a.triangles, a.materials, a.camera, a.light_triangles,
                                                       a.light_surface_areas, a.light_triangle_indices)]
        self.h = C.c_void_p(self.oracle.lib.ref_renderer_create(*[_ptr(x) for x in self._keep],
                                                                 C.c_int32(int(a.light_counts)), _ptr(self.seeds)))
        self.samples = 0

    def stage(self, idx_or_name):
        i = STAGES.index(idx_or_name) if isinstance(idx_or_name, str) else int(idx_or_name)
        self.oracle.lib.ref_renderer_stage(self.h, i)

    def run_sample(self):
        self.oracle.lib.ref_renderer_run_sample(self.h)
        self.samples += 1

    def buffer(self, name, dtype=np.uint8):
        """Zero-copy numpy view of a renderer buffer (valid until close())."""
        nbytes = C.c_uint64()
        p = self.oracle.lib.ref_renderer_buffer(self.h, name.encode(), C.byref(nbytes))
        if not p:
            raise KeyError(name)
        raw = (C.c_uint8 * nbytes.value).from_address(p)
        return np.frombuffer(raw, dtype=dtype)

    # conveniences in the reference's shapes (renderer.py:41-45)
    @property
    def summed_image(self):
        return self.buffer("summed_image", np.float32).reshape(self.height, self.width, 3)

    @property
    def summed_sample_weights(self):
        return self.buffer("summed_sample_weights", np.float32).reshape(self.height, self.width, 1)

    @property
    def summed_sample_counts(self):
        return self.buffer("summed_sample_counts", np.int32).reshape(self.height, self.width, 1)

    @property
    def unidirectional_image_buffer(self):
        return self.buffer("unidirectional", np.float32).reshape(self.height, self.width, 3)

    @property
    def stage_seconds(self):
        return dict(zip(STAGES, self.buffer("stage_seconds", np.float64).tolist()))

    def close(self):
        if self.h:
            self.oracle.lib.ref_renderer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
