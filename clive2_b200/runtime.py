"""ctypes binding of libclive_b200.so and the device-object surface the reference gets from `metalcompute`.

The reference talks to its GPU through five things (src/scene.py:29,74-82; src/renderer.py:27-37,113-250,318-352):
`metalcompute.Device()`, `dev.buffer(nbytes | array)`, `dev.kernel(text).function(name)`, `fn(thread_count, *buffers)`
and `metalcompute.release(obj)`.  `Device`, `Buffer`, `KernelLibrary` and `KernelFunction` below give the same five
with the same names and positional argument orders, on top of the C ABI in include/clive_b200.h.  Differences that
follow from having discrete device memory instead of Apple unified memory:
  * a `Buffer` is not a Python buffer; `buf.numpy(dtype)` downloads it (and waits for pending kernels);
  * launches are stream ordered; the returned handle's deletion does not need to block.

There is no CPU fallback: loading fails loudly when the CUDA library is missing, and creating a `Device` fails when
no B200-class GPU is visible.
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import struct_types as T

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libclive_b200.so")
_lib = None

TRAVERSE_REFERENCE_ORDER, TRAVERSE_ORDERED = 0, 1
WALK_AUTO, WALK_MEGAKERNEL, WALK_WAVEFRONT, WALK_PIPELINE = 0, 1, 2, 3  # clv_ctx_walk_mode
BUF_SUMMED_IMAGE, BUF_SUMMED_SAMPLE_COUNTS, BUF_SUMMED_SAMPLE_WEIGHTS, BUF_UNIDIRECTIONAL = 0, 1, 2, 3
BUF_SEEDS, BUF_ACCUM_PACKED, BUF_LIGHT_PATHS, BUF_CAMERA_PATHS, BUF_AGGREGATORS, BUF_PASS_WEIGHT = 4, 5, 6, 7, 8, 9


class CliveError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the role of `metalcompute.error`, render.py:39)."""


error = CliveError  # metalcompute-style alias

_vp, _u32, _i32, _sz, _int = C.c_void_p, C.c_uint32, C.c_int32, C.c_size_t, C.c_int
_pp = C.POINTER(C.c_void_p)

#: name -> argtypes (every function returns int status unless listed in _RESTYPES)
SIGNATURES = {
    "clv_device_count": [C.POINTER(C.c_int)],
    "clv_ctx_create": [_int, _pp],
    "clv_ctx_destroy": [_vp],
    "clv_ctx_synchronize": [_vp],
    "clv_ctx_stream": [_vp, _pp],
    "clv_mem_alloc": [_vp, _sz, _pp],
    "clv_mem_free": [_vp, _vp],
    "clv_mem_upload": [_vp, _vp, _vp, _sz],
    "clv_mem_download": [_vp, _vp, _vp, _sz],
    "clv_mem_fill32": [_vp, _vp, _u32, _sz],
    "clv_host_alloc": [_sz, _pp],
    "clv_host_free": [_vp],
    "clv_accel_create": [_vp, _vp, _i32, _vp, _i32, _pp],
    "clv_accel_destroy": [_vp],
    "clv_accel_stats": [_vp, C.POINTER(C.c_uint64)],
    "clv_generate_light_rays": [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "clv_generate_camera_rays": [_vp, _u32, _vp, _vp, _vp, _vp],
    "clv_generate_paths": [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "clv_reset_light_indices": [_vp, _u32, _vp, _vp, _vp, _vp, _vp],
    "clv_connect_paths": [_vp, _u32] + [_vp] * 12,
    "clv_adaptive_finalize_samples": [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp],
    "clv_light_sort": [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32],
    "clv_light_image_gather": [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp],
    "clv_light_image_splat": [_vp, _u32, _u32] + [_vp] * 9,
    "clv_renderer_create": [_vp, _vp, _vp, _vp, _vp, _u32, _pp],
    "clv_renderer_destroy": [_vp],
    "clv_renderer_run_sample": [_vp],
    "clv_renderer_run_sample_e2e": [_vp, _vp, _vp],
    "clv_renderer_reset": [_vp, _vp],
    "clv_renderer_tone_map": [_vp, C.c_float, _vp],
    "clv_renderer_run_sample_preview": [_vp, _vp, C.c_float, _vp],
    "clv_renderer_read": [_vp, _int, _vp, _sz],
    "clv_renderer_device_ptr": [_vp, _int, _pp, C.POINTER(_sz)],
    "clv_renderer_set_traverse_mode": [_vp, _int],
    "clv_renderer_set_path_records": [_vp, _int],
    "clv_set_stage_traverse_mode": [_int],
    "clv_traverse": [_vp, _vp, _vp, _u32, _int, _vp, _vp, _vp, _vp],
    "clv_traverse_device": [_vp, _vp, _vp, _u32, _int, _vp, _vp, _vp, _vp],
    "clv_visibility": [_vp, _vp, _vp, _vp, _u32, _int, _vp],
    "clv_counters_enable": [_vp, _int],
    "clv_counters_read": [_vp, C.POINTER(C.c_uint64), _int],
    "clv_launch_count": [_vp, C.POINTER(C.c_uint64)],
    "clv_ctx_pass_slices": [_vp, _int, C.POINTER(C.c_int)],
    "clv_ctx_walk_mode": [_vp, _int, C.POINTER(C.c_int)],
    "clv_ctx_set_knob": [_vp, C.c_char_p, _int, C.POINTER(C.c_int)],
    "clv_measure_read_bandwidth": [_vp, _sz, _int, C.POINTER(C.c_double)],
    "clv_renderer_kernel_timing": [_vp, _int],
    "clv_renderer_kernel_times": [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), _int],
    "clv_last_error": [],
    "clv_version": [],
}
_RESTYPES = {"clv_last_error": C.c_char_p, "clv_version": C.c_char_p}


class SceneDesc(C.Structure):  # clv_scene_desc
    _fields_ = [("materials", _vp), ("n_materials", _i32), ("camera", _vp), ("light_triangles", _vp),
                ("light_surface_areas", _vp), ("light_triangle_indices", _vp), ("n_lights", _i32)]


def load_library(path=None):
    """Load libclive_b200.so and declare every prototype.  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("CLV_LIBRARY") or _LIB_PATH  # CLV_LIBRARY: a differently configured build (tools/variants.sh)
    if not os.path.exists(p):
        raise CliveError(f"{p} is missing: build it with `python -m clive2_b200.build` (nvcc, sm_100a). "
                         "clive2_b200 has no CPU execution path.")
    lib = C.CDLL(p)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if path is None:
        _lib = lib
    return lib


def _check(status):
    if status != 0:
        raise CliveError(f"clive_b200 status {status}: {load_library().clv_last_error().decode(errors='replace')}")


def device_count():
    n = C.c_int(0)
    lib = load_library()
    if lib.clv_device_count(C.byref(n)) != 0:
        return 0
    return n.value


def _host_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Buffer:
    """Device allocation (the object `dev.buffer(...)` returns in the reference)."""

    def __init__(self, device, nbytes, dtype=None, shape=None):
        self.device = device
        self.nbytes = int(nbytes)
        self.dtype = np.dtype(dtype) if dtype is not None else None
        self.shape = shape
        self.version = 0  # bumped whenever the contents change: invalidates acceleration structures built from them
        p = C.c_void_p()
        _check(device.lib.clv_mem_alloc(device.ctx, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, array):
        a = np.ascontiguousarray(array)
        if a.nbytes > self.nbytes:
            raise ValueError("array larger than buffer")
        _check(self.device.lib.clv_mem_upload(self.device.ctx, self.ptr, _host_ptr(a), a.nbytes))
        self._changed()
        return self

    def _changed(self):
        self.version += 1
        self.device._forget_accels(self)

    def numpy(self, dtype=None, count=None):
        """Download (waits for pending launches).  `dtype` defaults to the dtype the buffer was created from."""
        dt = np.dtype(dtype) if dtype is not None else (self.dtype or np.dtype(np.uint8))
        n = self.nbytes // dt.itemsize if count is None else int(count)
        out = np.empty(n, dtype=dt)
        _check(self.device.lib.clv_mem_download(self.device.ctx, _host_ptr(out), self.ptr, out.nbytes))
        if dtype is None and self.shape is not None and count is None:
            out = out.reshape(self.shape)
        return out

    def fill32(self, value):
        _check(self.device.lib.clv_mem_fill32(self.device.ctx, self.ptr, int(value) & 0xFFFFFFFF, self.nbytes // 4))
        self._changed()

    @property
    def __cuda_array_interface__(self):
        """Zero-copy view for torch.as_tensor(..., device='cuda') (used for the NCCL reduce of the accumulators)."""
        return {"shape": (self.nbytes // 4,), "typestr": "<f4", "data": (self.ptr, False), "version": 3, "strides": None}

    def release(self):
        # drop the cache entry of an acceleration structure built from this buffer (the structure itself is reference
        # counted in the library: a renderer that still uses it keeps it alive until the renderer is released)
        self.device._forget_accels(self)
        if getattr(self, "ptr", None):
            self.device.lib.clv_mem_free(self.device.ctx, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class LazyBuffer(Buffer):
    """A Buffer whose device copy is made on first use of `.ptr`.  The scene's `boxes` / `triangles` are such buffers:
    the fused renderer traverses the acceleration structure (built from the host arrays, with its own compact device
    records) and never touches them, so a 10 M-triangle scene does not hold a second 1.3 GB copy of its Triangle[] in
    HBM; the stage-level kernels, which take them as arguments, trigger the upload."""

    def __init__(self, device, array):
        a = np.ascontiguousarray(array)
        self.device = device
        self.nbytes = max(int(a.nbytes), 16)
        self.dtype, self.shape = a.dtype, a.shape
        self.version = 0
        self._host = a
        self._ptr = None

    @property
    def ptr(self):
        if self._ptr is None and self._host is not None:
            p = C.c_void_p()
            _check(self.device.lib.clv_mem_alloc(self.device.ctx, self.nbytes, C.byref(p)))
            self._ptr = p.value
            if self._host.nbytes:
                _check(self.device.lib.clv_mem_upload(self.device.ctx, self._ptr, _host_ptr(self._host), self._host.nbytes))
        return self._ptr

    @ptr.setter
    def ptr(self, value):
        self._ptr = value
        if value is None:
            self._host = None

    @property
    def resident(self):
        return self._ptr is not None

    def host_array(self):
        """The host copy the buffer was created from (None once its contents were changed on the device)."""
        return self._host

    def upload(self, array):
        a = np.ascontiguousarray(array)
        if a.nbytes > self.nbytes:
            raise ValueError("array larger than buffer")
        if self._ptr is None:
            self._host = a.copy() if a.nbytes == self._host.nbytes else None
            if self._host is None:
                raise ValueError("partial upload into a lazy buffer that is not resident yet")
            self._changed()
            return self
        self._host = None
        return Buffer.upload(self, a)

    def fill32(self, value):
        _ = self.ptr
        self._host = None
        Buffer.fill32(self, value)

    def numpy(self, dtype=None, count=None):
        if self._ptr is None and self._host is not None:
            out = self._host.reshape(-1).view(np.uint8)
            dt = np.dtype(dtype) if dtype is not None else self.dtype
            out = out.view(dt).copy()
            if count is not None:
                out = out[:int(count)]
            elif dtype is None:
                out = out.reshape(self.shape)
            return out
        return Buffer.numpy(self, dtype, count)

    def release(self):
        self.device._forget_accels(self)
        if self._ptr:
            self.device.lib.clv_mem_free(self.device.ctx, self._ptr)
        self._ptr = None
        self._host = None


class _ForeignBuffer:
    """Device memory owned by someone else (e.g. a renderer's accumulators) exposed through the Buffer interface."""

    def __init__(self, device, ptr, nbytes, owner):
        self.device, self.ptr, self.nbytes, self._owner = device, ptr, int(nbytes), owner

    numpy = Buffer.numpy
    __cuda_array_interface__ = Buffer.__cuda_array_interface__
    dtype = None
    shape = None

    def release(self):
        self.ptr = None


class Accel:
    """Device acceleration structure built from the flattened BVH arrays (clv_accel_create)."""

    def __init__(self, device, boxes, triangles):
        self.device = device
        b = np.ascontiguousarray(boxes, dtype=T.Box)
        t = np.ascontiguousarray(triangles, dtype=T.Triangle)
        h = C.c_void_p()
        _check(device.lib.clv_accel_create(device.ctx, _host_ptr(b), len(b), _host_ptr(t), len(t), C.byref(h)))
        self.handle = h
        self.n_boxes, self.n_triangles = len(b), len(t)

    def stats(self):
        out = (C.c_uint64 * 12)()
        _check(self.device.lib.clv_accel_stats(self.handle, out))
        return {"node_bytes": int(out[0]), "triangle_bytes": int(out[1]), "nodes": int(out[2]), "triangles": int(out[3]),
                "wide_node_bytes": int(out[4]), "wide_nodes": int(out[5]), "closest_uses_wide": bool(out[6]),
                "occlusion_uses_wide": bool(out[7]), "large_scene_tuning": bool(out[8]), "trace_uses_wide": bool(out[9]),
                "binary_depth": int(out[10]), "wide_depth": int(out[11])}

    def release(self):
        if getattr(self, "handle", None):
            self.device.lib.clv_accel_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class _Handle:
    """What a kernel call returns.  In the reference `del handle` blocks until the GPU is done; here launches are
    stream ordered, so nothing needs to wait until a buffer is read."""

    def __init__(self, device):
        self._device = device

    def wait(self):
        self._device.synchronize()


class KernelFunction:
    """`kernel.function(name)` of the reference: call as fn(thread_count, *buffers) with the reference's positional
    order.  Arguments may be `Buffer`s, numpy arrays (uploaded for the call) or scalars."""

    def __init__(self, device, name):
        if name not in _DISPATCH:
            raise CliveError(f"unknown kernel function '{name}'")
        self.device, self.name = device, name

    def __call__(self, thread_count, *args):
        _DISPATCH[self.name](self.device, int(thread_count), *args)
        return _Handle(self.device)


class KernelLibrary:
    """`dev.kernel(source)` of the reference.  The Metal source text is irrelevant here: the eight entry points are
    precompiled sm_100a kernels."""

    NAMES = ("generate_paths", "connect_paths", "generate_camera_rays", "generate_light_rays",
             "adaptive_finalize_samples", "light_sort", "light_image_gather", "reset_light_indices")

    def __init__(self, device):
        self.device = device

    def function(self, name):
        return KernelFunction(self.device, name)


class Device:
    """`metalcompute.Device()` (scene.py:29): one GPU, one launch stream."""

    def __init__(self, index=None):
        self.lib = load_library()
        if index is None:
            index = int(os.environ.get("LOCAL_RANK", "0")) if device_count() > 1 else 0
        h = C.c_void_p()
        _check(self.lib.clv_ctx_create(int(index), C.byref(h)))
        self.ctx = h
        self.index = int(index)
        self._accels = {}

    # --- metalcompute surface -------------------------------------------------------------------------------------
    def buffer(self, data):
        if isinstance(data, (int, np.integer)):
            return Buffer(self, int(data))
        a = np.ascontiguousarray(data)
        buf = Buffer(self, max(a.nbytes, 16), dtype=a.dtype, shape=a.shape)
        if a.nbytes:
            buf.upload(a)
        return buf

    def kernel(self, source=None):
        return KernelLibrary(self)

    # --- extras ---------------------------------------------------------------------------------------------------
    def lazy_buffer(self, data):
        """Like `buffer(array)`, but the device copy is made when a kernel first asks for it (see LazyBuffer)."""
        return LazyBuffer(self, data)

    def accel(self, boxes, triangles):
        """Acceleration structure for a (boxes, triangles) pair.

        For a pair of Buffers it is cached — the reference's call sites pass both buffers to every traversal kernel —
        under the identity AND the content version of the two buffers: `upload()` / `fill32()` / `release()` of either
        drops the entry, so a kernel never traverses a stale BVH.  Host arrays are not cached (CPython reuses `id()`s and
        nothing says when an array's contents change): the caller gets a fresh `Accel` and owns it."""
        if isinstance(boxes, Buffer) and isinstance(triangles, Buffer):
            key = (id(boxes), id(triangles))
            ent = self._accels.get(key)
            if ent is not None:
                rb, rt, acc, versions = ent
                if rb() is boxes and rt() is triangles and versions == (boxes.version, triangles.version) and acc.handle:
                    return acc
                self._accels.pop(key)
                acc.release()
            # host copies when the buffers still have them (no 1.3 GB device-to-host copy at 10 M triangles)
            hb = boxes.host_array() if isinstance(boxes, LazyBuffer) else None
            ht = triangles.host_array() if isinstance(triangles, LazyBuffer) else None
            acc = Accel(self, boxes.numpy(T.Box) if hb is None else hb, triangles.numpy(T.Triangle) if ht is None else ht)
            self._accels[key] = (weakref.ref(boxes), weakref.ref(triangles), acc, (boxes.version, triangles.version))
            return acc
        b = boxes.numpy(T.Box) if isinstance(boxes, Buffer) else boxes
        t = triangles.numpy(T.Triangle) if isinstance(triangles, Buffer) else triangles
        return Accel(self, b, t)

    def _accel_for_call(self, boxes, triangles):
        """(accel, owned): `owned` structures were built for this one call and are released by the caller."""
        cached = isinstance(boxes, Buffer) and isinstance(triangles, Buffer)
        return self.accel(boxes, triangles), not cached

    def _forget_accels(self, buf):
        """Drop cache entries built from `buf` (and entries whose buffers are gone).  Dropping releases the cache's
        reference only: the library keeps a structure alive while a renderer uses it (clv_accel is reference counted)."""
        for key in list(getattr(self, "_accels", {})):
            rb, rt, acc, _ = self._accels[key]
            if rb() is buf or rt() is buf or rb() is None or rt() is None:
                self._accels.pop(key)
                acc.release()

    def synchronize(self):
        _check(self.lib.clv_ctx_synchronize(self.ctx))

    @property
    def stream(self):
        s = C.c_void_p()
        _check(self.lib.clv_ctx_stream(self.ctx, C.byref(s)))
        return s.value or 0

    def counters_enable(self, enabled=True):
        _check(self.lib.clv_counters_enable(self.ctx, int(bool(enabled))))

    def counters(self, reset=True, per_class=False):
        """Work counters of the instrumented kernels (see clv_counters_read).  Totals by default; with `per_class`
        a dict {"walk": {...}, "connect": {...}}."""
        out = (C.c_uint64 * 10)()
        _check(self.lib.clv_counters_read(self.ctx, out, int(bool(reset))))
        names = ("rays_ref", "rays_traced", "node_fetches", "tri_fetches", "vertices")
        classes = {cls: {n: int(out[5 * k + i]) for i, n in enumerate(names)} for k, cls in enumerate(("walk", "connect"))}
        if per_class:
            return classes
        return {n: classes["walk"][n] + classes["connect"][n] for n in names}

    def read_bandwidth(self, nbytes, iters=20):
        """Streaming-read GB/s of an nbytes scratch buffer (L2 peak when it fits L2, HBM read peak when far larger)."""
        g = C.c_double()
        _check(self.lib.clv_measure_read_bandwidth(self.ctx, int(nbytes), int(iters), C.byref(g)))
        return float(g.value)

    @property
    def launch_count(self):
        n = C.c_uint64()
        _check(self.lib.clv_launch_count(self.ctx, C.byref(n)))
        return int(n.value)

    @property
    def walk_mode(self):
        """Stages 1-4 of the fused pass: 0 automatic, 1 megakernel, 2 wavefront rounds (clv_ctx_walk_mode)."""
        out = C.c_int()
        _check(self.lib.clv_ctx_walk_mode(self.ctx, -1, C.byref(out)))
        return int(out.value)

    @walk_mode.setter
    def walk_mode(self, mode):
        _check(self.lib.clv_ctx_walk_mode(self.ctx, int(mode), None))

    def knob(self, name, value=-1):
        """Set (value >= 0) or query a scheduling knob of the ray kernels (clv_ctx_set_knob); returns the previous value."""
        prev = C.c_int()
        _check(self.lib.clv_ctx_set_knob(self.ctx, name.encode(), int(value), C.byref(prev)))
        return int(prev.value)

    @property
    def pass_slices(self):
        """Number of independent sample slices (streams) a fused pass is launched as (4 * slices + 1 kernels per
        pass); 0 = automatic (3 from one full wave of samples up, else 1)."""
        n = C.c_int()
        _check(self.lib.clv_ctx_pass_slices(self.ctx, -1, C.byref(n)))
        return int(n.value)

    @pass_slices.setter
    def pass_slices(self, slices):
        if int(slices) < 0:
            raise ValueError("pass_slices must be >= 0 (0 = automatic)")
        _check(self.lib.clv_ctx_pass_slices(self.ctx, int(slices), None))

    def traverse(self, accel, rays, mode=TRAVERSE_ORDERED):
        rays = np.ascontiguousarray(rays, dtype=T.Ray)
        n = len(rays)
        tri = np.empty(n, np.int32)
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        _check(self.lib.clv_traverse(self.ctx, accel.handle, _host_ptr(rays), n, mode, _host_ptr(tri), _host_ptr(t),
                                     _host_ptr(u), _host_ptr(v)))
        return tri, t, u, v

    def visibility(self, accel, a, b, mode=TRAVERSE_ORDERED):
        a = np.ascontiguousarray(a, dtype=T.Ray)
        b = np.ascontiguousarray(b, dtype=T.Ray)
        out = np.empty(len(a), np.int32)
        _check(self.lib.clv_visibility(self.ctx, accel.handle, _host_ptr(a), _host_ptr(b), len(a), mode, _host_ptr(out)))
        return out

    def light_image_splat(self, n_entries, n_pixels, light_paths, materials, pixel_indices, path_indices, ray_indices,
                          weights, shades, light_image, sum_weights):
        """Device-side replacement of sort + `light_bins()` + `light_image_gather` on the UNSORTED light entries written
        by `connect_paths` (clv_light_image_splat): same sums, atomic order."""
        A = _Args(self)
        _check(self.lib.clv_light_image_splat(self.ctx, int(n_entries), int(n_pixels), A.ptr(light_paths), A.ptr(materials),
                                              A.ptr(pixel_indices), A.ptr(path_indices), A.ptr(ray_indices),
                                              A.ptr(weights), A.ptr(shades), A.ptr(light_image), A.ptr(sum_weights)))
        A.done()

    def release(self):
        for ent in list(getattr(self, "_accels", {}).values()):
            ent[2].release()
        self._accels = {}
        if getattr(self, "ctx", None):
            self.lib.clv_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def release(obj):
    """`metalcompute.release(obj)` (renderer.py:250,318-352)."""
    if hasattr(obj, "release"):
        obj.release()


def set_stage_traverse_mode(mode):
    _check(load_library().clv_set_stage_traverse_mode(int(mode)))


# ---------------------------------------------------------------------------------------------------------------
# positional dispatch of the eight kernels (reference orders: renderer.py:115-123, 129-135, 140-150, 171-178, 181-196,
# 201-209, 219-230, 235-247)
# ---------------------------------------------------------------------------------------------------------------
class _Args:
    """Turns call arguments into device pointers; numpy arrays are uploaded into temporaries that live until the
    launch has been enqueued and the stream drained."""

    def __init__(self, device):
        self.device, self.temps = device, []

    def ptr(self, a):
        if isinstance(a, (Buffer, _ForeignBuffer)):
            return a.ptr
        if isinstance(a, np.ndarray) or np.isscalar(a):
            buf = self.device.buffer(np.atleast_1d(np.asarray(a)))
            self.temps.append(buf)
            return buf.ptr
        raise TypeError(f"cannot pass {type(a).__name__} to a kernel")

    def done(self):
        if self.temps:
            self.device.synchronize()
            for b in self.temps:
                b.release()


def _k_generate_light_rays(dev, n, light_triangles, surface_areas, light_triangle_indices, materials, rand, out, counts):
    A = _Args(dev)
    _check(dev.lib.clv_generate_light_rays(dev.ctx, n, A.ptr(light_triangles), A.ptr(surface_areas),
                                           A.ptr(np.asarray(light_triangle_indices, np.int32)
                                                 if isinstance(light_triangle_indices, np.ndarray)
                                                 else light_triangle_indices),
                                           A.ptr(materials), A.ptr(rand), A.ptr(out), A.ptr(counts)))
    A.done()


def _k_generate_camera_rays(dev, n, camera, rand, indices, out):
    A = _Args(dev)
    _check(dev.lib.clv_generate_camera_rays(dev.ctx, n, A.ptr(camera), A.ptr(rand), A.ptr(indices), A.ptr(out)))
    A.done()


def _k_generate_paths(dev, n, rays, boxes, triangles, materials, rand, out, paths, debug):
    A = _Args(dev)
    acc, owned = dev._accel_for_call(boxes, triangles)
    try:
        _check(dev.lib.clv_generate_paths(dev.ctx, n, A.ptr(rays), acc.handle, A.ptr(materials), A.ptr(rand), A.ptr(out),
                                          A.ptr(paths), A.ptr(debug)))
        A.done()
    finally:
        if owned:
            acc.release()  # waits for the launch


def _k_reset_light_indices(dev, n, pix, path, ray, w, shade):
    A = _Args(dev)
    _check(dev.lib.clv_reset_light_indices(dev.ctx, n, A.ptr(pix), A.ptr(path), A.ptr(ray), A.ptr(w), A.ptr(shade)))
    A.done()


def _k_connect_paths(dev, n, camera_paths, light_paths, triangles, materials, boxes, camera, aggregators, out, pix, path,
                     ray, w, shade):
    A = _Args(dev)
    acc, owned = dev._accel_for_call(boxes, triangles)
    try:
        _check(dev.lib.clv_connect_paths(dev.ctx, n, A.ptr(camera_paths), A.ptr(light_paths), acc.handle, A.ptr(materials),
                                         A.ptr(camera), A.ptr(aggregators), A.ptr(out), A.ptr(pix), A.ptr(path), A.ptr(ray),
                                         A.ptr(w), A.ptr(shade)))
        A.done()
    finally:
        if owned:
            acc.release()


def _k_adaptive_finalize_samples(dev, n, aggregators, camera, out, counts, bin_offsets, weights):
    A = _Args(dev)
    _check(dev.lib.clv_adaptive_finalize_samples(dev.ctx, n, A.ptr(aggregators), A.ptr(camera), A.ptr(out),
                                                 A.ptr(counts), A.ptr(bin_offsets), A.ptr(weights)))
    A.done()


def _k_light_sort(dev, n_threads, pix, path, ray, w, shade, stage, pass_of_stage, n, pairs_per_thread):
    A = _Args(dev)
    _check(dev.lib.clv_light_sort(dev.ctx, n_threads, A.ptr(pix), A.ptr(path), A.ptr(ray), A.ptr(w), A.ptr(shade),
                                  int(stage), int(pass_of_stage), int(n), int(pairs_per_thread)))
    A.done()


def _k_light_image_gather(dev, n, light_paths, materials, path_idx, ray_idx, bins, offset, weights, shades, light_image,
                          sum_weights):
    A = _Args(dev)
    if isinstance(bins, np.ndarray):
        bins = bins.astype(np.int32, copy=False)
    _check(dev.lib.clv_light_image_gather(dev.ctx, n, A.ptr(light_paths), A.ptr(materials), A.ptr(path_idx),
                                          A.ptr(ray_idx), A.ptr(bins), int(offset), A.ptr(weights), A.ptr(shades),
                                          A.ptr(light_image), A.ptr(sum_weights)))
    A.done()


_DISPATCH = {
    "generate_light_rays": _k_generate_light_rays,
    "generate_camera_rays": _k_generate_camera_rays,
    "generate_paths": _k_generate_paths,
    "reset_light_indices": _k_reset_light_indices,
    "connect_paths": _k_connect_paths,
    "adaptive_finalize_samples": _k_adaptive_finalize_samples,
    "light_sort": _k_light_sort,
    "light_image_gather": _k_light_image_gather,
}


# ---------------------------------------------------------------------------------------------------------------
# fused per-pass renderer handle (clv_renderer_*)
# ---------------------------------------------------------------------------------------------------------------
class FusedRenderer:
    """Thin owner of a `clv_renderer`: one call per sample pass, accumulators resident on the device."""
    #: kernels of one pass, in launch order
    KERNELS = ("k_fused_walks", "k_connect_enumerate", "k_connect_rays", "k_connect_finish", "k_fused_finalize")
    #: slots of clv_renderer_kernel_times (the walks are either the megakernel or the two wavefront kernels)
    TIMED = KERNELS + ("k_walk_trace", "k_walk_shade", "k_walk_pipeline")

    def __init__(self, device, accel, arrays, seeds, sample_ids=None):
        self.device, self.accel = device, accel
        self.width, self.height = int(arrays.pixel_width), int(arrays.pixel_height)
        self.n_pixels = self.width * self.height
        keep = [np.ascontiguousarray(arrays.materials, dtype=T.Material),
                np.ascontiguousarray(arrays.camera, dtype=T.Camera).reshape(1),
                np.ascontiguousarray(arrays.light_triangles, dtype=T.Triangle),
                np.ascontiguousarray(arrays.light_surface_areas, dtype=np.float32),
                np.ascontiguousarray(arrays.light_triangle_indices, dtype=np.int32)]
        desc = SceneDesc(_host_ptr(keep[0]), len(keep[0]), _host_ptr(keep[1]), _host_ptr(keep[2]), _host_ptr(keep[3]),
                         _host_ptr(keep[4]), len(keep[2]))
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32).reshape(self.n_pixels, 2)
        ids = None if sample_ids is None else np.ascontiguousarray(sample_ids, dtype=np.uint32)
        if ids is not None and len(ids) == 0:
            raise ValueError("empty shard: this rank owns no sample ids (frame smaller than world_size tiles?); "
                             "render with fewer ranks or smaller tiles")
        if ids is not None and len(np.unique(ids)) != len(ids):
            # every pixel's accumulator is updated without atomics by the one sample that owns it
            raise ValueError("sample_ids must not contain duplicates")
        self.n_local = self.n_pixels if ids is None else len(ids)
        h = C.c_void_p()
        _check(device.lib.clv_renderer_create(device.ctx, accel.handle, C.byref(desc), _host_ptr(seeds),
                                              None if ids is None else _host_ptr(ids), self.n_local, C.byref(h)))
        self.handle = h

    def run_sample(self):
        _check(self.device.lib.clv_renderer_run_sample(self.handle))

    def run_sample_e2e(self, h_seeds_ptr, h_accum_ptr):
        _check(self.device.lib.clv_renderer_run_sample_e2e(self.handle, h_seeds_ptr, h_accum_ptr))

    def set_traverse_mode(self, mode):
        _check(self.device.lib.clv_renderer_set_traverse_mode(self.handle, int(mode)))

    def tone_map(self, exposure, out=None):
        """uint8 (H,W,3) BGR preview computed on the device (clv_renderer_tone_map)."""
        if out is None:
            out = np.empty((self.height, self.width, 3), np.uint8)
        ptr = out.ptr if isinstance(out, PinnedArray) else _host_ptr(out)
        _check(self.device.lib.clv_renderer_tone_map(self.handle, float(exposure), ptr))
        return out.array if isinstance(out, PinnedArray) else out

    def run_sample_preview(self, h_seeds_ptr, exposure, h_image_ptr):
        _check(self.device.lib.clv_renderer_run_sample_preview(self.handle, h_seeds_ptr, float(exposure), h_image_ptr))

    def set_path_records(self, enabled=True):
        """Keep the reference-layout Path records of each pass (readable as BUF_LIGHT_PATHS / BUF_CAMERA_PATHS)."""
        _check(self.device.lib.clv_renderer_set_path_records(self.handle, int(bool(enabled))))

    def kernel_timing(self, enabled=True):
        _check(self.device.lib.clv_renderer_kernel_timing(self.handle, int(bool(enabled))))

    def kernel_times(self, reset=True):
        """(ms per kernel summed over the timed passes, number of passes)."""
        ms, n = (C.c_double * 8)(), C.c_uint64()
        _check(self.device.lib.clv_renderer_kernel_times(self.handle, ms, C.byref(n), int(bool(reset))))
        return {k: v for k, v in zip(self.TIMED, ms) if v > 0.0 or k in self.KERNELS[1:]}, int(n.value)

    def reset(self, seeds=None):
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32).reshape(self.n_pixels, 2)
        _check(self.device.lib.clv_renderer_reset(self.handle, None if s is None else _host_ptr(s)))

    def read(self, buffer_id, dtype, count):
        out = np.empty(count, dtype=dtype)
        _check(self.device.lib.clv_renderer_read(self.handle, int(buffer_id), _host_ptr(out), out.nbytes))
        return out

    def device_buffer(self, buffer_id):
        p, n = C.c_void_p(), C.c_size_t()
        _check(self.device.lib.clv_renderer_device_ptr(self.handle, int(buffer_id), C.byref(p), C.byref(n)))
        return _ForeignBuffer(self.device, p.value, n.value, self)

    def release(self):
        if getattr(self, "handle", None):
            self.device.lib.clv_renderer_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class PinnedArray:
    """Page-locked host array (clv_host_alloc) for the end-to-end path."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = C.c_void_p()
        _check(self.lib.clv_host_alloc(n, C.byref(p)))
        self.ptr = p.value
        raw = (C.c_uint8 * n).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=dt).reshape(shape)

    def release(self):
        if getattr(self, "ptr", None):
            self.array = None
            self.lib.clv_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
