"""Per-pass dispatcher: the `Renderer` surface of the reference (src/renderer.py:16-352) on the B200 runtime.

Two ways to run a sample pass, same results:

* **fused** (default): `run_sample()` is ONE call into the C ABI (`clv_renderer_run_sample`): five kernels on one
  stream, no host synchronisation, light image splatted with atomics instead of sorted, accumulators resident on the
  device.  `summed_image` & co. are read back only when asked for.
* **staged** (`fused=False`): the reference's eight stage methods, each launching the same-named kernel through
  `device.kernel().function(name)` with the reference's positional arguments and buffers (renderer.py:113-250),
  including the bitonic `light_sort` loop, the host-side `light_bins()` and the numpy accumulation of
  `process_images()`.  This is the seam the differential tests use: every stage's buffers can be compared with the
  oracle's.

`summed_image (H,W,3) f32`, `summed_sample_counts (H,W,1) i32`, `summed_sample_weights (H,W,1) f32`,
`unidirectional_image_buffer (H,W,3) f32`, `samples`, `image`, `unweighted_image`, `unidirectional_image` keep the
reference's meaning (renderer.py:41-45,293-316).
"""
import numpy as np

from . import runtime
from .camera import tone_map
from .constants import MAX_PATH_LENGTH, TONE_MAP_EXPOSURE
from .struct_types import Path, Ray, WeightAggregator


def next_power_of_two(n):
    """Smallest power of two >= n (renderer.py:11-13)."""
    return 1 << (n - 1).bit_length() if n > 0 else 1


def make_seeds(batch_size, seed=None):
    """uint32[batch][2] RNG state (renderer.py:86-87).  The reference draws from the unseeded global numpy RNG; a
    `seed` makes it reproducible.  Zero words are replaced by 1 (0 is the fixed point of xorshift32)."""
    if seed is None:
        s = np.random.randint(0, 2 ** 32, size=(batch_size, 2), dtype=np.uint32)
    else:
        s = np.random.RandomState(seed).randint(0, 2 ** 32, size=(batch_size, 2), dtype=np.uint32)
    s[s == 0] = 1
    return s


class Renderer:
    def __init__(self, scene, kernel_path=None, fused=True, seed=None, seeds=None, sample_ids=None,
                 traverse_mode=runtime.TRAVERSE_ORDERED, path_records=False):
        dev = scene.device
        self.device = dev
        self.scene = scene
        self.fused = bool(fused)
        self.pixel_width = scene.pixel_width
        self.pixel_height = scene.pixel_height
        self.batch_size = scene.pixel_width * scene.pixel_height
        self.samples = 0
        self._seed = seed
        self._initial_seeds = np.ascontiguousarray(seeds, np.uint32).reshape(self.batch_size, 2) if seeds is not None \
            else self.get_random_buffer()
        self.traverse_mode = traverse_mode
        if self.fused:
            if scene.arrays is None:
                raise ValueError("fused rendering needs scene.arrays (build the scene with clive2_b200.scene)")
            accel = dev.accel(scene.boxes, scene.triangles)
            self._fr = runtime.FusedRenderer(dev, accel, scene.arrays, self._initial_seeds, sample_ids)
            self._fr.set_traverse_mode(traverse_mode)
            if path_records:  # keep the reference's Path records readable (tests, tools); default: compact SoA records
                self._fr.set_path_records(True)
            self._cache = None
        else:
            if sample_ids is not None:
                raise ValueError("sample_ids (sharding) is a feature of the fused renderer")
            self._init_staged()

    # ---- shared helpers ------------------------------------------------------------------------------------------
    def get_random_buffer(self):
        return make_seeds(self.batch_size, self._seed)

    # ================================================================================================================
    # fused path
    # ================================================================================================================
    def _accum(self):
        if self._cache is None:
            a = self._fr.read(runtime.BUF_ACCUM_PACKED, np.float32, self.batch_size * 8)
            self._cache = a.reshape(self.pixel_height, self.pixel_width, 8)
        return self._cache

    def accumulator_buffer(self):
        """Device view of the packed accumulators float[H][W][8] (image3, weight, uni3, count): the payload of the
        multi-GPU reduce.  Invalidate host caches after modifying it in place."""
        self._cache = None
        return self._fr.device_buffer(runtime.BUF_ACCUM_PACKED)

    def run_sample(self):
        if self.fused:
            self._fr.run_sample()
            self._cache = None
        else:
            self.make_light_rays()
            self.make_camera_rays()
            self.trace_light_rays()
            self.trace_camera_rays()
            self.join_paths()
            self.finalize_samples()
            self.gather_light_image()
            self.process_images()
        self.samples += 1

    def run_sample_e2e(self, host_seeds, host_accum):
        """One pass end to end with HOST buffers (fused mode): upload `host_seeds` (uint32 (W*H,2), a fresh RNG state
        for this pass; None keeps the device-resident streams), run the pass, download the packed accumulators
        float32 (H,W,8) into `host_accum` and wait.  Both arguments are `runtime.PinnedArray`s (or None)."""
        if not self.fused:
            raise ValueError("run_sample_e2e is a fused-renderer call")
        self._fr.run_sample_e2e(None if host_seeds is None else host_seeds.ptr,
                                None if host_accum is None else host_accum.ptr)
        self._cache = None
        self.samples += 1

    def run_sample_preview(self, host_seeds, host_image):
        """The interactive loop body of the reference's render.py:31-37 with HOST buffers: upload seeds (or None), one
        pass, device tone map (exposure 4), download the uint8 (H,W,3) preview into `host_image` (PinnedArray)."""
        if not self.fused:
            raise ValueError("run_sample_preview is a fused-renderer call")
        self._fr.run_sample_preview(None if host_seeds is None else host_seeds.ptr, TONE_MAP_EXPOSURE, host_image.ptr)
        self._cache = None
        self.samples += 1

    def image_device(self, out=None):
        """`image` computed on the device (log-average tone map as camera.tone_map; equal up to +-1 per channel because
        only the order of the float64 sum of logs differs)."""
        if not self.fused:
            raise ValueError("image_device needs the fused renderer")
        return self._fr.tone_map(TONE_MAP_EXPOSURE, out)

    # ================================================================================================================
    # staged path: renderer.py:47-84 buffers and :113-278 stages
    # ================================================================================================================
    def _init_staged(self):
        dev, b = self.device, self.batch_size
        runtime.set_stage_traverse_mode(self.traverse_mode)
        k = dev.kernel()
        self.kernel = k
        self.trace_fn = k.function("generate_paths")
        self.join_fn = k.function("connect_paths")
        self.camera_ray_fn = k.function("generate_camera_rays")
        self.light_ray_fn = k.function("generate_light_rays")
        self.finalize_fn = k.function("adaptive_finalize_samples")
        self.light_sort_fn = k.function("light_sort")
        self.light_image_gather_fn = k.function("light_image_gather")
        self.light_reset_fn = k.function("reset_light_indices")

        res = (self.pixel_height, self.pixel_width)
        self.summed_image = np.zeros((*res, 3), dtype=np.float32)
        self.summed_sample_counts = np.zeros((*res, 1), dtype=np.int32)
        self.summed_sample_weights = np.zeros((*res, 1), dtype=np.float32)
        self.light_image = np.zeros((*res, 1), dtype=np.float32)
        self.unidirectional_image_buffer = np.zeros((*res, 3), dtype=np.float32)

        self.camera_ray_buffer = dev.buffer(b * Ray.itemsize)
        self.light_ray_buffer = dev.buffer(b * Ray.itemsize)
        self.rand_buffer = dev.buffer(self._initial_seeds)
        self.out_camera_image = dev.buffer(b * 16)
        self.out_camera_paths = dev.buffer(b * Path.itemsize)
        self.out_camera_debug_image = dev.buffer(b * 16)
        self.out_samples = dev.buffer(b * 16)
        sz = next_power_of_two(b * MAX_PATH_LENGTH) * 4
        self.out_light_indices = dev.buffer(sz)
        self.out_light_path_indices = dev.buffer(sz)
        self.out_light_ray_indices = dev.buffer(sz)
        self.out_light_weights = dev.buffer(sz)
        self.out_light_shade = dev.buffer(sz)
        self.weight_aggregators = dev.buffer(b * WeightAggregator.itemsize)
        self.finalized_samples = dev.buffer(b * 16)
        self.sample_counts = dev.buffer(b * 4)
        self.sample_weights = dev.buffer(b * 4)
        self.out_light_image = dev.buffer(b * 16)
        self.out_light_paths = dev.buffer(b * Path.itemsize)
        self.out_light_debug_image = dev.buffer(b * 16)
        self.assign_indices()

    def assign_indices(self):
        self.indices_buffer = self.device.buffer(np.arange(self.batch_size, dtype=np.uint32))
        self.summed_bins_buffer = self.device.buffer(np.arange(self.batch_size + 1, dtype=np.uint32))

    def light_bins(self):
        """Host binning of the sorted light entries (renderer.py:97-111)."""
        idx = self.out_light_indices.numpy(np.int32)
        bins = np.bincount(idx[idx >= 0], minlength=self.pixel_height * self.pixel_width)
        summed = self.device.buffer(np.insert(np.cumsum(bins), 0, 0).astype(np.uint32))
        offset = np.sum(idx < 0).astype(np.uint32)
        return summed, offset

    def make_light_rays(self):
        s = self.scene
        self.light_ray_fn(self.batch_size, s.light_triangles, s.light_surface_areas, s.light_triangle_indices,
                          s.materials, self.rand_buffer, self.light_ray_buffer, s.light_counts)

    def make_camera_rays(self):
        self.camera_ray_fn(self.batch_size, self.scene.camera, self.rand_buffer, self.indices_buffer,
                           self.camera_ray_buffer)

    def trace_camera_rays(self):
        s = self.scene
        self.trace_fn(self.batch_size, self.camera_ray_buffer, s.boxes, s.triangles, s.materials, self.rand_buffer,
                      self.out_camera_image, self.out_camera_paths, self.out_camera_debug_image)

    def trace_light_rays(self):
        s = self.scene
        self.trace_fn(self.batch_size, self.light_ray_buffer, s.boxes, s.triangles, s.materials, self.rand_buffer,
                      self.out_light_image, self.out_light_paths, self.out_light_debug_image)

    def join_paths(self):
        s = self.scene
        n = next_power_of_two(self.batch_size * MAX_PATH_LENGTH)
        self.light_reset_fn(n, self.out_light_indices, self.out_light_path_indices, self.out_light_ray_indices,
                            self.out_light_weights, self.out_light_shade)
        self.join_fn(self.batch_size, self.out_camera_paths, self.out_light_paths, s.triangles, s.materials, s.boxes,
                     s.camera, self.weight_aggregators, self.out_samples, self.out_light_indices,
                     self.out_light_path_indices, self.out_light_ray_indices, self.out_light_weights,
                     self.out_light_shade)

    def finalize_samples(self):
        self.finalize_fn(self.batch_size, self.weight_aggregators, self.scene.camera, self.finalized_samples,
                         self.sample_counts, self.summed_bins_buffer, self.sample_weights)

    def gather_light_image(self):
        n = next_power_of_two(self.batch_size * MAX_PATH_LENGTH)
        log_n = int(np.log2(n))
        pairs_per_thread = 4
        for stage in range(1, log_n + 1):
            for pass_of_stage in range(stage, 0, -1):
                self.light_sort_fn(n // 8, self.out_light_indices, self.out_light_path_indices,
                                   self.out_light_ray_indices, self.out_light_weights, self.out_light_shade,
                                   np.uint32(stage), np.uint32(pass_of_stage), np.uint32(n),
                                   np.uint32(pairs_per_thread))
        bins, offset = self.light_bins()
        self.light_image_gather_fn(self.batch_size, self.out_light_paths, self.scene.materials,
                                   self.out_light_path_indices, self.out_light_ray_indices, bins, offset,
                                   self.out_light_weights, self.out_light_shade, self.out_light_image,
                                   self.sample_weights)
        self.device.synchronize()
        runtime.release(bins)

    def process_images(self):
        h, w = self.pixel_height, self.pixel_width
        finalized = self.finalized_samples.numpy(np.float32).reshape(h, w, 4)[:, :, :3]
        light = self.out_light_image.numpy(np.float32).reshape(h, w, 4)[:, :, :3]
        counts = self.sample_counts.numpy(np.int32).reshape(h, w, 1)
        weights = self.sample_weights.numpy(np.float32).reshape(h, w, 1)
        with np.errstate(invalid="ignore", over="ignore"):
            image = light + finalized
        self.summed_image += np.nan_to_num(image, posinf=0, neginf=0)
        self.summed_sample_counts += counts
        self.summed_sample_weights += weights
        uni = self.out_camera_image.numpy(np.float32).reshape(h, w, 4)[:, :, :3]
        self.unidirectional_image_buffer += np.nan_to_num(uni, posinf=0, neginf=0)

    # ================================================================================================================
    # images (renderer.py:293-316)
    # ================================================================================================================
    def __getattr__(self, name):
        # fused mode: the four accumulators are views of the packed device buffer, fetched on demand
        if name in ("summed_image", "summed_sample_counts", "summed_sample_weights", "unidirectional_image_buffer") \
                and self.__dict__.get("fused"):
            a = self._accum()
            if name == "summed_image":
                return a[:, :, 0:3]
            if name == "summed_sample_weights":
                return a[:, :, 3:4]
            if name == "unidirectional_image_buffer":
                return a[:, :, 4:7]
            return a[:, :, 7:8].astype(np.int32)
        raise AttributeError(name)

    @property
    def image(self):
        with np.errstate(invalid="ignore", divide="ignore"):
            est = self.summed_image / self.summed_sample_weights
        return tone_map(np.nan_to_num(est, neginf=0, posinf=0), exposure=TONE_MAP_EXPOSURE)

    @property
    def unweighted_image(self):
        return tone_map(np.nan_to_num(self.summed_image, neginf=0, posinf=0), exposure=TONE_MAP_EXPOSURE)

    @property
    def unidirectional_image(self):
        with np.errstate(invalid="ignore", divide="ignore"):
            est = self.unidirectional_image_buffer / self.summed_sample_counts
        return tone_map(np.nan_to_num(est, neginf=0, posinf=0), exposure=TONE_MAP_EXPOSURE)

    def estimate(self):
        """float32 (H,W,3) radiance estimate `summed_image / summed_sample_weights` before tone mapping."""
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.nan_to_num(self.summed_image / self.summed_sample_weights, neginf=0, posinf=0)

    def release(self):
        if self.__dict__.get("fused") and self.__dict__.get("_fr") is not None:
            self._fr.release()
            self._fr = None
        for v in list(self.__dict__.values()):
            if isinstance(v, runtime.Buffer):
                v.release()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
