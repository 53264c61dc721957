/* clive_b200.h — C ABI of libclive_b200.so: the B200 (sm_100a) replacement for Clive2's Metal dispatch.
 *
 * The reference drives its GPU through the third-party `metalcompute` Python module (src/renderer.py:1,
 * src/scene.py:3): `Device()`, `dev.buffer(nbytes | array)`, `dev.kernel(msl).function(name)`,
 * `fn(thread_count, *buffers)` and `release()`.  That object API is the only thing between the reference's host code
 * and its eight Metal kernels, so this header is what a binding for the same path has to export:
 *
 *   section 1  record layouts                      = src/struct_types.py:4-85 == src/trace.metal:7-85
 *   section 2  context + device memory             = metalcompute.Device / .buffer / release (scene.py:29,74-82; renderer.py:47-80,318-352)
 *   section 3  acceleration structure              = the Box[]/Triangle[] pair emitted by bvh.py:329-389, re-laid-out for sm_100a
 *   section 4  the eight kernels, argument for argument (renderer.py:113-250)
 *   section 5  fused per-pass renderer             = Renderer.run_sample() (renderer.py:281-291) without host round trips
 *   section 6  parity / measurement hooks
 *
 * Conventions: every function returns 0 on success and a negative clv_status on failure; the message of the last
 * failure on the calling thread is available from clv_last_error().  No exceptions cross the boundary.  Host pointers
 * are borrowed for the duration of the call only.  "d_" parameters are device pointers obtained from clv_mem_alloc
 * (or any CUDA allocation on the context's device).  Kernel launches are asynchronous on the context's stream;
 * clv_mem_download and clv_ctx_synchronize wait for them.  A context is bound to one GPU and is not thread-safe;
 * different contexts are independent (one per GPU / per process).
 */
#ifndef CLIVE_B200_H
#define CLIVE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CLV_API __attribute__((visibility("default")))
#else
#define CLV_API
#endif

/* ---- 1. record layouts (float3 members occupy 16 bytes) ------------------------------------------------------ */
typedef struct clv_Ray { /* trace.metal:7-23, struct_types.py:4-22; 128 bytes */
    float origin[4], direction[4], inv_direction[4], color[4], normal[4];
    int32_t material, triangle;
    float c_importance, l_importance, tot_importance;
    int32_t hit_light, from_camera, hit_camera, pixel_idx;
    int32_t pad[3];
} clv_Ray;

typedef struct clv_Path { /* trace.metal:31-36; 1040 bytes */
    clv_Ray rays[8];
    int32_t length, from_camera, pad[2];
} clv_Path;

typedef struct clv_Box { /* trace.metal:38-44; 48 bytes.  right == 0: inner node, children at left, left+1 */
    float min[4], max[4];
    int32_t left, right, pad[2];
} clv_Box;

typedef struct clv_Triangle { /* trace.metal:47-59; 128 bytes */
    float v0[4], v1[4], v2[4], n0[4], n1[4], n2[4], normal[4];
    int32_t material, is_light, is_camera, pad;
} clv_Triangle;

typedef struct clv_Material { /* trace.metal:62-69; 48 bytes */
    float color[4], emission[4];
    int32_t type;
    float alpha, ior;
    int32_t pad;
} clv_Material;

typedef struct clv_Camera { /* trace.metal:72-85; 112 bytes */
    float center[4], focal_point[4], direction[4], dx[4], dy[4];
    int32_t pixel_width, pixel_height;
    float phys_width, phys_height, h_fov, v_fov;
    int32_t pad[2];
} clv_Camera;

typedef struct clv_WeightAggregator { /* trace.metal:25-29; 80 bytes (the reference allocates 128 per element) */
    float weights[3][3];
    float pad0[3];
    float total_contribution[4];
    float contrib_weight_sum;
    float pad1[3];
} clv_WeightAggregator;

typedef enum clv_status {
    CLV_OK = 0,
    CLV_ERR_INVALID = -1, /* bad argument */
    CLV_ERR_CUDA = -2,    /* CUDA runtime error, see clv_last_error() */
    CLV_ERR_NOMEM = -3,
    CLV_ERR_NO_DEVICE = -4
} clv_status;

CLV_API const char *clv_last_error(void);
CLV_API const char *clv_version(void);

/* ---- 2. context + device memory ------------------------------------------------------------------------------ */
typedef struct clv_ctx clv_ctx;

CLV_API int clv_device_count(int *count);
/* replaces metalcompute.Device() (scene.py:29): binds GPU `device`, creates the launch stream */
CLV_API int clv_ctx_create(int device, clv_ctx **out);
CLV_API int clv_ctx_destroy(clv_ctx *ctx);
/* replaces `del handle` after each dispatch (renderer.py:125 etc.): waits for everything launched so far */
CLV_API int clv_ctx_synchronize(clv_ctx *ctx);
/* the CUDA stream kernels are launched on, as a cudaStream_t cast to void* (for event timing by the caller) */
CLV_API int clv_ctx_stream(clv_ctx *ctx, void **stream);

/* replace dev.buffer(nbytes) / dev.buffer(array) / metalcompute.release (renderer.py:51-80; scene.py:74-82) */
CLV_API int clv_mem_alloc(clv_ctx *ctx, size_t nbytes, void **d_ptr); /* zero filled */
CLV_API int clv_mem_free(clv_ctx *ctx, void *d_ptr);
CLV_API int clv_mem_upload(clv_ctx *ctx, void *d_dst, const void *h_src, size_t nbytes);
CLV_API int clv_mem_download(clv_ctx *ctx, void *h_dst, const void *d_src, size_t nbytes); /* synchronises first */
CLV_API int clv_mem_fill32(clv_ctx *ctx, void *d_dst, uint32_t value, size_t count);
/* page-locked host staging memory for the end-to-end path */
CLV_API int clv_host_alloc(size_t nbytes, void **h_ptr);
CLV_API int clv_host_free(void *h_ptr);

/* ---- 3. acceleration structure --------------------------------------------------------------------------------- */
/* Built once from the flattened BVH of bvh.py:329-389 (borrowed host arrays).  Keeps a device copy of the reference
 * arrays (shading data, reference-order traversal) and the compact traversal layout. */
typedef struct clv_accel clv_accel;
CLV_API int clv_accel_create(clv_ctx *ctx, const clv_Box *h_boxes, int32_t n_boxes, const clv_Triangle *h_triangles,
                             int32_t n_triangles, clv_accel **out);
/* Reference counted: clv_renderer_create takes a reference and clv_renderer_destroy drops it, so destroying the
 * structure (or the scene that owns it) while a renderer still traverses it only defers the free. */
CLV_API int clv_accel_destroy(clv_accel *accel);
/* the compact layout: [0] bytes of binary node records, [1] bytes of triangle records, [2] binary node count,
 * [3] triangle count, [4] bytes of 4-wide node records, [5] 4-wide node count, [6]/[7] 1 if closest-hit / occlusion
 * searches of the ORDERED flavour traverse the 4-wide nodes, [8] 1 if the records exceed the cache-resident size and the
 * fused pass therefore runs its high-occupancy kernel flavours (same results, different register budget), [9] 1 if the
 * wavefront trace kernel steps over the 4-wide nodes for this structure, [10]/[11] inner levels of the binary / 4-wide
 * tree.  A tree deeper than the traversal stacks allow (40 binary / 21 4-wide levels) is refused by clv_accel_create with
 * CLV_ERR_INVALID — the kernels never drop a push (the reference's own 64-entry stack silently aborts the search,
 * trace.metal:149) */
CLV_API int clv_accel_stats(clv_accel *accel, uint64_t out[12]);

/* ---- 4. the eight kernels (trace.metal entry points), reference argument order, device pointers ----------------- */
/* generate_light_rays, trace.metal:1070-1124 / renderer.py:114-125 */
CLV_API int clv_generate_light_rays(clv_ctx *ctx, uint32_t n, const clv_Triangle *d_light_triangles,
                                    const float *d_surface_areas, const int32_t *d_light_triangle_indices,
                                    const clv_Material *d_materials, uint32_t *d_random_buffer, clv_Ray *d_out,
                                    const int32_t *d_counts);
/* generate_camera_rays, trace.metal:1020-1067 / renderer.py:128-136 */
CLV_API int clv_generate_camera_rays(clv_ctx *ctx, uint32_t n, const clv_Camera *d_camera, uint32_t *d_random_buffer,
                                     const uint32_t *d_indices, clv_Ray *d_out);
/* generate_paths, trace.metal:381-532 / renderer.py:139-166; `accel` stands for the (boxes, triangles) buffer pair */
CLV_API int clv_generate_paths(clv_ctx *ctx, uint32_t n, const clv_Ray *d_rays, const clv_accel *accel,
                               const clv_Material *d_materials, uint32_t *d_random_buffer, float *d_out /*float4[n]*/,
                               clv_Path *d_output_paths, float *d_float_debug /*float4[n]*/);
/* reset_light_indices, trace.metal:967-978 / renderer.py:171-179 */
CLV_API int clv_reset_light_indices(clv_ctx *ctx, uint32_t n, int32_t *d_light_pixel_indices,
                                    int32_t *d_light_path_indices, int32_t *d_light_ray_indices,
                                    float *d_light_weights, float *d_light_shade);
/* connect_paths, trace.metal:620-869 / renderer.py:181-197 */
CLV_API int clv_connect_paths(clv_ctx *ctx, uint32_t n, const clv_Path *d_camera_paths, const clv_Path *d_light_paths,
                              const clv_accel *accel, const clv_Material *d_materials, const clv_Camera *d_camera,
                              clv_WeightAggregator *d_weight_aggregators, float *d_out /*float4[n]*/,
                              int32_t *d_light_pixel_indices, int32_t *d_light_path_indices,
                              int32_t *d_light_ray_indices, float *d_light_weights, float *d_light_shade);
/* adaptive_finalize_samples, trace.metal:981-1018 / renderer.py:200-210 */
CLV_API int clv_adaptive_finalize_samples(clv_ctx *ctx, uint32_t n, const clv_WeightAggregator *d_weight_aggregators,
                                          const clv_Camera *d_camera, float *d_out /*float4[n]*/,
                                          uint32_t *d_sample_counts, const uint32_t *d_sample_bin_offsets,
                                          float *d_sample_weights);
/* light_sort (one bitonic step), trace.metal:872-934 / renderer.py:213-231 */
CLV_API int clv_light_sort(clv_ctx *ctx, uint32_t n_threads, int32_t *d_light_pixel_indices,
                           int32_t *d_light_path_indices, int32_t *d_light_ray_indices, float *d_light_weights,
                           float *d_light_shade, uint32_t stage, uint32_t pass_of_stage, uint32_t n,
                           uint32_t pairs_per_thread);
/* light_image_gather, trace.metal:937-964 / renderer.py:235-248 */
CLV_API int clv_light_image_gather(clv_ctx *ctx, uint32_t n, const clv_Path *d_light_paths,
                                   const clv_Material *d_materials, const int32_t *d_path_indices,
                                   const int32_t *d_ray_indices, const int32_t *d_bins, uint32_t offset,
                                   const float *d_weights, const float *d_shades, float *d_light_image /*float4[n]*/,
                                   float *d_sum_weights);
/* Device-side replacement of the host binning in Renderer.light_bins (renderer.py:97-111) followed by the gather:
 * scatters the unsorted light entries written by connect_paths straight into the light image with atomics, so the
 * 190-300 bitonic dispatches and the device->host->device round trip disappear.  Same sums, different float order. */
CLV_API int clv_light_image_splat(clv_ctx *ctx, uint32_t n_entries, uint32_t n_pixels, const clv_Path *d_light_paths,
                                  const clv_Material *d_materials, const int32_t *d_light_pixel_indices,
                                  const int32_t *d_path_indices, const int32_t *d_ray_indices, const float *d_weights,
                                  const float *d_shades, float *d_light_image /*float4[n_pixels]*/,
                                  float *d_sum_weights);

/* ---- 5. fused per-pass renderer --------------------------------------------------------------------------------- */
typedef struct clv_renderer clv_renderer;

typedef struct clv_scene_desc { /* borrowed host arrays = the Scene attribute bag of scene.py:107-134 */
    const clv_Material *materials;
    int32_t n_materials;
    const clv_Camera *camera; /* one record */
    const clv_Triangle *light_triangles;
    const float *light_surface_areas;
    const int32_t *light_triangle_indices;
    int32_t n_lights;
} clv_scene_desc;

/* seeds: uint32[W*H][2] (renderer.py:86-87), indexed by GLOBAL sample id.  sample_ids: the ids this renderer traces
 * each pass (NULL = all W*H, the reference's identity `indices`, renderer.py:90-94); a subset gives a tile/stripe
 * shard whose RNG streams equal those of the full render. */
CLV_API int clv_renderer_create(clv_ctx *ctx, const clv_accel *accel, const clv_scene_desc *scene,
                                const uint32_t *h_seeds, const uint32_t *h_sample_ids, uint32_t n_local,
                                clv_renderer **out);
CLV_API int clv_renderer_destroy(clv_renderer *r);
/* one sample pass (= Renderer.run_sample, renderer.py:281-291): five kernels on the context's stream, asynchronous,
 * accumulates on the device */
CLV_API int clv_renderer_run_sample(clv_renderer *r);
/* end-to-end variant with HOST buffers: uploads a fresh seed buffer uint32[W*H][2] for this pass (NULL = keep the
 * device-resident streams), runs the pass, then downloads the packed accumulators float[H][W][8] (NULL = skip) and
 * waits.  Page-locked memory from clv_host_alloc makes both copies asynchronous DMA. */
CLV_API int clv_renderer_run_sample_e2e(clv_renderer *r, const uint32_t *h_seeds, float *h_accum_packed);
CLV_API int clv_renderer_reset(clv_renderer *r, const uint32_t *h_seeds); /* zero accumulators, reload seeds */
/* Renderer.image (renderer.py:293-300 + camera.py:73-82) on the device: tone-mapped uint8 BGR [H][W][3] of the current
 * estimate summed_image / summed_sample_weights, downloaded into h_out_bgr (waits) */
CLV_API int clv_renderer_tone_map(clv_renderer *r, float exposure, uint8_t *h_out_bgr);
/* the interactive loop body of render.py:31-37 with HOST buffers: upload a seed buffer (NULL = keep the device
 * streams), run one pass, tone map, download the uint8 preview (waits) */
CLV_API int clv_renderer_run_sample_preview(clv_renderer *r, const uint32_t *h_seeds, float exposure, uint8_t *h_image_bgr);

typedef enum clv_buffer_id {
    CLV_BUF_SUMMED_IMAGE = 0,          /* float[H][W][3]   renderer.py:41  */
    CLV_BUF_SUMMED_SAMPLE_COUNTS = 1,  /* int32[H][W]      renderer.py:42  */
    CLV_BUF_SUMMED_SAMPLE_WEIGHTS = 2, /* float[H][W]      renderer.py:43  */
    CLV_BUF_UNIDIRECTIONAL = 3,        /* float[H][W][3]   renderer.py:45  */
    CLV_BUF_SEEDS = 4,                 /* uint32[W*H][2]                   */
    CLV_BUF_ACCUM_PACKED = 5,          /* float[H][W][8]: image3, weight, uni3, count — the accumulators as stored
                                          on the device, and the multi-GPU reduce payload */
    CLV_BUF_LIGHT_PATHS = 6,           /* Path[n_local]  (last pass)  renderer.py:79; needs clv_renderer_set_path_records */
    CLV_BUF_CAMERA_PATHS = 7,          /* Path[n_local]  (last pass)  renderer.py:58; needs clv_renderer_set_path_records */
    CLV_BUF_AGGREGATORS = 8,           /* WeightAggregator[W*H] (last pass, indexed by pixel) */
    CLV_BUF_PASS_WEIGHT = 9            /* float: running sum over passes and pixels of the per-pass sample weights */
} clv_buffer_id;
CLV_API int clv_renderer_read(clv_renderer *r, int buffer_id, void *h_dst, size_t nbytes); /* synchronises */
/* device pointer + size of a buffer (ids >= 4), e.g. to hand the packed accumulators to a collective in place */
CLV_API int clv_renderer_device_ptr(clv_renderer *r, int buffer_id, void **d_ptr, size_t *nbytes);

/* ---- 6. parity and measurement hooks ------------------------------------------------------------------------------ */
typedef enum clv_traverse_mode {
    CLV_TRAVERSE_REFERENCE_ORDER = 0, /* trace.metal:144-176 verbatim on the reference arrays */
    CLV_TRAVERSE_ORDERED = 1          /* compact layout, near child first, (t, rank) tie rule */
} clv_traverse_mode;
/* closest hit for n host rays; outputs are host arrays of n elements (triangle index in the REFERENCE numbering) */
CLV_API int clv_traverse(clv_ctx *ctx, const clv_accel *accel, const clv_Ray *h_rays, uint32_t n, int mode,
                         int32_t *h_tri, float *h_t, float *h_u, float *h_v);
/* the same kernel on device-resident rays/outputs, asynchronous (for timing the traversal alone) */
CLV_API int clv_traverse_device(clv_ctx *ctx, const clv_accel *accel, const clv_Ray *d_rays, uint32_t n, int mode,
                                int32_t *d_tri, float *d_t, float *d_u, float *d_v);
/* visibility_test (trace.metal:178-197) for n host vertex pairs a -> b */
CLV_API int clv_visibility(clv_ctx *ctx, const clv_accel *accel, const clv_Ray *h_a, const clv_Ray *h_b, uint32_t n,
                           int mode, int32_t *h_visible);
/* traversal flavour used by the stage-level entry points of section 4 (process-wide; default ORDERED) and by one
 * fused renderer (default ORDERED).  REFERENCE_ORDER visits boxes exactly like trace.metal:144-176. */
CLV_API int clv_set_stage_traverse_mode(int mode);
CLV_API int clv_renderer_set_traverse_mode(clv_renderer *r, int mode);
/* By default the fused renderer keeps the vertices of a pass in its own compact SoA records (64 B per vertex, coalesced);
 * with path records ON it writes and reads the reference's Path records (renderer.py:58,79) instead, so that
 * CLV_BUF_LIGHT_PATHS / CLV_BUF_CAMERA_PATHS can be read back and compared.  Results are the same bits either way. */
CLV_API int clv_renderer_set_path_records(clv_renderer *r, int enabled);
/* work counters accumulated by kernels launched while counting is enabled (instrumented variants; the timed kernels
 * carry no counter code).  Two classes of five words: out[0..4] random-walk / traversal kernels, out[5..9] connection
 * kernels; per class: rays = traverse_bvh invocations the reference would make, traversals actually started,
 * 64-byte node records fetched, 48-byte triangle records fetched, 128-byte path vertices written / staged */
CLV_API int clv_counters_enable(clv_ctx *ctx, int enabled);
CLV_API int clv_counters_read(clv_ctx *ctx, uint64_t out[10], int reset);
/* per-kernel device time of the fused pass: CUDA events recorded on the launch stream after every launch while
 * enabled (the pass then runs as one slice on one stream); `kernel_times` waits and returns the milliseconds accumulated
 * per kernel — slots: 0 k_fused_walks (megakernel walks), 1 k_connect_enumerate, 2 k_connect_rays, 3 k_connect_finish,
 * 4 k_fused_finalize, 5 k_walk_trace (all rounds), 6 k_walk_init + k_walk_shade (all rounds), 7 k_walk_pipeline — and the
 * number of passes they cover */
#define CLV_TIMED_KERNEL_SLOTS 8
CLV_API int clv_renderer_kernel_timing(clv_renderer *r, int enabled);
CLV_API int clv_renderer_kernel_times(clv_renderer *r, double ms_out[CLV_TIMED_KERNEL_SLOTS], uint64_t *passes_out, int reset);
/* Stages 1-4 of the fused pass (ray generation + both random walks; renderer.py:281-285) run either as ONE megakernel,
 * a thread per sample (k_fused_walks), or as wavefront rounds: k_walk_init, then 12 (+ 4) x (k_walk_trace: persistent warps
 * that fetch rays from the round's queue as their lanes finish; k_walk_shade: the bounce of every traced ray; a search
 * that outlasts its round is suspended and resumed first thing in the next one, up to 4 times per sample).  Same
 * results, bit for bit.  mode 0 = automatic (wavefront for scenes whose traversal records exceed the caches when a pass of
 * the renderer has at least 600 Ki samples; env
 * CLV_WALK_MODE), 1 = megakernel, 2 = wavefront rounds, 3 = wavefront as ONE pipelined persistent kernel (k_walk_pipeline:
 * warps trace rays from a queue and, when they hold none, shade traced samples from a second queue — no round
 * boundaries); < 0 only queries. */
CLV_API int clv_ctx_walk_mode(clv_ctx *ctx, int mode, int *current_out);
/* Scheduling knobs of the two ray kernels, for measurement sweeps (none changes a result): "walk_refill_below" /
 * "connect_refill_below" = lanes still holding a ray below which a warp of k_walk_trace / k_connect_rays fetches new
 * rays (1..32); "walk_schedule" / "connect_schedule" = vote weight w, an inner-node step wins when
 * lanes_at_inner * w >= lanes_at_leaf * 4 (connect_schedule 0 = round form; defaults: 4 where the leaves are tested by
 * the whole warp — k_walk_trace and the large-scene flavour of k_connect_rays — 8 for the cache-resident flavour;
 * setting connect_schedule sets both flavours, a query returns the cache-resident one); "connect_dynamic_fetch" = 1: k_connect_rays as
 * persistent warps that fetch tasks with one atomicAdd per refill, 0: static 512-task chunks per warp; "walk_grid_divisor" =
 * k_walk_trace's persistent grid is the resident blocks / divisor; "walk_tail_budget" = iterations a warp of k_walk_trace
 * keeps stepping after its round's queue ran dry before it suspends the searches it still holds into continuation records
 * (0 = never; default 64); "finish_cooperative", "pipe_shade_every": see DESIGN.md section 5.  value < 0 only queries. */
CLV_API int clv_ctx_set_knob(clv_ctx *ctx, const char *name, int value, int *previous);
/* streaming-read bandwidth (GB/s) of a scratch buffer of nbytes read iters times: the L2 read peak for nbytes well
 * below the 126 MB L2 (the bound of the cache-resident configurations), the HBM read peak for nbytes >> L2 */
CLV_API int clv_measure_read_bandwidth(clv_ctx *ctx, size_t nbytes, int iters, double *gbs_out);
/* number of kernel launches issued on this context so far */
CLV_API int clv_launch_count(clv_ctx *ctx, uint64_t *count);
/* The fused pass runs the walk and connection kernels as `slices` independent slices of the samples on `slices`
 * streams and joins them before the finalize kernel (4 * slices + 1 launches per pass): blocks of one slice's kernel
 * fill the SMs that drain at the end of another's.  1..8 = fixed; 0 = automatic (the default; env CLV_PASS_SLICES):
 * 3 slices when a pass has at least one full wave of samples (1024 per SM), else 1.  Results do not depend on it (the
 * atomic summation order of the light image aside).  `slices` < 0 only queries the configured value. */
CLV_API int clv_ctx_pass_slices(clv_ctx *ctx, int slices, int *current_out);

#ifdef __cplusplus
}
#endif
#endif /* CLIVE_B200_H */
