/* ref_glue.c — the reference's per-pass dispatcher (src/renderer.py:16-291: buffers, stage order, bitonic schedule, host
 * binning, accumulation) restated in C over the eight kernel entry points `ref_<kernel>(n, buffers...)`.
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Linked twice: with oracle/trace_ref.c (the C restatement of trace.metal ->
 * _build/libclive_oracle*.so) and with oracle/ref_build/ref_driver.cpp (the reference's own trace.metal compiled as C++
 * -> _ref/libclive_ref*.so), so that both run the very same host glue around their kernels. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ref_types.h"

/* the eight kernels, reference positional argument order, one call = all thread ids 0..n-1 */
void ref_generate_light_rays(uint32_t n, const Triangle *light_triangles, const float *surface_areas,
                             const int32_t *light_triangle_indices, const Material *materials, uint32_t *random_buffer,
                             Ray *out, const int32_t *counts);
void ref_generate_camera_rays(uint32_t n, const Camera *camera, uint32_t *random_buffer, const uint32_t *indices, Ray *out);
void ref_generate_paths(uint32_t n, const Ray *rays, const Box *boxes, const Triangle *triangles, const Material *materials,
                        uint32_t *random_buffer, float4 *out, Path *output_paths, float4 *float_debug);
void ref_reset_light_indices(uint32_t n, int32_t *light_pixel_indices, int32_t *light_path_indices,
                             int32_t *light_ray_indices, float *light_weights, float *light_shade);
void ref_connect_paths(uint32_t n, const Path *camera_paths, const Path *light_paths, const Triangle *triangles,
                       const Material *materials, const Box *boxes, const Camera *camera,
                       WeightAggregator *weight_aggregators, float4 *out, int32_t *light_pixel_indices,
                       int32_t *light_path_indices, int32_t *light_ray_indices, float *light_weights, float *light_shade);
void ref_adaptive_finalize_samples(uint32_t n, const WeightAggregator *weight_aggregators, const Camera *camera,
                                   float4 *out, uint32_t *sample_counts, const uint32_t *sample_bin_offsets,
                                   float *sample_weights);
void ref_light_sort(uint32_t n_threads, int32_t *light_pixel_indices, int32_t *light_path_indices,
                    int32_t *light_ray_indices, float *light_weights, float *light_shade, uint32_t stage,
                    uint32_t pass_of_stage, uint32_t n, uint32_t pairs_per_thread);
void ref_light_image_gather(uint32_t n, const Path *light_paths, const Material *materials, const int32_t *path_indices,
                            const int32_t *ray_indices, const int32_t *bins, const uint32_t *offset, const float *weights,
                            const float *shades, float4 *light_image, float *sum_weights);

/* ------------------------------------------------------------------------------------------------------------
 * the per-pass dispatcher: renderer.py:16-291 restated (buffers, stage order, host binning, accumulation)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    /* scene (borrowed pointers, owned by the caller) */
    const Box *boxes;
    const Triangle *triangles;
    const Material *materials;
    const Camera *camera;
    const Triangle *light_triangles;
    const float *light_surface_areas;
    const int32_t *light_triangle_indices;
    int32_t light_count;
    /* renderer.py:47-80 */
    int width, height;
    uint32_t batch, n_light;
    Ray *camera_rays, *light_rays;
    uint32_t *indices, *rand, *summed_bins_identity;
    float4 *out_camera_image, *out_camera_debug, *out_samples, *finalized_samples, *out_light_image, *out_light_debug;
    Path *camera_paths, *light_paths;
    int32_t *light_pixel_indices, *light_path_indices, *light_ray_indices;
    float *light_weights, *light_shade;
    WeightAggregator *aggregators;
    uint32_t *sample_counts;
    float *sample_weights;
    int32_t *bins;
    /* renderer.py:41-45 */
    float *summed_image;            /* H*W*3 */
    int32_t *summed_sample_counts;  /* H*W */
    float *summed_sample_weights;   /* H*W */
    float *unidirectional;          /* H*W*3 */
    int samples;
    double stage_seconds[8];
} RefRenderer;

static uint32_t next_power_of_two(uint32_t n) { /* renderer.py:11-13 */
    uint32_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

EXPORT RefRenderer *ref_renderer_create(const Box *boxes, const Triangle *triangles, const Material *materials,
                                        const Camera *camera, const Triangle *light_triangles,
                                        const float *light_surface_areas, const int32_t *light_triangle_indices,
                                        int32_t light_count, const uint32_t *seeds /* [batch][2] */) {
    RefRenderer *r = (RefRenderer *)calloc(1, sizeof *r);
    r->boxes = boxes; r->triangles = triangles; r->materials = materials; r->camera = camera;
    r->light_triangles = light_triangles; r->light_surface_areas = light_surface_areas;
    r->light_triangle_indices = light_triangle_indices; r->light_count = light_count;
    r->width = camera->pixel_width; r->height = camera->pixel_height;
    uint32_t b = r->batch = (uint32_t)(r->width * r->height);
    uint32_t n = r->n_light = next_power_of_two(b * 8u);
    r->camera_rays = calloc(b, sizeof(Ray)); r->light_rays = calloc(b, sizeof(Ray));
    r->indices = malloc(4u * b); r->rand = malloc(8u * (size_t)b);
    r->summed_bins_identity = malloc(4u * ((size_t)b + 1));
    for (uint32_t i = 0; i < b; i++) r->indices[i] = i;          /* renderer.py:90-94 */
    for (uint32_t i = 0; i <= b; i++) r->summed_bins_identity[i] = i;
    memcpy(r->rand, seeds, 8u * (size_t)b);
    r->out_camera_image = calloc(b, 16); r->out_camera_debug = calloc(b, 16); r->out_samples = calloc(b, 16);
    r->finalized_samples = calloc(b, 16); r->out_light_image = calloc(b, 16); r->out_light_debug = calloc(b, 16);
    r->camera_paths = calloc(b, sizeof(Path)); r->light_paths = calloc(b, sizeof(Path));
    r->light_pixel_indices = calloc(n, 4); r->light_path_indices = calloc(n, 4); r->light_ray_indices = calloc(n, 4);
    r->light_weights = calloc(n, 4); r->light_shade = calloc(n, 4);
    r->aggregators = calloc(b, sizeof(WeightAggregator));
    r->sample_counts = calloc(b, 4); r->sample_weights = calloc(b, 4);
    r->bins = calloc((size_t)b + 2 * (size_t)r->width + 8, 4);
    r->summed_image = calloc((size_t)b * 3, 4); r->summed_sample_counts = calloc(b, 4);
    r->summed_sample_weights = calloc(b, 4); r->unidirectional = calloc((size_t)b * 3, 4);
    return r;
}

EXPORT void ref_renderer_destroy(RefRenderer *r) {
    if (!r) return;
    free(r->camera_rays); free(r->light_rays); free(r->indices); free(r->rand); free(r->summed_bins_identity);
    free(r->out_camera_image); free(r->out_camera_debug); free(r->out_samples); free(r->finalized_samples);
    free(r->out_light_image); free(r->out_light_debug); free(r->camera_paths); free(r->light_paths);
    free(r->light_pixel_indices); free(r->light_path_indices); free(r->light_ray_indices); free(r->light_weights);
    free(r->light_shade); free(r->aggregators); free(r->sample_counts); free(r->sample_weights); free(r->bins);
    free(r->summed_image); free(r->summed_sample_counts); free(r->summed_sample_weights); free(r->unidirectional);
    free(r);
}

/* stage ids match renderer.py:281-291 */
enum { ST_LIGHT_RAYS = 0, ST_CAMERA_RAYS, ST_TRACE_LIGHT, ST_TRACE_CAMERA, ST_JOIN, ST_FINALIZE, ST_GATHER, ST_PROCESS };

static inline float nan_to_num0(float x) { return (isnan(x) || isinf(x)) ? 0.0f : x; } /* renderer.py:271,278 */

EXPORT void ref_renderer_stage(RefRenderer *r, int stage) {
    double t0 = now_s();
    uint32_t b = r->batch, n = r->n_light;
    switch (stage) {
    case ST_LIGHT_RAYS: /* renderer.py:113-125 */
        ref_generate_light_rays(b, r->light_triangles, r->light_surface_areas, r->light_triangle_indices, r->materials,
                                r->rand, r->light_rays, &r->light_count);
        break;
    case ST_CAMERA_RAYS: /* :127-136 */
        ref_generate_camera_rays(b, r->camera, r->rand, r->indices, r->camera_rays);
        break;
    case ST_TRACE_LIGHT: /* :153-166 */
        ref_generate_paths(b, r->light_rays, r->boxes, r->triangles, r->materials, r->rand, r->out_light_image,
                           r->light_paths, r->out_light_debug);
        break;
    case ST_TRACE_CAMERA: /* :138-151 */
        ref_generate_paths(b, r->camera_rays, r->boxes, r->triangles, r->materials, r->rand, r->out_camera_image,
                           r->camera_paths, r->out_camera_debug);
        break;
    case ST_JOIN: /* :168-197 */
        ref_reset_light_indices(n, r->light_pixel_indices, r->light_path_indices, r->light_ray_indices,
                                r->light_weights, r->light_shade);
        ref_connect_paths(b, r->camera_paths, r->light_paths, r->triangles, r->materials, r->boxes, r->camera,
                          r->aggregators, r->out_samples, r->light_pixel_indices, r->light_path_indices,
                          r->light_ray_indices, r->light_weights, r->light_shade);
        break;
    case ST_FINALIZE: /* :199-210 */
        ref_adaptive_finalize_samples(b, r->aggregators, r->camera, r->finalized_samples, r->sample_counts,
                                      r->summed_bins_identity, r->sample_weights);
        break;
    case ST_GATHER: { /* :212-250 with the host binning of :96-111 */
        uint32_t log_n = 0;
        while ((1u << log_n) < n) log_n++;
        for (uint32_t st = 1; st <= log_n; st++)
            for (uint32_t ps = st; ps >= 1; ps--)
                ref_light_sort(n / 8, r->light_pixel_indices, r->light_path_indices, r->light_ray_indices,
                               r->light_weights, r->light_shade, st, ps, n, 4);
        /* light_bins(): bincount of keys >= 0 (minlength batch), exclusive cumsum, offset = number of negatives */
        size_t nb = (size_t)b + 2 * (size_t)r->width + 8;
        memset(r->bins, 0, 4 * nb);
        uint32_t offset = 0;
        for (uint32_t i = 0; i < n; i++) {
            int32_t k = r->light_pixel_indices[i];
            if (k < 0) offset++;
            else if ((size_t)k + 1 < nb) r->bins[k + 1]++;
        }
        for (size_t i = 1; i < nb; i++) r->bins[i] += r->bins[i - 1];
        ref_light_image_gather(b, r->light_paths, r->materials, r->light_path_indices, r->light_ray_indices, r->bins,
                               &offset, r->light_weights, r->light_shade, r->out_light_image, r->sample_weights);
        break;
    }
    case ST_PROCESS: /* :252-278 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)b; i++) {
            const float *fin = &r->finalized_samples[i].x, *li = &r->out_light_image[i].x, *uni = &r->out_camera_image[i].x;
            for (int ch = 0; ch < 3; ch++) {
                r->summed_image[3 * i + ch] += nan_to_num0(li[ch] + fin[ch]);
                r->unidirectional[3 * i + ch] += nan_to_num0(uni[ch]);
            }
            r->summed_sample_counts[i] += (int32_t)r->sample_counts[i];
            r->summed_sample_weights[i] += r->sample_weights[i];
        }
        break;
    default: break;
    }
    if (stage >= 0 && stage < 8) r->stage_seconds[stage] += now_s() - t0;
}

EXPORT void ref_renderer_run_sample(RefRenderer *r) { /* renderer.py:280-291 */
    for (int s = 0; s < 8; s++) ref_renderer_stage(r, s);
    r->samples++;
}

/* buffer access for differential tests: returns a pointer into the renderer and its size in bytes */
EXPORT void *ref_renderer_buffer(RefRenderer *r, const char *name, uint64_t *nbytes) {
    uint64_t b = r->batch, n = r->n_light;
#define BUF(s, p, sz) if (!strcmp(name, s)) { *nbytes = (sz); return (void *)(p); }
    BUF("camera_rays", r->camera_rays, b * sizeof(Ray)) BUF("light_rays", r->light_rays, b * sizeof(Ray))
    BUF("rand", r->rand, b * 8) BUF("indices", r->indices, b * 4)
    BUF("out_camera_image", r->out_camera_image, b * 16) BUF("out_light_image", r->out_light_image, b * 16)
    BUF("out_samples", r->out_samples, b * 16) BUF("finalized_samples", r->finalized_samples, b * 16)
    BUF("camera_paths", r->camera_paths, b * sizeof(Path)) BUF("light_paths", r->light_paths, b * sizeof(Path))
    BUF("light_pixel_indices", r->light_pixel_indices, n * 4) BUF("light_path_indices", r->light_path_indices, n * 4)
    BUF("light_ray_indices", r->light_ray_indices, n * 4) BUF("light_weights", r->light_weights, n * 4)
    BUF("light_shade", r->light_shade, n * 4) BUF("aggregators", r->aggregators, b * sizeof(WeightAggregator))
    BUF("sample_counts", r->sample_counts, b * 4) BUF("sample_weights", r->sample_weights, b * 4)
    BUF("summed_image", r->summed_image, b * 12) BUF("summed_sample_counts", r->summed_sample_counts, b * 4)
    BUF("summed_sample_weights", r->summed_sample_weights, b * 4) BUF("unidirectional", r->unidirectional, b * 12)
    BUF("stage_seconds", r->stage_seconds, sizeof r->stage_seconds)
#undef BUF
    *nbytes = 0;
    return NULL;
}

