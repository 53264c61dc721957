// ref_driver.cpp — compiles the reference's src/trace.metal, UNMODIFIED and where it lies, as C++ and exposes its eight
// kernels behind the same C entry points as oracle/trace_ref.c (ref_generate_paths(n, ...) = `for id in range(n)` over
// the kernel, OpenMP over ids).  TEST INFRASTRUCTURE: pins the oracle (tests/test_ref_pin.py) and is the CPU arm of
// bench.py (cpu_baseline.kind = "reference").  Built by `make -C oracle ref` into oracle/_ref/ (git-ignored).
//
//   g++ -x c++ -std=c++17 -O3 -fopenmp -fsingle-precision-constant -ffp-contract=off -ftrivial-auto-var-init=zero
//       -Wno-attributes -Ioracle/ref_build -DTRACE_METAL='"/root/reference/src/trace.metal"' ref_driver.cpp
//
// The Metal attributes ([[buffer(n)]], [[thread_position_in_grid]]) are unknown to g++ and ignored (-Wno-attributes),
// so `id` is an ordinary trailing argument.  One reference behaviour is not reproduced: generate_light_rays reads one
// element past its three light arrays when the RNG returns exactly 1.0 (trace.metal:1087-1089, probability ~3e-8 per
// draw); the wrappers pass copies padded with a duplicate of the last element, i.e. the out-of-bounds read returns
// what oracle/trace_ref.c's documented clamp returns.
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <metal_stdlib> /* ref_build/metal_stdlib: after the system headers (it may define sin(x)... as macros) */

#ifndef TRACE_METAL
#error "pass -DTRACE_METAL='\"/root/reference/src/trace.metal\"'"
#endif
#include TRACE_METAL

#define EXPORT extern "C" __attribute__((visibility("default")))

static_assert(sizeof(Ray) == 128 && sizeof(Path) == 1040 && sizeof(Box) == 48 && sizeof(Triangle) == 128 &&
                  sizeof(Material) == 48 && sizeof(Camera) == 112 && sizeof(WeightAggregator) == 80,
              "the MSL structs must keep the layouts of struct_types.py");

EXPORT void ref_generate_paths(uint32_t n, const Ray *rays, const Box *boxes, const Triangle *triangles,
                               const Material *materials, uint32_t *random_buffer, float4 *out, Path *output_paths,
                               float4 *float_debug) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t id = 0; id < (int64_t)n; id++)
        generate_paths(rays, boxes, triangles, materials, random_buffer, out, output_paths, float_debug, (uint)id);
}

EXPORT void ref_connect_paths(uint32_t n, const Path *camera_paths, const Path *light_paths, const Triangle *triangles,
                              const Material *materials, const Box *boxes, const Camera *camera,
                              WeightAggregator *weight_aggregators, float4 *out, int *light_pixel_indices,
                              int *light_path_indices, int *light_ray_indices, float *light_weights, float *light_shade) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t id = 0; id < (int64_t)n; id++)
        connect_paths(camera_paths, light_paths, triangles, materials, boxes, camera, weight_aggregators, out,
                      light_pixel_indices, light_path_indices, light_ray_indices, light_weights, light_shade, (uint)id);
}

EXPORT void ref_light_sort(uint32_t n_threads, int32_t *light_pixel_indices, int32_t *light_path_indices,
                           int32_t *light_ray_indices, float *light_weights, float *light_shade, uint32_t stage,
                           uint32_t pass_of_stage, uint32_t n, uint32_t pairs_per_thread) {
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n_threads; id++)
        light_sort(light_pixel_indices, light_path_indices, light_ray_indices, light_weights, light_shade, stage,
                   pass_of_stage, n, pairs_per_thread, (uint)id);
}

EXPORT void ref_light_image_gather(uint32_t n, const Path *light_paths, const Material *materials,
                                   const int32_t *path_indices, const int32_t *ray_indices, const int32_t *bins,
                                   const uint32_t *offset, const float *weights, const float *shades,
                                   float4 *light_image, float *sum_weights) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t id = 0; id < (int64_t)n; id++)
        light_image_gather(light_paths, materials, path_indices, ray_indices, bins, *offset, weights, shades,
                           light_image, sum_weights, (uint)id);
}

EXPORT void ref_reset_light_indices(uint32_t n, int32_t *light_pixel_indices, int32_t *light_path_indices,
                                    int32_t *light_ray_indices, float *light_weights, float *light_shade) {
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n; id++)
        reset_light_indices(light_pixel_indices, light_path_indices, light_ray_indices, light_weights, light_shade, (uint)id);
}

EXPORT void ref_adaptive_finalize_samples(uint32_t n, const WeightAggregator *weight_aggregators, const Camera *camera,
                                          float4 *out, uint32_t *sample_counts, const uint32_t *sample_bin_offsets,
                                          float *sample_weights) {
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n; id++)
        adaptive_finalize_samples(weight_aggregators, camera, out, sample_counts, sample_bin_offsets, sample_weights, (uint)id);
}

EXPORT void ref_generate_camera_rays(uint32_t n, const Camera *camera, uint32_t *random_buffer, const uint32_t *indices,
                                     Ray *out) {
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n; id++)
        generate_camera_rays(camera, random_buffer, const_cast<uint32_t *>(indices), out, (uint)id);
}

EXPORT void ref_generate_light_rays(uint32_t n, const Triangle *light_triangles, const float *surface_areas,
                                    const int32_t *light_triangle_indices, const Material *materials,
                                    uint32_t *random_buffer, Ray *out, const int32_t *counts) {
    const int c = counts[0];
    if (c <= 0) return;
    // one duplicate of the last element behind each array: see the header comment
    std::vector<Triangle> lt(light_triangles, light_triangles + c);
    std::vector<float> sa(surface_areas, surface_areas + c);
    std::vector<int32_t> li(light_triangle_indices, light_triangle_indices + c);
    lt.push_back(lt.back());
    sa.push_back(sa.back());
    li.push_back(li.back());
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n; id++)
        generate_light_rays(lt.data(), sa.data(), li.data(), materials, random_buffer, out, counts, (uint)id);
}

// ---- parity hooks: the reference's traverse_bvh / visibility_test helpers on caller-supplied rays -------------------
EXPORT void ref_traverse(uint32_t n, const Ray *rays, const Box *boxes, const Triangle *triangles, int32_t *tri, float *t,
                         float *u, float *v) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t id = 0; id < (int64_t)n; id++) {
        Ray ray = rays[id];
        int best_i = -1;
        float best_t = INFINITY, bu = 0.0f, bv = 0.0f;
        traverse_bvh(ray, boxes, triangles, best_i, best_t, bu, bv);
        tri[id] = best_i; t[id] = best_t; u[id] = bu; v[id] = bv;
    }
}

EXPORT void ref_visibility(uint32_t n, const Ray *a, const Ray *b, const Box *boxes, const Triangle *triangles,
                           int32_t *visible) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t id = 0; id < (int64_t)n; id++) visible[id] = visibility_test(a[id], b[id], boxes, triangles) ? 1 : 0;
}

EXPORT void ref_math_probe(uint32_t n, const float *x, int which, float *y) {
    for (uint32_t i = 0; i < n; i++) {
        switch (which) {
        case 0: y[i] = sin(x[i]); break;
        case 1: y[i] = cos(x[i]); break;
        case 2: y[i] = acos(x[i]); break;
        case 3: y[i] = atan(x[i]); break;
        default: y[i] = exp(x[i]); break;
        }
    }
}

EXPORT uint32_t ref_xorshift32(uint32_t seed, float *as_float) {
    unsigned int s = seed;
    float f = xorshift_random(s);
    if (as_float) *as_float = f;
    return s;
}

// the compiled reference carries no work counters
EXPORT void ref_counters_reset(void) {}
EXPORT void ref_counters_get(uint64_t out[3]) { out[0] = out[1] = out[2] = 0; }

EXPORT void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
EXPORT int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
EXPORT int ref_uses_shared_math(void) {
#ifdef CLV_SHARED_MATH
    return 1;
#else
    return 0;
#endif
}
EXPORT int ref_is_compiled_reference(void) { return 1; }
