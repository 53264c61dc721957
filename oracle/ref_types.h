/* ref_types.h — record layouts of the reference (src/trace.metal:7-85, src/struct_types.py), shared by the oracle's
 * kernels (trace_ref.c) and its per-pass dispatcher (ref_glue.c).  TEST INFRASTRUCTURE, NOT PRODUCT. */
#ifndef CLV_REF_TYPES_H
#define CLV_REF_TYPES_H
#include <stdint.h>

/* ------------------------------------------------------------------------------------------------------------
 * record layouts (trace.metal:7-85); float3 occupies 16 bytes
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { float x, y, z, w; } float3; /* w is padding */
typedef struct { float x, y, z, w; } float4;
typedef struct { float x, y; } float2;

typedef struct {
    float3 origin, direction, inv_direction, color, normal;
    int32_t material, triangle;
    float c_importance, l_importance, tot_importance;
    int32_t hit_light, from_camera, hit_camera, pixel_idx;
    int32_t pad[3];
} Ray;

typedef struct {
    float weights[3][3];
    float pad_[3];
    float3 total_contribution; /* MSL aligns float3 to 16: offset 48 */
    float contrib_weight_sum;  /* offset 64; sizeof rounds up to 80 (trace.metal:25-29) */
    float pad2_[3];
} WeightAggregator;

typedef struct {
    Ray rays[8];
    int32_t length, from_camera, pad[2];
} Path;

typedef struct {
    float3 min, max;
    int32_t left, right, pad[2];
} Box;

typedef struct {
    float3 v0, v1, v2, n0, n1, n2, normal;
    int32_t material, is_light, is_camera, pad;
} Triangle;

typedef struct {
    float3 color, emission;
    int32_t type;
    float alpha, ior;
    int32_t transmissive;
} Material;

typedef struct {
    float3 center, focal_point, direction, dx, dy;
    int32_t pixel_width, pixel_height;
    float phys_width, phys_height, h_fov, v_fov;
    int32_t pad[2];
} Camera;

_Static_assert(sizeof(Ray) == 128, "Ray");
_Static_assert(sizeof(Path) == 1040, "Path");
_Static_assert(sizeof(Box) == 48, "Box");
_Static_assert(sizeof(Triangle) == 128, "Triangle");
_Static_assert(sizeof(Material) == 48, "Material");
_Static_assert(sizeof(Camera) == 112, "Camera");
_Static_assert(sizeof(WeightAggregator) == 80, "WeightAggregator");

#define EXPORT __attribute__((visibility("default")))
#endif /* CLV_REF_TYPES_H */
