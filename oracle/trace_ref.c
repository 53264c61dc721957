/* trace_ref.c — CPU ORACLE for the Clive2 BDPT hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A C99 + OpenMP restatement of the eight Metal compute kernels in the reference's src/trace.metal and of the host
 * glue in src/renderer.py that wires them into one sample pass.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; nothing under clive2_b200/ does.
 *
 * PARITY STATUS: **pinned to the reference itself.**  The reference's kernels are Metal Shading Language and ship no
 * tests or known-answer vectors, but the unmodified src/trace.metal compiles as C++ behind a small
 * metal_stdlib shim (oracle/ref_build/, `make ref` -> oracle/_ref/).  tests/test_ref_pin.py checks that this file and
 * that compiled reference write the same bytes into every buffer after every stage of every pass (Cornell scene, GGX /
 * transmission scene, mesh scene; glibc and shared-math builds), and tests/golden/ref_kernels.npz holds vectors the
 * compiled reference produced (generator committed: tests/golden/make_ref_golden.py) for machines without the
 * reference tree.  The record layouts, scene arrays, BVH and camera that feed the kernels are checked byte for byte
 * against outputs of the reference's own Python (tests/golden/make_golden.py).  Every function cites the lines it
 * restates.
 *
 * Arithmetic model: IEEE-754 binary32 everywhere (MSL has no double; unsuffixed literals are float), one rounding per
 * source-level operation, no fused multiply-add (build with -ffp-contract=off, no -ffast-math), operations associated
 * exactly as the MSL expression trees.  Elementary functions come from glibc libm, or from
 * clive2_b200/csrc/clv_math.h when built with -DCLV_SHARED_MATH (bit-exact differential testing against the CUDA
 * kernels, which always use that header).
 *
 * Deliberate, documented deviations from the MSL text:
 *   A2  light_index is clamped to count-1 (the reference reads out of bounds when the RNG returns exactly 1.0);
 *   A4  locals the MSL leaves uninitialised (new_ray/next_ray, p_ratios/p_values, unused Path slots) start as zero;
 *   min()/max() on NaN return the non-NaN operand (fminf/fmaxf).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef CLV_SHARED_MATH
#include "../clive2_b200/csrc/clv_math.h"
#define M_SIN clv_sinf
#define M_COS clv_cosf
#define M_ACOS clv_acosf
#define M_ATAN clv_atanf
#define M_EXP clv_expf
#else
#define M_SIN sinf
#define M_COS cosf
#define M_ACOS acosf
#define M_ATAN atanf
#define M_EXP expf
#endif

#define PI 3.14159265359f /* trace.metal:4 */
#define DELTA 0.0001f     /* trace.metal:5 */

#include "ref_types.h"


/* ------------------------------------------------------------------------------------------------------------
 * float3 algebra with MSL operator semantics (component-wise, one rounding per operation)
 * ---------------------------------------------------------------------------------------------------------- */
static inline float3 v3(float x, float y, float z) { float3 r = {x, y, z, 0.0f}; return r; }
static inline float3 vadd(float3 a, float3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 vsub(float3 a, float3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 vmul(float3 a, float3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float3 vscale(float3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline float3 vdivs(float3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline float3 vneg(float3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float3 vrcp(float3 a) { return v3(1.0f / a.x, 1.0f / a.y, 1.0f / a.z); }
static inline float vdot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float3 vcross(float3 a, float3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float vlength(float3 a) { return sqrtf(vdot(a, a)); }
static inline float3 vnormalize(float3 a) { return vdivs(a, vlength(a)); }

/* work counters: one "ray" is one traverse_bvh invocation */
typedef struct { uint64_t rays, node_visits, tri_tests; } Counters;
static Counters g_counters;
EXPORT void ref_counters_reset(void) { memset(&g_counters, 0, sizeof g_counters); }
EXPORT void ref_counters_get(uint64_t out[3]) {
    out[0] = g_counters.rays; out[1] = g_counters.node_visits; out[2] = g_counters.tri_tests;
}
static void counters_merge(const Counters *c) {
#pragma omp critical(clv_counters)
    {
        g_counters.rays += c->rays;
        g_counters.node_visits += c->node_visits;
        g_counters.tri_tests += c->tri_tests;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * RNG (trace.metal:87-93)
 * ---------------------------------------------------------------------------------------------------------- */
static inline float xorshift_random(uint32_t *seed) {
    uint32_t s = *seed;
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    *seed = s;
    return (float)s / (float)0xFFFFFFFF; /* the divisor rounds to 2^32: an exact scaling, 1.0 is reachable */
}
EXPORT uint32_t ref_xorshift32(uint32_t seed, float *as_float) {
    float f = xorshift_random(&seed);
    if (as_float) *as_float = f;
    return seed;
}

/* ------------------------------------------------------------------------------------------------------------
 * intersection and traversal (trace.metal:106-197)
 * ---------------------------------------------------------------------------------------------------------- */
static inline void ray_box_intersect(const Ray *ray, const Box *box, int *hit, float *t) { /* :106-115 */
    float3 t0s = vmul(vsub(box->min, ray->origin), ray->inv_direction);
    float3 t1s = vmul(vsub(box->max, ray->origin), ray->inv_direction);
    float3 tmin = v3(fminf(t0s.x, t1s.x), fminf(t0s.y, t1s.y), fminf(t0s.z, t1s.z));
    float3 tmax = v3(fmaxf(t0s.x, t1s.x), fmaxf(t0s.y, t1s.y), fmaxf(t0s.z, t1s.z));
    float tmin_final = fmaxf(fmaxf(tmin.x, tmin.y), fmaxf(tmin.z, 0.0f));
    float tmax_final = fminf(fminf(tmax.x, tmax.y), fminf(tmax.z, *t));
    *hit = tmin_final <= tmax_final;
    *t = tmin_final;
}

static inline void ray_triangle_intersect(const Ray *ray, const Triangle *tri, int *hit, float *t_out, float *u,
                                          float *v) { /* :117-142 */
    float3 edge1 = vsub(tri->v1, tri->v0);
    float3 edge2 = vsub(tri->v2, tri->v0);
    float3 h = vcross(ray->direction, edge2);
    float a = vdot(edge1, h);
    float f = 1.0f / a;
    float3 s = vsub(ray->origin, tri->v0);
    *u = f * vdot(s, h);
    if (*u < 0 || *u > 1) { *hit = 0; return; }
    float3 q = vcross(s, edge1);
    *v = f * vdot(ray->direction, q);
    if (*v < 0 || *u + *v > 1) { *hit = 0; return; }
    float t = f * vdot(edge2, q);
    if (t > DELTA) { *hit = 1; *t_out = t; } else { *hit = 0; }
}

static void traverse_bvh(const Ray *ray, const Box *boxes, const Triangle *triangles, int *best_i, float *best_t,
                         float *u_out, float *v_out, Counters *cnt) { /* :144-176 */
    int stack[64];
    int sp = 0;
    stack[sp++] = 0;
    cnt->rays++;
    while (sp > 0 && sp < 64) {
        int box_id = stack[--sp];
        Box box = boxes[box_id];
        int hit = 0;
        float t = INFINITY;
        float u = 0.0f, v = 0.0f;
        cnt->node_visits++;
        ray_box_intersect(ray, &box, &hit, &t);
        if (hit && t < *best_t) {
            if (box.right == 0) {
                stack[sp++] = box.left;
                stack[sp++] = box.left + 1;
            } else {
                for (int i = box.left; i < box.right; i++) {
                    hit = 0;
                    t = INFINITY;
                    cnt->tri_tests++;
                    ray_triangle_intersect(ray, &triangles[i], &hit, &t, &u, &v);
                    if (hit && t < *best_t) {
                        *best_i = i;
                        *best_t = t;
                        *u_out = u;
                        *v_out = v;
                    }
                }
            }
        }
    }
}

static int visibility_test(const Ray *a, const Ray *b, const Box *boxes, const Triangle *triangles,
                           Counters *cnt) { /* :178-197 */
    Ray test_ray;
    memset(&test_ray, 0, sizeof test_ray);
    test_ray.origin = a->origin;
    float3 direction = vnormalize(vsub(b->origin, a->origin));
    test_ray.direction = direction;
    test_ray.inv_direction = vrcp(direction);
    test_ray.triangle = a->triangle;
    int best_i = -1;
    float best_t = INFINITY;
    float u = 0.0f, v = 0.0f;
    traverse_bvh(&test_ray, boxes, triangles, &best_i, &best_t, &u, &v, cnt);
    if (best_i == -1) return 0;
    if (best_i == a->triangle) return 0;
    if (best_i == b->triangle) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * sampling and BSDFs (trace.metal:200-379)
 * ---------------------------------------------------------------------------------------------------------- */
static void orthonormal(float3 n, float3 *x, float3 *y) { /* :200-211 */
    float3 v;
    if (fabsf(n.x) <= fabsf(n.y) && fabsf(n.x) <= fabsf(n.z)) v = v3(1, 0, 0);
    else if (fabsf(n.y) <= fabsf(n.z)) v = v3(0, 1, 0);
    else v = v3(0, 0, 1);
    *x = vnormalize(vsub(v, vscale(n, vdot(v, n))));
    *y = vnormalize(vcross(n, *x));
}

static float3 random_hemisphere_cosine(float3 xa, float3 ya, float3 za, float2 r) { /* :213-217 */
    float theta = M_ACOS(sqrtf(r.x));
    float phi = 2 * PI * r.y;
    float st = M_SIN(theta), ct = M_COS(theta), sp = M_SIN(phi), cp = M_COS(phi);
    return vnormalize(vadd(vadd(vscale(xa, st * cp), vscale(ya, st * sp)), vscale(za, ct)));
}

static float3 random_hemisphere_uniform(float3 xa, float3 ya, float3 za, float2 r) { /* :219-224 */
    float z = r.x;
    float rr = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    float phi = 2 * PI * r.y;
    float sp = M_SIN(phi), cp = M_COS(phi);
    return vnormalize(vadd(vadd(vscale(xa, rr * cp), vscale(ya, rr * sp)), vscale(za, z)));
}

static float3 GGX_sample(float3 n, float2 r, float alpha) { /* :226-233 */
    float3 x, y;
    orthonormal(n, &x, &y);
    float theta = 2 * PI * r.x;
    float phi = M_ATAN(alpha * sqrtf(r.y) / sqrtf(1.0f - r.y));
    float sp = M_SIN(phi), cp = M_COS(phi), st = M_SIN(theta), ct = M_COS(theta);
    return vnormalize(vadd(vadd(vscale(x, sp * ct), vscale(y, sp * st)), vscale(n, cp)));
}

static float3 specular_reflection(float3 i, float3 m) { /* :235-237 */
    return vnormalize(vsub(vscale(m, 2 * vdot(i, m)), i));
}

static float3 GGX_transmit(float3 i, float3 m, float ni, float no) { /* :243-248, eta (not eta^2) as written */
    float cosTheta_i = vdot(i, m);
    float eta = ni / no;
    float cosTheta_t = sqrtf(1 + eta * (cosTheta_i * cosTheta_i - 1));
    return vnormalize(vsub(vscale(m, eta * cosTheta_i - cosTheta_t), vscale(i, eta)));
}

static float3 specular_transmit_half_direction(float3 i, float3 o, float ni, float no) { /* :250-252 */
    return vnormalize(vadd(vscale(o, no), vscale(i, ni)));
}

static float degreve_fresnel(float3 i, float3 m, float ni, float nt) { /* :254-264 */
    float cosTheta_i = fabsf(vdot(i, m));
    float eta = ni / nt;
    float sinTheta_t2 = eta * eta * (1.0f - cosTheta_i * cosTheta_i);
    if (sinTheta_t2 >= 1.0f) return 1.0f;
    float cosTheta_t = sqrtf(1.0f - sinTheta_t2);
    float r_parallel = (nt * cosTheta_i - ni * cosTheta_t) / (nt * cosTheta_i + ni * cosTheta_t);
    float r_perpendicular = (ni * cosTheta_i - nt * cosTheta_t) / (ni * cosTheta_i + nt * cosTheta_t);
    return 0.5f * (r_parallel * r_parallel + r_perpendicular * r_perpendicular);
}

static float GGX_G1(float3 v, float3 m, float alpha) { /* :266-271 (n is unused there) */
    float mv = vdot(m, v);
    float sin2 = 1.0f - mv * mv;
    float tan2 = sin2 / (mv * mv);
    return 2.0f / (1.0f + sqrtf(1.0f + alpha * alpha * tan2));
}

static float GGX_G(float3 i, float3 o, float3 m, float3 n, float alpha) { /* :273-277 */
    if (vdot(i, m) * vdot(i, n) <= 0.0f) return 0.0f;
    if (vdot(o, m) * vdot(o, n) <= 0.0f) return 0.0f;
    return GGX_G1(i, m, alpha) * GGX_G1(o, m, alpha);
}

static float GGX_D(float3 m, float3 n, float alpha) { /* :279-288 */
    if (alpha == 0.0f) return 1.0f;
    float alpha2 = alpha * alpha;
    float cosTheta = vdot(m, n);
    float cosTheta2 = cosTheta * cosTheta;
    float denom = cosTheta2 * (alpha2 - 1.0f) + 1.0f;
    return alpha2 / (PI * denom * denom);
}

static float reflect_jacobian(float3 m, float3 o) { return 1.0f / (4.0f * fabsf(vdot(m, o))); } /* :290-292 */

static float transmit_jacobian(float3 i, float3 o, float3 m, float ni, float no) { /* :294-301 (m unused there) */
    (void)m;
    float3 h = specular_transmit_half_direction(i, o, ni, no);
    float cosTheta_i = vdot(i, h);
    float cosTheta_o = vdot(o, h);
    float numerator = no * no * fabsf(cosTheta_o);
    float s = ni * cosTheta_i + no * cosTheta_o;
    float denominator = s * s;
    return numerator / denominator;
}

static float GGX_BRDF_reflect(float3 i, float3 o, float3 m, float3 n, float ni, float no, float alpha) { /* :303-309 */
    float D = GGX_D(m, n, alpha);
    float G = GGX_G(i, o, m, n, alpha);
    float F = degreve_fresnel(i, m, ni, no);
    return (D * G * F) / (4.0f * fabsf(vdot(i, m)));
}

static float GGX_BRDF_transmit(float3 i, float3 o, float3 m, float3 n, float ni, float no, float alpha) { /* :311-328 */
    float3 h = specular_transmit_half_direction(i, o, ni, no);
    float D = GGX_D(m, n, alpha);
    float G = GGX_G(i, o, m, n, alpha);
    float F = degreve_fresnel(i, m, ni, no);
    float im = vdot(i, h);
    float om = vdot(o, h);
    float in = vdot(i, n);
    float on = vdot(o, n);
    float coeff = (im * om) / (in * on);
    float num = no * no * D * G * (1.0f - F);
    float s = ni * im + no * om;
    float denom = s * s;
    return coeff * num / denom;
}

static float3 sample_normal(const Triangle *t, float u, float v) { /* :330-332 */
    return vnormalize(vadd(vadd(vscale(t->n0, 1 - u - v), vscale(t->n1, u)), vscale(t->n2, v)));
}

static void diffuse_bounce(float3 wi, float3 n, int from_camera, float2 roll, float3 *wo, float *f, float *c_p,
                           float *l_p) { /* :334-346 */
    float3 x, y;
    orthonormal(n, &x, &y);
    *wo = random_hemisphere_cosine(x, y, n, roll);
    *f = fabsf(vdot(n, *wo)) / PI;
    if (from_camera) {
        *c_p = fabsf(vdot(n, *wo)) / PI;
        *l_p = fabsf(vdot(n, wi)) / PI;
    } else {
        *c_p = fabsf(vdot(n, wi)) / PI;
        *l_p = fabsf(vdot(n, *wo)) / PI;
    }
}

static void reflect_bounce(float3 wi, float3 n, float3 m, float ni, float no, float alpha, int from_camera, float3 *wo,
                           float *f, float *c_p, float *l_p) { /* :348-362 */
    *wo = specular_reflection(wi, m);
    *f = GGX_BRDF_reflect(wi, *wo, m, n, ni, no, alpha);
    float pf = degreve_fresnel(wi, m, ni, no);
    float pm = fabsf(vdot(m, n)) * GGX_D(m, n, alpha);
    if (from_camera) {
        *c_p = pf * pm * reflect_jacobian(m, *wo);
        *l_p = pf * pm * reflect_jacobian(m, wi);
    } else {
        *c_p = pf * pm * reflect_jacobian(m, wi);
        *l_p = pf * pm * reflect_jacobian(m, *wo);
    }
}

static void transmit_bounce(float3 wi, float3 n, float3 m, float ni, float no, float alpha, int from_camera, float3 *wo,
                            float *f, float *c_p, float *l_p) { /* :364-379 */
    *wo = GGX_transmit(wi, m, ni, no);
    *f = GGX_BRDF_transmit(wi, *wo, m, n, ni, no, alpha);
    float pf = 1.0f - degreve_fresnel(wi, m, ni, no);
    float pm = fabsf(vdot(m, n)) * GGX_D(m, n, alpha);
    if (from_camera) {
        *c_p = pf * pm * transmit_jacobian(wi, *wo, m, ni, no);
        *l_p = pf * pm * transmit_jacobian(*wo, wi, vneg(m), no, ni);
    } else {
        *c_p = pf * pm * transmit_jacobian(*wo, wi, vneg(m), no, ni);
        *l_p = pf * pm * transmit_jacobian(wi, *wo, m, ni, no);
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * kernel: generate_paths (trace.metal:381-532)
 * ---------------------------------------------------------------------------------------------------------- */
static void generate_paths_one(uint32_t id, const Ray *rays, const Box *boxes, const Triangle *triangles,
                               const Material *materials, uint32_t *random_buffer, float4 *out, Path *output_paths,
                               float4 *float_debug, Counters *cnt) {
    Path path;
    memset(&path, 0, sizeof path);
    path.length = 0;
    Ray ray, new_ray, next_ray;
    memset(&new_ray, 0, sizeof new_ray);
    memset(&next_ray, 0, sizeof next_ray);
    ray = rays[id];
    path.from_camera = ray.from_camera;
    out[id] = (float4){0, 0, 0, 0};

    uint32_t seed0 = random_buffer[2 * id];
    uint32_t seed1 = random_buffer[2 * id + 1];

    if (path.from_camera == 0) new_ray.l_importance = 1.0f / (2.0f * PI);
    else new_ray.c_importance = ray.c_importance;

    for (int i = 0; i < 6; i++) {
        int best_i = -1;
        float best_t = INFINITY;
        float u = 0.0f, v = 0.0f;
        traverse_bvh(&ray, boxes, triangles, &best_i, &best_t, &u, &v, cnt);
        if (best_i == -1) break;

        Triangle triangle = triangles[best_i];
        Material material = materials[triangle.material];

        float3 n;
        float ni, no;
        float alpha = material.alpha;
        float3 sampled_normal = sample_normal(&triangle, u, v);
        float side = vdot(vneg(ray.direction), triangle.normal);
        if (side > 0) {
            n = sampled_normal; ni = 1.0f; no = material.ior;
        } else if (side < 0) {
            n = vneg(sampled_normal); ni = material.ior; no = 1.0f;
        } else {
            break;
        }

        new_ray.origin = vadd(ray.origin, vscale(ray.direction, best_t));
        new_ray.material = triangle.material;
        new_ray.triangle = best_i;

        if (triangle.is_light && vdot(ray.direction, triangle.normal) < 0.0f) new_ray.hit_light = best_i;
        else new_ray.hit_light = -1;
        if (triangle.is_camera) new_ray.hit_camera = best_i;
        else new_ray.hit_camera = -1;

        float3 wi = vneg(ray.direction);

        float rand_x_a = xorshift_random(&seed0);
        float rand_y_a = xorshift_random(&seed1);
        float2 roll_a = {rand_x_a, rand_y_a};
        float rand_x_b = xorshift_random(&seed0);
        float rand_y_b = xorshift_random(&seed1);
        float2 roll_b = {rand_x_b, rand_y_b};

        float3 wo = v3(0, 0, 0);
        float f = 1.0f, c_p = 1.0f, l_p = 1.0f;

        float3 m = GGX_sample(n, roll_a, alpha);
        if (vdot(wi, m) < 0.0f) break;
        if (vdot(m, n) < 0.0f) break;
        new_ray.normal = n;

        float fresnel = degreve_fresnel(wi, m, ni, no);
        if (material.type == 0) {
            diffuse_bounce(wi, n, path.from_camera, roll_b, &wo, &f, &c_p, &l_p);
        } else if (material.type == 1) {
            if (roll_b.x <= fresnel) reflect_bounce(wi, n, m, ni, no, alpha, path.from_camera, &wo, &f, &c_p, &l_p);
            else transmit_bounce(wi, n, m, ni, no, alpha, path.from_camera, &wo, &f, &c_p, &l_p);
        } else if (material.type == 2) {
            if (roll_b.x <= fresnel) reflect_bounce(wi, n, m, ni, no, alpha, path.from_camera, &wo, &f, &c_p, &l_p);
            else diffuse_bounce(wi, n, path.from_camera, roll_b, &wo, &f, &c_p, &l_p);
        } else {
            reflect_bounce(wi, n, m, ni, no, alpha, path.from_camera, &wo, &f, &c_p, &l_p);
        }

        float wi_n = vdot(wi, triangle.normal), wo_n = vdot(wo, triangle.normal);
        if (wi_n > 0.0f && wo_n > 0.0f) new_ray.color = vmul(vscale(ray.color, f), material.color);      /* external reflection */
        else if (wi_n < 0.0f && wo_n > 0.0f) new_ray.color = vmul(vscale(ray.color, f), material.color); /* egress */
        else new_ray.color = vscale(ray.color, f);                                                       /* internal, ingress */

        new_ray.direction = wo;
        new_ray.inv_direction = vrcp(wo);

        if (path.from_camera) {
            next_ray.c_importance = c_p;
            ray.l_importance = l_p;
            new_ray.tot_importance = ray.tot_importance * new_ray.c_importance;
        } else {
            next_ray.l_importance = l_p;
            ray.c_importance = c_p;
            new_ray.tot_importance = ray.tot_importance * new_ray.l_importance;
        }

        if (f == 0.0f) break;

        path.rays[i] = ray;
        path.length = i + 1;

        ray = new_ray;
        new_ray = next_ray;
    }

    output_paths[id] = path;
    float_debug[id] = (float4){100.0f, 100.0f, 100.0f, 100.0f};

    for (int i = 0; i < path.length; i++) {
        if (path.rays[i].hit_light >= 0) {
            float3 c = vdivs(path.rays[i - 1].color, path.rays[i].tot_importance);
            out[id] = (float4){c.x, c.y, c.z, 1.0f};
            break;
        }
    }

    random_buffer[2 * id] = seed0;
    random_buffer[2 * id + 1] = seed1;
}

EXPORT void ref_generate_paths(uint32_t n, const Ray *rays, const Box *boxes, const Triangle *triangles,
                               const Material *materials, uint32_t *random_buffer, float4 *out, Path *output_paths,
                               float4 *float_debug) {
#pragma omp parallel
    {
        Counters cnt = {0, 0, 0};
#pragma omp for schedule(dynamic, 64)
        for (int64_t id = 0; id < (int64_t)n; id++)
            generate_paths_one((uint32_t)id, rays, boxes, triangles, materials, random_buffer, out, output_paths,
                               float_debug, &cnt);
        counters_merge(&cnt);
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * connection helpers (trace.metal:534-617)
 * ---------------------------------------------------------------------------------------------------------- */
static float cosine_geometry_term(const Ray *a, const Ray *b) { /* :539-544 */
    float dist = vlength(vsub(b->origin, a->origin));
    float cos_a = fabsf(vdot(a->direction, a->normal));
    float cos_b = fabsf(vdot(b->direction, b->normal));
    return cos_a * cos_b / (dist * dist);
}

static const Ray *get_ray(const Path *camera_path, const Path *light_path, int t, int s, int i) { /* :546-549 */
    if (i < s) return &light_path->rays[i];
    return &camera_path->rays[t + s - i - 1];
}

static float3 pixel_center(const Camera *camera, int x, int y) { /* :551-562 (no +0.5: the pixel corner) */
    float x_normalized = ((float)x - 0.5f * (float)camera->pixel_width) / (float)camera->pixel_width;
    float y_normalized = ((float)y - 0.5f * (float)camera->pixel_height) / (float)camera->pixel_height;
    float3 x_vector = vscale(camera->dx, x_normalized * camera->phys_width);
    float3 y_vector = vscale(camera->dy, y_normalized * camera->phys_height);
    return vadd(vadd(camera->center, x_vector), y_vector);
}

static float gaussian_weight(float3 p, float3 q, float sigma) { /* :564-567 */
    float dist = vlength(vsub(p, q));
    return M_EXP(-dist * dist / (2.0f * sigma * sigma));
}

static void world_ray_to_camera_ray(const Box *boxes, const Triangle *triangles, const Material *materials,
                                    const Camera *camera, const Ray *world_ray, int *pixel_idx, Ray *camera_ray,
                                    Counters *cnt) { /* :569-617 */
    if (materials[triangles[world_ray->triangle].material].type > 0) return;

    Ray test_ray;
    memset(&test_ray, 0, sizeof test_ray);
    test_ray.origin = world_ray->origin;
    test_ray.direction = vnormalize(vsub(camera->focal_point, world_ray->origin));
    if (vdot(test_ray.direction, camera->direction) > 0.0f) return;
    test_ray.inv_direction = vrcp(test_ray.direction);
    test_ray.triangle = world_ray->triangle;
    test_ray.normal = world_ray->normal;

    int best_i = -1;
    float best_t = INFINITY;
    float u = 0.0f, v = 0.0f;
    traverse_bvh(&test_ray, boxes, triangles, &best_i, &best_t, &u, &v, cnt);
    if (best_i == -1) return;
    if (!triangles[best_i].is_camera) return;

    float3 camera_point = vadd(test_ray.origin, vscale(test_ray.direction, best_t));
    float x = vdot(vsub(camera_point, camera->center), camera->dx);
    float y = vdot(vsub(camera_point, camera->center), camera->dy);
    int pixel_x = (int)roundf((x / camera->phys_width + 0.5f) * (float)camera->pixel_width);
    int pixel_y = (int)roundf((y / camera->phys_height + 0.5f) * (float)camera->pixel_height);

    *pixel_idx = pixel_y * camera->pixel_width + pixel_x;

    camera_ray->origin = camera_point;
    camera_ray->direction = vnormalize(vsub(camera->focal_point, camera_point));
    camera_ray->inv_direction = vrcp(camera_ray->direction);
    camera_ray->normal = camera->direction;
    camera_ray->material = 7;
    camera_ray->color = v3(1.0f, 1.0f, 1.0f);
    camera_ray->triangle = best_i;
    camera_ray->tot_importance = 1.0f;
    camera_ray->hit_light = -1;
    camera_ray->hit_camera = best_i;
}

/* ------------------------------------------------------------------------------------------------------------
 * kernel: connect_paths (trace.metal:620-869)
 * ---------------------------------------------------------------------------------------------------------- */
static void connect_paths_one(uint32_t id, const Path *camera_paths, const Path *light_paths, const Triangle *triangles,
                              const Material *materials, const Box *boxes, const Camera *camera,
                              WeightAggregator *weight_aggregators, float4 *out, int32_t *light_pixel_indices,
                              int32_t *light_path_indices, int32_t *light_ray_indices, float *light_weights,
                              float *light_shade, Counters *cnt) {
    Path camera_path = camera_paths[id];
    Path light_path = light_paths[id];
    const Ray cached_camera_zero = camera_path.rays[0];

    out[id] = (float4){0, 0, 0, 0};
    Camera c = camera[0];

    WeightAggregator aggregator;
    memset(&aggregator, 0, sizeof aggregator);
    aggregator.total_contribution = v3(0, 0, 0);
    int pixel_idx = cached_camera_zero.pixel_idx;
    int light_pixel_idx = -1;
    int total_pixels = c.pixel_width * c.pixel_height;
    float contrib_weight_sum = 0.0f;

    for (int t = 1; t < camera_path.length + 1; t++) {
        for (int s = 0; s < light_path.length + 1; s++) {
            if (t + s < 2) continue;

            Ray light_ray;
            memset(&light_ray, 0, sizeof light_ray);
            light_ray.triangle = -1;
            Ray camera_ray;
            memset(&camera_ray, 0, sizeof camera_ray);
            camera_ray.triangle = -1;
            camera_path.rays[0] = cached_camera_zero;
            float3 dir_l_to_c = v3(0, 0, 0);
            light_pixel_idx = -1;

            if (s == 0) {
                camera_ray = camera_path.rays[t - 1];
                if (camera_ray.hit_light < 0) continue;
            } else if (t == 1) {
                light_ray = light_path.rays[s - 1];
                world_ray_to_camera_ray(boxes, triangles, materials, &c, &light_ray, &light_pixel_idx,
                                        &camera_path.rays[0], cnt);
                if (light_pixel_idx == -1) continue;
                camera_ray = camera_path.rays[0];
                dir_l_to_c = vnormalize(vsub(camera_ray.origin, light_ray.origin));
            } else {
                camera_ray = camera_path.rays[t - 1];
                light_ray = light_path.rays[s - 1];
                if (materials[light_ray.material].type > 0) continue;
                if (materials[camera_ray.material].type > 0) continue;
                dir_l_to_c = vnormalize(vsub(camera_ray.origin, light_ray.origin));
                if (vdot(light_ray.normal, dir_l_to_c) < DELTA) continue;
                if (vdot(camera_ray.normal, vneg(dir_l_to_c)) < DELTA) continue;
                if (!visibility_test(&light_ray, &camera_ray, boxes, triangles, cnt)) continue;
            }

            float p_ratios[32];
            float p_values[32];
            memset(p_ratios, 0, sizeof p_ratios);
            memset(p_values, 0, sizeof p_values);

            for (int i = 0; i < s + t; i++) { /* :709-735 */
                float num, denom;
                if (i == 0) {
                    const Ray *a = get_ray(&camera_path, &light_path, t, s, 0);
                    const Ray *b = get_ray(&camera_path, &light_path, t, s, 1);
                    num = a->l_importance;
                    denom = a->c_importance * cosine_geometry_term(a, b);
                } else if (i == s + t - 1) {
                    const Ray *a = get_ray(&camera_path, &light_path, t, s, s + t - 1);
                    const Ray *b = get_ray(&camera_path, &light_path, t, s, s + t - 2);
                    num = a->l_importance * cosine_geometry_term(a, b);
                    denom = a->c_importance;
                } else {
                    const Ray *a = get_ray(&camera_path, &light_path, t, s, i - 1);
                    const Ray *b = get_ray(&camera_path, &light_path, t, s, i);
                    const Ray *cc = get_ray(&camera_path, &light_path, t, s, i + 1);
                    num = b->l_importance * cosine_geometry_term(a, b);
                    denom = b->c_importance * cosine_geometry_term(b, cc);
                }
                p_ratios[i] = num / denom;
            }

            float prior_camera_importance = camera_ray.tot_importance;
            float prior_light_importance = (s == 0) ? 1.0f : light_ray.tot_importance;
            float p_s = prior_camera_importance * prior_light_importance;

            float p_i = p_s;
            for (int i = s; i < s + t + 1; i++) { /* :746-749; reads p_ratios[s+t] (never written, here 0) */
                p_values[i + 1] = p_ratios[i] * p_i;
                p_i = p_values[i + 1];
            }
            p_i = p_s;
            for (int i = s - 1; i >= 0; i--) {
                p_values[i] = p_i / p_ratios[i];
                p_i = p_values[i];
            }
            p_values[s] = p_s;

            for (int i = 0; i < s + t; i++) {
                if (materials[get_ray(&camera_path, &light_path, t, s, i)->material].type > 0) {
                    p_values[i] = 0.0f;
                    p_values[i + 1] = 0.0f;
                }
            }
            p_values[s + t] = 0.0f;

            float sum = 0.0f;
            for (int i = 0; i < s + t + 1; i++) sum += p_values[i];

            float w;
            if (p_values[s] > 0.0f && sum > 0.0f) w = p_values[s] / sum;
            else continue;

            float3 color = v3(1.0f, 1.0f, 1.0f);
            float g = 1.0f;
            float new_light_f = 1.0f;
            float new_camera_f = 1.0f;

            if (s == 0) {
                float3 prior_color = camera_path.rays[t - 2].color;
                float3 emission = materials[camera_ray.material].emission;
                color = vmul(prior_color, emission);
            } else if (t == 1) {
                int prior_light_ind = (s - 2 > 0) ? s - 2 : 0;
                float3 prior_color = light_path.rays[prior_light_ind].color;
                if (s > 1) new_light_f = fabsf(vdot(dir_l_to_c, light_ray.normal)) / PI;
                color = vmul(vscale(prior_color, new_light_f), materials[light_ray.material].color);
                g = cosine_geometry_term(&light_ray, &camera_ray);
            } else {
                float3 prior_camera_color = camera_path.rays[t - 2].color;
                Material camera_material = materials[camera_ray.material];
                new_camera_f = fabsf(vdot(vneg(dir_l_to_c), camera_ray.normal)) / PI;
                float3 camera_color = vmul(vscale(prior_camera_color, new_camera_f), camera_material.color);
                float3 light_color;
                if (s == 1) {
                    light_color = materials[light_ray.material].emission;
                } else {
                    float3 prior_light_color = light_path.rays[s - 2].color;
                    Material light_material = materials[light_ray.material];
                    new_light_f = fabsf(vdot(dir_l_to_c, light_ray.normal)) / PI;
                    light_color = vmul(vscale(prior_light_color, new_light_f), light_material.color);
                }
                color = vmul(camera_color, light_color);
                g = cosine_geometry_term(&camera_ray, &light_ray);
            }
            if (t != 1) {
                aggregator.total_contribution = vadd(aggregator.total_contribution, vdivs(vscale(color, w * g), p_s));
                contrib_weight_sum += w;
            } else {
                int slot = (int)id + s * total_pixels;
                light_pixel_indices[slot] = light_pixel_idx;
                light_path_indices[slot] = (int)id;
                light_ray_indices[slot] = s - 1;
                light_weights[slot] = w;
                light_shade[slot] = new_light_f * g / p_s;
            }
        }
    }

    /* 3x3 splat weights (:827-864) */
    float weight_sum = 0.0f;
    float pixel_phys_width = c.phys_width / (float)c.pixel_width;
    float pixel_phys_height = c.phys_height / (float)c.pixel_height;
    float sigma = 0.5f * sqrtf(pixel_phys_width * pixel_phys_width + pixel_phys_height * pixel_phys_height);

    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) aggregator.weights[i][j] = 0.0f;

    for (int i = -1; i < 2; i++) {
        for (int j = -1; j < 2; j++) {
            int new_sample_x = (pixel_idx % c.pixel_width) + i;
            int new_sample_y = (pixel_idx / c.pixel_width) + j;
            if (new_sample_x < 0 || new_sample_x >= c.pixel_width || new_sample_y < 0 || new_sample_y >= c.pixel_height)
                continue;
            int new_sample_index = new_sample_y * c.pixel_width + new_sample_x;
            if (new_sample_index < 0 || new_sample_index >= c.pixel_width * c.pixel_height) continue;
            float weight = gaussian_weight(pixel_center(&c, new_sample_x, new_sample_y), camera_paths[id].rays[0].origin,
                                           sigma);
            aggregator.weights[i + 1][j + 1] = weight;
            weight_sum += weight;
        }
    }
    if (weight_sum != 0.0f)
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) aggregator.weights[i][j] = aggregator.weights[i][j] / weight_sum;

    aggregator.contrib_weight_sum = contrib_weight_sum;

    out[id] = (float4){aggregator.total_contribution.x, aggregator.total_contribution.y,
                       aggregator.total_contribution.z, 1.0f};
    weight_aggregators[id] = aggregator;
}

EXPORT void ref_connect_paths(uint32_t n, const Path *camera_paths, const Path *light_paths, const Triangle *triangles,
                              const Material *materials, const Box *boxes, const Camera *camera,
                              WeightAggregator *weight_aggregators, float4 *out, int32_t *light_pixel_indices,
                              int32_t *light_path_indices, int32_t *light_ray_indices, float *light_weights,
                              float *light_shade) {
#pragma omp parallel
    {
        Counters cnt = {0, 0, 0};
#pragma omp for schedule(dynamic, 64)
        for (int64_t id = 0; id < (int64_t)n; id++)
            connect_paths_one((uint32_t)id, camera_paths, light_paths, triangles, materials, boxes, camera,
                              weight_aggregators, out, light_pixel_indices, light_path_indices, light_ray_indices,
                              light_weights, light_shade, &cnt);
        counters_merge(&cnt);
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * kernel: light_sort — one bitonic compare-exchange step (trace.metal:872-934)
 * ---------------------------------------------------------------------------------------------------------- */
EXPORT void ref_light_sort(uint32_t n_threads, int32_t *light_pixel_indices, int32_t *light_path_indices,
                           int32_t *light_ray_indices, float *light_weights, float *light_shade, uint32_t stage,
                           uint32_t passOfStage, uint32_t n, uint32_t pairs_per_thread) {
#pragma omp parallel for schedule(static)
    for (int64_t idl = 0; idl < (int64_t)n_threads; idl++) {
        uint32_t id = (uint32_t)idl;
        for (uint32_t p = 0; p < pairs_per_thread; ++p) {
            uint32_t global_pair_id = id * pairs_per_thread + p;
            uint32_t pairDistance = 1u << (passOfStage - 1);
            uint32_t blockWidth = 1u << stage;
            uint32_t leftId = (global_pair_id / pairDistance) * pairDistance * 2u + (global_pair_id % pairDistance);
            uint32_t rightId = leftId + pairDistance;
            if (rightId >= n || leftId >= n) continue;
            int ascending = ((global_pair_id & (blockWidth >> 1)) == 0u);
            int32_t lp = light_pixel_indices[leftId];
            int32_t rp = light_pixel_indices[rightId];
            if ((ascending && lp > rp) || (!ascending && lp < rp)) {
                int32_t ti;
                float tf;
                light_pixel_indices[leftId] = rp; light_pixel_indices[rightId] = lp;
                ti = light_path_indices[leftId]; light_path_indices[leftId] = light_path_indices[rightId]; light_path_indices[rightId] = ti;
                ti = light_ray_indices[leftId]; light_ray_indices[leftId] = light_ray_indices[rightId]; light_ray_indices[rightId] = ti;
                tf = light_weights[leftId]; light_weights[leftId] = light_weights[rightId]; light_weights[rightId] = tf;
                tf = light_shade[leftId]; light_shade[leftId] = light_shade[rightId]; light_shade[rightId] = tf;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * kernel: light_image_gather (trace.metal:937-964)
 * ---------------------------------------------------------------------------------------------------------- */
EXPORT void ref_light_image_gather(uint32_t n, const Path *light_paths, const Material *materials,
                                   const int32_t *path_indices, const int32_t *ray_indices, const int32_t *bins,
                                   const uint32_t *offset_p, const float *weights, const float *shades,
                                   float4 *light_image, float *sum_weights) {
    uint32_t offset = *offset_p;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t id = 0; id < (int64_t)n; id++) {
        int start_idx = bins[id];
        int end_idx = bins[id + 1];
        float3 total = v3(0, 0, 0);
        float weight_sum = 0.0f;
        for (uint32_t i = (uint32_t)start_idx + offset; i < (uint32_t)end_idx + offset; i++) {
            int path_idx = path_indices[i];
            int ray_idx = ray_indices[i];
            const Path *path = &light_paths[path_idx];
            const Ray *ray = &path->rays[ray_idx];
            const Ray *prior_ray = &path->rays[(ray_idx - 1 > 0) ? ray_idx - 1 : 0];
            const Material *mat = &materials[ray->material];
            total = vadd(total, vmul(vscale(prior_ray->color, weights[i] * shades[i]), mat->color));
            weight_sum += weights[i];
        }
        light_image[id] = (float4){total.x, total.y, total.z, 1.0f};
        sum_weights[id] += weight_sum;
    }
}

/* kernel: reset_light_indices (trace.metal:967-978) */
EXPORT void ref_reset_light_indices(uint32_t n, int32_t *light_pixel_indices, int32_t *light_path_indices,
                                    int32_t *light_ray_indices, float *light_weights, float *light_shade) {
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n; id++) {
        light_pixel_indices[id] = -1;
        light_path_indices[id] = 0;
        light_ray_indices[id] = 0;
        light_weights[id] = 0.0f;
        light_shade[id] = 0.0f;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * kernel: adaptive_finalize_samples (trace.metal:981-1018)
 * ---------------------------------------------------------------------------------------------------------- */
EXPORT void ref_adaptive_finalize_samples(uint32_t n, const WeightAggregator *weight_aggregators,
                                          const Camera *camera_buffer, float4 *out, uint32_t *sample_counts,
                                          const uint32_t *sample_bin_offsets, float *sample_weights) {
    Camera camera = camera_buffer[0];
#pragma omp parallel for schedule(static)
    for (int64_t idl = 0; idl < (int64_t)n; idl++) {
        uint32_t id = (uint32_t)idl;
        float3 total_sample = v3(0, 0, 0);
        float weight_sum = 0.0f;
        for (int i = -1; i < 2; i++) {
            for (int j = -1; j < 2; j++) {
                int sample_x = (int)(id % (uint32_t)camera.pixel_width) + i;
                int sample_y = (int)(id / (uint32_t)camera.pixel_width) + j;
                if (sample_x < 0 || sample_x >= camera.pixel_width || sample_y < 0 || sample_y >= camera.pixel_height)
                    continue;
                int sample_index = sample_y * camera.pixel_width + sample_x;
                if (sample_index < 0 || sample_index >= camera.pixel_width * camera.pixel_height) continue;
                for (uint32_t k = sample_bin_offsets[sample_index]; k < sample_bin_offsets[sample_index + 1]; k++) {
                    const WeightAggregator *wa = &weight_aggregators[k];
                    float weight = wa->weights[1 - i][1 - j];
                    total_sample = vadd(total_sample, vscale(wa->total_contribution, weight));
                    weight_sum += weight * wa->contrib_weight_sum;
                }
            }
        }
        sample_counts[id] = sample_bin_offsets[id + 1] - sample_bin_offsets[id];
        out[id] = (float4){total_sample.x, total_sample.y, total_sample.z, 1.0f};
        sample_weights[id] = weight_sum;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * kernel: generate_camera_rays (trace.metal:1020-1067)
 * ---------------------------------------------------------------------------------------------------------- */
EXPORT void ref_generate_camera_rays(uint32_t n, const Camera *camera, uint32_t *random_buffer, const uint32_t *indices,
                                     Ray *out) {
    Camera c = camera[0];
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n; id++) {
        Ray ray;
        memset(&ray, 0, sizeof ray);
        uint32_t seed0 = random_buffer[2 * id];
        uint32_t seed1 = random_buffer[2 * id + 1];
        float x_offset = xorshift_random(&seed0);
        float y_offset = xorshift_random(&seed1);
        int pixel_idx = (int)indices[id];
        int pixel_x = pixel_idx % c.pixel_width;
        int pixel_y = pixel_idx / c.pixel_width;
        float x_normalized = ((float)pixel_x + x_offset - 0.5f * (float)c.pixel_width) / (float)c.pixel_width;
        float y_normalized = ((float)pixel_y + y_offset - 0.5f * (float)c.pixel_height) / (float)c.pixel_height;
        float3 x_vector = vscale(vscale(c.dx, x_normalized), c.phys_width);
        float3 y_vector = vscale(vscale(c.dy, y_normalized), c.phys_height);
        float3 origin = vadd(vadd(c.center, x_vector), y_vector);
        float3 direction = vnormalize(vsub(c.focal_point, origin));
        ray.origin = origin;
        ray.direction = direction;
        ray.normal = c.direction;
        ray.inv_direction = vrcp(direction);
        ray.color = v3(1.0f, 1.0f, 1.0f);
        ray.material = 7;
        ray.triangle = -1;
        ray.hit_light = -1;
        ray.hit_camera = -1;
        ray.from_camera = 1;
        ray.c_importance = 1.0f / (c.phys_width * c.phys_height);
        ray.l_importance = 1.0f;
        ray.tot_importance = ray.c_importance;
        ray.pixel_idx = pixel_idx;
        out[id] = ray;
        random_buffer[2 * id] = seed0;
        random_buffer[2 * id + 1] = seed1;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * kernel: generate_light_rays (trace.metal:1070-1124)
 * ---------------------------------------------------------------------------------------------------------- */
EXPORT void ref_generate_light_rays(uint32_t n, const Triangle *light_triangles, const float *surface_areas,
                                    const int32_t *light_triangle_indices, const Material *materials,
                                    uint32_t *random_buffer, Ray *out, const int32_t *counts) {
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n; id++) {
        Ray ray;
        memset(&ray, 0, sizeof ray);
        ray.from_camera = 0;
        ray.hit_light = -1;
        ray.hit_camera = -1;
        uint32_t seed0 = random_buffer[2 * id];
        uint32_t seed1 = random_buffer[2 * id + 1];

        int light_count = counts[0];
        int light_index = (int)(xorshift_random(&seed0) * (float)light_count);
        if (light_index > light_count - 1) light_index = light_count - 1; /* deviation A2 */
        Triangle light_triangle = light_triangles[light_index];
        float surface_area = surface_areas[light_index];

        float u = xorshift_random(&seed0);
        float v = xorshift_random(&seed1);
        if (u + v > 1.0f) {
            u = 1.0f - u;
            v = 1.0f - v;
        }
        float w = 1.0f - u - v;

        ray.normal = light_triangle.normal;
        ray.origin = vadd(vadd(vadd(vscale(light_triangle.v0, u), vscale(light_triangle.v1, v)),
                               vscale(light_triangle.v2, w)),
                          vscale(ray.normal, DELTA));
        float3 x, y;
        orthonormal(ray.normal, &x, &y);
        float rand_x = xorshift_random(&seed0);
        float rand_y = xorshift_random(&seed1);
        float2 roll = {rand_x, rand_y};
        ray.direction = random_hemisphere_uniform(x, y, ray.normal, roll);
        ray.inv_direction = vrcp(ray.direction);

        ray.material = light_triangle.material;
        Material material = materials[ray.material];
        ray.color = material.emission;
        ray.triangle = light_triangle_indices[light_index];
        ray.c_importance = 1.0f;
        ray.l_importance = 1.0f / ((float)light_count * surface_area);
        ray.tot_importance = ray.l_importance;

        out[id] = ray;
        random_buffer[2 * id] = seed0;
        random_buffer[2 * id + 1] = seed1;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * parity hooks: closest hit for arbitrary rays, and the visibility / projection predicates
 * ---------------------------------------------------------------------------------------------------------- */
EXPORT void ref_traverse(uint32_t n, const Ray *rays, const Box *boxes, const Triangle *triangles, int32_t *tri,
                         float *t, float *u, float *v) {
#pragma omp parallel
    {
        Counters cnt = {0, 0, 0};
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)n; i++) {
            int best_i = -1;
            float best_t = INFINITY, uu = 0.0f, vv = 0.0f;
            traverse_bvh(&rays[i], boxes, triangles, &best_i, &best_t, &uu, &vv, &cnt);
            tri[i] = best_i;
            t[i] = best_t;
            u[i] = uu;
            v[i] = vv;
        }
        counters_merge(&cnt);
    }
}

EXPORT void ref_visibility(uint32_t n, const Ray *a, const Ray *b, const Box *boxes, const Triangle *triangles,
                           int32_t *visible) {
#pragma omp parallel
    {
        Counters cnt = {0, 0, 0};
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)n; i++) visible[i] = visibility_test(&a[i], &b[i], boxes, triangles, &cnt);
        counters_merge(&cnt);
    }
}

/* math probes so tests can compare clv_math.h with libm through either build */
EXPORT void ref_math_probe(uint32_t n, const float *x, int which, float *y) {
    for (uint32_t i = 0; i < n; i++) {
        switch (which) {
        case 0: y[i] = M_SIN(x[i]); break;
        case 1: y[i] = M_COS(x[i]); break;
        case 2: y[i] = M_ACOS(x[i]); break;
        case 3: y[i] = M_ATAN(x[i]); break;
        default: y[i] = M_EXP(x[i]); break;
        }
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline is meant to use all host cores */
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
