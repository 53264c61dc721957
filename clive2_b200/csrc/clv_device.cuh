/* clv_device.cuh — per-sample device code of the BDPT hot path (sm_100a), shared by every kernel in clv_kernels.cu.
 *
 * Everything here is `__host__ __device__`: the product only ever runs the device instantiation
 * (libclive_b200.so has no host execution path and fails when no GPU is present); tests/hostemu compiles the same
 * functions for the host so the kernel logic can be diffed against the oracle in a container without a GPU.
 *
 * Arithmetic contract: IEEE binary32, one rounding per source operation, NO fused multiply-add
 * (the translation unit is compiled with -fmad=false), correctly rounded div/sqrt (nvcc defaults), elementary
 * functions from clv_math.h.  With that contract every buffer written by these kernels equals, bit for bit, what
 * oracle/trace_ref.c built with -DCLV_SHARED_MATH writes for the same inputs.
 *
 * What is restated from the reference (src/trace.metal) and what is new:
 *   records          trace.metal:7-85            same layouts (include/clive_b200.h)
 *   RNG              trace.metal:87-93           same integer stream
 *   slab / triangle  trace.metal:106-142         same operations on a compact layout (child boxes stored in the parent,
 *                                                v0/e1/e2 per triangle with e1,e2 pre-subtracted in fp32 = same bits)
 *   traversal        trace.metal:144-176         REFERENCE mode = same visiting order; ORDERED mode = near child first
 *                                                with a (t, reference-DFS-rank) tie rule, same result
 *   shadow rays      trace.metal:178-197         target triangle tested first, early exit on the first occluder
 *   projection rays  trace.metal:569-617         sensor triangles tested first, traversal skipped when they miss
 *   BSDFs            trace.metal:200-379         same formulas
 *   random walk      trace.metal:381-532         vertices streamed to the Path record as they are produced
 *   connection       trace.metal:620-869         per-sample MIS terms staged once (shared memory), per-strategy work
 *                                                reduced to the connecting edge
 */
#ifndef CLV_DEVICE_CUH
#define CLV_DEVICE_CUH

#include <math.h>
#include <stdint.h>

#include "../../include/clive_b200.h"
#include "clv_math.h"

#if defined(__CUDACC__)
#define CLV_HD __host__ __device__ __forceinline__
#define CLV_HDM __host__ __device__ __forceinline__ /* member functions */
#else
#define CLV_HD static inline
#define CLV_HDM inline
#endif

#define CLV_PI 3.14159265359f /* trace.metal:4 */
#define CLV_DELTA 0.0001f     /* trace.metal:5 */
#define CLV_INF (clv_bits_to_float(0x7f800000u))
#define CLV_MAX_BOUNCES 6     /* trace.metal:407 */
/* Traversal stacks (thread-private local memory).  The reference's stack has 64 int entries (trace.metal:145) and its
 * search silently stops when it fills; here a search defers at most ONE sibling per level of the binary tree (three per
 * level of the 4-wide collapse), pruned children are never pushed, and clv_accel_create REFUSES a tree deeper than
 * these bounds (CLV_ERR_INVALID) instead of dropping pushes: the `sp < CLV_*STACK` tests below are memory-safety guards
 * that never fire on an accepted structure.  Closest-hit entries are (t_entry, ref) packed into one 64-bit word: half
 * the local-memory instructions of two 32-bit arrays, 320 B per thread instead of 768 B.  (The 10 M-triangle tree of
 * BASELINE config 4 is 29 binary levels deep.) */
/* which node layout the ORDERED flavour traverses: the 4-wide collapse of the BVH (1) or the binary nodes (0),
 * separately for closest-hit searches (random walks) and occlusion queries (connection rays) */
#ifndef CLV_WIDE_CLOSEST
#define CLV_WIDE_CLOSEST 0
#endif
#ifndef CLV_WIDE_OCCLUSION
#define CLV_WIDE_OCCLUSION 1
#endif
#define CLV_STACK (CLV_WIDE_CLOSEST ? 96 : 40) /* closest-hit: 64-bit (t, ref) entries */
#define CLV_OCC_STACK 64                       /* occlusion queries over the 4-wide nodes: 32-bit refs */
#define CLV_MAX_BINARY_DEPTH (CLV_WIDE_CLOSEST ? CLV_OCC_STACK / 3 : CLV_STACK) /* deepest accepted tree (inner levels) */
#define CLV_MAX_WIDE_DEPTH (CLV_OCC_STACK / 3)

namespace clv {

/* ---- vectors ------------------------------------------------------------------------------------------------ */
struct alignas(16) V4 { float x, y, z, w; };
struct V3 { float x, y, z; };

CLV_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
CLV_HD V3 xyz(const V4 &a) { return v3(a.x, a.y, a.z); }
CLV_HD V4 pad0(const V3 &a) { V4 r; r.x = a.x; r.y = a.y; r.z = a.z; r.w = 0.0f; return r; }
CLV_HD V4 v4(float x, float y, float z, float w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
CLV_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
CLV_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
CLV_HD V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
CLV_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
CLV_HD V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
CLV_HD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
CLV_HD V3 rcp(V3 a) { return v3(1.0f / a.x, 1.0f / a.y, 1.0f / a.z); }
CLV_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CLV_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
CLV_HD float length(V3 a) { return sqrtf(dot(a, a)); }
CLV_HD V3 normalize(V3 a) { return a / length(a); }

/* ---- records (same bytes as include/clive_b200.h, 16-byte aligned so that copies compile to 128-bit accesses) -- */
struct alignas(16) Ray {
    V4 origin, direction, inv_direction, color, normal;
    int32_t material, triangle;
    float c_importance, l_importance, tot_importance;
    int32_t hit_light, from_camera, hit_camera, pixel_idx;
    int32_t pad[3];
};
struct alignas(16) Path {
    Ray rays[8];
    int32_t length, from_camera, pad[2];
};
struct alignas(16) Box { V4 min, max; int32_t left, right, pad[2]; };
struct alignas(16) Triangle {
    V4 v0, v1, v2, n0, n1, n2, normal;
    int32_t material, is_light, is_camera, pad;
};
struct alignas(16) Material { V4 color, emission; int32_t type; float alpha, ior; int32_t pad; };
struct alignas(16) Camera {
    V4 center, focal_point, direction, dx, dy;
    int32_t pixel_width, pixel_height;
    float phys_width, phys_height, h_fov, v_fov;
    int32_t pad[2];
};
struct alignas(16) WeightAggregator {
    float weights[3][3];
    float pad0[3];
    V4 total_contribution;
    float contrib_weight_sum;
    float pad1[3];
};
static_assert(sizeof(Ray) == sizeof(clv_Ray) && sizeof(Ray) == 128, "Ray");
static_assert(sizeof(Path) == sizeof(clv_Path) && sizeof(Path) == 1040, "Path");
static_assert(sizeof(Box) == sizeof(clv_Box) && sizeof(Box) == 48, "Box");
static_assert(sizeof(Triangle) == sizeof(clv_Triangle) && sizeof(Triangle) == 128, "Triangle");
static_assert(sizeof(Material) == sizeof(clv_Material) && sizeof(Material) == 48, "Material");
static_assert(sizeof(Camera) == sizeof(clv_Camera) && sizeof(Camera) == 112, "Camera");
static_assert(sizeof(WeightAggregator) == sizeof(clv_WeightAggregator) && sizeof(WeightAggregator) == 80, "WA");

/* L2 eviction hints (measured option): BVH node / triangle records are re-read by every ray while the per-sample
 * records (path vertices, walk state, hit records: GBs per pass) stream through once; with CLV_EVICT_HINTS the former are
 * loaded with evict_last and the latter stored with the streaming hint, so that they do not push each other out of L2 */
#ifndef CLV_EVICT_HINTS
#define CLV_EVICT_HINTS 0
#endif
/* read-only 16-byte load through the non-coherent path */
CLV_HD V4 ld4(const V4 *p) {
#if defined(__CUDA_ARCH__)
#if CLV_EVICT_HINTS
    float4 f;
    unsigned long long policy;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "l"(p), "l"(policy));
    return v4(f.x, f.y, f.z, f.w);
#else
    float4 f = __ldg(reinterpret_cast<const float4 *>(p));
    return v4(f.x, f.y, f.z, f.w);
#endif
#else
    return *p;
#endif
}
/* 16-byte store of a record that is written once and read once much later (path vertices) */
CLV_HD void st4_stream(V4 *p, V4 v) {
#if defined(__CUDA_ARCH__) && CLV_EVICT_HINTS
    __stcs(reinterpret_cast<float4 *>(p), make_float4(v.x, v.y, v.z, v.w));
#else
    *p = v;
#endif
}
/* Software prefetch of traversal records into L1 (device only; measured options, profiles/r02_notes.md).
 * CLV_PREFETCH_LEAF: a leaf step first requests every 128-byte line of its triangle range, so that the per-triangle loads
 * of its loop overlap instead of exposing one memory round trip each.  CLV_PREFETCH_NEXT: the record a search will read
 * next (inner node or first lines of a leaf) is requested the moment it is chosen — the lane may wait several warp
 * iterations for its kind of step to win the vote. */
#ifndef CLV_PREFETCH_LEAF
#define CLV_PREFETCH_LEAF 0
#endif
#ifndef CLV_PREFETCH_NEXT
#define CLV_PREFETCH_NEXT 0
#endif
CLV_HD void prefetch_l1(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
/* every 128-byte line of the triangle records [first, first + count) */
CLV_HD void prefetch_leaf(const V4 *tris, int first, int count) {
#if defined(__CUDA_ARCH__)
    const char *p = (const char *)(tris + 3 * (size_t)first);
    const char *end = p + 48 * count;
    p = (const char *)((size_t)p & ~(size_t)127);
    for (; p < end; p += 128) prefetch_l1(p);
#else
    (void)tris; (void)first; (void)count;
#endif
}
/* the record behind a child ref: one binary node (64 B) / 4-wide node (128 B), or a leaf's triangle range */
template <bool WIDE>
CLV_HD void prefetch_ref(const V4 *nodes, const V4 *tris, int ref) {
#if defined(__CUDA_ARCH__)
    if (ref >= 0) prefetch_l1(nodes + (WIDE ? 8 : 4) * (size_t)ref);
    else prefetch_leaf(tris, (~ref) >> 4, (~ref) & 15);
#else
    (void)nodes; (void)tris; (void)ref;
#endif
}
/* 16-byte load / store that bypass L1 (served by L2, the point of coherence between SMs): for records one SM writes and
 * another reads INSIDE one kernel (walk state and hit records of the pipelined walks) */
CLV_HD V4 ldcg4(const V4 *p) {
#if defined(__CUDA_ARCH__)
    float4 f = __ldcg(reinterpret_cast<const float4 *>(p));
    return v4(f.x, f.y, f.z, f.w);
#else
    return *p;
#endif
}
CLV_HD void stcg4(V4 *p, V4 v) {
#if defined(__CUDA_ARCH__)
    __stcg(reinterpret_cast<float4 *>(p), make_float4(v.x, v.y, v.z, v.w));
#else
    *p = v;
#endif
}
CLV_HD int f2i(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}
CLV_HD float i2f(int i) { return clv_bits_to_float((uint32_t)i); }
/* one closest-hit stack entry: entry distance of a deferred child in the high word, its ref in the low word */
typedef unsigned long long StackEntry;
CLV_HD StackEntry stack_entry(float t, int ref) { return ((StackEntry)(uint32_t)f2i(t) << 32) | (StackEntry)(uint32_t)ref; }
CLV_HD float entry_t(StackEntry e) { return i2f((int)(uint32_t)(e >> 32)); }
CLV_HD int entry_ref(StackEntry e) { return (int)(uint32_t)e; }

/* ---- work counters (instrumented kernel variants only; the timed variants compile all of this out) -------------- */
struct Counters {
    unsigned long long rays_ref;     /* traverse_bvh invocations the reference would make */
    unsigned long long rays_traced;  /* traversals this implementation actually starts */
    unsigned long long node_fetches; /* 64-byte inner-node records fetched */
    unsigned long long tri_fetches;  /* 48-byte triangle records fetched */
    unsigned long long vertices;     /* 128-byte path vertices written (walks) or staged (connection) */
};
#define CLV_COUNTER_WORDS 5
enum { COUNTERS_WALK = 0, COUNTERS_CONNECT = 1, COUNTERS_CLASSES = 2 };
template <bool COUNT> struct Cnt;
template <> struct Cnt<false> {
    CLV_HDM void ray_ref() {} CLV_HDM void ray_traced() {} CLV_HDM void node() {} CLV_HDM void tri() {}
    CLV_HDM void vertex(int) {}
};
template <> struct Cnt<true> {
    unsigned int rr = 0, rt = 0, nf = 0, tf = 0, vx = 0;
    CLV_HDM void ray_ref() { rr++; } CLV_HDM void ray_traced() { rt++; } CLV_HDM void node() { nf++; } CLV_HDM void tri() { tf++; }
    CLV_HDM void vertex(int n) { vx += (unsigned)n; }
};

/* ---- RNG (trace.metal:87-93) -------------------------------------------------------------------------------- */
CLV_HD float xorshift_random(uint32_t &seed) {
    seed ^= seed << 13;
    seed ^= seed >> 17;
    seed ^= seed << 5;
    return (float)seed / 4294967296.0f; /* (float)0xFFFFFFFF rounds to 2^32 */
}

/* ---- acceleration structure view ------------------------------------------------------------------------------ *
 * nodes: 4 x V4 per inner node  [lmin.xyz lmax.x] [lmax.yz rmin.xy] [rmin.z rmax.xyz] [lref rref - -]
 *        a ref >= 0 is an inner node index, a ref < 0 is a leaf: ~ref = (first_triangle << 4) | count, count 1..15
 * tris : 3 x V4 per triangle in REFERENCE order [v0.xyz rank] [e1.xyz -] [e2.xyz -], e1 = v1 - v0, e2 = v2 - v0 (fp32),
 *        rank = position of the triangle in the reference's traversal order with no pruning (trace.metal:158-161)
 */
struct AccelView {
    const V4 *nodes;
    const V4 *wnodes;  /* 4-wide nodes of the ORDERED flavour (clv_accel_build.h WideRec), 8 x V4 each */
    int32_t wroot_ref;
    const V4 *tris;
    /* shading data of every triangle, 4 x V4 = 64 bytes (one aligned fetch of two sectors instead of the 80 bytes a hit
     * touches in the reference's 128-byte Triangle record): [n0.xyz material] [n1.xyz is_light] [n2.xyz is_camera]
     * [normal.xyz -], integers as bits */
    const V4 *shade;
    V4 root_min, root_max;
    int32_t root_ref;
    int32_t n_camera_tris;
    const int32_t *camera_tris; /* indices of is_camera triangles */
};

CLV_HD void slab_test(V3 o, V3 inv, V3 bmin, V3 bmax, bool &hit, float &t_entry) { /* trace.metal:106-115, t = INF */
    V3 t0s = (bmin - o) * inv;
    V3 t1s = (bmax - o) * inv;
    float tminx = fminf(t0s.x, t1s.x), tminy = fminf(t0s.y, t1s.y), tminz = fminf(t0s.z, t1s.z);
    float tmaxx = fmaxf(t0s.x, t1s.x), tmaxy = fmaxf(t0s.y, t1s.y), tmaxz = fmaxf(t0s.z, t1s.z);
    float tmin_final = fmaxf(fmaxf(tminx, tminy), fmaxf(tminz, 0.0f));
    float tmax_final = fminf(fminf(tmaxx, tmaxy), fminf(tmaxz, CLV_INF));
    hit = tmin_final <= tmax_final;
    t_entry = tmin_final;
}

/* one 4-wide node: the reference's slab test on each of its up to four boxes.  t[k] = entry distance of slot k if the
 * ray passes it with t_entry <= t_limit, +INF otherwise (also for empty slots); ref[k] = the slot's child ref. */
CLV_HD void wide_test(const V4 *n, V3 o, V3 inv, float t_limit, float t[4], int ref[4]) {
    V4 mnx = ld4(n), mny = ld4(n + 1), mnz = ld4(n + 2), mxx = ld4(n + 3), mxy = ld4(n + 4), mxz = ld4(n + 5), rf = ld4(n + 6);
    const float bmnx[4] = {mnx.x, mnx.y, mnx.z, mnx.w}, bmny[4] = {mny.x, mny.y, mny.z, mny.w}, bmnz[4] = {mnz.x, mnz.y, mnz.z, mnz.w};
    const float bmxx[4] = {mxx.x, mxx.y, mxx.z, mxx.w}, bmxy[4] = {mxy.x, mxy.y, mxy.z, mxy.w}, bmxz[4] = {mxz.x, mxz.y, mxz.z, mxz.w};
    const float rfk[4] = {rf.x, rf.y, rf.z, rf.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        bool hit;
        float te;
        slab_test(o, inv, v3(bmnx[k], bmny[k], bmnz[k]), v3(bmxx[k], bmxy[k], bmxz[k]), hit, te);
        ref[k] = f2i(rfk[k]);
        t[k] = (ref[k] != -1 && hit && te <= t_limit) ? te : CLV_INF;
    }
}
/* order the four slots by descending t (empty/missed slots, t = +INF, first; the nearest one last) */
CLV_HD void wide_sort(float t[4], int ref[4]) {
#define CLV_CSWAP(a, b) { bool sw = t[a] < t[b]; float tt = sw ? t[b] : t[a]; int rr = sw ? ref[b] : ref[a]; \
                          t[b] = sw ? t[a] : t[b]; ref[b] = sw ? ref[a] : ref[b]; t[a] = tt; ref[a] = rr; }
    CLV_CSWAP(0, 1) CLV_CSWAP(2, 3) CLV_CSWAP(0, 2) CLV_CSWAP(1, 3) CLV_CSWAP(1, 2)
#undef CLV_CSWAP
}

/* Moeller-Trumbore exactly as trace.metal:117-142 with edge1/edge2 already formed */
CLV_HD bool tri_test(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float &t_out, float &u, float &v) {
    V3 h = cross(d, e2);
    float a = dot(e1, h);
    float f = 1.0f / a;
    V3 s = o - v0;
    u = f * dot(s, h);
    if (u < 0 || u > 1) return false;
    V3 q = cross(s, e1);
    v = f * dot(d, q);
    if (v < 0 || u + v > 1) return false;
    float t = f * dot(e2, q);
    if (t > CLV_DELTA) { t_out = t; return true; }
    return false;
}

enum { TRAVERSE_REFERENCE = 0, TRAVERSE_ORDERED = 1 };

/* Closest hit.  Result contract = trace.metal:144-176: the first triangle, in the reference's visiting order, that
 * attains the minimum t among the triangles whose boxes pass `hit && t_entry < best_t`.
 *   REFERENCE: visits right child first like the reference (children tested when the parent is expanded, the deferred
 *              one re-checked against best_t when popped = the same predicate the reference evaluates at pop time).
 *   ORDERED  : near child first; exact ties in t resolved by the reference rank, boxes kept while t_entry <= best_t so
 *              tie candidates are still reached.
 * `stop_below`/`stop_skip`: early exit for predicate rays — as soon as a triangle other than `stop_skip` is accepted
 * with t < stop_below the search ends (the caller only needs to know that such a hit exists).  Pass -INF to disable.
 */
template <int MODE, bool COUNT>
CLV_HD void closest_hit(const AccelView &A, V3 o, V3 d, V3 inv, int &best_i, float &best_t, float &bu, float &bv,
                        Cnt<COUNT> &cnt, float stop_below, int stop_skip) {
    /* Control flow is "while-while" with a single back edge per loop and no `continue`: lanes of a warp that sit at
     * inner nodes step together until every lane has reached a leaf (or finished), then all leaves are processed
     * together.  (A loop with several back edges lets the lanes drift apart for good: the first version of this
     * function ran its slab tests with 5 of 32 lanes active, profiles/r01_c2_ncu_v1.md.) */
    StackEntry st[CLV_STACK];
    int sp = 0;
    int best_rank = 0x7fffffff;
    cnt.ray_traced();

    bool hit;
    float te;
    slab_test(o, inv, xyz(A.root_min), xyz(A.root_max), hit, te);
    if (!(hit && te < best_t)) return;
    const bool wide = (MODE == TRAVERSE_ORDERED) && CLV_WIDE_CLOSEST; /* ORDERED walks the 4-wide collapse of the tree */
    int cur = wide ? A.wroot_ref : A.root_ref;
    bool have = true; /* `cur` holds a node that still has to be processed */

    /* one inner node: test both child boxes, continue with the nearer surviving child, keep the other */
    auto inner_step = [&]() {
        if (wide) {
            /* up to four children at once: continue with the nearest that survives, keep the others farthest first */
            float tk[4];
            int rk[4];
            wide_test(A.wnodes + 8 * (size_t)cur, o, inv, best_t, tk, rk);
            cnt.node();
            wide_sort(tk, rk);
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (tk[k] < CLV_INF && sp < CLV_STACK) { st[sp] = stack_entry(tk[k], rk[k]); sp++; }
            have = tk[3] < CLV_INF;
            cur = rk[3];
            while (!have && sp > 0) {
                sp--;
                const StackEntry e = st[sp];
                if (entry_t(e) <= best_t) { cur = entry_ref(e); have = true; }
            }
            return;
        }
        const V4 *n = A.nodes + 4 * (size_t)cur;
        V4 n0 = ld4(n), n1 = ld4(n + 1), n2 = ld4(n + 2), n3 = ld4(n + 3);
        cnt.node();
        bool hl, hr;
        float tl, tr;
        slab_test(o, inv, v3(n0.x, n0.y, n0.z), v3(n0.w, n1.x, n1.y), hl, tl);
        slab_test(o, inv, v3(n1.z, n1.w, n2.x), v3(n2.y, n2.z, n2.w), hr, tr);
        int lref = f2i(n3.x), rref = f2i(n3.y);
        bool go_l, go_r, push;
        int push_ref;
        float push_t;
        if (MODE == TRAVERSE_REFERENCE) {
            /* the reference pushes left then right and pops right first; the deferred left child is re-checked
             * against best_t when popped = the predicate the reference evaluates at pop time */
            go_r = hr && tr < best_t;
            go_l = hl && tl < best_t;
            push = go_r && hl;
            push_ref = lref;
            push_t = tl;
            cur = go_r ? rref : lref;
            have = go_r || go_l;
        } else {
            go_l = hl && tl <= best_t;
            go_r = hr && tr <= best_t;
            bool left_first = tl < tr; /* tie: right first, the reference's order */
            push = go_l && go_r;
            push_ref = left_first ? rref : lref;
            push_t = left_first ? tr : tl;
            cur = (go_l && (left_first || !go_r)) ? lref : rref;
            have = go_l || go_r;
        }
        if (push && sp < CLV_STACK) { st[sp] = stack_entry(push_t, push_ref); sp++; }
        while (!have && sp > 0) {
            sp--;
            const StackEntry e = st[sp];
            const float t = entry_t(e);
            if ((MODE == TRAVERSE_REFERENCE) ? (t < best_t) : (t <= best_t)) { cur = entry_ref(e); have = true; }
        }
    };
    /* one leaf: test its triangles */
    auto leaf_step = [&]() {
        int code = ~cur;
        int first = code >> 4, count = code & 15;
        bool stop = false;
        for (int i = first; i < first + count; i++) {
            const V4 *tp = A.tris + 3 * (size_t)i;
            V4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
            cnt.tri();
            float t, u, v;
            if (tri_test(o, d, xyz(a), xyz(b), xyz(c), t, u, v)) {
                bool accept;
                if (MODE == TRAVERSE_REFERENCE) accept = t < best_t;
                /* a "hit" at t = +inf (ray parallel to the triangle's plane: f = 1/0, u and v NaN, which pass the
                 * reference's range checks) is never accepted by the reference's strict t < best_t: no tie at INF */
                else accept = t < best_t || (t == best_t && t < CLV_INF && f2i(a.w) < best_rank);
                if (accept) {
                    best_i = i; best_t = t; bu = u; bv = v;
                    best_rank = f2i(a.w);
                    stop = stop || (i != stop_skip && t < stop_below);
                }
            }
        }
        have = false;
        if (stop) sp = 0; /* early exit of a predicate ray: nothing left to look at */
        while (!have && sp > 0) {
            sp--;
            const StackEntry e = st[sp];
            const float t = entry_t(e);
            if ((MODE == TRAVERSE_REFERENCE) ? (t < best_t) : (t <= best_t)) { cur = entry_ref(e); have = true; }
        }
    };
    /* while-while: lanes at inner nodes step together until every lane has reached a leaf (or finished), then all
     * leaves are processed together.  (Measured alternatives inside this kernel — a majority vote between node and
     * leaf steps as in k_connect_rays, bounded bursts of node steps — were slower: profiles/r01_notes.md.) */
    while (have) {
        while (have && cur >= 0) inner_step();
        if (have) leaf_step();
    }
}

/* ---- resumable closest hit (wavefront walks) ------------------------------------------------------------------------ *
 * The ORDERED search of closest_hit over the binary nodes, cut into single steps — one inner node or one whole leaf — so
 * that a warp can hold 32 rays at different stages, run ONE kind of step per iteration for the lanes that wait for it
 * (majority vote) and hand a finished lane the next ray of the queue (k_walk_trace).  Same operations per ray in the same
 * order as closest_hit<TRAVERSE_ORDERED>, hence the same (best_i, best_t, u, v) bits; tests/hostemu checks both against
 * the oracle.  Scalars only (registers); the stack of packed (t, ref) entries is a separate local array. */
struct HitRay {
    V3 o, d, inv;
    float best_t, bu, bv;
    int best_i, best_rank;
    int cur, sp;
};
enum { HIT_RUNNING = 0, HIT_DONE = 1 };

/* stack entries a resumable search needs: one deferred sibling per binary level, three per 4-wide level */
#define CLV_HIT_STACK(WIDE) ((WIDE) ? CLV_OCC_STACK : CLV_STACK)
template <bool COUNT, bool WIDE = false>
CLV_HD int hit_begin(const AccelView &A, HitRay &r, Cnt<COUNT> &cnt) {
    cnt.ray_traced();
    r.best_i = -1; r.best_t = CLV_INF; r.bu = 0.0f; r.bv = 0.0f; r.best_rank = 0x7fffffff;
    r.sp = 0;
    bool hit;
    float te;
    slab_test(r.o, r.inv, xyz(A.root_min), xyz(A.root_max), hit, te);
    if (!(hit && te < r.best_t)) return HIT_DONE;
    r.cur = WIDE ? A.wroot_ref : A.root_ref;
    return HIT_RUNNING;
}
/* next deferred child that can still hold a closer (or tying) hit, or the end of the search */
CLV_HD int hit_pop(HitRay &r, const StackEntry *st) {
    while (r.sp > 0) {
        r.sp--;
        const StackEntry e = st[r.sp];
        if (entry_t(e) <= r.best_t) { r.cur = entry_ref(e); return HIT_RUNNING; }
    }
    return HIT_DONE;
}
template <bool COUNT>
CLV_HD int hit_inner_step(const AccelView &A, HitRay &r, StackEntry *st, Cnt<COUNT> &cnt) {
    const V4 *n = A.nodes + 4 * (size_t)r.cur;
    V4 n0 = ld4(n), n1 = ld4(n + 1), n2 = ld4(n + 2), n3 = ld4(n + 3);
    cnt.node();
    bool hl, hr;
    float tl, tr;
    slab_test(r.o, r.inv, v3(n0.x, n0.y, n0.z), v3(n0.w, n1.x, n1.y), hl, tl);
    slab_test(r.o, r.inv, v3(n1.z, n1.w, n2.x), v3(n2.y, n2.z, n2.w), hr, tr);
    const int lref = f2i(n3.x), rref = f2i(n3.y);
    const bool go_l = hl && tl <= r.best_t, go_r = hr && tr <= r.best_t;
    const bool left_first = tl < tr; /* tie: right first, the reference's order */
    if (go_l && go_r && r.sp < CLV_STACK) { st[r.sp] = stack_entry(left_first ? tr : tl, left_first ? rref : lref); r.sp++; }
    int status = HIT_RUNNING;
    if (go_l || go_r) r.cur = (go_l && (left_first || !go_r)) ? lref : rref;
    else status = hit_pop(r, st);
    if (CLV_PREFETCH_NEXT && status == HIT_RUNNING) prefetch_ref<false>(A.nodes, A.tris, r.cur);
    return status;
}
/* the same step over the 4-wide collapse of the tree (clv_accel_build.h): up to four child boxes per record, continue with
 * the nearest survivor, keep the others farthest first.  Half the dependent record fetches per ray of the binary form. */
template <bool COUNT>
CLV_HD int hit_inner_step_wide(const AccelView &A, HitRay &r, StackEntry *st, Cnt<COUNT> &cnt) {
    float tk[4];
    int rk[4];
    wide_test(A.wnodes + 8 * (size_t)r.cur, r.o, r.inv, r.best_t, tk, rk);
    cnt.node();
    wide_sort(tk, rk);
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (tk[k] < CLV_INF && r.sp < CLV_HIT_STACK(true)) { st[r.sp] = stack_entry(tk[k], rk[k]); r.sp++; }
    int status = HIT_RUNNING;
    if (tk[3] < CLV_INF) r.cur = rk[3];
    else status = hit_pop(r, st);
    if (CLV_PREFETCH_NEXT && status == HIT_RUNNING) prefetch_ref<true>(A.wnodes, A.tris, r.cur);
    return status;
}
template <bool COUNT, bool WIDE = false>
CLV_HD int hit_leaf_step(const AccelView &A, HitRay &r, const StackEntry *st, Cnt<COUNT> &cnt) {
    const int code = ~r.cur;
    const int first = code >> 4, count = code & 15;
    if (CLV_PREFETCH_LEAF && !CLV_PREFETCH_NEXT) prefetch_leaf(A.tris, first, count);
    for (int i = first; i < first + count; i++) {
        const V4 *tp = A.tris + 3 * (size_t)i;
        V4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
        cnt.tri();
        float t, u, v;
        if (tri_test(r.o, r.d, xyz(a), xyz(b), xyz(c), t, u, v) &&
            (t < r.best_t || (t == r.best_t && t < CLV_INF && f2i(a.w) < r.best_rank))) {
            r.best_i = i; r.best_t = t; r.bu = u; r.bv = v;
            r.best_rank = f2i(a.w);
        }
    }
    const int status = hit_pop(r, st);
    if (CLV_PREFETCH_NEXT && status == HIT_RUNNING) prefetch_ref<WIDE>(WIDE ? A.wnodes : A.nodes, A.tris, r.cur);
    return status;
}

/* ---- leaves tested by the whole warp (k_walk_trace, k_connect_rays) ------------------------------------------------- *
 * A ray of the 10 M-triangle scene meets 7 node records but 19 triangle records, and a leaf step that lets every lane loop
 * over its own leaf runs as long as the longest leaf of the warp (up to 15 triangles) with only the lanes at leaves busy.
 * Here the (ray, triangle) pairs of all lanes at leaves are dealt to ALL 32 lanes: pair p belongs to the lane whose
 * inclusive prefix sum of leaf sizes first exceeds p (binary search by shuffles), the tester borrows that lane's ray by
 * shuffles and runs the SAME tri_test on the SAME operands, so every pair yields the bits it yields in the sequential
 * loop; a leaf's result is the lexicographic minimum (t, rank) of its accepted pairs and the running best — the
 * sequential loop's result, whatever the order (ranks are unique; t = +INF is never accepted, as there).
 * `W` supplies the warp: lane(), ballot(pred), shfl(value, source lane), all 32 lanes calling together — the hardware
 * intrinsics in the kernels (DeviceWarp, clive_b200.cu), 32 host threads in lockstep in tests/hostemu. */
CLV_HD int clv_popc(unsigned v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
CLV_HD int clv_ffs(unsigned v) { /* 1-based position of the lowest set bit, 0 for none */
#if defined(__CUDA_ARCH__)
    return __ffs((int)v);
#else
    return __builtin_ffs((int)v);
#endif
}
struct LeafDeal { int incl, excl, first_minus_excl, total; };
template <class W>
CLV_HD LeafDeal leaf_deal(const W &w, bool at_leaf, int cur) {
    const int code = ~cur;
    LeafDeal D;
    const int count = at_leaf ? (code & 15) : 0;
    /* prefix sum of the leaf sizes (< 16: four bits) from four ballots — the shuffles of a scan would run on the LSU pipe,
     * which is what bounds these kernels */
    const unsigned below = (1u << w.lane()) - 1u;
    D.excl = 0;
    D.total = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const unsigned has = w.ballot(((count >> b) & 1) != 0);
        D.excl += clv_popc(has & below) << b;
        D.total += clv_popc(has) << b;
    }
    D.incl = D.excl + count;
    D.first_minus_excl = (code >> 4) - D.excl;
    return D;
}
/* the lane that owns pair p: the first one whose inclusive sum exceeds p (lanes without a leaf own nothing) */
template <class W>
CLV_HD int leaf_owner(const W &w, const LeafDeal &D, int p) {
    int owner = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const int v = w.shfl(D.incl, owner + step - 1);
        if (v <= p) owner += step;
    }
    return owner;
}
template <class W>
CLV_HD V3 shfl3(const W &w, V3 a, int src) { return v3(w.shfl(a.x, src), w.shfl(a.y, src), w.shfl(a.z, src)); }
/* closest hit: all 32 lanes call; returns the lane's status after its leaf (HIT_RUNNING for lanes that held none) */
template <bool COUNT, class W>
CLV_HD int hit_leaf_step_warp(const W &w, const AccelView &A, HitRay &r, const StackEntry *st, bool at_leaf, Cnt<COUNT> &cnt) {
    const LeafDeal D = leaf_deal(w, at_leaf, r.cur);
    for (int base = 0; base < D.total; base += 32) {
        const int p = base + w.lane();
        const int owner = leaf_owner(w, D, p);
        const int i = w.shfl(D.first_minus_excl, owner) + p;
        const V3 o = shfl3(w, r.o, owner), d = shfl3(w, r.d, owner);
        bool hit = false;
        float t = CLV_INF, u = 0.0f, v = 0.0f;
        int rank = 0;
        if (p < D.total) {
            const V4 *tp = A.tris + 3 * (size_t)i;
            const V4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
            cnt.tri();
            rank = f2i(a.w);
            hit = tri_test(o, d, xyz(a), xyz(b), xyz(c), t, u, v) && t < CLV_INF;
        }
        /* accepted pairs are rare (a ray has ONE closest hit): hand each to its owner in turn */
        unsigned hits = w.ballot(hit);
        while (hits) {
            const int src = clv_ffs(hits) - 1;
            hits &= hits - 1u;
            const int own_s = w.shfl(owner, src);
            const float t_s = w.shfl(t, src), u_s = w.shfl(u, src), v_s = w.shfl(v, src);
            const int rank_s = w.shfl(rank, src), i_s = w.shfl(i, src);
            if (w.lane() == own_s && (t_s < r.best_t || (t_s == r.best_t && rank_s < r.best_rank))) {
                r.best_i = i_s; r.best_t = t_s; r.bu = u_s; r.bv = v_s; r.best_rank = rank_s;
            }
        }
    }
    return at_leaf ? hit_pop(r, st) : HIT_RUNNING;
}

/* ---- continuation of a closest-hit search across wavefront rounds ------------------------------------------------------ *
 * A round of the wavefront walks cannot end before its longest ray, and fewer than 1 % of the rays — the ones still running
 * 64 iterations after every other ray of the round is done — held 20 % of the trace time (profiles/r02_notes.md).  Such a
 * search is suspended: its scalars and the live part of its stack go to a continuation record and its hit record says
 * PENDING.  The records of round r are the FIRST work of round r + 1 — trace resumes each search exactly where it stopped
 * (long searches start early instead of forming the next tail), shade then treats the sample like any other of that round —
 * so the sample simply waits one round.  The search itself is unchanged — same steps in the same order — and so is its result.
 * Record: 4 x V4 header [o.xyz best_t] [d.xyz best_rank] [bu bv best_i cur] [sp slot - -], then the stack entries. */
#define CLV_HIT_PENDING (-2)
#define CLV_WALK_MAX_LAG 4              /* rounds a sample may fall behind; the walks run CLV_WALK_ROUNDS + this many rounds */
#define CLV_HIT_RECORD_V4 (4 + (CLV_STACK > CLV_OCC_STACK ? CLV_STACK : CLV_OCC_STACK) / 2) /* record stride, 16-byte units */
CLV_HD void hit_save(V4 *rec, uint32_t slot, const HitRay &r, const StackEntry *st) {
    stcg4(rec + 0, v4(r.o.x, r.o.y, r.o.z, r.best_t));
    stcg4(rec + 1, v4(r.d.x, r.d.y, r.d.z, i2f(r.best_rank)));
    stcg4(rec + 2, v4(r.bu, r.bv, i2f(r.best_i), i2f(r.cur)));
    stcg4(rec + 3, v4(i2f(r.sp), i2f((int)slot), 0.0f, 0.0f));
    StackEntry *dst = reinterpret_cast<StackEntry *>(rec + 4);
    for (int k = 0; k < r.sp; k++) dst[k] = st[k];
}
CLV_HD uint32_t hit_record_slot(const V4 *rec) { return (uint32_t)f2i(ldcg4(rec + 3).y); }
CLV_HD uint32_t hit_restore(const V4 *rec, HitRay &r, StackEntry *st) { /* returns the sample slot */
    const V4 a = ldcg4(rec + 0), b = ldcg4(rec + 1), c = ldcg4(rec + 2), d = ldcg4(rec + 3);
    r.o = xyz(a); r.best_t = a.w;
    r.d = xyz(b); r.best_rank = f2i(b.w);
    r.inv = rcp(r.d);
    r.bu = c.x; r.bv = c.y; r.best_i = f2i(c.z); r.cur = f2i(c.w);
    r.sp = f2i(d.x);
    const StackEntry *src = reinterpret_cast<const StackEntry *>(rec + 4);
    for (int k = 0; k < r.sp; k++) st[k] = src[k];
    return (uint32_t)f2i(d.y);
}

/* ---- occlusion query (predicate rays of the connection stage) ---------------------------------------------------- *
 * Both predicates of connect_paths ask whether ONE known triangle is the closest hit of a ray: visibility_test
 * (trace.metal:178-197: closest hit must be b's triangle) and world_ray_to_camera_ray (:569-617: closest hit must be a
 * sensor triangle).  Intersecting that triangle first gives its (t, rank); the answer is then "no other triangle has a
 * lexicographically smaller (t, rank)", an any-hit search bounded by t from the start.  Unlike the closest-hit search
 * the bound never shrinks, so the stack holds node refs only and a box is pruned when it is pushed.
 * The search is resumable (one round = walk down to the next leaf, test it, pop) so that a warp can keep all lanes
 * busy by handing a finished lane the next ray of its task list (k_connect_rays). */
struct OccRay { /* scalars only, so that they live in registers; the node stack is a separate local array */
    V3 o, d, inv;
    float t_bound;
    int rank_bound, skip;
    int cur, sp;
};
enum { OCC_RUNNING = 0, OCC_CLEAR = 1, OCC_BLOCKED = 2 };


template <bool COUNT>
CLV_HD int occ_begin(const AccelView &A, OccRay &r, Cnt<COUNT> &cnt) {
    cnt.ray_traced();
    bool hit;
    float te;
    slab_test(r.o, r.inv, xyz(A.root_min), xyz(A.root_max), hit, te);
    /* a ray that misses the scene box finds nothing in the reference either: its target is then not "the closest hit" */
    if (!hit) return OCC_BLOCKED;
    r.cur = CLV_WIDE_OCCLUSION ? A.wroot_ref : A.root_ref;
    r.sp = 0;
    return OCC_RUNNING;
}

template <bool COUNT>
CLV_HD int occ_round(const AccelView &A, OccRay &r, int *stack, Cnt<COUNT> &cnt) {
    int cur = r.cur, sp = r.sp;
    int status = OCC_RUNNING;
    while (status == OCC_RUNNING && cur >= 0) {
#if CLV_WIDE_OCCLUSION
        {
            float tk[4];
            int rk[4];
            wide_test(A.wnodes + 8 * (size_t)cur, r.o, r.inv, r.t_bound, tk, rk);
            cnt.node();
            /* any order finds an occluder; take the nearest surviving child, keep the rest */
            /* (selects, not tk[near]: an index computed at run time would put both arrays in local memory) */
            int near = 0, rn = rk[0];
            float tn = tk[0];
#pragma unroll
            for (int k = 1; k < 4; k++) {
                const bool closer = tk[k] < tn;
                tn = closer ? tk[k] : tn; rn = closer ? rk[k] : rn; near = closer ? k : near;
            }
            const bool any = tn < CLV_INF;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k != near && tk[k] < CLV_INF && sp < CLV_OCC_STACK) stack[sp++] = rk[k];
            if (any) cur = rn;
            else if (sp > 0) cur = stack[--sp];
            else status = OCC_CLEAR;
            continue;
        }
#endif
        const V4 *n = A.nodes + 4 * (size_t)cur;
        V4 n0 = ld4(n), n1 = ld4(n + 1), n2 = ld4(n + 2), n3 = ld4(n + 3);
        cnt.node();
        bool hl, hr;
        float tl, tr;
        slab_test(r.o, r.inv, v3(n0.x, n0.y, n0.z), v3(n0.w, n1.x, n1.y), hl, tl);
        slab_test(r.o, r.inv, v3(n1.z, n1.w, n2.x), v3(n2.y, n2.z, n2.w), hr, tr);
        int lref = f2i(n3.x), rref = f2i(n3.y);
        bool go_l = hl && tl <= r.t_bound, go_r = hr && tr <= r.t_bound;
        bool left_first = tl < tr;
        if (go_l && go_r && sp < CLV_OCC_STACK) stack[sp++] = left_first ? rref : lref;
        if (go_l || go_r) cur = (go_l && (left_first || !go_r)) ? lref : rref;
        else if (sp > 0) cur = stack[--sp];
        else status = OCC_CLEAR;
    }
    if (status == OCC_RUNNING) {
        int code = ~cur;
        int first = code >> 4, count = code & 15;
        for (int i = first; i < first + count; i++) {
            const V4 *tp = A.tris + 3 * (size_t)i;
            V4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
            cnt.tri();
            float t, u, v;
            if (i != r.skip && tri_test(r.o, r.d, xyz(a), xyz(b), xyz(c), t, u, v) &&
                (t < r.t_bound || (t == r.t_bound && f2i(a.w) < r.rank_bound)))
                status = OCC_BLOCKED;
        }
        if (status == OCC_RUNNING) {
            if (sp > 0) cur = stack[--sp];
            else status = OCC_CLEAR;
        }
    }
    r.cur = cur;
    r.sp = sp;
    return status;
}

/* the two halves of a round as separate steps, for the majority-vote schedule of k_connect_rays: one inner node ... */
template <bool COUNT>
CLV_HD int occ_inner_step(const AccelView &A, OccRay &r, int *stack, Cnt<COUNT> &cnt) {
#if CLV_WIDE_OCCLUSION
    {
        float tk[4];
        int rk[4];
        wide_test(A.wnodes + 8 * (size_t)r.cur, r.o, r.inv, r.t_bound, tk, rk);
        cnt.node();
        /* the nearest surviving child by selects: tk[near] with a run-time index would put both arrays in local memory
         * (it did: 2 x STL.128 + 5 LDL per step, 45 % of the L1 wavefronts of k_connect_rays) */
        int near = 0, rn = rk[0];
        float tn = tk[0];
#pragma unroll
        for (int k = 1; k < 4; k++) {
            const bool closer = tk[k] < tn;
            tn = closer ? tk[k] : tn; rn = closer ? rk[k] : rn; near = closer ? k : near;
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k != near && tk[k] < CLV_INF && r.sp < CLV_OCC_STACK) stack[r.sp++] = rk[k];
        if (tn < CLV_INF) r.cur = rn;
        else if (r.sp > 0) r.cur = stack[--r.sp];
        else return OCC_CLEAR;
        if (CLV_PREFETCH_NEXT) prefetch_ref<true>(A.wnodes, A.tris, r.cur);
        return OCC_RUNNING;
    }
#endif
    const V4 *n = A.nodes + 4 * (size_t)r.cur;
    V4 n0 = ld4(n), n1 = ld4(n + 1), n2 = ld4(n + 2), n3 = ld4(n + 3);
    cnt.node();
    bool hl, hr;
    float tl, tr;
    slab_test(r.o, r.inv, v3(n0.x, n0.y, n0.z), v3(n0.w, n1.x, n1.y), hl, tl);
    slab_test(r.o, r.inv, v3(n1.z, n1.w, n2.x), v3(n2.y, n2.z, n2.w), hr, tr);
    int lref = f2i(n3.x), rref = f2i(n3.y);
    bool go_l = hl && tl <= r.t_bound, go_r = hr && tr <= r.t_bound;
    bool left_first = tl < tr;
    if (go_l && go_r && r.sp < CLV_OCC_STACK) stack[r.sp++] = left_first ? rref : lref;
    if (go_l || go_r) r.cur = (go_l && (left_first || !go_r)) ? lref : rref;
    else if (r.sp > 0) r.cur = stack[--r.sp];
    else return OCC_CLEAR;
    if (CLV_PREFETCH_NEXT) prefetch_ref<false>(A.nodes, A.tris, r.cur);
    return OCC_RUNNING;
}
/* ... or one whole leaf */
template <bool COUNT>
CLV_HD int occ_leaf_step(const AccelView &A, OccRay &r, const int *stack, Cnt<COUNT> &cnt) {
    int code = ~r.cur;
    int first = code >> 4, count = code & 15;
    if (CLV_PREFETCH_LEAF && !CLV_PREFETCH_NEXT) prefetch_leaf(A.tris, first, count);
    bool blocked = false;
    for (int i = first; i < first + count; i++) {
        const V4 *tp = A.tris + 3 * (size_t)i;
        V4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
        cnt.tri();
        float t, u, v;
        if (i != r.skip && tri_test(r.o, r.d, xyz(a), xyz(b), xyz(c), t, u, v) &&
            (t < r.t_bound || (t == r.t_bound && f2i(a.w) < r.rank_bound)))
            blocked = true;
    }
    if (blocked) return OCC_BLOCKED;
    if (r.sp > 0) {
        r.cur = stack[--r.sp];
        if (CLV_PREFETCH_NEXT) prefetch_ref<CLV_WIDE_OCCLUSION != 0>(CLV_WIDE_OCCLUSION ? A.wnodes : A.nodes, A.tris, r.cur);
        return OCC_RUNNING;
    }
    return OCC_CLEAR;
}

/* ... the same for the occlusion queries (k_connect_rays): a leaf blocks when any of its pairs does */
template <bool COUNT, class W>
CLV_HD int occ_leaf_step_warp(const W &w, const AccelView &A, OccRay &r, const int *stack, bool at_leaf, Cnt<COUNT> &cnt) {
    const LeafDeal D = leaf_deal(w, at_leaf, r.cur);
    bool blocked = false;
    for (int base = 0; base < D.total; base += 32) {
        const int p = base + w.lane();
        const int owner = leaf_owner(w, D, p);
        const int i = w.shfl(D.first_minus_excl, owner) + p;
        const V3 o = shfl3(w, r.o, owner), d = shfl3(w, r.d, owner);
        const float t_bound = w.shfl(r.t_bound, owner);
        const int rank_bound = w.shfl(r.rank_bound, owner), skip = w.shfl(r.skip, owner);
        bool blocks = false;
        if (p < D.total) {
            const V4 *tp = A.tris + 3 * (size_t)i;
            const V4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
            cnt.tri();
            float t, u, v;
            blocks = i != skip && tri_test(o, d, xyz(a), xyz(b), xyz(c), t, u, v) &&
                     (t < t_bound || (t == t_bound && f2i(a.w) < rank_bound));
        }
        /* my pairs of this round sit in lanes [excl - base, incl - base) */
        const unsigned blocking = w.ballot(blocks);
        const int lo = D.excl - base > 0 ? D.excl - base : 0, hi = D.incl - base < 32 ? D.incl - base : 32;
        if (hi > lo && (blocking & ((hi >= 32 ? 0xffffffffu : (1u << hi) - 1u) & ~((1u << lo) - 1u)))) blocked = true;
    }
    if (!at_leaf) return OCC_RUNNING;
    if (blocked) return OCC_BLOCKED;
    if (r.sp > 0) { r.cur = stack[--r.sp]; return OCC_RUNNING; }
    return OCC_CLEAR;
}

/* direct test of one triangle of the reference numbering; also returns its reference rank */
CLV_HD bool test_one(const AccelView &A, int tri, V3 o, V3 d, float &t, int &rank) {
    const V4 *tp = A.tris + 3 * (size_t)tri;
    V4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
    float u, v;
    rank = f2i(a.w);
    /* t = +inf (parallel ray, NaN barycentrics) passes the reference's triangle test but can never be its closest hit */
    return tri_test(o, d, xyz(a), xyz(b), xyz(c), t, u, v) && t < CLV_INF;
}

/* visibility_test (trace.metal:178-197): the closest hit from a toward b must be b's triangle. */
template <int MODE, bool COUNT>
CLV_HD bool visibility_test(const AccelView &A, V3 a_origin, int a_triangle, V3 b_origin, int b_triangle,
                            Cnt<COUNT> &cnt) {
    V3 direction = normalize(b_origin - a_origin);
    V3 inv = rcp(direction);
    cnt.ray_ref();
    if (MODE == TRAVERSE_ORDERED) {
        if (b_triangle < 0 || b_triangle == a_triangle) return false;
        OccRay r;
        r.o = a_origin; r.d = direction; r.inv = inv; r.skip = b_triangle;
        if (!test_one(A, b_triangle, a_origin, direction, r.t_bound, r.rank_bound)) return false;
        int stack[CLV_OCC_STACK];
        int st = occ_begin<COUNT>(A, r, cnt);
        while (st == OCC_RUNNING) st = occ_round<COUNT>(A, r, stack, cnt);
        return st == OCC_CLEAR;
    }
    int best_i = -1;
    float best_t = CLV_INF, u = 0.0f, v = 0.0f;
    closest_hit<MODE, COUNT>(A, a_origin, direction, inv, best_i, best_t, u, v, cnt, -CLV_INF, -1);
    if (best_i == -1) return false;
    if (best_i == a_triangle) return false;
    return best_i == b_triangle;
}

/* ---- sampling and BSDFs (trace.metal:200-379) ------------------------------------------------------------------- */
CLV_HD void orthonormal(V3 n, V3 &x, V3 &y) {
    V3 v;
    if (fabsf(n.x) <= fabsf(n.y) && fabsf(n.x) <= fabsf(n.z)) v = v3(1, 0, 0);
    else if (fabsf(n.y) <= fabsf(n.z)) v = v3(0, 1, 0);
    else v = v3(0, 0, 1);
    x = normalize(v - n * dot(v, n));
    y = normalize(cross(n, x));
}
CLV_HD V3 random_hemisphere_cosine(V3 xa, V3 ya, V3 za, float rx, float ry) {
    float theta = clv_acosf(sqrtf(rx));
    float phi = 2 * CLV_PI * ry;
    float st = clv_sinf(theta), ct = clv_cosf(theta), sp = clv_sinf(phi), cp = clv_cosf(phi);
    return normalize((xa * (st * cp) + ya * (st * sp)) + za * ct);
}
CLV_HD V3 random_hemisphere_uniform(V3 xa, V3 ya, V3 za, float rx, float ry) {
    float z = rx;
    float rr = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    float phi = 2 * CLV_PI * ry;
    float sp = clv_sinf(phi), cp = clv_cosf(phi);
    return normalize((xa * (rr * cp) + ya * (rr * sp)) + za * z);
}
CLV_HD V3 GGX_sample(V3 n, float rx, float ry, float alpha) {
    V3 x, y;
    orthonormal(n, x, y);
    float theta = 2 * CLV_PI * rx;
    float phi = clv_atanf(alpha * sqrtf(ry) / sqrtf(1.0f - ry));
    float sp = clv_sinf(phi), cp = clv_cosf(phi), st = clv_sinf(theta), ct = clv_cosf(theta);
    return normalize((x * (sp * ct) + y * (sp * st)) + n * cp);
}
CLV_HD V3 specular_reflection(V3 i, V3 m) { return normalize(m * (2 * dot(i, m)) - i); }
CLV_HD V3 GGX_transmit(V3 i, V3 m, float ni, float no) { /* eta, not eta^2, as the reference has it */
    float cos_i = dot(i, m);
    float eta = ni / no;
    float cos_t = sqrtf(1 + eta * (cos_i * cos_i - 1));
    return normalize(m * (eta * cos_i - cos_t) - i * eta);
}
CLV_HD V3 transmit_half(V3 i, V3 o, float ni, float no) { return normalize(o * no + i * ni); }
CLV_HD float degreve_fresnel(V3 i, V3 m, float ni, float nt) {
    float cos_i = fabsf(dot(i, m));
    float eta = ni / nt;
    float sin_t2 = eta * eta * (1.0f - cos_i * cos_i);
    if (sin_t2 >= 1.0f) return 1.0f;
    float cos_t = sqrtf(1.0f - sin_t2);
    float r_par = (nt * cos_i - ni * cos_t) / (nt * cos_i + ni * cos_t);
    float r_perp = (ni * cos_i - nt * cos_t) / (ni * cos_i + nt * cos_t);
    return 0.5f * (r_par * r_par + r_perp * r_perp);
}
CLV_HD float GGX_G1(V3 v, V3 m, float alpha) {
    float mv = dot(m, v);
    float sin2 = 1.0f - mv * mv;
    float tan2 = sin2 / (mv * mv);
    return 2.0f / (1.0f + sqrtf(1.0f + alpha * alpha * tan2));
}
CLV_HD float GGX_G(V3 i, V3 o, V3 m, V3 n, float alpha) {
    if (dot(i, m) * dot(i, n) <= 0.0f) return 0.0f;
    if (dot(o, m) * dot(o, n) <= 0.0f) return 0.0f;
    return GGX_G1(i, m, alpha) * GGX_G1(o, m, alpha);
}
CLV_HD float GGX_D(V3 m, V3 n, float alpha) {
    if (alpha == 0.0f) return 1.0f;
    float alpha2 = alpha * alpha;
    float c = dot(m, n);
    float c2 = c * c;
    float denom = c2 * (alpha2 - 1.0f) + 1.0f;
    return alpha2 / (CLV_PI * denom * denom);
}
CLV_HD float reflect_jacobian(V3 m, V3 o) { return 1.0f / (4.0f * fabsf(dot(m, o))); }
CLV_HD float transmit_jacobian(V3 i, V3 o, float ni, float no) {
    V3 h = transmit_half(i, o, ni, no);
    float ci = dot(i, h), co = dot(o, h);
    float numerator = no * no * fabsf(co);
    float s = ni * ci + no * co;
    return numerator / (s * s);
}
CLV_HD float GGX_BRDF_reflect(V3 i, V3 o, V3 m, V3 n, float ni, float no, float alpha) {
    float D = GGX_D(m, n, alpha), G = GGX_G(i, o, m, n, alpha), F = degreve_fresnel(i, m, ni, no);
    return (D * G * F) / (4.0f * fabsf(dot(i, m)));
}
CLV_HD float GGX_BRDF_transmit(V3 i, V3 o, V3 m, V3 n, float ni, float no, float alpha) {
    V3 h = transmit_half(i, o, ni, no);
    float D = GGX_D(m, n, alpha), G = GGX_G(i, o, m, n, alpha), F = degreve_fresnel(i, m, ni, no);
    float im = dot(i, h), om = dot(o, h), in = dot(i, n), on = dot(o, n);
    float coeff = (im * om) / (in * on);
    float num = no * no * D * G * (1.0f - F);
    float s = ni * im + no * om;
    return coeff * num / (s * s);
}
CLV_HD void diffuse_bounce(V3 wi, V3 n, bool from_camera, float rx, float ry, V3 &wo, float &f, float &c_p, float &l_p) {
    V3 x, y;
    orthonormal(n, x, y);
    wo = random_hemisphere_cosine(x, y, n, rx, ry);
    f = fabsf(dot(n, wo)) / CLV_PI;
    float po = fabsf(dot(n, wo)) / CLV_PI, pi_ = fabsf(dot(n, wi)) / CLV_PI;
    if (from_camera) { c_p = po; l_p = pi_; } else { c_p = pi_; l_p = po; }
}
CLV_HD void reflect_bounce(V3 wi, V3 n, V3 m, float ni, float no, float alpha, bool from_camera, V3 &wo, float &f,
                           float &c_p, float &l_p) {
    wo = specular_reflection(wi, m);
    f = GGX_BRDF_reflect(wi, wo, m, n, ni, no, alpha);
    float pf = degreve_fresnel(wi, m, ni, no);
    float pm = fabsf(dot(m, n)) * GGX_D(m, n, alpha);
    float po = pf * pm * reflect_jacobian(m, wo), pi_ = pf * pm * reflect_jacobian(m, wi);
    if (from_camera) { c_p = po; l_p = pi_; } else { c_p = pi_; l_p = po; }
}
CLV_HD void transmit_bounce(V3 wi, V3 n, V3 m, float ni, float no, float alpha, bool from_camera, V3 &wo, float &f,
                            float &c_p, float &l_p) {
    wo = GGX_transmit(wi, m, ni, no);
    f = GGX_BRDF_transmit(wi, wo, m, n, ni, no, alpha);
    float pf = 1.0f - degreve_fresnel(wi, m, ni, no);
    float pm = fabsf(dot(m, n)) * GGX_D(m, n, alpha);
    float fwd = pf * pm * transmit_jacobian(wi, wo, ni, no), rev = pf * pm * transmit_jacobian(wo, wi, no, ni);
    if (from_camera) { c_p = fwd; l_p = rev; } else { c_p = rev; l_p = fwd; }
}

/* ---- ray generation --------------------------------------------------------------------------------------------- */
CLV_HD Ray zero_ray() {
    Ray r;
    r.origin = r.direction = r.inv_direction = r.color = r.normal = v4(0, 0, 0, 0);
    r.material = r.triangle = 0;
    r.c_importance = r.l_importance = r.tot_importance = 0.0f;
    r.hit_light = r.from_camera = r.hit_camera = r.pixel_idx = 0;
    r.pad[0] = r.pad[1] = r.pad[2] = 0;
    return r;
}

/* generate_camera_rays, trace.metal:1020-1067 */
/* ... from the two jitter draws (the wavefront walks draw them up front, in the reference's order, and build the ray
 * when the sample's light walk is over) */
CLV_HD Ray camera_ray_from_jitter(const Camera &c, int pixel_idx, float x_offset, float y_offset) {
    Ray ray = zero_ray();
    int pixel_x = pixel_idx % c.pixel_width;
    int pixel_y = pixel_idx / c.pixel_width;
    float x_n = ((float)pixel_x + x_offset - 0.5f * (float)c.pixel_width) / (float)c.pixel_width;
    float y_n = ((float)pixel_y + y_offset - 0.5f * (float)c.pixel_height) / (float)c.pixel_height;
    V3 x_vec = (xyz(c.dx) * x_n) * c.phys_width;
    V3 y_vec = (xyz(c.dy) * y_n) * c.phys_height;
    V3 origin = (xyz(c.center) + x_vec) + y_vec;
    V3 direction = normalize(xyz(c.focal_point) - origin);
    ray.origin = pad0(origin);
    ray.direction = pad0(direction);
    ray.normal = pad0(xyz(c.direction));
    ray.inv_direction = pad0(rcp(direction));
    ray.color = v4(1.0f, 1.0f, 1.0f, 0.0f);
    ray.material = 7;
    ray.triangle = -1;
    ray.hit_light = -1;
    ray.hit_camera = -1;
    ray.from_camera = 1;
    ray.c_importance = 1.0f / (c.phys_width * c.phys_height);
    ray.l_importance = 1.0f;
    ray.tot_importance = ray.c_importance;
    ray.pixel_idx = pixel_idx;
    return ray;
}
CLV_HD Ray make_camera_ray(const Camera &c, int pixel_idx, uint32_t &seed0, uint32_t &seed1) {
    float x_offset = xorshift_random(seed0);
    float y_offset = xorshift_random(seed1);
    return camera_ray_from_jitter(c, pixel_idx, x_offset, y_offset);
}

/* generate_light_rays, trace.metal:1070-1124 (light_index clamped to count-1: the reference reads out of bounds
 * when the RNG returns exactly 1.0) */
CLV_HD Ray make_light_ray(const Triangle *light_triangles, const float *surface_areas,
                          const int32_t *light_triangle_indices, const Material *materials, int light_count,
                          uint32_t &seed0, uint32_t &seed1) {
    Ray ray = zero_ray();
    ray.hit_light = -1;
    ray.hit_camera = -1;
    int light_index = (int)(xorshift_random(seed0) * (float)light_count);
    if (light_index > light_count - 1) light_index = light_count - 1;
    const Triangle *lt = light_triangles + light_index;
    float surface_area = surface_areas[light_index];
    float u = xorshift_random(seed0);
    float v = xorshift_random(seed1);
    if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
    float w = 1.0f - u - v;
    V3 n = xyz(ld4(&lt->normal));
    V3 origin = ((xyz(ld4(&lt->v0)) * u + xyz(ld4(&lt->v1)) * v) + xyz(ld4(&lt->v2)) * w) + n * CLV_DELTA;
    V3 x, y;
    orthonormal(n, x, y);
    float rand_x = xorshift_random(seed0);
    float rand_y = xorshift_random(seed1);
    V3 direction = random_hemisphere_uniform(x, y, n, rand_x, rand_y);
    ray.normal = pad0(n);
    ray.origin = pad0(origin);
    ray.direction = pad0(direction);
    ray.inv_direction = pad0(rcp(direction));
    ray.material = lt->material;
    ray.color = ld4(&materials[ray.material].emission); /* whole 16-byte slot, like the oracle's struct copy */
    ray.triangle = light_triangle_indices[light_index];
    ray.c_importance = 1.0f;
    ray.l_importance = 1.0f / ((float)light_count * surface_area);
    ray.tot_importance = ray.l_importance;
    return ray;
}

/* ---- where stored path vertices live ------------------------------------------------------------------------------ *
 * Two layouts behind one interface.  AoS = the reference's Path records (1040 B per subpath, 128 B per vertex): what the
 * stage-level ABI reads and writes.  SoA = the fused renderer's own compact layout: 64 B per vertex in four 16-byte
 * parts, part-major over the samples ([(subpath*6 + vertex)*4 + part][n_samples]), so that a warp reads or writes 512
 * contiguous bytes per part; only what the connection stage consumes is kept (the direction survives as |dir.n|):
 *   part 0: origin.xyz, tot_importance      part 1: normal.xyz, |direction . normal|
 *   part 2: color.xyz, c_importance         part 3: l_importance, material, triangle, hit_light (int bits)
 * Values are the same bits in both layouts; k_connect_finish missed L1 on 90 % of its AoS reads and moved twice the path
 * data through DRAM (profiles/r01_notes.md). */
enum { SUB_LIGHT = 0, SUB_CAMERA = 1 };

struct AosPaths {
    const Path *paths[2]; /* this sample's two records */
    CLV_HDM int length(int sub) const { return paths[sub]->length; }
    CLV_HDM V3 origin(int sub, int k) const { return xyz(paths[sub]->rays[k].origin); }
    CLV_HDM V3 normal(int sub, int k) const { return xyz(paths[sub]->rays[k].normal); }
    CLV_HDM V3 color(int sub, int k) const { return xyz(paths[sub]->rays[k].color); }
    CLV_HDM float abs_cos(int sub, int k) const {
        return fabsf(dot(xyz(paths[sub]->rays[k].direction), xyz(paths[sub]->rays[k].normal)));
    }
    CLV_HDM float limp(int sub, int k) const { return paths[sub]->rays[k].l_importance; }
    CLV_HDM float cimp(int sub, int k) const { return paths[sub]->rays[k].c_importance; }
    CLV_HDM float tot(int sub, int k) const { return paths[sub]->rays[k].tot_importance; }
    CLV_HDM int material(int sub, int k) const { return paths[sub]->rays[k].material; }
    CLV_HDM int triangle(int sub, int k) const { return paths[sub]->rays[k].triangle; }
    CLV_HDM bool hit_light(int sub, int k) const { return paths[sub]->rays[k].hit_light >= 0; }
    CLV_HDM int pixel_idx0() const { return paths[SUB_CAMERA]->rays[0].pixel_idx; }
};
CLV_HD AosPaths aos_paths(const Path *camera_paths, const Path *light_paths, uint32_t slot) {
    AosPaths P;
    P.paths[SUB_LIGHT] = light_paths + slot;
    P.paths[SUB_CAMERA] = camera_paths + slot;
    return P;
}

struct SoaPaths {
    const V4 *v;        /* [(sub*6 + k)*4 + part][n] */
    const int32_t *len; /* [2][n] */
    uint32_t n, slot;
    int32_t pixel0;     /* the pixel this sample belongs to (its global id) */
    CLV_HDM V4 part(int sub, int k, int part_) const { return ld4(v + ((size_t)((sub * 6 + k) * 4 + part_) * n + slot)); }
    CLV_HDM int length(int sub) const { return len[(size_t)sub * n + slot]; }
    CLV_HDM V3 origin(int sub, int k) const { return xyz(part(sub, k, 0)); }
    CLV_HDM V3 normal(int sub, int k) const { return xyz(part(sub, k, 1)); }
    CLV_HDM V3 color(int sub, int k) const { return xyz(part(sub, k, 2)); }
    CLV_HDM float abs_cos(int sub, int k) const { return part(sub, k, 1).w; }
    CLV_HDM float limp(int sub, int k) const { return part(sub, k, 3).x; }
    CLV_HDM float cimp(int sub, int k) const { return part(sub, k, 2).w; }
    CLV_HDM float tot(int sub, int k) const { return part(sub, k, 0).w; }
    CLV_HDM int material(int sub, int k) const { return f2i(part(sub, k, 3).y); }
    CLV_HDM int triangle(int sub, int k) const { return f2i(part(sub, k, 3).z); }
    CLV_HDM bool hit_light(int sub, int k) const { return f2i(part(sub, k, 3).w) >= 0; }
    /* an empty camera path leaves a zeroed slot 0 behind, whose pixel index is 0 (see random_walk) */
    CLV_HDM int pixel_idx0() const { return length(SUB_CAMERA) > 0 ? pixel0 : 0; }
};

/* where a random walk puts its vertices */
struct AosSink {
    Path *path;
    CLV_HDM void store(int i, const Ray &ray) const { path->rays[i] = ray; }
    CLV_HDM void patch(int i, bool from_camera, float value) const {
        if (from_camera) path->rays[i].l_importance = value;
        else path->rays[i].c_importance = value;
    }
    CLV_HDM void finish(int length, bool from_camera) const;
};
struct SoaSink {
    V4 *v;
    int32_t *len;
    uint32_t n, slot;
    int sub;
    CLV_HDM V4 *part(int k, int part_) const { return v + ((size_t)((sub * 6 + k) * 4 + part_) * n + slot); }
    CLV_HDM void store(int i, const Ray &ray) const {
        st4_stream(part(i, 0), v4(ray.origin.x, ray.origin.y, ray.origin.z, ray.tot_importance));
        st4_stream(part(i, 1), v4(ray.normal.x, ray.normal.y, ray.normal.z, fabsf(dot(xyz(ray.direction), xyz(ray.normal)))));
        st4_stream(part(i, 2), v4(ray.color.x, ray.color.y, ray.color.z, ray.c_importance));
        st4_stream(part(i, 3), v4(ray.l_importance, i2f(ray.material), i2f(ray.triangle), i2f(ray.hit_light)));
    }
    CLV_HDM void patch(int i, bool from_camera, float value) const {
        if (from_camera) part(i, 3)->x = value;
        else part(i, 2)->w = value;
    }
    CLV_HDM void finish(int length, bool from_camera) const;
};

/* ---- random walk: generate_paths, trace.metal:381-532 ----------------------------------------------------------- *
 * Vertices are written to their slot as they are produced (the reference builds the Path in thread memory and copies
 * 1040 bytes at the end).  Slots >= length keep whatever they held; nothing downstream reads them.  A path of length 0
 * gets a zeroed rays[0], which is what connect_paths' splat-weight code then reads (the reference reads uninitialised
 * memory there; the oracle reads zeros).
 *
 * The walk is written as begin / bounce / finish around the closest-hit search so that two drivers share every
 * operation: `random_walk` (one thread runs the whole walk: the stage-level ABI and the megakernel k_fused_walks) and the
 * wavefront kernels (k_walk_trace traces the ray of every live walk, k_walk_shade runs `walk_bounce` on the hits).
 */
struct WalkRegs { /* what a walk carries from one bounce to the next */
    V3 o, d;        /* the ray about to be traced (inv_direction = rcp(d), the same bits the reference stores) */
    V3 rc;          /* its colour = throughput up to this vertex */
    float ray_tot;  /* tot_importance of the vertex the ray leaves */
    float carry_imp;
    int ray_hit_light;
    V3 last_stored_color;
    V4 uni;         /* unidirectional estimate, trace.metal:523-528 */
    bool uni_found;
    int length;
};

/* vertex 0 = the generated ray: stored (speculatively, like every vertex: slot i only counts once length > i) */
template <typename SINK>
CLV_HD void walk_begin(const Ray &ray, const SINK &sink, WalkRegs &W) {
    const bool from_camera = ray.from_camera != 0;
    sink.store(0, ray);
    W.o = xyz(ray.origin); W.d = xyz(ray.direction); W.rc = xyz(ray.color);
    W.ray_tot = ray.tot_importance;
    W.carry_imp = from_camera ? ray.c_importance : 1.0f / (2.0f * CLV_PI);
    W.ray_hit_light = ray.hit_light;
    W.last_stored_color = v3(0, 0, 0);
    W.uni = v4(0, 0, 0, 0);
    W.uni_found = false;
    W.length = 0;
}

/* Bounce i of a walk given the closest hit of its ray (trace.metal:414-516).  Returns false when the walk ends here
 * (miss, side test, microfacet rejection, f == 0); otherwise vertex i is confirmed (its reverse pdf patched in), vertex
 * i + 1 is stored when the loop will trace another ray (i + 1 < 6), and W holds that ray. */
template <bool COUNT, typename SINK>
CLV_HD bool walk_bounce(WalkRegs &W, int i, bool from_camera, int best_i, float best_t, float u, float v, const AccelView &A,
                        const Material *materials, uint32_t &seed0, uint32_t &seed1, const SINK &sink, Cnt<COUNT> &cnt) {
    if (best_i == -1) return false;
    const V3 o = W.o, d = W.d, rc = W.rc;
    const float ray_tot = W.ray_tot;

    const V4 *tri = A.shade + 4 * (size_t)best_i;
    const V4 s0 = ld4(tri), s1 = ld4(tri + 1), s2 = ld4(tri + 2), s3 = ld4(tri + 3);
    V3 tn = xyz(s3);
    V3 n0 = xyz(s0), n1 = xyz(s1), n2 = xyz(s2);
    int tri_material = f2i(s0.w), tri_is_light = f2i(s1.w), tri_is_camera = f2i(s2.w);
    const Material *mat = materials + tri_material;
    V3 mat_color = xyz(ld4(&mat->color));
    int mat_type = mat->type;
    float alpha = mat->alpha, ior = mat->ior;

    V3 n;
    float ni, no;
    V3 sampled_normal = normalize((n0 * (1 - u - v) + n1 * u) + n2 * v);
    float side = dot(-d, tn);
    if (side > 0) { n = sampled_normal; ni = 1.0f; no = ior; }
    else if (side < 0) { n = -sampled_normal; ni = ior; no = 1.0f; }
    else return false;

    Ray nr = zero_ray();
    if (from_camera) nr.c_importance = W.carry_imp; else nr.l_importance = W.carry_imp;
    nr.origin = pad0(o + d * best_t);
    nr.material = tri_material;
    nr.triangle = best_i;
    nr.hit_light = (tri_is_light && dot(d, tn) < 0.0f) ? best_i : -1;
    nr.hit_camera = tri_is_camera ? best_i : -1;

    V3 wi = -d;
    float rand_x_a = xorshift_random(seed0);
    float rand_y_a = xorshift_random(seed1);
    float rand_x_b = xorshift_random(seed0);
    float rand_y_b = xorshift_random(seed1);

    V3 wo = v3(0, 0, 0);
    float f = 1.0f, c_p = 1.0f, l_p = 1.0f;
    V3 m = GGX_sample(n, rand_x_a, rand_y_a, alpha);
    if (dot(wi, m) < 0.0f) return false;
    if (dot(m, n) < 0.0f) return false;
    nr.normal = pad0(n);

    float fresnel = degreve_fresnel(wi, m, ni, no);
    /* material dispatch of :474-487 resolved to one of three bounce kinds first, so that each bounce function is
     * instantiated once (the walk kernel is larger than the instruction cache) */
    int kind; /* 0 diffuse, 1 reflect, 2 transmit */
    if (mat_type == 0) kind = 0;
    else if (mat_type == 1) kind = (rand_x_b <= fresnel) ? 1 : 2;
    else if (mat_type == 2) kind = (rand_x_b <= fresnel) ? 1 : 0;
    else kind = 1;
    if (kind == 0) diffuse_bounce(wi, n, from_camera, rand_x_b, rand_y_b, wo, f, c_p, l_p);
    else if (kind == 1) reflect_bounce(wi, n, m, ni, no, alpha, from_camera, wo, f, c_p, l_p);
    else transmit_bounce(wi, n, m, ni, no, alpha, from_camera, wo, f, c_p, l_p);

    float wi_n = dot(wi, tn), wo_n = dot(wo, tn);
    if ((wi_n > 0.0f && wo_n > 0.0f) || (wi_n < 0.0f && wo_n > 0.0f)) nr.color = pad0((rc * f) * mat_color);
    else nr.color = pad0(rc * f);
    nr.direction = pad0(wo);
    nr.inv_direction = pad0(rcp(wo));

    float next_carry;
    if (from_camera) {
        next_carry = c_p;
        nr.tot_importance = ray_tot * nr.c_importance;
    } else {
        next_carry = l_p;
        nr.tot_importance = ray_tot * nr.l_importance;
    }
    if (f == 0.0f) return false;

    /* path.rays[i] = ray with the reverse-direction pdf the bounce produced (:499-507, :512) */
    sink.patch(i, from_camera, from_camera ? l_p : c_p);
    cnt.vertex(1);
    W.length = i + 1;
    if (!W.uni_found && W.ray_hit_light >= 0) { /* trace.metal:523-528 (i >= 1 here: generated rays have -1) */
        V3 c = W.last_stored_color / ray_tot;
        W.uni = v4(c.x, c.y, c.z, 1.0f);
        W.uni_found = true;
    }
    W.last_stored_color = rc;

    /* the next vertex goes to its slot BEFORE its bounce is known to succeed (slot i + 1 only counts once length > i + 1;
     * the one field that bounce still changes is patched then).  That ends the live range of the 32-word record here
     * instead of across the traversal, which is what bounds the walk kernels' registers and therefore their occupancy. */
    if (i + 1 < CLV_MAX_BOUNCES) sink.store(i + 1, nr);
    W.o = xyz(nr.origin); W.d = wo; W.rc = xyz(nr.color);
    W.ray_tot = nr.tot_importance;
    W.ray_hit_light = nr.hit_light;
    W.carry_imp = next_carry;
    return true;
}

template <int MODE, bool COUNT, typename SINK>
CLV_HD void random_walk(Ray ray, const AccelView &A, const Material *materials, uint32_t &seed0, uint32_t &seed1,
                        const SINK &sink, V4 *out_uni, Cnt<COUNT> &cnt) {
    const bool from_camera = ray.from_camera != 0;
    WalkRegs W;
    walk_begin(ray, sink, W);
    /* the first ray is traced with the inv_direction its record carries (a caller of the stage-level ABI may pass any),
     * later ones with rcp(direction), which is what the bounce stores into the record (trace.metal:497) */
    const V3 inv0 = xyz(ray.inv_direction);
    for (int i = 0; i < CLV_MAX_BOUNCES; i++) {
        int best_i = -1;
        float best_t = CLV_INF, u = 0.0f, v = 0.0f;
        cnt.ray_ref();
        closest_hit<MODE, COUNT>(A, W.o, W.d, i == 0 ? inv0 : rcp(W.d), best_i, best_t, u, v, cnt, -CLV_INF, -1);
        if (!walk_bounce<COUNT>(W, i, from_camera, best_i, best_t, u, v, A, materials, seed0, seed1, sink, cnt)) break;
    }
    sink.finish(W.length, from_camera);
    *out_uni = W.uni;
}
CLV_HDM void AosSink::finish(int length, bool from_camera) const {
    if (length == 0) path->rays[0] = zero_ray();
    path->length = length;
    path->from_camera = from_camera ? 1 : 0;
    path->pad[0] = 0;
    path->pad[1] = 0;
}
CLV_HDM void SoaSink::finish(int length, bool) const {
    if (length == 0) store(0, zero_ray());
    len[(size_t)sub * n + slot] = length;
}

/* ---- connection helpers (trace.metal:539-617) ---------------------------------------------------------------------- */
CLV_HD float abs_cos(const Ray *r) { return fabsf(dot(xyz(r->direction), xyz(r->normal))); }
CLV_HD float geometry_from(float cos_a, float cos_b, V3 oa, V3 ob) { /* cosine_geometry_term, :539-544 */
    float dist = length(ob - oa);
    return cos_a * cos_b / (dist * dist);
}
CLV_HD V3 pixel_center(const Camera &c, int x, int y) { /* :551-562 — the pixel corner, as the reference has it */
    float x_n = ((float)x - 0.5f * (float)c.pixel_width) / (float)c.pixel_width;
    float y_n = ((float)y - 0.5f * (float)c.pixel_height) / (float)c.pixel_height;
    V3 x_vec = xyz(c.dx) * (x_n * c.phys_width);
    V3 y_vec = xyz(c.dy) * (y_n * c.phys_height);
    return (xyz(c.center) + x_vec) + y_vec;
}
CLV_HD float gaussian_weight(V3 p, V3 q, float sigma) {
    float dist = length(p - q);
    return clv_expf(-dist * dist / (2.0f * sigma * sigma));
}

/* per-thread array living in shared memory (device: element k of thread at base[k * stride]) or on the stack (host) */
struct TArr {
    float *base;
    int stride;
    CLV_HDM float &operator[](int k) const { return base[k * stride]; }
};
/* scratch layout of one sample in the connection stage (floats) */
enum { L_LIMP = 0, L_CIMP = 6, L_COS = 12, L_TOT = 18, C_LIMP = 24, C_CIMP = 30, C_COS = 36, C_TOT = 42, G_LL = 48, G_CC = 53,
       P_BACK = 58 /* p_values[0..s-1] of the strategy being evaluated */ };
#define CLV_CONNECT_SCRATCH 64
/* per-sample record of the t = 1 projections (12 words): [s-1] = t of the nearest sensor-triangle hit of light vertex
 * s-1, [6 + s-1] = index of that triangle (int bits) */
enum { T_SENSOR = 0, CAM_TRI = 6 };
#define CLV_SENSOR_WORDS 12
#define CLV_ENUM_SCRATCH 36 /* origin + normal of up to 6 camera vertices, staged by connect_enumerate */

/* light-image sink of the t = 1 strategies */
struct LightSink {
    /* staged mode (reference arrays, trace.metal:817-823) */
    int32_t *pixel_indices, *path_indices, *ray_indices;
    float *weights, *shade;
    /* splat mode: per-pass light image float4[n_pixels] (xyz = colour, w = sum of weights) */
    V4 *splat_image;
    int n_pixels;
};

/* one 16-byte reduction per light-image splat (sm_90+: atomicAdd on float4, one RED.128 instead of four RED.32); the four
 * components are added independently, exactly as four scalar atomics would */
CLV_HD void atomic_add4(V4 *p, V4 v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(reinterpret_cast<float4 *>(p), make_float4(v.x, v.y, v.z, v.w));
#else
    float *f = &p->x;
#pragma omp atomic
    f[0] += v.x;
#pragma omp atomic
    f[1] += v.y;
#pragma omp atomic
    f[2] += v.z;
#pragma omp atomic
    f[3] += v.w;
#endif
}
CLV_HD void atomic_add(float *p, float v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
#pragma omp atomic
    *p += v;
#endif
}

/* ---- connect_paths, trace.metal:620-869, in three phases per sample ---------------------------------------------------
 *   enumerate  the (s,t) strategies that need a predicate ray (t = 1 projections, regular joins) after the reference's
 *              cheap rejections (:577-586, :676-688)
 *   rays       one occlusion query per candidate strategy -> a visibility bit per strategy
 *   stage      the per-vertex MIS ingredients, once per sample (l/c importance, |dir.n|, tot importance, subpath edge
 *              geometry terms), and the s = 0 strategies, which need no ray (:662-666)
 *   finish     for the surviving strategies IN THE REFERENCE'S ORDER (t outer, s inner): MIS weight (:709-776) from the
 *              staged terms with the same operations in the same order, contribution (:778-823), then the 3x3 splat
 *              weights (:827-864)
 * The reference interleaves the three per strategy and recomputes every edge of the full path for each of its up to 41
 * strategies.  Splitting them makes the ray phase a dense list of independent rays (any lane of the warp can take any
 * ray) and the finish phase a loop over survivors only.
 * Strategy index q = (t-1)*7 + s, t = 1..6, s = 0..6; ascending q is the reference's loop order.
 */
template <typename PV>
struct ConnectCtx {
    PV P;
    int len_c, len_l;
    unsigned spec_l, spec_c; /* bit i: vertex i has a specular material */
    bool spec_sensor;        /* material of the rewritten sensor vertex (:611) is specular */
    TArr S;                  /* staged MIS terms, CLV_CONNECT_SCRATCH floats */
    TArr Sens;               /* projection record, CLV_SENSOR_WORDS words */
};

/* nearest sensor triangle hit by a ray (lexicographic (t, rank) minimum over the is_camera triangles) */
CLV_HD bool sensor_hit(const AccelView &A, V3 origin, V3 direction, float &t_best, int &rank_best, int &tri_best) {
    t_best = CLV_INF;
    rank_best = 0x7fffffff;
    tri_best = -1;
    for (int k = 0; k < A.n_camera_tris; k++) {
        float t;
        int rank, tri = A.camera_tris[k];
        if (test_one(A, tri, origin, direction, t, rank) && (t < t_best || (t == t_best && rank < rank_best))) {
            t_best = t; rank_best = rank; tri_best = tri;
        }
    }
    return tri_best >= 0;
}

/* phase 1a: stage the MIS terms of one sample and find the s = 0 strategies (camera subpath ends on an emitter) */
template <bool COUNT, typename PV>
CLV_HD void connect_stage(const PV &P, const Material *materials, TArr S, TArr Sens, ConnectCtx<PV> &X,
                          unsigned long long &valid0, Cnt<COUNT> &cnt) {
    const int len_c = P.length(SUB_CAMERA), len_l = P.length(SUB_LIGHT);
    X.P = P; X.len_c = len_c; X.len_l = len_l; X.S = S; X.Sens = Sens;
    unsigned spec_l = 0, spec_c = 0;
    cnt.vertex(len_l + len_c);
    for (int i = 0; i < len_l; i++) {
        S[L_LIMP + i] = P.limp(SUB_LIGHT, i);
        S[L_CIMP + i] = P.cimp(SUB_LIGHT, i);
        S[L_COS + i] = P.abs_cos(SUB_LIGHT, i);
        S[L_TOT + i] = P.tot(SUB_LIGHT, i);
        if (materials[P.material(SUB_LIGHT, i)].type > 0) spec_l |= 1u << i;
    }
    for (int i = 0; i + 1 < len_l; i++)
        S[G_LL + i] = geometry_from(S[L_COS + i], S[L_COS + i + 1], P.origin(SUB_LIGHT, i), P.origin(SUB_LIGHT, i + 1));
    valid0 = 0ull;
    for (int j = 0; j < len_c; j++) {
        S[C_LIMP + j] = P.limp(SUB_CAMERA, j);
        S[C_CIMP + j] = P.cimp(SUB_CAMERA, j);
        S[C_COS + j] = P.abs_cos(SUB_CAMERA, j);
        S[C_TOT + j] = P.tot(SUB_CAMERA, j);
        if (materials[P.material(SUB_CAMERA, j)].type > 0) spec_c |= 1u << j;
        if (j >= 1 && P.hit_light(SUB_CAMERA, j)) valid0 |= 1ull << (j * 7); /* t = j + 1, s = 0 (t + s >= 2) */
    }
    for (int j = 0; j + 1 < len_c; j++)
        S[G_CC + j] = geometry_from(S[C_COS + j], S[C_COS + j + 1], P.origin(SUB_CAMERA, j), P.origin(SUB_CAMERA, j + 1));
    X.spec_l = spec_l; X.spec_c = spec_c;
    X.spec_sensor = materials[7].type > 0;
}

/* phase 1b: the strategies that need a predicate ray, after the reference's cheap rejections */
template <int MODE, bool COUNT, typename PV>
CLV_HD unsigned long long connect_enumerate(const PV &P, const AccelView &A, const Material *materials, const Camera &c,
                                            TArr Sens, TArr Cam /* CLV_ENUM_SCRATCH floats */, Cnt<COUNT> &cnt) {
    const int len_c = P.length(SUB_CAMERA), len_l = P.length(SUB_LIGHT);
    unsigned long long cand = 0ull;
    unsigned spec_c = 0;
    /* camera vertices are needed once per light vertex: read each from global memory once (the per-thread working set
     * of path records is far larger than L1, so re-reads went to L2/DRAM: 1.26 GB per launch at C2 before this) */
    for (int j = 0; j < len_c; j++) {
        if (materials[P.material(SUB_CAMERA, j)].type > 0) spec_c |= 1u << j;
        V3 o = P.origin(SUB_CAMERA, j), nn = P.normal(SUB_CAMERA, j);
        Cam[6 * j + 0] = o.x; Cam[6 * j + 1] = o.y; Cam[6 * j + 2] = o.z;
        Cam[6 * j + 3] = nn.x; Cam[6 * j + 4] = nn.y; Cam[6 * j + 5] = nn.z;
    }
    for (int s = 1; s < len_l + 1; s++) {
        V3 lig_o = P.origin(SUB_LIGHT, s - 1);
        /* t = 1: world_ray_to_camera_ray up to its traversal (:577-586) */
        if (len_c >= 1 && !(materials[f2i(ld4(A.shade + 4 * (size_t)P.triangle(SUB_LIGHT, s - 1)).w)].type > 0)) {
            V3 direction = normalize(xyz(c.focal_point) - lig_o);
            if (!(dot(direction, xyz(c.direction)) > 0.0f)) {
                cnt.ray_ref();
                bool ok = true;
                if (MODE == TRAVERSE_ORDERED) {
                    /* the closest hit has to be a sensor triangle: if the ray misses all of them no traversal can
                     * make the strategy valid */
                    float ts;
                    int rank, tri;
                    ok = sensor_hit(A, lig_o, direction, ts, rank, tri);
                    if (ok) { Sens[T_SENSOR + s - 1] = ts; Sens[CAM_TRI + s - 1] = i2f(tri); }
                }
                if (ok) cand |= 1ull << s;
            }
        }
        /* t >= 2: regular joins up to visibility_test (:676-688) */
        if (materials[P.material(SUB_LIGHT, s - 1)].type > 0) continue;
        V3 lig_n = P.normal(SUB_LIGHT, s - 1);
        for (int t = 2; t < len_c + 1; t++) {
            if ((spec_c >> (t - 1)) & 1u) continue;
            const int j = t - 1;
            V3 dir_l_to_c = normalize(v3(Cam[6 * j + 0], Cam[6 * j + 1], Cam[6 * j + 2]) - lig_o);
            if (dot(lig_n, dir_l_to_c) < CLV_DELTA) continue;
            if (dot(v3(Cam[6 * j + 3], Cam[6 * j + 4], Cam[6 * j + 5]), -dir_l_to_c) < CLV_DELTA) continue;
            cnt.ray_ref();
            cand |= 1ull << ((t - 1) * 7 + s);
        }
    }
    return cand;
}

/* Start the predicate ray of strategy q of the sample whose paths are (cr, lr) and whose scratch is S (which may belong
 * to another lane of the warp).  Returns OCC_RUNNING with `ray` set up, or the final status. */
template <int MODE, bool COUNT, typename PV>
CLV_HD int connect_ray_begin(int q, const PV &P, TArr Sens, const AccelView &A, const Camera &c, OccRay &ray,
                             Cnt<COUNT> &cnt) {
    const int t = q / 7 + 1, s = q % 7;
    V3 lig_o = P.origin(SUB_LIGHT, s - 1);
    if (t == 1) {
        V3 direction = normalize(xyz(c.focal_point) - lig_o);
        if (MODE == TRAVERSE_ORDERED) {
            int tri = f2i(Sens[CAM_TRI + s - 1]);
            ray.o = lig_o; ray.d = direction; ray.inv = rcp(direction);
            ray.t_bound = Sens[T_SENSOR + s - 1];
            ray.rank_bound = f2i(ld4(A.tris + 3 * (size_t)tri).w);
            ray.skip = tri;
            return occ_begin<COUNT>(A, ray, cnt);
        }
        int best_i = -1;
        float best_t = CLV_INF, u = 0.0f, v = 0.0f;
        closest_hit<MODE, COUNT>(A, lig_o, direction, rcp(direction), best_i, best_t, u, v, cnt, -CLV_INF, -1);
        if (best_i == -1 || !f2i(ld4(A.shade + 4 * (size_t)best_i + 2).w)) return OCC_BLOCKED;
        Sens[T_SENSOR + s - 1] = best_t;
        Sens[CAM_TRI + s - 1] = i2f(best_i);
        return OCC_CLEAR;
    }
    V3 cam_o = P.origin(SUB_CAMERA, t - 1);
    int a_triangle = P.triangle(SUB_LIGHT, s - 1), b_triangle = P.triangle(SUB_CAMERA, t - 1);
    if (MODE == TRAVERSE_ORDERED) {
        V3 direction = normalize(cam_o - lig_o);
        if (b_triangle < 0 || b_triangle == a_triangle) return OCC_BLOCKED;
        ray.o = lig_o; ray.d = direction; ray.inv = rcp(direction); ray.skip = b_triangle;
        if (!test_one(A, b_triangle, lig_o, direction, ray.t_bound, ray.rank_bound)) return OCC_BLOCKED;
        return occ_begin<COUNT>(A, ray, cnt);
    }
    /* reference-order flavour: visibility_test (:178-197) as written, one closest-hit search */
    V3 direction = normalize(cam_o - lig_o);
    int best_i = -1;
    float best_t = CLV_INF, u = 0.0f, v = 0.0f;
    closest_hit<MODE, COUNT>(A, lig_o, direction, rcp(direction), best_i, best_t, u, v, cnt, -CLV_INF, -1);
    if (best_i == -1 || best_i == a_triangle) return OCC_BLOCKED;
    return best_i == b_triangle ? OCC_CLEAR : OCC_BLOCKED;
}

/* one surviving strategy q = (t-1)*7 + s of the sample behind X: its MIS weight (:693-776) and contribution (:778-823).
 * t != 1: returns add = true with `term` = w*g*color/p_s and `w_out` = w, to be added to the sample's totals IN THE
 * REFERENCE'S ORDER (t outer, s inner) by the caller; t == 1: the light-image entry is emitted here (splat or staged
 * arrays) and add stays false.  PB: 6 floats of scratch private to the evaluating thread (p_values[0..s-1]). */
template <bool SPLAT, typename PV>
CLV_HD void connect_strategy(const ConnectCtx<PV> &X, int q, TArr PB, uint32_t entry_id, const Material *materials,
                             const Camera &c, const LightSink &sink, V3 &term, float &w_out, bool &add) {
    const PV &P = X.P;
    TArr S = X.S;
    const unsigned spec_l = X.spec_l, spec_c = X.spec_c;
    const int total_pixels = c.pixel_width * c.pixel_height;
    term = v3(0, 0, 0);
    w_out = 0.0f;
    add = false;
    {
        const int t = q / 7 + 1, s = q % 7;

        V3 cam_n = v3(0, 0, 0), lig_o = v3(0, 0, 0), lig_n = v3(0, 0, 0), dir_l_to_c = v3(0, 0, 0);
        float cam_tot, g_conn = 0.0f;
        int cam_material;
        int light_pixel_idx = -1;
        bool sensor_vertex = false; /* camera endpoint is the rewritten vertex 0 (:607-616) */

        if (s == 0) {
            cam_tot = S[C_TOT + t - 1];
            cam_material = P.material(SUB_CAMERA, t - 1);
        } else if (t == 1) {
            lig_o = P.origin(SUB_LIGHT, s - 1);
            lig_n = P.normal(SUB_LIGHT, s - 1);
            V3 direction = normalize(xyz(c.focal_point) - lig_o);
            V3 camera_point = lig_o + direction * X.Sens[T_SENSOR + s - 1];
            float x = dot(camera_point - xyz(c.center), xyz(c.dx));
            float y = dot(camera_point - xyz(c.center), xyz(c.dy));
            int pixel_x = (int)roundf((x / c.phys_width + 0.5f) * (float)c.pixel_width);
            int pixel_y = (int)roundf((y / c.phys_height + 0.5f) * (float)c.pixel_height);
            light_pixel_idx = pixel_y * c.pixel_width + pixel_x;
            sensor_vertex = true;
            cam_n = xyz(c.direction);
            float cam_cos = fabsf(dot(normalize(xyz(c.focal_point) - camera_point), cam_n));
            cam_tot = 1.0f;
            cam_material = 7;
            dir_l_to_c = normalize(camera_point - lig_o);
            g_conn = geometry_from(S[L_COS + s - 1], cam_cos, lig_o, camera_point);
        } else {
            V3 cam_o = P.origin(SUB_CAMERA, t - 1);
            cam_n = P.normal(SUB_CAMERA, t - 1);
            lig_o = P.origin(SUB_LIGHT, s - 1);
            lig_n = P.normal(SUB_LIGHT, s - 1);
            dir_l_to_c = normalize(cam_o - lig_o);
            cam_tot = S[C_TOT + t - 1];
            cam_material = P.material(SUB_CAMERA, t - 1);
            g_conn = geometry_from(S[L_COS + s - 1], S[C_COS + t - 1], lig_o, cam_o);
        }

        /* MIS weight (:693-776).  Full-path vertex i < s is light vertex i, otherwise camera vertex j = k-1-i (:546-549).
         * p_values[i] = p_i: p_s at i = s, divided down the light side by the ratios p_ratios[i] (:751-755), multiplied
         * up the camera side (:745-749); entries i and i+1 of every specular vertex i and entry k are zero (:759-766);
         * the sum runs over ascending i (:768-770).  Same operations in the same order as the reference, written as
         * three short loops over the vertices that exist instead of fixed-size arrays. */
        const int k = s + t; /* number of vertices */
        unsigned specmask = spec_l & ((1u << s) - 1u);
        for (int j = 0; j < t; j++) {
            bool sp = (j == 0 && sensor_vertex) ? X.spec_sensor : (((spec_c >> j) & 1u) != 0);
            if (sp) specmask |= 1u << (k - 1 - j);
        }
        const unsigned zmask = specmask | (specmask << 1) | (1u << k);

        float prior_light_importance = (s == 0) ? 1.0f : S[L_TOT + s - 1];
        float p_s = cam_tot * prior_light_importance;

        float p_i = p_s;
        for (int i = s - 1; i >= 0; i--) { /* light side, p_ratios[i] = num / denom of :711-734 */
            float num = (i == 0) ? S[L_LIMP] : S[L_LIMP + i] * S[G_LL + i - 1];
            float denom = S[L_CIMP + i] * ((i == s - 1) ? g_conn : S[G_LL + i]);
            p_i = p_i / (num / denom);
            PB[i] = p_i;
        }
        float sum = 0.0f;
        for (int i = 0; i < s; i++) sum += ((zmask >> i) & 1u) ? 0.0f : PB[i];
        const float p_at_s = ((zmask >> s) & 1u) ? 0.0f : p_s;
        sum += p_at_s;
        p_i = p_s;
        for (int i = s; i <= k - 2; i++) { /* camera side */
            const int j = k - 1 - i;
            float num = (i == 0) ? S[C_LIMP + j] : S[C_LIMP + j] * ((i == s) ? g_conn : S[G_CC + j]);
            float denom = S[C_CIMP + j] * S[G_CC + j - 1];
            p_i = (num / denom) * p_i;
            sum += ((zmask >> (i + 1)) & 1u) ? 0.0f : p_i;
        }

        /* one back edge only: a `continue` here would let the lanes that skip drift apart from the rest for good */
        if (p_at_s > 0.0f && sum > 0.0f) {
        float w = p_at_s / sum;

        V3 color;
        float g = 1.0f, new_light_f = 1.0f;
        if (s == 0) {
            V3 prior_color = P.color(SUB_CAMERA, t - 2);
            V3 emission = xyz(ld4(&materials[cam_material].emission));
            color = prior_color * emission;
        } else if (t == 1) {
            int prior_light_ind = (s - 2 > 0) ? s - 2 : 0;
            V3 prior_color = P.color(SUB_LIGHT, prior_light_ind);
            if (s > 1) new_light_f = fabsf(dot(dir_l_to_c, lig_n)) / CLV_PI;
            color = (prior_color * new_light_f) * xyz(ld4(&materials[P.material(SUB_LIGHT, s - 1)].color));
            g = g_conn; /* cosine_geometry_term(light_ray, camera_ray) */
        } else {
            V3 prior_camera_color = P.color(SUB_CAMERA, t - 2);
            float new_camera_f = fabsf(dot(-dir_l_to_c, cam_n)) / CLV_PI;
            V3 camera_color = (prior_camera_color * new_camera_f) * xyz(ld4(&materials[cam_material].color));
            V3 light_color;
            if (s == 1) {
                light_color = xyz(ld4(&materials[P.material(SUB_LIGHT, s - 1)].emission));
            } else {
                V3 prior_light_color = P.color(SUB_LIGHT, s - 2);
                new_light_f = fabsf(dot(dir_l_to_c, lig_n)) / CLV_PI;
                light_color = (prior_light_color * new_light_f) * xyz(ld4(&materials[P.material(SUB_LIGHT, s - 1)].color));
            }
            color = camera_color * light_color;
            g = g_conn; /* cosine_geometry_term(camera_ray, light_ray): symmetric, same bits */
        }
        if (t != 1) {
            term = (color * (w * g)) / p_s;
            w_out = w;
            add = true;
        } else {
            float shade = new_light_f * g / p_s;
            if (SPLAT) {
                /* light_image_gather (:937-964) applied on the spot: weights * shades * prior.color * mat.color */
                if (light_pixel_idx >= 0 && light_pixel_idx < sink.n_pixels) {
                    int prior_light_ind = (s - 2 > 0) ? s - 2 : 0;
                    V3 contrib = (P.color(SUB_LIGHT, prior_light_ind) * (w * shade)) *
                                 xyz(ld4(&materials[P.material(SUB_LIGHT, s - 1)].color));
                    atomic_add4(&sink.splat_image[light_pixel_idx], v4(contrib.x, contrib.y, contrib.z, w));
                }
            } else {
                size_t e = (size_t)entry_id + (size_t)s * (size_t)total_pixels;
                sink.pixel_indices[e] = light_pixel_idx;
                sink.path_indices[e] = (int32_t)entry_id;
                sink.ray_indices[e] = s - 1;
                sink.weights[e] = w;
                sink.shade[e] = shade;
            }
        }
        }
    }
}

/* finish: MIS + contributions of the surviving strategies, splat weights, aggregator */
/* 3x3 splat weights (:827-864) and the aggregator record of one sample */
template <typename PV>
CLV_HD void connect_write_aggregator(const PV &P, V3 total_contribution, float contrib_weight_sum, uint32_t agg_slot,
                                     const Camera &c, WeightAggregator *weight_aggregators, V4 *out) {
    WeightAggregator agg;
    for (int i = 0; i < 3; i++) { agg.pad0[i] = 0.0f; agg.pad1[i] = 0.0f; }
    const int pixel_idx = P.pixel_idx0();
    const V3 sensor_origin = P.origin(SUB_CAMERA, 0);
    float weight_sum = 0.0f;
    float ppw = c.phys_width / (float)c.pixel_width;
    float pph = c.phys_height / (float)c.pixel_height;
    float sigma = 0.5f * sqrtf(ppw * ppw + pph * pph);
    float wts[3][3];
#pragma unroll
    for (int i = -1; i < 2; i++) {
#pragma unroll
        for (int j = -1; j < 2; j++) {
            wts[i + 1][j + 1] = 0.0f;
            int x = (pixel_idx % c.pixel_width) + i;
            int y = (pixel_idx / c.pixel_width) + j;
            if (x < 0 || x >= c.pixel_width || y < 0 || y >= c.pixel_height) continue;
            float weight = gaussian_weight(pixel_center(c, x, y), sensor_origin, sigma);
            wts[i + 1][j + 1] = weight;
            weight_sum += weight;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) agg.weights[i][j] = (weight_sum != 0.0f) ? wts[i][j] / weight_sum : wts[i][j];
    agg.total_contribution = pad0(total_contribution);
    agg.contrib_weight_sum = contrib_weight_sum;
    if (out) out[agg_slot] = v4(total_contribution.x, total_contribution.y, total_contribution.z, 1.0f);
    weight_aggregators[agg_slot] = agg;
}

/* the sequential form: one thread evaluates its sample's survivors one after the other (host build, the stage-level ABI's
 * Path records; k_connect_finish distributes the same evaluations over the lanes of a warp) */
template <bool SPLAT, typename PV>
CLV_HD void connect_finish(const ConnectCtx<PV> &X, unsigned long long survivors, uint32_t entry_id, uint32_t agg_slot,
                           const Material *materials, const Camera &c, WeightAggregator *weight_aggregators, V4 *out,
                           const LightSink &sink, unsigned warp_lanes /* device: lanes of the warp that make this call */) {
    V3 total_contribution = v3(0, 0, 0);
    float contrib_weight_sum = 0.0f;
    TArr PB;
    PB.base = &X.S[P_BACK];
    PB.stride = X.S.stride;
    while (survivors) {
        int q;
#if defined(__CUDA_ARCH__)
        q = __ffsll((long long)survivors) - 1;
#else
        q = __builtin_ctzll(survivors);
#endif
        survivors &= survivors - 1;
        V3 term;
        float w;
        bool add;
        connect_strategy<SPLAT>(X, q, PB, entry_id, materials, c, sink, term, w, add);
        /* one back edge only: a `continue` here would let the lanes that skip drift apart from the rest for good */
        if (add) {
            total_contribution = total_contribution + term;
            contrib_weight_sum += w;
        }
    }
#if defined(__CUDA_ARCH__)
    /* lanes leave the strategy loop at different times; without this the splat-weight code below ran with 2.3 of 32
     * lanes active (ncu, profiles/r01_notes.md) */
    __syncwarp(warp_lanes);
#endif
    connect_write_aggregator(X.P, total_contribution, contrib_weight_sum, agg_slot, c, weight_aggregators, out);
}

/* the three phases for one sample, rays traced one after the other by the calling thread (host build; also the shape
 * of the computation the warp-cooperative kernel distributes) */
template <int MODE, bool SPLAT, bool COUNT, typename PV>
CLV_HD void connect_sample(const PV &P, uint32_t entry_id, uint32_t agg_slot, const AccelView &A, const Material *materials,
                           const Camera &c, WeightAggregator *weight_aggregators, V4 *out, const LightSink &sink, TArr S,
                           Cnt<COUNT> &cnt) {
    ConnectCtx<PV> X;
    unsigned long long valid0;
    float sens_words[CLV_SENSOR_WORDS];
    TArr Sens;
    Sens.base = sens_words;
    Sens.stride = 1;
    connect_stage<COUNT>(P, materials, S, Sens, X, valid0, cnt);
    float cam_words[CLV_ENUM_SCRATCH];
    TArr Cam;
    Cam.base = cam_words;
    Cam.stride = 1;
    unsigned long long cand = connect_enumerate<MODE, COUNT>(P, A, materials, c, Sens, Cam, cnt);
    unsigned long long visible = 0ull, m = cand;
    while (m) {
#if defined(__CUDA_ARCH__)
        int q = __ffsll((long long)m) - 1;
#else
        int q = __builtin_ctzll(m);
#endif
        m &= m - 1;
        OccRay ray;
        int stack[CLV_OCC_STACK];
        int st = connect_ray_begin<MODE, COUNT>(q, P, Sens, A, c, ray, cnt);
        while (st == OCC_RUNNING) st = occ_round<COUNT>(A, ray, stack, cnt);
        if (st == OCC_CLEAR) visible |= 1ull << q;
    }
#if defined(__CUDA_ARCH__)
    const unsigned warp_lanes = __activemask();
#else
    const unsigned warp_lanes = 0u;
#endif
    connect_finish<SPLAT>(X, valid0 | visible, entry_id, agg_slot, materials, c, weight_aggregators, out, sink, warp_lanes);
}

/* adaptive_finalize_samples for one pixel, trace.metal:981-1018, with general sample bins */
CLV_HD void finalize_pixel(uint32_t id, const WeightAggregator *aggs, const Camera &camera, const uint32_t *bin_offsets,
                           V3 &total_sample, float &weight_sum, uint32_t &count) {
    total_sample = v3(0, 0, 0);
    weight_sum = 0.0f;
    int px = (int)(id % (uint32_t)camera.pixel_width), py = (int)(id / (uint32_t)camera.pixel_width);
    for (int i = -1; i < 2; i++) {
        for (int j = -1; j < 2; j++) {
            int sx = px + i, sy = py + j;
            if (sx < 0 || sx >= camera.pixel_width || sy < 0 || sy >= camera.pixel_height) continue;
            int sample_index = sy * camera.pixel_width + sx;
            uint32_t k0 = bin_offsets ? bin_offsets[sample_index] : (uint32_t)sample_index;
            uint32_t k1 = bin_offsets ? bin_offsets[sample_index + 1] : (uint32_t)sample_index + 1;
            for (uint32_t k = k0; k < k1; k++) {
                const WeightAggregator *wa = aggs + k;
                float weight = wa->weights[1 - i][1 - j];
                total_sample = total_sample + xyz(wa->total_contribution) * weight;
                weight_sum += weight * wa->contrib_weight_sum;
            }
        }
    }
    count = bin_offsets ? bin_offsets[id + 1] - bin_offsets[id] : 1u;
}

CLV_HD float nan_to_num0(float x) { return (x != x || fabsf(x) == CLV_INF) ? 0.0f : x; } /* renderer.py:271,278 */

/* ---- fused per-pass pieces (Renderer.run_sample without host round trips) --------------------------------------- */
struct Accum { float image[3], weight, uni[3], count; }; /* CLV_BUF_ACCUM_PACKED, 32 bytes per pixel */
static_assert(sizeof(Accum) == 32, "Accum");

/* stages 1-4 of renderer.py:281-285 for one sample: both ray generators and both random walks, RNG streams consumed in
 * the reference's order (light gen, camera gen, light walk, camera walk), then the unidirectional part of
 * process_images (renderer.py:273-278).  `gid` = global sample id = pixel; this pixel's accumulator is touched by
 * exactly one sample per pass, so the read-modify-write needs no atomics. */
/* the fused renderer's path storage: SoA records (default) or the reference's Path records */
struct PathStore {
    Path *light_paths, *camera_paths; /* AoS, or NULL */
    V4 *soa;                          /* SoA vertex parts, or NULL */
    int32_t *soa_len;
    uint32_t n;
};
template <int MODE, bool COUNT, bool SOA>
CLV_HD void fused_walk_sample(uint32_t slot, uint32_t gid, const Triangle *light_triangles, const float *surface_areas,
                              const int32_t *light_triangle_indices, int light_count, const Material *materials,
                              const Camera &c, const AccelView &A, uint32_t &seed0, uint32_t &seed1, const PathStore &store,
                              Accum *accum, Cnt<COUNT> &cnt) {
    Ray lray = make_light_ray(light_triangles, surface_areas, light_triangle_indices, materials, light_count, seed0, seed1);
    Ray cray = make_camera_ray(c, (int)gid, seed0, seed1);
    V4 uni;
    /* one copy of the walk code for both subpaths (two inlined copies made instruction fetch the top stall reason) */
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int which = 0; which < 2; which++) { /* which = SUB_LIGHT, then SUB_CAMERA */
        if (SOA) {
            SoaSink sink;
            sink.v = store.soa; sink.len = store.soa_len; sink.n = store.n; sink.slot = slot; sink.sub = which;
            random_walk<MODE, COUNT>(which ? cray : lray, A, materials, seed0, seed1, sink, &uni, cnt);
        } else {
            AosSink sink;
            sink.path = (which ? store.camera_paths : store.light_paths) + slot;
            random_walk<MODE, COUNT>(which ? cray : lray, A, materials, seed0, seed1, sink, &uni, cnt);
        }
    }
    Accum *a = accum + gid;
    a->uni[0] += nan_to_num0(uni.x);
    a->uni[1] += nan_to_num0(uni.y);
    a->uni[2] += nan_to_num0(uni.z);
    a->count += 1.0f;
}

/* ---- wavefront form of stages 1-4 (k_walk_init / k_walk_trace / k_walk_shade) ------------------------------------------ *
 * A sample runs its light walk, then its camera walk (the camera walk's RNG state is where the light walk left it, so the
 * two cannot overlap).  Its walk state lives in global memory between kernels, 16-byte parts, part-major over the samples:
 *   s0: origin.xyz, tot_importance of the vertex the ray leaves     s1: direction.xyz, carried importance
 *   s2: colour.xyz, flags (bit 0 camera walk, bits 1-3 bounce index, bit 4 unidirectional estimate found)
 *   s3: last stored colour.xyz, hit_light of the current vertex (camera walk only: feeds the unidirectional estimate)
 *   hit: closest hit of the current ray (triangle, t, u, v)         jitter: the two sensor-jitter draws of the camera ray
 * k_walk_trace reads s0/s1 (32 B per ray) and writes hit (16 B); k_walk_shade reads and writes everything else.
 * Round r of a pass traces ray number r of every sample that still has one (at most 12: 6 bounces per walk). */
struct WalkState {
    V4 *s0, *s1, *s2, *s3, *hit;
    float *jitter; /* [n][2] */
    uint8_t *lag;  /* [n]: rounds the sample has fallen behind because a search of its was suspended */
};
enum { WS_CAMERA = 1, WS_BOUNCE_SHIFT = 1, WS_UNI_FOUND = 16 };
#define CLV_WALK_ROUNDS (2 * CLV_MAX_BOUNCES)
#define CLV_WALK_LAUNCH_ROUNDS (CLV_WALK_ROUNDS + CLV_WALK_MAX_LAG) /* with suspended searches a sample may need more rounds */

template <bool SOA> struct SinkOf;
template <> struct SinkOf<true> {
    typedef SoaSink type;
    static CLV_HDM SoaSink make(const PathStore &st, uint32_t slot, int which) {
        SoaSink k;
        k.v = st.soa; k.len = st.soa_len; k.n = st.n; k.slot = slot; k.sub = which;
        return k;
    }
};
template <> struct SinkOf<false> {
    typedef AosSink type;
    static CLV_HDM AosSink make(const PathStore &st, uint32_t slot, int which) {
        AosSink k;
        k.path = (which ? st.camera_paths : st.light_paths) + slot;
        return k;
    }
};

CLV_HD void walk_state_save(const WalkState &ws, uint32_t slot, const WalkRegs &W, int which, int bounce) {
    stcg4(ws.s0 + slot, v4(W.o.x, W.o.y, W.o.z, W.ray_tot));
    stcg4(ws.s1 + slot, v4(W.d.x, W.d.y, W.d.z, W.carry_imp));
    stcg4(ws.s2 + slot, v4(W.rc.x, W.rc.y, W.rc.z, i2f(which | (bounce << WS_BOUNCE_SHIFT) | (W.uni_found ? WS_UNI_FOUND : 0))));
    if (which) stcg4(ws.s3 + slot, v4(W.last_stored_color.x, W.last_stored_color.y, W.last_stored_color.z, i2f(W.ray_hit_light)));
}

/* start of a sample's pass: both generators in the reference's order (light rays, then camera rays: renderer.py:281-282),
 * light walk ready for its first ray */
template <bool SOA>
CLV_HD void wavefront_init_sample(uint32_t slot, const Triangle *light_triangles, const float *surface_areas,
                                  const int32_t *light_triangle_indices, int light_count, const Material *materials,
                                  uint32_t &seed0, uint32_t &seed1, const PathStore &store, const WalkState &ws) {
    Ray lray = make_light_ray(light_triangles, surface_areas, light_triangle_indices, materials, light_count, seed0, seed1);
    ws.lag[slot] = 0;
    ws.jitter[2 * (size_t)slot] = xorshift_random(seed0); /* make_camera_ray's two draws */
    ws.jitter[2 * (size_t)slot + 1] = xorshift_random(seed1);
    WalkRegs W;
    walk_begin(lray, SinkOf<SOA>::make(store, slot, SUB_LIGHT), W);
    walk_state_save(ws, slot, W, SUB_LIGHT, 0);
}

/* one bounce of a sample given the hit of its current ray: WALK_NEXT while the sample has another ray to trace, WALK_WAIT
 * when the search for this hit was suspended (the sample waits a round and the search is resumed), else WALK_OVER */
enum { WALK_OVER = 0, WALK_NEXT = 1, WALK_WAIT = 2 };
template <bool COUNT, bool SOA>
CLV_HD int wavefront_shade_sample(uint32_t slot, uint32_t gid, const Material *materials, const Camera &c, const AccelView &A,
                                   uint32_t &seed0, uint32_t &seed1, const PathStore &store, const WalkState &ws, Accum *accum,
                                   Cnt<COUNT> &cnt) {
    const V4 h = ldcg4(ws.hit + slot);
    if (f2i(h.x) == CLV_HIT_PENDING) return WALK_WAIT;
    const V4 a0 = ldcg4(ws.s0 + slot), a1 = ldcg4(ws.s1 + slot), a2 = ldcg4(ws.s2 + slot);
    const int flags = f2i(a2.w);
    const int which = flags & WS_CAMERA, bounce = (flags >> WS_BOUNCE_SHIFT) & 7;
    WalkRegs W;
    W.o = xyz(a0); W.ray_tot = a0.w;
    W.d = xyz(a1); W.carry_imp = a1.w;
    W.rc = xyz(a2);
    W.uni_found = (flags & WS_UNI_FOUND) != 0;
    W.uni = v4(0, 0, 0, 0);
    W.length = bounce; /* every earlier bounce of this walk succeeded */
    W.last_stored_color = v3(0, 0, 0);
    W.ray_hit_light = -1;
    if (which) {
        const V4 a3 = ldcg4(ws.s3 + slot);
        W.last_stored_color = xyz(a3);
        W.ray_hit_light = f2i(a3.w);
    }
    const bool had_uni = W.uni_found;
    bool more;
    {
        typename SinkOf<SOA>::type sink = SinkOf<SOA>::make(store, slot, which);
        more = walk_bounce<COUNT>(W, bounce, which != 0, f2i(h.x), h.y, h.z, h.w, A, materials, seed0, seed1, sink, cnt);
        more = more && bounce + 1 < CLV_MAX_BOUNCES;
        if (!more) sink.finish(W.length, which != 0);
    }
    if (which && W.uni_found && !had_uni) {
        /* renderer.py:273-278: the estimate is added to this pixel's accumulator the moment it is known; a pass that
         * finds none adds 0, which leaves the sum's bits unchanged */
        Accum *a = accum + gid;
        a->uni[0] += nan_to_num0(W.uni.x);
        a->uni[1] += nan_to_num0(W.uni.y);
        a->uni[2] += nan_to_num0(W.uni.z);
    }
    if (more) {
        walk_state_save(ws, slot, W, which, bounce + 1);
        return WALK_NEXT;
    }
    if (!which) { /* light walk over: the camera walk starts with the RNG state the light walk leaves */
        Ray cray = camera_ray_from_jitter(c, (int)gid, ws.jitter[2 * (size_t)slot], ws.jitter[2 * (size_t)slot + 1]);
        walk_begin(cray, SinkOf<SOA>::make(store, slot, SUB_CAMERA), W);
        walk_state_save(ws, slot, W, SUB_CAMERA, 0);
        return WALK_NEXT;
    }
    accum[gid].count += 1.0f;
    return WALK_OVER;
}

/* stages 6-8 for one pixel: adaptive_finalize_samples with identity bins, the light image the connect kernel
 * splatted this pass (cleared here for the next one), and process_images' accumulation (renderer.py:253-271).
 * Returns this pixel's sample weight of the pass. */
CLV_HD float fused_finalize_pixel(uint32_t id, const WeightAggregator *aggs, const Camera &c, V4 *light_image,
                                  Accum *accum) {
    V3 fin;
    float wsum;
    uint32_t count;
    finalize_pixel(id, aggs, c, nullptr, fin, wsum, count);
    V4 li = light_image[id];
    light_image[id] = v4(0.0f, 0.0f, 0.0f, 0.0f);
    Accum *a = accum + id;
    a->image[0] += nan_to_num0(li.x + fin.x);
    a->image[1] += nan_to_num0(li.y + fin.y);
    a->image[2] += nan_to_num0(li.z + fin.z);
    float wpass = wsum + li.w;
    a->weight += wpass;
    return wpass;
}

} /* namespace clv */
#endif /* CLV_DEVICE_CUH */
