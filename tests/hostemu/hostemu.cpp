/* hostemu.cpp — TEST INFRASTRUCTURE.  Compiles the per-sample kernel bodies of clive2_b200/csrc/clv_device.cuh for the
 * HOST (g++, no CUDA) so their logic can be diffed against oracle/trace_ref.c in a container without a GPU.
 *
 * Nothing in clive2_b200/ loads this library; the product's only execution path is the CUDA one in
 * libclive_b200.so.  The `-m gpu` tests repeat the same comparisons on the real kernels.
 * Build: tests/hostemu/__init__.py (g++ -O2 -fopenmp -ffp-contract=off: same arithmetic contract as -fmad=false).
 */
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../clive2_b200/csrc/clv_accel_build.h"

using namespace clv;

#define EXPORT extern "C" __attribute__((visibility("default")))

struct EmuAccel {
    HostAccel h;
    std::vector<Triangle> ref;
    AccelView view;
};

EXPORT EmuAccel *emu_accel_create(const Box *boxes, int32_t n_boxes, const Triangle *tris, int32_t n_tris) {
    EmuAccel *a = new EmuAccel();
    if (build_host_accel(boxes, n_boxes, tris, n_tris, a->h)) { delete a; return nullptr; }
    a->ref.assign(tris, tris + n_tris);
    a->view.nodes = &a->h.nodes.data()->a;
    a->view.wnodes = a->h.wide.empty() ? nullptr : a->h.wide.data()->q;
    a->view.wroot_ref = a->h.wide_root_ref;
    a->view.tris = a->h.tris.data();
    a->view.shade = a->h.shade.data();
    a->view.root_min = boxes[0].min;
    a->view.root_max = boxes[0].max;
    a->view.root_ref = a->h.root_ref;
    a->view.n_camera_tris = (int32_t)a->h.camera_tris.size();
    a->view.camera_tris = a->h.camera_tris.data();
    return a;
}
EXPORT void emu_accel_destroy(EmuAccel *a) { delete a; }
EXPORT void emu_accel_info(EmuAccel *a, int64_t out[3]) {
    out[0] = (int64_t)a->h.nodes.size(); out[1] = (int64_t)a->h.tris.size() / 3; out[2] = a->h.root_ref;
}

static unsigned long long g_cnt[5];
EXPORT void emu_counters(unsigned long long out[5], int reset) {
    for (int i = 0; i < 5; i++) out[i] = g_cnt[i];
    if (reset) memset(g_cnt, 0, sizeof g_cnt);
}
static void merge(const Cnt<true> &c) {
#pragma omp critical(emu_cnt)
    { g_cnt[0] += c.rr; g_cnt[1] += c.rt; g_cnt[2] += c.nf; g_cnt[3] += c.tf; g_cnt[4] += c.vx; }
}

EXPORT void emu_generate_light_rays(uint32_t n, const Triangle *lt, const float *areas, const int32_t *lti,
                                    const Material *mats, uint32_t *rnd, Ray *out, const int32_t *counts) {
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n; id++) {
        uint32_t s0 = rnd[2 * id], s1 = rnd[2 * id + 1];
        out[id] = make_light_ray(lt, areas, lti, mats, counts[0], s0, s1);
        rnd[2 * id] = s0; rnd[2 * id + 1] = s1;
    }
}
EXPORT void emu_generate_camera_rays(uint32_t n, const Camera *cam, uint32_t *rnd, const uint32_t *indices, Ray *out) {
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n; id++) {
        uint32_t s0 = rnd[2 * id], s1 = rnd[2 * id + 1];
        out[id] = make_camera_ray(cam[0], (int)indices[id], s0, s1);
        rnd[2 * id] = s0; rnd[2 * id + 1] = s1;
    }
}

template <int MODE>
static void gen_paths(uint32_t n, const Ray *rays, const EmuAccel *a, const Material *mats, uint32_t *rnd, V4 *out,
                      Path *paths, V4 *debug) {
#pragma omp parallel
    {
        Cnt<true> cnt;
#pragma omp for schedule(dynamic, 64)
        for (int64_t id = 0; id < (int64_t)n; id++) {
            uint32_t s0 = rnd[2 * id], s1 = rnd[2 * id + 1];
            V4 uni;
            AosSink sink;
            sink.path = paths + id;
            random_walk<MODE, true>(rays[id], a->view, mats, s0, s1, sink, &uni, cnt);
            out[id] = uni;
            if (debug) debug[id] = v4(100.0f, 100.0f, 100.0f, 100.0f);
            rnd[2 * id] = s0; rnd[2 * id + 1] = s1;
        }
        merge(cnt);
    }
}
EXPORT void emu_generate_paths(uint32_t n, const Ray *rays, const EmuAccel *a, const Material *mats, uint32_t *rnd,
                               V4 *out, Path *paths, V4 *debug, int mode) {
    if (mode == TRAVERSE_REFERENCE) gen_paths<TRAVERSE_REFERENCE>(n, rays, a, mats, rnd, out, paths, debug);
    else gen_paths<TRAVERSE_ORDERED>(n, rays, a, mats, rnd, out, paths, debug);
}

template <int MODE, bool SPLAT>
static void connect(uint32_t n, const Path *cp, const Path *lp, const EmuAccel *a, const Material *mats,
                    const Camera *cam, WeightAggregator *aggs, V4 *out, LightSink sink) {
#pragma omp parallel
    {
        Cnt<true> cnt;
#pragma omp for schedule(dynamic, 64)
        for (int64_t id = 0; id < (int64_t)n; id++) {
            float scratch[CLV_CONNECT_SCRATCH];
            TArr S;
            S.base = scratch; S.stride = 1;
            connect_sample<MODE, SPLAT, true>(aos_paths(cp, lp, (uint32_t)id), (uint32_t)id, (uint32_t)id, a->view, mats,
                                              cam[0], aggs, out, sink, S, cnt);
        }
        merge(cnt);
    }
}
EXPORT void emu_connect_paths(uint32_t n, const Path *cp, const Path *lp, const EmuAccel *a, const Material *mats,
                              const Camera *cam, WeightAggregator *aggs, V4 *out, int32_t *pix, int32_t *path,
                              int32_t *ray, float *w, float *shade, V4 *splat_image, int mode) {
    LightSink sink;
    memset(&sink, 0, sizeof sink);
    sink.pixel_indices = pix; sink.path_indices = path; sink.ray_indices = ray; sink.weights = w; sink.shade = shade;
    sink.splat_image = splat_image;
    sink.n_pixels = cam[0].pixel_width * cam[0].pixel_height;
    bool splat = splat_image != nullptr;
    if (mode == TRAVERSE_REFERENCE) {
        if (splat) connect<TRAVERSE_REFERENCE, true>(n, cp, lp, a, mats, cam, aggs, out, sink);
        else connect<TRAVERSE_REFERENCE, false>(n, cp, lp, a, mats, cam, aggs, out, sink);
    } else {
        if (splat) connect<TRAVERSE_ORDERED, true>(n, cp, lp, a, mats, cam, aggs, out, sink);
        else connect<TRAVERSE_ORDERED, false>(n, cp, lp, a, mats, cam, aggs, out, sink);
    }
}

EXPORT void emu_adaptive_finalize_samples(uint32_t n, const WeightAggregator *aggs, const Camera *cam, V4 *out,
                                          uint32_t *counts, const uint32_t *bins, float *weights) {
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)n; id++) {
        V3 total;
        float wsum;
        uint32_t count;
        finalize_pixel((uint32_t)id, aggs, cam[0], bins, total, wsum, count);
        counts[id] = count;
        out[id] = v4(total.x, total.y, total.z, 1.0f);
        weights[id] = wsum;
    }
}

template <int MODE>
static void traverse(uint32_t n, const Ray *rays, const EmuAccel *a, int32_t *tri, float *t, float *u, float *v) {
#pragma omp parallel
    {
        Cnt<true> cnt;
#pragma omp for schedule(dynamic, 256)
        for (int64_t id = 0; id < (int64_t)n; id++) {
            int best_i = -1;
            float best_t = CLV_INF, bu = 0.0f, bv = 0.0f;
            cnt.ray_ref();
            closest_hit<MODE, true>(a->view, xyz(rays[id].origin), xyz(rays[id].direction), xyz(rays[id].inv_direction),
                                    best_i, best_t, bu, bv, cnt, -CLV_INF, -1);
            tri[id] = best_i; t[id] = best_t; u[id] = bu; v[id] = bv;
        }
        merge(cnt);
    }
}
EXPORT void emu_traverse(uint32_t n, const Ray *rays, const EmuAccel *a, int32_t *tri, float *t, float *u, float *v,
                         int mode) {
    if (mode == TRAVERSE_REFERENCE) traverse<TRAVERSE_REFERENCE>(n, rays, a, tri, t, u, v);
    else traverse<TRAVERSE_ORDERED>(n, rays, a, tri, t, u, v);
}

/* per-ray work of the ORDERED closest-hit search: node records and triangle records fetched (tools/lane_model.py) */
EXPORT void emu_traverse_counts(uint32_t n, const Ray *rays, const EmuAccel *a, uint32_t *nodes, uint32_t *tris) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t id = 0; id < (int64_t)n; id++) {
        Cnt<true> cnt;
        int best_i = -1;
        float best_t = CLV_INF, u = 0.0f, v = 0.0f;
        closest_hit<TRAVERSE_ORDERED, true>(a->view, xyz(rays[id].origin), xyz(rays[id].direction),
                                            xyz(rays[id].inv_direction), best_i, best_t, u, v, cnt, -CLV_INF, -1);
        nodes[id] = cnt.nf;
        tris[id] = cnt.tf;
    }
}

/* ---- the warp-cooperative leaf steps (hit_leaf_step_warp / occ_leaf_step_warp) on 32 host threads in lockstep ----------- *
 * The kernels run them with shuffles and ballots; here a "warp" is an OpenMP team of 32 threads and every warp primitive
 * is a write to a shared slot between two barriers, so the very same source runs (and is compared, leaf by leaf, with the
 * per-lane sequential loop it replaces). */
#include <omp.h>
struct HostWarpShared { uint32_t slot[32]; };
struct HostWarp {
    HostWarpShared *s;
    int lane_;
    int lane() const { return lane_; }
    unsigned ballot(bool p) const {
        s->slot[lane_] = p ? 1u : 0u;
#pragma omp barrier
        unsigned m = 0;
        for (int k = 0; k < 32; k++) m |= (s->slot[k] & 1u) << k;
#pragma omp barrier
        return m;
    }
    uint32_t shfl_bits(uint32_t v, int src) const {
        s->slot[lane_] = v;
#pragma omp barrier
        const uint32_t r = s->slot[src & 31];
#pragma omp barrier
        return r;
    }
    int shfl(int v, int src) const { return (int)shfl_bits((uint32_t)v, src); }
    float shfl(float v, int src) const { return i2f((int)shfl_bits((uint32_t)f2i(v), src)); }
};
static bool same_bits(float a, float b) { return f2i(a) == f2i(b); }

/* closest hit: rays in groups of 32; per group every lane walks down to its next leaf on its own, then the leaves of the
 * group are tested (a) lane by lane with hit_leaf_step and (b) together with hit_leaf_step_warp; states must agree bit
 * for bit after every leaf.  out: [0] leaf states compared, [1] mismatches, [2] largest number of pairs in one step,
 * [3] final hits that differ from closest_hit<ORDERED>; returns 0, or -1 when no team of 32 threads could be had */
EXPORT int emu_coop_hit_leaves(uint32_t n, const Ray *rays, const EmuAccel *a, int wide, int64_t out[4]) {
    int64_t compared = 0, mismatches = 0, max_pairs = 0, wrong_hits = 0;
    bool team_ok = true;
    omp_set_dynamic(0);
    for (uint32_t g0 = 0; g0 < n; g0 += 32) {
        HostWarpShared shared;
        int64_t c_cmp = 0, c_mis = 0, c_max = 0, c_wrong = 0;
#pragma omp parallel num_threads(32) reduction(+ : c_cmp, c_mis, c_wrong) reduction(max : c_max)
        {
            if (omp_get_num_threads() != 32) {
#pragma omp single
                team_ok = false;
            } else {
                HostWarp w{&shared, omp_get_thread_num()};
                Cnt<true> cnt;
                const uint32_t id = g0 + (uint32_t)w.lane();
                HitRay ray;
                StackEntry st[CLV_HIT_STACK(true) > CLV_HIT_STACK(false) ? CLV_HIT_STACK(true) : CLV_HIT_STACK(false)];
                bool have = false;
                if (id < n) {
                    ray.o = xyz(rays[id].origin); ray.d = xyz(rays[id].direction); ray.inv = rcp(ray.d);
                    have = (wide ? hit_begin<true, true>(a->view, ray, cnt) : hit_begin<true, false>(a->view, ray, cnt)) == HIT_RUNNING;
                }
                for (;;) {
                    while (have && ray.cur >= 0) /* own inner steps down to the next leaf */
                        have = (wide ? hit_inner_step_wide<true>(a->view, ray, st, cnt) : hit_inner_step<true>(a->view, ray, st, cnt)) == HIT_RUNNING;
                    const unsigned at_leaf_lanes = w.ballot(have);
                    if (!at_leaf_lanes) break;
                    HitRay seq = ray;
                    int status_seq = HIT_RUNNING;
                    if (have) status_seq = hit_leaf_step<true>(a->view, seq, st, cnt);
                    const int pairs = have ? ((~ray.cur) & 15) : 0;
                    int total = 0;
                    for (int k = 0; k < 32; k++) total += w.shfl(pairs, k);
                    const int status_warp = hit_leaf_step_warp<true>(w, a->view, ray, st, have, cnt);
                    if (have) {
                        c_cmp++;
                        if (total > c_max) c_max = total;
                        if (status_warp != status_seq || ray.best_i != seq.best_i || !same_bits(ray.best_t, seq.best_t) ||
                            !same_bits(ray.bu, seq.bu) || !same_bits(ray.bv, seq.bv) || ray.best_rank != seq.best_rank ||
                            ray.sp != seq.sp || (status_warp == HIT_RUNNING && ray.cur != seq.cur)) c_mis++;
                        have = status_warp == HIT_RUNNING;
                    } else if (status_warp != HIT_RUNNING) c_mis++;
                }
                if (id < n) {
                    int best_i = -1;
                    float best_t = CLV_INF, bu = 0.0f, bv = 0.0f;
                    closest_hit<TRAVERSE_ORDERED, true>(a->view, xyz(rays[id].origin), xyz(rays[id].direction), rcp(xyz(rays[id].direction)),
                                                        best_i, best_t, bu, bv, cnt, -CLV_INF, -1);
                    if (best_i != ray.best_i || !same_bits(best_t, ray.best_t) || !same_bits(bu, ray.bu) || !same_bits(bv, ray.bv)) c_wrong++;
                }
            }
        }
        if (!team_ok) return -1;
        compared += c_cmp; mismatches += c_mis; wrong_hits += c_wrong;
        if (c_max > max_pairs) max_pairs = c_max;
    }
    out[0] = compared; out[1] = mismatches; out[2] = max_pairs; out[3] = wrong_hits;
    return 0;
}

/* occlusion queries: ray id runs from rays[id].origin along rays[id].direction, bounded by (t_bound[id], rank_bound[id]),
 * skipping triangle skip[id]; same comparison, occ_leaf_step against occ_leaf_step_warp.  out as above ([3]: final
 * verdicts that differ between the two) */
EXPORT int emu_coop_occ_leaves(uint32_t n, const Ray *rays, const float *t_bound, const int32_t *rank_bound, const int32_t *skip,
                               const EmuAccel *a, int64_t out[4]) {
    int64_t compared = 0, mismatches = 0, max_pairs = 0, wrong = 0;
    bool team_ok = true;
    omp_set_dynamic(0);
    for (uint32_t g0 = 0; g0 < n; g0 += 32) {
        HostWarpShared shared;
        int64_t c_cmp = 0, c_mis = 0, c_max = 0, c_wrong = 0;
#pragma omp parallel num_threads(32) reduction(+ : c_cmp, c_mis, c_wrong) reduction(max : c_max)
        {
            if (omp_get_num_threads() != 32) {
#pragma omp single
                team_ok = false;
            } else {
                HostWarp w{&shared, omp_get_thread_num()};
                Cnt<true> cnt;
                const uint32_t id = g0 + (uint32_t)w.lane();
                OccRay ray, solo;
                int stack[CLV_OCC_STACK], solo_stack[CLV_OCC_STACK];
                bool have = false;
                int verdict = OCC_CLEAR, solo_verdict = OCC_CLEAR;
                if (id < n) {
                    ray.o = xyz(rays[id].origin); ray.d = xyz(rays[id].direction); ray.inv = rcp(ray.d);
                    ray.t_bound = t_bound[id]; ray.rank_bound = rank_bound[id]; ray.skip = skip[id];
                    solo = ray;
                    const int st0 = occ_begin<true>(a->view, ray, cnt);
                    have = st0 == OCC_RUNNING;
                    if (!have) verdict = st0;
                    /* the same query on its own, per-lane steps only */
                    int st1 = occ_begin<true>(a->view, solo, cnt);
                    while (st1 == OCC_RUNNING)
                        st1 = solo.cur >= 0 ? occ_inner_step<true>(a->view, solo, solo_stack, cnt) : occ_leaf_step<true>(a->view, solo, solo_stack, cnt);
                    solo_verdict = st1;
                }
                for (;;) {
                    while (have && ray.cur >= 0) {
                        const int st1 = occ_inner_step<true>(a->view, ray, stack, cnt);
                        have = st1 == OCC_RUNNING;
                        if (!have) verdict = st1;
                    }
                    if (!w.ballot(have)) break;
                    OccRay seq = ray;
                    int seq_stack[CLV_OCC_STACK];
                    int status_seq = OCC_RUNNING;
                    if (have) {
                        for (int k = 0; k < ray.sp; k++) seq_stack[k] = stack[k];
                        status_seq = occ_leaf_step<true>(a->view, seq, seq_stack, cnt);
                    }
                    const int pairs = have ? ((~ray.cur) & 15) : 0;
                    int total = 0;
                    for (int k = 0; k < 32; k++) total += w.shfl(pairs, k);
                    const int status_warp = occ_leaf_step_warp<true>(w, a->view, ray, stack, have, cnt);
                    if (have) {
                        c_cmp++;
                        if (total > c_max) c_max = total;
                        if (status_warp != status_seq || (status_warp == OCC_RUNNING && (ray.cur != seq.cur || ray.sp != seq.sp))) c_mis++;
                        have = status_warp == OCC_RUNNING;
                        if (!have) verdict = status_warp;
                    } else if (status_warp != OCC_RUNNING) c_mis++;
                }
                if (id < n && verdict != solo_verdict) c_wrong++;
            }
        }
        if (!team_ok) return -1;
        compared += c_cmp; mismatches += c_mis; wrong += c_wrong;
        if (c_max > max_pairs) max_pairs = c_max;
    }
    out[0] = compared; out[1] = mismatches; out[2] = max_pairs; out[3] = wrong;
    return 0;
}

EXPORT void emu_visibility(uint32_t n, const Ray *ra, const Ray *rb, const EmuAccel *a, int32_t *visible, int mode) {
#pragma omp parallel
    {
        Cnt<true> cnt;
#pragma omp for schedule(dynamic, 256)
        for (int64_t id = 0; id < (int64_t)n; id++) {
            bool vis;
            if (mode == TRAVERSE_REFERENCE)
                vis = visibility_test<TRAVERSE_REFERENCE, true>(a->view, xyz(ra[id].origin), ra[id].triangle,
                                                               xyz(rb[id].origin), rb[id].triangle, cnt);
            else
                vis = visibility_test<TRAVERSE_ORDERED, true>(a->view, xyz(ra[id].origin), ra[id].triangle,
                                                             xyz(rb[id].origin), rb[id].triangle, cnt);
            visible[id] = vis ? 1 : 0;
        }
        merge(cnt);
    }
}

/* one fused pass exactly as clv_renderer_run_sample launches it: walks -> connect (splat) -> finalize+accumulate.
 * soa != NULL selects the renderer's compact SoA vertex records (soa: V4[2*6*4*n_local], soa_len: int32[2*n_local]),
 * otherwise the reference's Path records are used. */
/* stages 1-4 the way k_walk_init / k_walk_trace / k_walk_shade run them: rounds over a queue of live samples, each ray
 * traced with the resumable single-step search (hit_begin / hit_inner_step / hit_leaf_step), each hit shaded by
 * wavefront_shade_sample, survivors appended to the next round's queue */
template <bool SOA>
static void wavefront_walks(uint32_t n_local, const uint32_t *sample_ids, const Triangle *lt, const float *areas,
                            const int32_t *lti, int light_count, const Material *mats, const Camera *cam, const EmuAccel *a,
                            uint32_t *seeds, PathStore store, Accum *accum, Cnt<true> &cnt, bool wide, int suspend_after) {
    std::vector<V4> state((size_t)n_local * 5);
    std::vector<float> jitter((size_t)n_local * 2);
    std::vector<uint8_t> lag(n_local);
    WalkState ws;
    ws.s0 = state.data(); ws.s1 = ws.s0 + n_local; ws.s2 = ws.s1 + n_local; ws.s3 = ws.s2 + n_local; ws.hit = ws.s3 + n_local;
    ws.jitter = jitter.data();
    ws.lag = lag.data();
    /* continuation records (hit_save / hit_restore): a search that has not ended after `suspend_after` steps is suspended
     * and resumed by the next round, the way k_walk_trace treats the searches that outlast a round */
    std::vector<V4> records[2];
    std::vector<uint32_t> queue(n_local), next;
    for (uint32_t id = 0; id < n_local; id++) {
        uint32_t s0 = seeds[2 * (size_t)id], s1 = seeds[2 * (size_t)id + 1];
        wavefront_init_sample<SOA>(id, lt, areas, lti, light_count, mats, s0, s1, store, ws);
        seeds[2 * (size_t)id] = s0; seeds[2 * (size_t)id + 1] = s1;
        queue[id] = id;
    }
    const int rounds = suspend_after > 0 ? CLV_WALK_LAUNCH_ROUNDS : CLV_WALK_ROUNDS;
    for (int round = 0; round < rounds; round++) {
        /* a later round's queue is in whatever order the warps of the previous one appended to it: reverse it here so that
         * a dependence on the order would show */
        if (round & 1) std::reverse(queue.begin(), queue.end());
        std::vector<V4> &rec_out = records[(round + 1) & 1];
        const std::vector<V4> &rec_in = records[round & 1];
        rec_out.clear();
        /* the round's work list: the searches the previous round suspended, then the queue */
        const size_t n_resume = rec_in.size() / CLV_HIT_RECORD_V4, n_work = n_resume + queue.size();
        for (size_t w = 0; w < n_work; w++) { /* k_walk_trace */
            HitRay ray;
            StackEntry st[CLV_HIT_STACK(true) > CLV_HIT_STACK(false) ? CLV_HIT_STACK(true) : CLV_HIT_STACK(false)];
            int status = HIT_RUNNING;
            uint32_t slot;
            if (w < n_resume) {
                slot = hit_restore(rec_in.data() + w * CLV_HIT_RECORD_V4, ray, st);
            } else {
                slot = queue[w - n_resume];
                ray.o = xyz(ws.s0[slot]); ray.d = xyz(ws.s1[slot]); ray.inv = rcp(ray.d);
                cnt.ray_ref();
                status = wide ? hit_begin<true, true>(a->view, ray, cnt) : hit_begin<true, false>(a->view, ray, cnt);
            }
            int steps = 0;
            bool suspended = false;
            while (status == HIT_RUNNING) {
                if (suspend_after > 0 && steps == suspend_after && ws.lag[slot] < CLV_WALK_MAX_LAG) {
                    const size_t rec = rec_out.size() / CLV_HIT_RECORD_V4;
                    rec_out.resize(rec_out.size() + CLV_HIT_RECORD_V4);
                    hit_save(rec_out.data() + rec * CLV_HIT_RECORD_V4, slot, ray, st);
                    ws.hit[slot] = v4(i2f(CLV_HIT_PENDING), 0.0f, 0.0f, 0.0f);
                    ws.lag[slot]++;
                    suspended = true;
                    break;
                }
                if (ray.cur < 0) status = hit_leaf_step<true>(a->view, ray, st, cnt);
                else status = wide ? hit_inner_step_wide<true>(a->view, ray, st, cnt) : hit_inner_step<true>(a->view, ray, st, cnt);
                steps++;
            }
            if (!suspended) ws.hit[slot] = v4(i2f(ray.best_i), ray.best_t, ray.bu, ray.bv);
        }
        next.clear();
        for (size_t w = 0; w < n_work; w++) { /* k_walk_shade */
            const uint32_t slot = w < n_resume ? hit_record_slot(rec_in.data() + w * CLV_HIT_RECORD_V4) : queue[w - n_resume];
            uint32_t gid = sample_ids ? sample_ids[slot] : slot;
            uint32_t s0 = seeds[2 * (size_t)slot], s1 = seeds[2 * (size_t)slot + 1];
            const int more = wavefront_shade_sample<true, SOA>(slot, gid, mats, cam[0], a->view, s0, s1, store, ws, accum, cnt);
            if (more == WALK_NEXT) next.push_back(slot); /* WALK_WAIT: its new record is on the next round's list already */
            seeds[2 * (size_t)slot] = s0; seeds[2 * (size_t)slot + 1] = s1;
        }
        queue.swap(next);
    }
    if (!queue.empty() || !records[rounds & 1].empty()) abort(); /* every sample ends within the launched rounds */
}

template <int MODE, bool SOA>
static void fused_pass(uint32_t n_local, const uint32_t *sample_ids, const Triangle *lt, const float *areas,
                       const int32_t *lti, int light_count, const Material *mats, const Camera *cam, const EmuAccel *a,
                       uint32_t *seeds, PathStore store, WeightAggregator *aggs, V4 *light_image, Accum *accum, int wavefront) {
    uint32_t n_pixels = (uint32_t)(cam[0].pixel_width * cam[0].pixel_height);
    if (wavefront) {
        Cnt<true> cnt;
        wavefront_walks<SOA>(n_local, sample_ids, lt, areas, lti, light_count, mats, cam, a, seeds, store, accum, cnt, (wavefront & 3) == 2, wavefront >> 8);
        merge(cnt);
    }
#pragma omp parallel
    {
        Cnt<true> cnt;
#pragma omp for schedule(dynamic, 64)
        for (int64_t id = 0; id < (int64_t)n_local; id++) {
            if (wavefront) continue;
            uint32_t gid = sample_ids ? sample_ids[id] : (uint32_t)id;
            uint32_t s0 = seeds[2 * id], s1 = seeds[2 * id + 1];
            fused_walk_sample<MODE, true, SOA>((uint32_t)id, gid, lt, areas, lti, light_count, mats, cam[0], a->view, s0, s1,
                                               store, accum, cnt);
            seeds[2 * id] = s0; seeds[2 * id + 1] = s1;
        }
        LightSink sink;
        memset(&sink, 0, sizeof sink);
        sink.splat_image = light_image;
        sink.n_pixels = (int)n_pixels;
#pragma omp for schedule(dynamic, 64)
        for (int64_t id = 0; id < (int64_t)n_local; id++) {
            uint32_t gid = sample_ids ? sample_ids[id] : (uint32_t)id;
            float scratch[CLV_CONNECT_SCRATCH];
            TArr S;
            S.base = scratch; S.stride = 1;
            if (SOA) {
                SoaPaths P;
                P.v = store.soa; P.len = store.soa_len; P.n = store.n; P.slot = (uint32_t)id; P.pixel0 = (int32_t)gid;
                connect_sample<MODE, true, true>(P, (uint32_t)id, gid, a->view, mats, cam[0], aggs, nullptr, sink, S, cnt);
            } else {
                connect_sample<MODE, true, true>(aos_paths(store.camera_paths, store.light_paths, (uint32_t)id), (uint32_t)id,
                                                 gid, a->view, mats, cam[0], aggs, nullptr, sink, S, cnt);
            }
        }
#pragma omp for schedule(static)
        for (int64_t id = 0; id < (int64_t)n_pixels; id++) fused_finalize_pixel((uint32_t)id, aggs, cam[0], light_image, accum);
        merge(cnt);
    }
}
EXPORT void emu_fused_pass(uint32_t n_local, const uint32_t *sample_ids, const Triangle *lt, const float *areas,
                           const int32_t *lti, int light_count, const Material *mats, const Camera *cam,
                           const EmuAccel *a, uint32_t *seeds, Path *light_paths, Path *camera_paths,
                           WeightAggregator *aggs, V4 *light_image, Accum *accum, int mode, V4 *soa, int32_t *soa_len,
                           int wavefront) {
    PathStore store;
    memset(&store, 0, sizeof store);
    store.light_paths = light_paths; store.camera_paths = camera_paths; store.soa = soa; store.soa_len = soa_len;
    store.n = n_local;
#define GO(M, SO) fused_pass<M, SO>(n_local, sample_ids, lt, areas, lti, light_count, mats, cam, a, seeds, store, aggs, light_image, accum, M == TRAVERSE_ORDERED ? wavefront : 0)
    if (soa) { if (mode == TRAVERSE_REFERENCE) GO(TRAVERSE_REFERENCE, true); else GO(TRAVERSE_ORDERED, true); }
    else { if (mode == TRAVERSE_REFERENCE) GO(TRAVERSE_REFERENCE, false); else GO(TRAVERSE_ORDERED, false); }
#undef GO
}
