/* warp_sim.cpp — PLANNING AID (host only, never shipped): replays real closest-hit searches of a frame under different
 * warp schedules and reports the lane utilisation each schedule would reach, before a kernel is written for it.
 *
 * A ray's ORDERED closest-hit search is a fixed sequence of steps whatever the schedule: inner-node steps ("I", one 64-byte
 * node record, two slab tests) and leaf steps ("L k", k triangle records / tests).  trace_rays() records that sequence with
 * the device code's own slab_test / tri_test (clv_device.cuh) and checks the result against closest_hit<ORDERED>.
 * simulate() then runs 32-lane warps over a queue of such traces under a schedule and counts issue slots:
 *   utilisation = sum over lanes of useful step cost / (32 x issue slots of the warp).
 * Costs are SASS instruction counts of the step bodies (C_I, C_T) plus per-iteration control and per-ray fetch overheads.
 */
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../clive2_b200/csrc/clv_accel_build.h"

using namespace clv;
#define EXPORT extern "C" __attribute__((visibility("default")))

struct SimAccel {
    HostAccel h;
    std::vector<Triangle> ref;
    AccelView view;
};
EXPORT SimAccel *sim_accel_create(const Box *boxes, int32_t n_boxes, const Triangle *tris, int32_t n_tris) {
    SimAccel *a = new SimAccel();
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
EXPORT void sim_accel_destroy(SimAccel *a) { delete a; }

/* step codes: 0 = inner node, k > 0 = leaf of k triangles */
struct Traces {
    std::vector<uint8_t> steps;
    std::vector<uint64_t> first; /* n + 1 offsets */
};

static void trace_one(const AccelView &A, V3 o, V3 d, V3 inv, std::vector<uint8_t> &out, int &best_i_out) {
    int st_ref[256];
    float st_t[256];
    int sp = 0, best_i = -1, best_rank = 0x7fffffff;
    float best_t = CLV_INF;
    bool hit;
    float te;
    slab_test(o, inv, xyz(A.root_min), xyz(A.root_max), hit, te);
    best_i_out = -1;
    if (!(hit && te < best_t)) return;
    int cur = A.root_ref;
    bool have = true;
    auto pop = [&]() {
        have = false;
        while (!have && sp > 0) { sp--; if (st_t[sp] <= best_t) { cur = st_ref[sp]; have = true; } }
    };
    while (have) {
        if (cur >= 0) {
            out.push_back(0);
            const V4 *n = A.nodes + 4 * (size_t)cur;
            V4 n0 = n[0], n1 = n[1], n2 = n[2], n3 = n[3];
            bool hl, hr;
            float tl, tr;
            slab_test(o, inv, v3(n0.x, n0.y, n0.z), v3(n0.w, n1.x, n1.y), hl, tl);
            slab_test(o, inv, v3(n1.z, n1.w, n2.x), v3(n2.y, n2.z, n2.w), hr, tr);
            int lref = f2i(n3.x), rref = f2i(n3.y);
            bool go_l = hl && tl <= best_t, go_r = hr && tr <= best_t, left_first = tl < tr;
            if (go_l && go_r) { st_ref[sp] = left_first ? rref : lref; st_t[sp] = left_first ? tr : tl; sp++; }
            if (go_l || go_r) cur = (go_l && (left_first || !go_r)) ? lref : rref;
            else pop();
        } else {
            int code = ~cur, first = code >> 4, count = code & 15;
            out.push_back((uint8_t)count);
            for (int i = first; i < first + count; i++) {
                const V4 *tp = A.tris + 3 * (size_t)i;
                float t, u, v;
                if (tri_test(o, d, xyz(tp[0]), xyz(tp[1]), xyz(tp[2]), t, u, v) &&
                    (t < best_t || (t == best_t && t < CLV_INF && f2i(tp[0].w) < best_rank))) {
                    best_i = i; best_t = t; best_rank = f2i(tp[0].w);
                }
            }
            pop();
        }
    }
    best_i_out = best_i;
}

EXPORT Traces *sim_trace_rays(const SimAccel *a, const Ray *rays, uint32_t n, int *mismatches) {
    Traces *T = new Traces();
    T->first.push_back(0);
    int bad = 0;
    for (uint32_t i = 0; i < n; i++) {
        int bi;
        trace_one(a->view, xyz(rays[i].origin), xyz(rays[i].direction), xyz(rays[i].inv_direction), T->steps, bi);
        T->first.push_back(T->steps.size());
        Cnt<false> cnt;
        int ref_i = -1;
        float bt = CLV_INF, u = 0, v = 0;
        closest_hit<TRAVERSE_ORDERED, false>(a->view, xyz(rays[i].origin), xyz(rays[i].direction), xyz(rays[i].inv_direction),
                                             ref_i, bt, u, v, cnt, -CLV_INF, -1);
        if (ref_i != bi) bad++;
    }
    *mismatches = bad;
    return T;
}
EXPORT void sim_traces_destroy(Traces *T) { delete T; }
/* leaves longer than `max_leaf` become a synthetic inner node (same box: both halves are always visited) over two halves,
 * recursively — what clv_accel_build.h leaf_ref() does for leaves longer than 15 */
static void split_leaf(std::vector<uint8_t> &out, int k, int max_leaf) {
    if (k <= max_leaf) { out.push_back((uint8_t)k); return; }
    out.push_back(0);
    split_leaf(out, k / 2, max_leaf);
    split_leaf(out, k - k / 2, max_leaf);
}
EXPORT Traces *sim_traces_split_leaves(const Traces *T, int max_leaf) {
    Traces *R = new Traces();
    R->first.push_back(0);
    for (size_t r = 0; r + 1 < T->first.size(); r++) {
        for (uint64_t i = T->first[r]; i < T->first[r + 1]; i++) {
            if (T->steps[i] == 0) R->steps.push_back(0);
            else split_leaf(R->steps, T->steps[i], max_leaf);
        }
        R->first.push_back(R->steps.size());
    }
    return R;
}
EXPORT void sim_trace_stats(const Traces *T, double out[4]) {
    uint64_t n = T->first.size() - 1, inner = 0, leaves = 0, tris = 0;
    for (uint8_t s : T->steps) { if (s == 0) inner++; else { leaves++; tris += s; } }
    out[0] = (double)inner / n; out[1] = (double)leaves / n; out[2] = (double)tris / n; out[3] = (double)n;
}

struct Costs { double c_inner, c_tri, c_leaf_setup, c_iter, c_fetch, c_pop; };

/* schedules */
enum { S_SYNC_WHILE = 0,   /* no refill: 32 rays start together, while-while, warp ends with its slowest ray (k_fused_walks) */
       S_REFILL_WHILE = 1, /* refill below R lanes, while-while rounds (k_connect_rays schedule 0) */
       S_REFILL_VOTE = 2,  /* refill below R, per iteration ONE kind of step by majority vote, whole leaves */
       S_STEP_BOTH = 3,    /* refill below R, per iteration: inner block (lanes at inner nodes) then triangle block of at
                              most B triangles (lanes in leaves) — both run when anyone needs them */
       S_STEP_VOTE = 4,    /* like 3 but only the block more lanes wait for (weighted) runs */
       S_COOP_LEAF = 5,    /* vote; a leaf step spreads ALL pending triangles of the lanes at leaves over the 32 lanes
                              (ray data by shuffle, results by warp reduction): ceil(total / 32) rounds of
                              (c_tri + c_coop) slots; tri_batch = c_coop (extra instructions per cooperative test) */
};

struct Lane { const uint8_t *p, *end; int tri_left; bool have; };

/* K rays per lane: a lane runs the voted kind of step on whichever of its rays waits for it (refill below `refill_below`
 * of the 32*K slots); extra_iter = issue slots per iteration for picking / swapping the ray */
EXPORT void sim_run_multi(const Traces *T, const uint32_t *order, uint32_t n, int K, int refill_below, double vote_weight,
                          const Costs *C, double extra_iter, int chunk, double out[4]) {
    double useful = 0.0, slots = 0.0, iters = 0.0;
    uint64_t warps = 0;
    for (uint32_t base = 0; base < n; base += chunk) {
        const uint32_t end = std::min<uint32_t>(n, base + chunk);
        uint32_t next = base;
        std::vector<Lane> L(32 * K);
        for (auto &l : L) l.have = false;
        warps++;
        for (;;) {
            int holding = 0;
            for (auto &l : L) holding += l.have;
            if (next < end && holding < refill_below) {
                bool any = false;
                for (auto &l : L) if (!l.have && next < end) {
                    uint32_t r = order[next++];
                    l.p = T->steps.data() + T->first[r]; l.end = T->steps.data() + T->first[r + 1];
                    l.have = l.p < l.end; useful += C->c_fetch; any = true;
                }
                if (any) slots += C->c_fetch;
            } else if (holding == 0) break;
            iters += 1;
            slots += C->c_iter + extra_iter;
            int ni = 0, nl = 0;
            for (int lane = 0; lane < 32; lane++) {
                bool ci = false, cl = false;
                for (int k = 0; k < K; k++) { Lane &l = L[lane * K + k]; if (l.have) { if (*l.p == 0) ci = true; else cl = true; } }
                ni += ci; nl += cl;
            }
            const bool inner = ni * vote_weight >= nl;
            int m = 0, longest = 0;
            double u = 0.0;
            for (int lane = 0; lane < 32; lane++) {
                for (int k = 0; k < K; k++) {
                    Lane &l = L[lane * K + k];
                    if (!l.have) continue;
                    if (inner && *l.p == 0) { m++; u += C->c_inner; l.p++; if (l.p >= l.end) l.have = false; break; }
                    if (!inner && *l.p != 0) { int take = *l.p; m++; longest = std::max(longest, take); u += take * C->c_tri + C->c_leaf_setup; l.p++; if (l.p >= l.end) l.have = false; break; }
                }
            }
            if (m) { slots += inner ? C->c_inner : longest * C->c_tri + C->c_leaf_setup; useful += u; }
        }
    }
    out[0] = useful / (32.0 * slots); out[1] = slots / std::max<uint64_t>(1, warps); out[2] = iters / std::max<uint64_t>(1, warps); out[3] = slots;
}

EXPORT void sim_run(const Traces *T, const uint32_t *order, uint32_t n, int schedule, int refill_below, int tri_batch,
                    double vote_weight, const Costs *C, int chunk, double out[4]) {
    /* rays order[0..n) are dealt to warps in chunks of `chunk` consecutive queue entries (chunk = 32 for S_SYNC_WHILE) */
    double useful = 0.0, slots = 0.0, iters = 0.0;
    uint64_t warps = 0;
    if (schedule == S_SYNC_WHILE) chunk = 32;
    for (uint32_t base = 0; base < n; base += chunk) {
        const uint32_t end = std::min<uint32_t>(n, base + chunk);
        uint32_t next = base;
        int rb = schedule == S_SYNC_WHILE ? 33 : refill_below; /* refill threshold of this warp */
        Lane L[32];
        for (int k = 0; k < 32; k++) L[k].have = false;
        warps++;
        auto fetch = [&](int k) {
            uint32_t r = order[next++];
            L[k].p = T->steps.data() + T->first[r];
            L[k].end = T->steps.data() + T->first[r + 1];
            L[k].tri_left = 0;
            L[k].have = L[k].p < L[k].end;
            useful += C->c_fetch;
        };
        for (;;) {
            int holding = 0;
            for (int k = 0; k < 32; k++) holding += L[k].have;
            if (next < end && holding < rb) {
                bool any = false;
                for (int k = 0; k < 32 && next < end; k++) if (!L[k].have) { fetch(k); any = true; }
                if (any) slots += C->c_fetch;
                if (schedule == S_SYNC_WHILE) rb = 0; /* one fill only */
            } else if (holding == 0) break;
            iters += 1;
            slots += C->c_iter;
            /* lane needs: inner step if *p == 0 and no triangles pending; triangles otherwise */
            auto at_inner = [&](int k) { return L[k].have && L[k].tri_left == 0 && *L[k].p == 0; };
            auto at_leaf = [&](int k) { return L[k].have && (L[k].tri_left > 0 || *L[k].p != 0); };
            auto advance = [&](int k) { L[k].p++; if (L[k].p >= L[k].end) L[k].have = false; };
            auto inner_block = [&]() {
                int m = 0;
                for (int k = 0; k < 32; k++) if (at_inner(k)) { m++; advance(k); }
                if (m) { slots += C->c_inner; useful += m * C->c_inner; }
                return m;
            };
            auto leaf_block = [&](int batch) { /* up to `batch` triangles per lane (0 = whole leaf) */
                int longest = 0, m = 0;
                double u = 0.0;
                for (int k = 0; k < 32; k++) if (at_leaf(k)) {
                    if (L[k].tri_left == 0) L[k].tri_left = *L[k].p;
                    int take = batch > 0 ? std::min(batch, L[k].tri_left) : L[k].tri_left;
                    L[k].tri_left -= take;
                    longest = std::max(longest, take);
                    u += take * C->c_tri + C->c_leaf_setup;
                    m++;
                    if (L[k].tri_left == 0) advance(k);
                }
                if (m) { slots += longest * C->c_tri + C->c_leaf_setup; useful += u; }
                return m;
            };
            if (schedule == S_SYNC_WHILE || schedule == S_REFILL_WHILE) {
                /* while (any at inner) inner step; then leaves */
                while (inner_block()) {}
                leaf_block(0);
            } else if (schedule == S_REFILL_VOTE) {
                int ni = 0, nl = 0;
                for (int k = 0; k < 32; k++) { ni += at_inner(k); nl += at_leaf(k); }
                if (ni * vote_weight >= nl) inner_block(); else leaf_block(0);
            } else if (schedule == S_COOP_LEAF) {
                int ni = 0, nl = 0, total = 0;
                for (int k = 0; k < 32; k++) { ni += at_inner(k); if (at_leaf(k)) { nl++; total += *L[k].p; } }
                /* compare lanes kept busy: an inner step occupies ni lanes, a cooperative leaf step min(32, total) */
                if (ni * vote_weight >= std::min(32, total) && ni > 0) inner_block();
                else if (total > 0) {
                    const int rounds = (total + 31) / 32;
                    slots += rounds * (C->c_tri + (double)tri_batch) + C->c_leaf_setup + 12.0 /* scan */;
                    useful += total * C->c_tri;
                    for (int k = 0; k < 32; k++) if (at_leaf(k)) advance(k);
                }
            } else if (schedule == S_STEP_BOTH) {
                inner_block();
                leaf_block(tri_batch);
            } else {
                int ni = 0, nl = 0;
                for (int k = 0; k < 32; k++) { ni += at_inner(k); nl += at_leaf(k); }
                if (ni * vote_weight >= nl) inner_block(); else leaf_block(tri_batch);
            }
        }
    }
    out[0] = useful / (32.0 * slots);
    out[1] = slots / std::max<uint64_t>(1, warps);
    out[2] = iters / std::max<uint64_t>(1, warps);
    out[3] = slots;
}
