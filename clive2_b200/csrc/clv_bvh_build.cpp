/* clv_bvh_build.cpp — native host builder of the reference's BVH.
 *
 * Reproduces, bit for bit, what clive2_b200/bvh.py (the numpy specification pinned against the reference's own
 * src/bvh.py:132-191, 288-389) emits with a stable sort: greedy top-down SAH object split on centroid-sorted triangles,
 * leaves of <= max_members triangles, breadth-first flat layout with "right == 0 => inner node, children at left and
 * left + 1".  The reference needs ~55 us per triangle and an O(n^2) flatten (minutes at 1 M triangles, tens of minutes
 * at 10 M); this one is O(n log^2 n) and runs subtrees as OpenMP tasks.
 *
 * Reference behaviours that shape the tree and are therefore kept:
 *   - arithmetic in the soup's own dtype T (float32 for the Cornell-only scene, float64 once a mesh joins): centroid
 *     (min + max) / 2, box extents, surface terms 2 * ((sx*sy + sy*sz) + sz*sx); the products with the integer counts
 *     are formed in float64 (numpy promotes float32 * int64 to float64);
 *   - the left-count factor is the 0-based index k (one less than the true left size), the right one (n-1) - k
 *     (bvh.py:150-154);
 *   - first minimum inside an axis, strict improvement across axes x -> y -> z;
 *   - a node becomes a leaf when it has <= max_members triangles or when the reference's work stack held more than
 *     max_depth entries after the node was popped (bvh.py:294) — that stack size equals the number of LEFT turns on the
 *     path from the root, which is what makes subtrees independent here;
 *   - child boxes are the min/max over the members' boxes; the flat records narrow them to float32.
 * Build: g++ -O3 -fopenmp -ffp-contract=off (no FMA contraction: one rounding per operation, like numpy).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <queue>
#include <vector>

namespace {

struct FlatBox { /* struct_types.Box, 48 bytes */
    float min[4], max[4];
    int32_t left, right, pad[2];
};
static_assert(sizeof(FlatBox) == 48, "Box");

struct Node {
    int64_t first, count; /* range in the permutation array */
    int64_t left = -1, right = -1;
    double lo[3], hi[3];  /* exact values of type T widened to double (lossless for float and double) */
};

template <typename T>
struct Builder {
    static const int64_t PARALLEL_AXES_MIN = 1 << 15; /* nodes at least this large sweep their three axes as tasks */
    static const int64_t PARALLEL_SORT_MIN = 1 << 17; /* ranges at least this large are sorted as two task halves + merge */
    const T *mins, *maxes; /* [n][3] */
    int64_t n;
    int max_members, max_depth;
    std::vector<T> centers;   /* [n][3] */
    std::vector<int64_t> perm; /* current order of all triangles; a node owns a contiguous range */
    std::vector<Node> nodes;

    static T half_area_x2(const T lo[3], const T hi[3]) {
        T s0 = hi[0] - lo[0], s1 = hi[1] - lo[1], s2 = hi[2] - lo[2];
        T a = s0 * s1;
        T b = s1 * s2;
        T c = s2 * s0;
        T sum = (a + b) + c;
        return (T)2 * sum;
    }

    struct Scratch {
        std::vector<int64_t> order[3];
        std::vector<T> left_area[3], right_area[3];
        std::vector<int64_t> merge_buffer[3];
    };

    /* stable sort of ord[0, count) by centroid along `axis`.  Large ranges are cut in two halves sorted as OpenMP tasks
     * and merged stably: the same permutation std::stable_sort gives (a stable merge of stably sorted halves IS the
     * stable sort), so the tree does not depend on the thread count. */
    void sort_axis(int64_t *ord, int64_t count, int axis, int64_t *buffer) {
        const T *c = centers.data();
        auto less = [c, axis](int64_t a, int64_t b) { return c[3 * a + axis] < c[3 * b + axis]; };
        if (count < PARALLEL_SORT_MIN) {
            std::stable_sort(ord, ord + count, less);
            return;
        }
        const int64_t half = count / 2;
#pragma omp task default(shared) firstprivate(ord, half, axis, buffer)
        sort_axis(ord, half, axis, buffer);
#pragma omp task default(shared) firstprivate(ord, half, count, axis, buffer)
        sort_axis(ord + half, count - half, axis, buffer + half);
#pragma omp taskwait
        std::merge(ord, ord + half, ord + half, ord + count, buffer, less); /* stable: ties take the left range first */
        std::copy(buffer, buffer + count, ord);
    }

    /* SAH sweep along one axis of perm[first, first+count): cost and position of the axis' first minimum (km < 0: none) */
    void sweep_axis(int64_t first, int64_t count, int axis, Scratch &S, double &m_out, int64_t &km_out) {
        std::vector<int64_t> &ord = S.order[axis];
        ord.assign(perm.begin() + first, perm.begin() + first + count);
        if (count >= PARALLEL_SORT_MIN) S.merge_buffer[axis].resize(count);
        sort_axis(ord.data(), count, axis, S.merge_buffer[axis].data());
        std::vector<T> &left_area = S.left_area[axis], &right_area = S.right_area[axis];
        left_area.resize(count);
        right_area.resize(count);
        T lo[3], hi[3];
        for (int64_t k = 0; k < count; k++) { /* prefix boxes */
            const T *mn = mins + 3 * ord[k], *mx = maxes + 3 * ord[k];
            for (int d = 0; d < 3; d++) {
                lo[d] = (k == 0) ? mn[d] : std::min(lo[d], mn[d]);
                hi[d] = (k == 0) ? mx[d] : std::max(hi[d], mx[d]);
            }
            left_area[k] = half_area_x2(lo, hi);
        }
        for (int64_t k = count - 1; k >= 0; k--) { /* suffix boxes */
            const T *mn = mins + 3 * ord[k], *mx = maxes + 3 * ord[k];
            for (int d = 0; d < 3; d++) {
                lo[d] = (k == count - 1) ? mn[d] : std::min(lo[d], mn[d]);
                hi[d] = (k == count - 1) ? mx[d] : std::max(hi[d], mx[d]);
            }
            right_area[k] = half_area_x2(lo, hi);
        }
        /* cost[k] = left_area[k] * k + right_area[k+1] * ((n-1) - k), k = 0 .. n-2; first minimum (numpy argmin
         * returns the first NaN if there is one) */
        double m = 0.0;
        int64_t km = -1;
        for (int64_t k = 0; k + 1 < count; k++) {
            double cost = (double)left_area[k] * (double)k + (double)right_area[k + 1] * (double)((count - 1) - k);
            if (cost != cost) { m = cost; km = k; break; }
            if (km < 0 || cost < m) { m = cost; km = k; }
        }
        m_out = m;
        km_out = km;
    }

    /* best SAH object split of perm[first, first+count); returns n_left (0 = no valid split) and leaves the chosen
     * order in perm.  The three axes of a large node are swept as parallel tasks (near the root there are no sibling
     * subtrees to keep the other threads busy); the choice between them is made afterwards in x, y, z order. */
    int64_t split(int64_t first, int64_t count, Scratch &S) {
        double m[3];
        int64_t km[3];
        if (count >= PARALLEL_AXES_MIN) {
            for (int axis = 0; axis < 3; axis++) {
#pragma omp task default(shared) firstprivate(axis, first, count)
                sweep_axis(first, count, axis, S, m[axis], km[axis]);
            }
#pragma omp taskwait
        } else {
            for (int axis = 0; axis < 3; axis++) sweep_axis(first, count, axis, S, m[axis], km[axis]);
        }
        double best_cost = std::numeric_limits<double>::infinity();
        int best_axis = -1;
        int64_t best_left = 0;
        for (int axis = 0; axis < 3; axis++)
            if (km[axis] >= 0 && m[axis] < best_cost) { best_cost = m[axis]; best_axis = axis; best_left = km[axis] + 1; }
        if (best_axis < 0) return 0;
        std::copy(S.order[best_axis].begin(), S.order[best_axis].end(), perm.begin() + first);
        return best_left;
    }

    void bounds(int64_t first, int64_t count, double lo[3], double hi[3]) const {
        T l[3], h[3];
        for (int64_t k = 0; k < count; k++) {
            const T *mn = mins + 3 * perm[first + k], *mx = maxes + 3 * perm[first + k];
            for (int d = 0; d < 3; d++) {
                l[d] = (k == 0) ? mn[d] : std::min(l[d], mn[d]);
                h[d] = (k == 0) ? mx[d] : std::max(h[d], mx[d]);
            }
        }
        for (int d = 0; d < 3; d++) { lo[d] = (double)l[d]; hi[d] = (double)h[d]; }
    }

    int64_t new_node(int64_t first, int64_t count) {
        int64_t id;
#pragma omp critical(clv_bvh_nodes)
        {
            id = (int64_t)nodes.size();
            nodes.emplace_back();
        }
        return id;
    }

    /* builds the subtree of node `id` (range and box already set) */
    void build(int64_t id, int left_turns) {
        int64_t first, count;
#pragma omp critical(clv_bvh_nodes)
        { first = nodes[id].first; count = nodes[id].count; }
        if (count <= max_members || left_turns > max_depth) return;
        Scratch S;
        int64_t n_left = split(first, count, S);
        if (n_left <= 0 || n_left >= count) return;
        Node l, r;
        l.first = first; l.count = n_left;
        r.first = first + n_left; r.count = count - n_left;
        bounds(l.first, l.count, l.lo, l.hi);
        bounds(r.first, r.count, r.lo, r.hi);
        int64_t li, ri;
#pragma omp critical(clv_bvh_nodes)
        {
            li = (int64_t)nodes.size();
            nodes.push_back(l);
            ri = (int64_t)nodes.size();
            nodes.push_back(r);
            nodes[id].left = li;
            nodes[id].right = ri;
        }
        if (count > 8192) {
#pragma omp task default(shared) firstprivate(li, left_turns)
            build(li, left_turns + 1);
#pragma omp task default(shared) firstprivate(ri, left_turns)
            build(ri, left_turns);
#pragma omp taskwait
        } else {
            build(li, left_turns + 1);
            build(ri, left_turns);
        }
    }
};

template <typename T>
int64_t run(const T *mins, const T *maxes, int64_t n, int max_members, int max_depth, FlatBox *out_boxes,
            int64_t box_capacity, int64_t *out_order) {
    Builder<T> B;
    B.mins = mins; B.maxes = maxes; B.n = n; B.max_members = max_members; B.max_depth = max_depth;
    B.centers.resize((size_t)n * 3);
    for (int64_t i = 0; i < n * 3; i++) B.centers[i] = (mins[i] + maxes[i]) / (T)2;
    B.perm.resize(n);
    for (int64_t i = 0; i < n; i++) B.perm[i] = i;
    B.nodes.reserve((size_t)std::max<int64_t>(16, n / 2));
    Node root;
    root.first = 0; root.count = n;
    B.bounds(0, n, root.lo, root.hi);
    if (n == 0) for (int d = 0; d < 3; d++) { root.lo[d] = INFINITY; root.hi[d] = -INFINITY; }
    B.nodes.push_back(root);
#pragma omp parallel
#pragma omp single
    B.build(0, 0);

    /* breadth-first flatten (bvh.py:329-389) */
    if ((int64_t)B.nodes.size() > box_capacity) return -(int64_t)B.nodes.size();
    std::queue<int64_t> q;
    q.push(0);
    int64_t bi = 0, ti = 0;
    while (!q.empty()) {
        int64_t id = q.front();
        q.pop();
        const Node &nd = B.nodes[id];
        FlatBox &bx = out_boxes[bi];
        memset(&bx, 0, sizeof bx);
        for (int d = 0; d < 3; d++) { bx.min[d] = (float)nd.lo[d]; bx.max[d] = (float)nd.hi[d]; }
        if (nd.left >= 0) {
            bx.left = (int32_t)(bi + (int64_t)q.size() + 1);
            bx.right = 0;
            q.push(nd.left);
            q.push(nd.right);
        } else {
            bx.left = (int32_t)ti;
            for (int64_t k = 0; k < nd.count; k++) out_order[ti + k] = B.perm[nd.first + k];
            ti += nd.count;
            bx.right = (int32_t)ti;
        }
        bi++;
    }
    return bi;
}

} /* namespace */

/* mins/maxes: [n][3] of float (is_double = 0) or double (is_double = 1).  out_boxes: capacity records; out_order: n
 * entries.  Returns the number of boxes written, or -(needed) when the capacity is too small. */
extern "C" __attribute__((visibility("default")))
int64_t clv_bvh_build(const void *mins, const void *maxes, int is_double, int64_t n, int max_members, int max_depth,
                      void *out_boxes, int64_t box_capacity, int64_t *out_order) {
    if (is_double) return run<double>((const double *)mins, (const double *)maxes, n, max_members, max_depth,
                                      (FlatBox *)out_boxes, box_capacity, out_order);
    return run<float>((const float *)mins, (const float *)maxes, n, max_members, max_depth, (FlatBox *)out_boxes,
                      box_capacity, out_order);
}

/* struct_types.Triangle records for soup members `order` (bvh.py:361-370 of the reference, bvh.emit_triangles here):
 * corners / smoothed corner normals [n][3][3] and face normals [n][3] of float (is_double = 0) or double, narrowed to
 * float32 by a plain conversion (round to nearest even, what numpy's astype does); material int32, emitter / camera
 * bool bytes.  A memory-bound gather: one OpenMP loop. */
namespace {
struct FlatTriangle { /* struct_types.Triangle, 128 bytes */
    float v[3][4], n[3][4], normal[4];
    int32_t material, is_light, is_camera, pad;
};
static_assert(sizeof(FlatTriangle) == 128, "Triangle");

template <typename T>
void emit(const int64_t *order, int64_t count, const T *corners, const T *smooth, const T *face_normals,
          const int32_t *material, const uint8_t *emitter, const uint8_t *camera, FlatTriangle *out) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < count; k++) {
        const int64_t i = order[k];
        FlatTriangle t;
        memset(&t, 0, sizeof t);
        for (int c = 0; c < 3; c++)
            for (int d = 0; d < 3; d++) {
                t.v[c][d] = (float)corners[9 * i + 3 * c + d];
                t.n[c][d] = (float)smooth[9 * i + 3 * c + d];
            }
        for (int d = 0; d < 3; d++) t.normal[d] = (float)face_normals[3 * i + d];
        t.material = material[i];
        t.is_light = emitter[i] ? 1 : 0;
        t.is_camera = camera[i] ? 1 : 0;
        out[k] = t;
    }
}
} /* namespace */

extern "C" __attribute__((visibility("default")))
void clv_emit_triangles(const int64_t *order, int64_t count, int is_double, const void *corners, const void *smooth,
                        const void *face_normals, const int32_t *material, const uint8_t *emitter,
                        const uint8_t *camera, void *out) {
    if (is_double) emit<double>(order, count, (const double *)corners, (const double *)smooth, (const double *)face_normals,
                                material, emitter, camera, (FlatTriangle *)out);
    else emit<float>(order, count, (const float *)corners, (const float *)smooth, (const float *)face_normals, material,
                     emitter, camera, (FlatTriangle *)out);
}
