/* clv_bvh_build.cpp — native host builder of the reference's BVH.
 *
 * Reproduces, bit for bit, what clive2_b200/bvh.py (the numpy specification pinned against the reference's own
 * src/bvh.py:132-191, 288-389) emits with a stable sort: greedy top-down SAH object split on centroid-sorted triangles,
 * leaves of <= max_members triangles, breadth-first flat layout with "right == 0 => inner node, children at left and
 * left + 1".  The reference needs ~55 us per triangle and an O(n^2) flatten (minutes at 1 M triangles, tens of minutes
 * at 10 M); this one sorts the triangles three times at the root and then only partitions (O(n log n)), and runs
 * subtrees, the three sweeps of large nodes and the halves of the root's sorts as OpenMP tasks.
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

    struct Keyed { T key; int64_t idx; }; /* centroid coordinate next to the index: sorts and sweeps read memory in order */
    typedef std::vector<Keyed> Order;

    /* One node's work: its triangles in three orders, each the STABLE sort of the node's current order P by the centroid
     * coordinate of that axis (what bvh.py's `argsort(kind="stable")` of the node's member list gives).  P itself is the
     * order along `cur_axis`, the axis the parent was split on (the root's P is the input order). */
    struct Work {
        int64_t id, first, count;
        int left_turns, cur_axis;
        Order K[3];
    };
    std::vector<int64_t> pos;   /* [n] position of a triangle in the order of the split axis of the node being split */
    std::vector<uint8_t> side;  /* [n] 0 = goes left, 1 = goes right */

    /* stable sort of a[0, count) by key.  Large ranges are cut in two halves sorted as OpenMP tasks and merged stably:
     * the same permutation std::stable_sort gives (a stable merge of stably sorted halves IS the stable sort), so the
     * tree does not depend on the thread count. */
    static void sort_keyed(Keyed *a, int64_t count, Keyed *buffer) {
        auto less = [](const Keyed &x, const Keyed &y) { return x.key < y.key; };
        if (count < PARALLEL_SORT_MIN) {
            std::stable_sort(a, a + count, less);
            return;
        }
        const int64_t half = count / 2;
#pragma omp task default(shared) firstprivate(a, half, buffer)
        sort_keyed(a, half, buffer);
#pragma omp task default(shared) firstprivate(a, half, count, buffer)
        sort_keyed(a + half, count - half, buffer + half);
#pragma omp taskwait
        std::merge(a, a + half, a + half, a + count, buffer, less); /* stable: ties take the left range first */
        std::copy(buffer, buffer + count, a);
    }

    /* SAH sweep along one order: cost and position of its first minimum (km < 0: none) */
    void sweep(const Order &ord, double &m_out, int64_t &km_out) const {
        const int64_t count = (int64_t)ord.size();
        std::vector<T> left_area(count), right_area(count);
        T lo[3], hi[3];
        for (int64_t k = 0; k < count; k++) { /* prefix boxes */
            const T *mn = mins + 3 * ord[k].idx, *mx = maxes + 3 * ord[k].idx;
            for (int d = 0; d < 3; d++) {
                lo[d] = (k == 0) ? mn[d] : std::min(lo[d], mn[d]);
                hi[d] = (k == 0) ? mx[d] : std::max(hi[d], mx[d]);
            }
            left_area[k] = half_area_x2(lo, hi);
        }
        for (int64_t k = count - 1; k >= 0; k--) { /* suffix boxes */
            const T *mn = mins + 3 * ord[k].idx, *mx = maxes + 3 * ord[k].idx;
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

    void bounds_of(const Keyed *a, int64_t count, double lo[3], double hi[3]) const {
        T l[3], h[3];
        for (int64_t k = 0; k < count; k++) {
            const T *mn = mins + 3 * a[k].idx, *mx = maxes + 3 * a[k].idx;
            for (int d = 0; d < 3; d++) {
                l[d] = (k == 0) ? mn[d] : std::min(l[d], mn[d]);
                h[d] = (k == 0) ? mx[d] : std::max(h[d], mx[d]);
            }
        }
        for (int d = 0; d < 3; d++) { lo[d] = (double)l[d]; hi[d] = (double)h[d]; }
    }

    /* `src` (one of the node's orders, axis != split axis) cut into the members of the two children, each again the
     * stable sort of the CHILD's current order by this axis.  A stable partition keeps the sort; what changes is the
     * tie order: equal keys must follow the child's current order, which is the order along the split axis (`pos`),
     * no longer the node's own.  When the node was itself split along the same axis as its parent the two agree. */
    void partition(const Order &src, int64_t n_left, bool fix_ties, Order &left, Order &right) const {
        const int64_t count = (int64_t)src.size();
        left.resize(n_left);
        right.resize(count - n_left);
        int64_t nl = 0, nr = 0;
        for (int64_t k = 0; k < count; k++) {
            if (side[src[k].idx]) right[nr++] = src[k];
            else left[nl++] = src[k];
        }
        if (!fix_ties) return;
        for (Order *o : {&left, &right}) {
            Keyed *a = o->data();
            const int64_t m = (int64_t)o->size();
            for (int64_t k = 0; k < m;) {
                int64_t e = k + 1;
                while (e < m && !(a[k].key < a[e].key) && !(a[e].key < a[k].key)) e++;
                if (e - k > 1) {
                    const int64_t *p = pos.data();
                    std::sort(a + k, a + e, [p](const Keyed &x, const Keyed &y) { return p[x.idx] < p[y.idx]; });
                }
                k = e;
            }
        }
    }

    void write_leaf(const Work &w) {
        /* the leaf keeps its triangles in the node's current order (the root's is the input order, already in perm) */
        if (w.cur_axis < 0) return;
        const Order &p = w.K[w.cur_axis];
        for (int64_t k = 0; k < w.count; k++) perm[w.first + k] = p[k].idx;
    }

    /* builds the subtree of the node described by `w` (range and box already set in nodes[w.id]); owns and frees w */
    void build(Work *wp) {
        Work &w = *wp;
        if (w.count <= max_members || w.left_turns > max_depth) { write_leaf(w); delete wp; return; }
        double m[3];
        int64_t km[3];
        if (w.count >= PARALLEL_AXES_MIN) { /* near the root there are no sibling subtrees to keep the other threads busy */
            for (int axis = 0; axis < 3; axis++) {
#pragma omp task default(shared) firstprivate(axis)
                sweep(w.K[axis], m[axis], km[axis]);
            }
#pragma omp taskwait
        } else {
            for (int axis = 0; axis < 3; axis++) sweep(w.K[axis], m[axis], km[axis]);
        }
        /* first minimum inside an axis, strict improvement across axes x -> y -> z */
        double best_cost = std::numeric_limits<double>::infinity();
        int split_axis = -1;
        int64_t n_left = 0;
        for (int axis = 0; axis < 3; axis++)
            if (km[axis] >= 0 && m[axis] < best_cost) { best_cost = m[axis]; split_axis = axis; n_left = km[axis] + 1; }
        if (split_axis < 0 || n_left <= 0 || n_left >= w.count) { write_leaf(w); delete wp; return; }

        const Order &along = w.K[split_axis];
        for (int64_t k = 0; k < w.count; k++) {
            pos[along[k].idx] = k;
            side[along[k].idx] = k < n_left ? 0 : 1;
        }
        Work *l = new Work(), *r = new Work();
        l->first = w.first; l->count = n_left;
        r->first = w.first + n_left; r->count = w.count - n_left;
        l->left_turns = w.left_turns + 1; r->left_turns = w.left_turns;
        l->cur_axis = r->cur_axis = split_axis;
        for (int axis = 0; axis < 3; axis++) {
            if (axis == split_axis) {
                l->K[axis].assign(along.begin(), along.begin() + n_left);
                r->K[axis].assign(along.begin() + n_left, along.end());
            } else {
                partition(w.K[axis], n_left, split_axis != w.cur_axis, l->K[axis], r->K[axis]);
            }
            Order().swap(w.K[axis]); /* the parent's copy is not needed any more */
        }
        Node ln, rn;
        ln.first = l->first; ln.count = l->count;
        rn.first = r->first; rn.count = r->count;
        bounds_of(l->K[split_axis].data(), l->count, ln.lo, ln.hi);
        bounds_of(r->K[split_axis].data(), r->count, rn.lo, rn.hi);
#pragma omp critical(clv_bvh_nodes)
        {
            l->id = (int64_t)nodes.size();
            nodes.push_back(ln);
            r->id = (int64_t)nodes.size();
            nodes.push_back(rn);
            nodes[w.id].left = l->id;
            nodes[w.id].right = r->id;
        }
        const bool parallel = w.count > 8192;
        delete wp;
        if (parallel) {
#pragma omp task default(shared) firstprivate(l)
            build(l);
#pragma omp task default(shared) firstprivate(r)
            build(r);
#pragma omp taskwait
        } else {
            build(l);
            build(r);
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
    B.pos.assign((size_t)n, 0);
    B.side.assign((size_t)n, 0);
    B.nodes.reserve((size_t)std::max<int64_t>(16, n / 2));
    Node root;
    root.first = 0; root.count = n;
    typename Builder<T>::Work *w = new typename Builder<T>::Work();
    w->id = 0; w->first = 0; w->count = n; w->left_turns = 0; w->cur_axis = -1;
    for (int axis = 0; axis < 3; axis++) {
        w->K[axis].resize((size_t)n);
        for (int64_t i = 0; i < n; i++) { w->K[axis][i].key = B.centers[3 * i + axis]; w->K[axis][i].idx = i; }
    }
    if (n == 0) for (int d = 0; d < 3; d++) { root.lo[d] = INFINITY; root.hi[d] = -INFINITY; }
    else B.bounds_of(w->K[0].data(), n, root.lo, root.hi);
    B.nodes.push_back(root);
#pragma omp parallel
#pragma omp single
    {
        /* the only sorts of the whole build: the root's three orders (stable, so ties keep the input order) */
        for (int axis = 0; axis < 3; axis++) {
#pragma omp task default(shared) firstprivate(axis)
            {
                typename Builder<T>::Order buffer(n >= Builder<T>::PARALLEL_SORT_MIN ? (size_t)n : 0);
                Builder<T>::sort_keyed(w->K[axis].data(), n, buffer.data());
            }
        }
#pragma omp taskwait
        B.build(w);
    }

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
