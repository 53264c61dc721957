/* clv_accel_build.h — host side of clv_accel_create: reference Box[]/Triangle[] (bvh.py:329-389 encoding) ->
 * compact traversal layout described at clv_device.cuh `AccelView`.  Plain C++, no CUDA: included by clive_b200.cu
 * (product) and by tests/hostemu (CPU check of the kernel logic).
 */
#ifndef CLV_ACCEL_BUILD_H
#define CLV_ACCEL_BUILD_H

#include <cstdio>
#include <cstring>
#include <utility>
#include <vector>

#include "clv_device.cuh"

namespace clv {

struct NodeRec { V4 a, b, c, d; };
static_assert(sizeof(NodeRec) == 64, "NodeRec");

/* 4-wide node of the ORDERED flavour: 8 x V4 = [minx x4][miny x4][minz x4][maxx x4][maxy x4][maxz x4][refs x4][-],
 * empty slots have ref -1.  Obtained by collapsing the binary nodes (a slot that is an inner node is replaced by its two
 * children, largest box first, until four slots are used): every box kept is one of the reference's boxes, the boxes
 * skipped contain the ones kept, so a ray that passes a kept box also passes the skipped one (the slab test is monotone
 * in the box bounds) and the set of triangles a ray can reach is unchanged. */
struct WideRec { V4 q[8]; };
static_assert(sizeof(WideRec) == 128, "WideRec");

struct HostAccel {
    std::vector<NodeRec> nodes;       /* inner nodes; reference inner boxes first (in reference order), then synthetic */
    std::vector<WideRec> wide;        /* collapsed 4-wide nodes */
    int32_t wide_root_ref;
    std::vector<V4> tris;             /* 3 per triangle, reference order */
    std::vector<V4> shade;            /* 4 per triangle: shading normals, face normal, material and flags (AccelView) */
    std::vector<int32_t> camera_tris; /* indices of is_camera triangles */
    int32_t root_ref;
    int32_t binary_depth, wide_depth; /* inner nodes on the longest root-to-leaf path: what bounds the traversal stacks */
};

/* inner nodes on the longest root-to-leaf path (iterative: a degenerate tree can be 10^5 levels deep) */
template <class CHILDREN>
inline int32_t tree_depth(int32_t root_ref, CHILDREN children) {
    if (root_ref < 0) return 0;
    std::vector<std::pair<int32_t, int32_t>> stack;
    stack.push_back({root_ref, 1});
    int32_t deepest = 0;
    while (!stack.empty()) {
        auto [ref, depth] = stack.back();
        stack.pop_back();
        if (depth > deepest) deepest = depth;
        int32_t kids[4];
        int n = children(ref, kids);
        for (int k = 0; k < n; k++)
            if (kids[k] >= 0) stack.push_back({kids[k], depth + 1});
    }
    return deepest;
}

inline NodeRec pack_node(const Box &l, const Box &r, int32_t lref, int32_t rref) {
    NodeRec n;
    n.a = v4(l.min.x, l.min.y, l.min.z, l.max.x);
    n.b = v4(l.max.y, l.max.z, r.min.x, r.min.y);
    n.c = v4(r.min.z, r.max.x, r.max.y, r.max.z);
    float lf, rf;
    memcpy(&lf, &lref, 4);
    memcpy(&rf, &rref, 4);
    n.d = v4(lf, rf, 0.0f, 0.0f);
    return n;
}

/* a leaf range as a ref; ranges longer than 15 become a chain of synthetic inner nodes that carry the leaf's own box
 * on both sides (the slab test of the leaf is simply repeated), first half on the side the reference order visits
 * first (the right one) */
inline int32_t leaf_ref(std::vector<NodeRec> &nodes, const Box &bx, int32_t first, int32_t last) {
    int32_t count = last - first;
    if (count <= 15) return ~((first << 4) | count);
    int32_t mid = first + count / 2;
    size_t me = nodes.size();
    nodes.push_back(NodeRec());
    int32_t right = leaf_ref(nodes, bx, first, mid);
    int32_t left = leaf_ref(nodes, bx, mid, last);
    nodes[me] = pack_node(bx, bx, left, right);
    return (int32_t)me;
}

struct WideSlot { float mn[3], mx[3]; int32_t ref; };
inline void slots_of(const NodeRec &n, WideSlot out[2]) {
    out[0].mn[0] = n.a.x; out[0].mn[1] = n.a.y; out[0].mn[2] = n.a.z; out[0].mx[0] = n.a.w; out[0].mx[1] = n.b.x; out[0].mx[2] = n.b.y;
    out[1].mn[0] = n.b.z; out[1].mn[1] = n.b.w; out[1].mn[2] = n.c.x; out[1].mx[0] = n.c.y; out[1].mx[1] = n.c.z; out[1].mx[2] = n.c.w;
    memcpy(&out[0].ref, &n.d.x, 4);
    memcpy(&out[1].ref, &n.d.y, 4);
}
inline int32_t wide_of(HostAccel &H, int32_t bin_ref) {
    if (bin_ref < 0) return bin_ref; /* leaf refs are shared */
    WideSlot slots[4];
    int n = 2;
    slots_of(H.nodes[bin_ref], slots);
    while (n < 4) {
        int pick = -1;
        float best = -1.0f;
        for (int k = 0; k < n; k++) {
            if (slots[k].ref < 0) continue;
            float sx = slots[k].mx[0] - slots[k].mn[0], sy = slots[k].mx[1] - slots[k].mn[1], sz = slots[k].mx[2] - slots[k].mn[2];
            float area = sx * sy + sy * sz + sz * sx;
            if (!(area <= best)) { best = area; pick = k; }
        }
        if (pick < 0) break;
        WideSlot kids[2];
        slots_of(H.nodes[slots[pick].ref], kids);
        slots[pick] = kids[0];
        slots[n++] = kids[1];
    }
    size_t me = H.wide.size();
    H.wide.push_back(WideRec());
    int32_t refs[4] = {-1, -1, -1, -1};
    for (int k = 0; k < n; k++) refs[k] = wide_of(H, slots[k].ref);
    WideRec w;
    for (int c = 0; c < 3; c++) {
        float mn[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0};
        for (int k = 0; k < n; k++) { mn[k] = slots[k].mn[c]; mx[k] = slots[k].mx[c]; }
        w.q[c] = v4(mn[0], mn[1], mn[2], mn[3]);
        w.q[3 + c] = v4(mx[0], mx[1], mx[2], mx[3]);
    }
    float rf[4];
    memcpy(rf, refs, 16);
    w.q[6] = v4(rf[0], rf[1], rf[2], rf[3]);
    w.q[7] = v4(0.0f, 0.0f, 0.0f, 0.0f);
    H.wide[me] = w;
    return (int32_t)me;
}
/* Order of the binary node records in memory.  0: the reference's breadth-first box order (bvh.py:329-389), synthetic
 * leaf-chain nodes last — the two children of a node are neighbours, a node and its grandchildren are not.  1: depth-first
 * pre-order, the subtree the reference visits LAST (left) directly behind its parent: a subtree is one contiguous range,
 * so a search that descends stays inside ever smaller address ranges (sector / page / TLB locality).  Refs are renumbered;
 * boxes, leaves and therefore every search result are unchanged. */
#ifndef CLV_BINARY_NODE_ORDER
#define CLV_BINARY_NODE_ORDER 0
#endif
inline void renumber_depth_first(HostAccel &H) {
    if (H.root_ref < 0) return;
    const size_t n = H.nodes.size();
    std::vector<int32_t> new_index(n, -1);
    std::vector<int32_t> stack;
    stack.push_back(H.root_ref);
    int32_t next = 0;
    while (!stack.empty()) {
        const int32_t ref = stack.back();
        stack.pop_back();
        new_index[ref] = next++;
        int32_t l, r;
        memcpy(&l, &H.nodes[ref].d.x, 4);
        memcpy(&r, &H.nodes[ref].d.y, 4);
        if (r >= 0) stack.push_back(r);
        if (l >= 0) stack.push_back(l); /* popped first: the left subtree follows its parent */
    }
    std::vector<NodeRec> out((size_t)next);
    for (size_t i = 0; i < n; i++) {
        if (new_index[i] < 0) continue; /* unreachable (never the case for a well-formed tree) */
        NodeRec rec = H.nodes[i];
        int32_t kids[2];
        memcpy(&kids[0], &rec.d.x, 4);
        memcpy(&kids[1], &rec.d.y, 4);
        for (int k = 0; k < 2; k++)
            if (kids[k] >= 0) kids[k] = new_index[kids[k]];
        memcpy(&rec.d.x, &kids[0], 4);
        memcpy(&rec.d.y, &kids[1], 4);
        out[new_index[i]] = rec;
    }
    H.root_ref = new_index[H.root_ref];
    H.nodes.swap(out);
}

inline void build_wide(HostAccel &H) {
    H.wide.clear();
    H.wide_root_ref = wide_of(H, H.root_ref);
}

/* returns NULL on success, else a description of what is wrong with the input */
inline const char *build_host_accel(const Box *boxes, int32_t n_boxes, const Triangle *tris, int32_t n_tris,
                                    HostAccel &H) {
    static thread_local char why[160];
    if (n_boxes <= 0 || n_tris <= 0) return "empty scene";
    if (n_tris >= (1 << 27)) return "more than 2^27 triangles";
    /* validate the encoding of bvh.py:345-349 before trusting indices on the device */
    for (int32_t b = 0; b < n_boxes; b++) {
        const Box &bx = boxes[b];
        if (bx.right == 0) {
            if (!(bx.left > b && bx.left + 1 < n_boxes)) {
                snprintf(why, sizeof why, "box %d: bad child index %d", b, bx.left);
                return why;
            }
        } else if (!(bx.left >= 0 && bx.left < bx.right && bx.right <= n_tris)) {
            snprintf(why, sizeof why, "box %d: bad triangle range [%d,%d)", b, bx.left, bx.right);
            return why;
        }
    }
    std::vector<int32_t> inner_index(n_boxes, -1);
    int32_t n_inner = 0;
    for (int32_t b = 0; b < n_boxes; b++)
        if (boxes[b].right == 0) inner_index[b] = n_inner++;
    H.nodes.assign(n_inner, NodeRec());
    auto ref_of = [&](int32_t b) -> int32_t {
        const Box &bx = boxes[b];
        return bx.right == 0 ? inner_index[b] : leaf_ref(H.nodes, bx, bx.left, bx.right);
    };
    for (int32_t b = 0; b < n_boxes; b++) {
        const Box &bx = boxes[b];
        if (bx.right != 0) continue;
        int32_t l = bx.left, r = bx.left + 1;
        int32_t lref = ref_of(l), rref = ref_of(r);
        H.nodes[inner_index[b]] = pack_node(boxes[l], boxes[r], lref, rref);
    }
    H.root_ref = ref_of(0);

    /* reference visiting rank of every triangle: DFS from box 0, right child popped first (trace.metal:158-161) */
    std::vector<int32_t> rank(n_tris, 0x7fffffff);
    {
        std::vector<int32_t> stack;
        stack.push_back(0);
        int32_t next = 0;
        while (!stack.empty()) {
            int32_t b = stack.back();
            stack.pop_back();
            const Box &bx = boxes[b];
            if (bx.right == 0) { stack.push_back(bx.left); stack.push_back(bx.left + 1); }
            else for (int32_t i = bx.left; i < bx.right; i++) rank[i] = next++;
        }
    }
#if CLV_BINARY_NODE_ORDER == 1
    renumber_depth_first(H);
#endif
    build_wide(H);
    H.binary_depth = tree_depth(H.root_ref, [&](int32_t ref, int32_t kids[4]) {
        memcpy(&kids[0], &H.nodes[ref].d.x, 4);
        memcpy(&kids[1], &H.nodes[ref].d.y, 4);
        return 2;
    });
    H.wide_depth = tree_depth(H.wide_root_ref, [&](int32_t ref, int32_t kids[4]) {
        memcpy(kids, &H.wide[ref].q[6], 16); /* leaves and empty slots are negative */
        return 4;
    });
    /* the traversal stacks hold one deferred sibling per binary level / three per 4-wide level (clv_device.cuh): a deeper
     * tree is refused here, loudly, instead of silently dropping pushes in the kernels.  (The reference's own 64-entry
     * stack silently aborts the search in that case, trace.metal:149.) */
    if (H.binary_depth > CLV_MAX_BINARY_DEPTH || H.wide_depth > CLV_MAX_WIDE_DEPTH) {
        snprintf(why, sizeof why, "BVH too deep for the traversal stacks: %d binary levels (max %d), %d 4-wide levels (max %d)",
                 H.binary_depth, CLV_MAX_BINARY_DEPTH, H.wide_depth, CLV_MAX_WIDE_DEPTH);
        return why;
    }
    H.tris.resize((size_t)n_tris * 3);
    H.shade.resize((size_t)n_tris * 4);
    H.camera_tris.clear();
    for (int32_t i = 0; i < n_tris; i++) {
        const Triangle &t = tris[i];
        float rk, fm, fl, fc;
        memcpy(&rk, &rank[i], 4);
        memcpy(&fm, &t.material, 4);
        memcpy(&fl, &t.is_light, 4);
        memcpy(&fc, &t.is_camera, 4);
        H.shade[4 * (size_t)i + 0] = v4(t.n0.x, t.n0.y, t.n0.z, fm);
        H.shade[4 * (size_t)i + 1] = v4(t.n1.x, t.n1.y, t.n1.z, fl);
        H.shade[4 * (size_t)i + 2] = v4(t.n2.x, t.n2.y, t.n2.z, fc);
        H.shade[4 * (size_t)i + 3] = v4(t.normal.x, t.normal.y, t.normal.z, 0.0f);
        H.tris[3 * (size_t)i + 0] = v4(t.v0.x, t.v0.y, t.v0.z, rk);
        H.tris[3 * (size_t)i + 1] = v4(t.v1.x - t.v0.x, t.v1.y - t.v0.y, t.v1.z - t.v0.z, 0.0f);
        H.tris[3 * (size_t)i + 2] = v4(t.v2.x - t.v0.x, t.v2.y - t.v0.y, t.v2.z - t.v0.z, 0.0f);
        if (t.is_camera) H.camera_tris.push_back(i);
    }
    return nullptr;
}

} /* namespace clv */
#endif
