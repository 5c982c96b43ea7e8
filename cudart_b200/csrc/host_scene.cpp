// cudart_b200/csrc/host_scene.cpp - host scene preparation (no CUDA).
//
// build_reference_bvh() rebuilds the tree of the reference's constructBVH (bvh.cpp:65-323) bit for
// bit, with the reference's undefined behaviour pinned the way SURVEY.md D1 defines it: every float3
// the reference reads uninitialised (centroid bounds bvh.cpp:133, bucket bounds :205, the four split
// bounds :231) starts at (0,0,0).  The quirks that shape the tree are kept on purpose:
//   - "surface area" is the box VOLUME dx*dy*dz (bvh.cpp:334-337, D2), NaN costs included;
//   - the lower partition [start,mid) is built (and its triangles emitted) BEFORE the upper one, yet
//     the upper one becomes child 0 / index+1 in the flattened array (bvh.cpp:302-304 under g++'s
//     right-to-left argument evaluation, D3) - spelled out here instead of relying on evaluation order;
//   - only leaves copy splitAxis, and a leaf's is 0, so every flattened node has axis 0 (D4);
//   - the permutation comes from libstdc++'s std::partition / std::nth_element / std::min_element
//     applied to a contiguous array with the reference's predicates (D26).
// tests/test_bvh_build.py compares the result node by node with the reference's own dump.
#include "host_scene.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <type_traits>

namespace crt {
namespace {

inline F3 lo3(const F3& a, const F3& b) { return {a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z}; }
inline F3 hi3(const F3& a, const F3& b) { return {a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z}; }
inline float axis_of(const F3& v, int d) { return d == 0 ? v.x : (d == 1 ? v.y : v.z); }

struct PrimBox {          // bvh.cpp:18-30
    int tri;
    F3 mn, mx, center;
};

struct BuildNode {                      // every field is written by build(): the pool is raw, uninitialised storage
    F3 mn, mx;
    int child0, child1;                 // child0 = upper partition [mid,end), child1 = lower [start,mid)
    uint32_t first, count;              // leaf payload (count == 0: interior node)
    uint32_t size;                      // nodes in this subtree (what lets flatten() place subtrees independently)
};

// Subtrees of at least this many nodes / primitives get a thread of their own in the top kForkLevels levels of a tree
// (at most 16 threads per tree).  CRT_BVH_FORK_MIN overrides the threshold: the tests force forking on the small golden
// scenes, and switch it off on large ones, to pin that no tree depends on the number of threads.
inline uint32_t fork_min() {
    static const uint32_t v = [] { const char* e = getenv("CRT_BVH_FORK_MIN"); return e ? (uint32_t)std::max(8, atoi(e)) : (1u << 14); }();
    return v;
}
constexpr int kForkLevels = 4;
// Above the forks only one thread (then two, four, eight) works on a whole range: the top levels scan their range - bounds,
// bins - in kScanChunks contiguous chunks on threads of their own and fold the chunk results IN ORDER with the same
// min / max the sequential scan uses.  `a < b ? a : b` keeps the LAST of several equal minima (that is what tells -0 from
// +0), and folding chunk results in order keeps the last chunk's last one, so for finite coordinates the result has the
// bits of the sequential scan whatever the chunking; the builders fall back to the sequential scan when a coordinate is not
// finite (a NaN does not fold that way).  The chunk count is fixed, not the machine's.
constexpr uint32_t kScanChunks = 16;
inline uint32_t scan_min() { return 8 * fork_min(); }          // ranges of >= 131 072 primitives by default
template <class F>
void chunked(uint32_t lo, uint32_t hi, F fn) {                  // fn(chunk, begin, end)
    const uint64_t n = hi - lo;
    auto bound = [=](uint32_t c) { return lo + (uint32_t)(n * c / kScanChunks); };
    std::vector<std::unique_ptr<Worker>> th;
    for (uint32_t c = 1; c < kScanChunks; c++) th.emplace_back(new Worker([=, &fn] { fn(c, bound(c), bound(c + 1)); }));
    std::exception_ptr first;
    try { fn(0u, bound(0), bound(1)); } catch (...) { first = std::current_exception(); }
    for (auto& t : th) {
        try { t->join(); } catch (...) { if (!first) first = std::current_exception(); }
    }
    if (first) std::rethrow_exception(first);
}
inline bool finite3(const F3& v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }

struct FreeDeleter { void operator()(void* p) const { std::free(p); } };
template <class T>
std::unique_ptr<T[], FreeDeleter> raw_array(size_t n) {          // uninitialised storage for trivially copyable records
    static_assert(std::is_trivially_copyable<T>::value && std::is_trivially_destructible<T>::value, "raw_array");
    void* p = std::malloc(std::max<size_t>(n, 1) * sizeof(T));
    if (!p) throw std::bad_alloc();
    return std::unique_ptr<T[], FreeDeleter>((T*)p);
}

// The tree is defined by the reference's SEQUENTIAL recursion (std::partition / std::nth_element per range, D26), but
// the two halves of a split touch disjoint ranges of the primitive array and a leaf's triangles end up exactly where
// the partitions left them (the reference emits the lower half first, so its `orderedTriangles` is the final array
// order).  Large subtrees are therefore built on their own threads ("parallel host" flavour): same
// nodes, same order, bit for bit - the node pool is pre-sized and handed out by an atomic counter, and `flatten`
// lays the nodes out afterwards from the tree structure alone.
struct Builder {
    const RefTriangle* src;
    std::vector<PrimBox> prims;
    std::unique_ptr<BuildNode[], FreeDeleter> pool;
    std::atomic<int> next_node{0};

    // "volume as area" (bvh.cpp:334-337)
    static float box_measure(const F3& mn, const F3& mx) {
        F3 d{mx.x - mn.x, mx.y - mn.y, mx.z - mn.z};
        return d.x * d.y * d.z;
    }
    // bvh.cpp:339-346
    static int widest_axis(const F3& mn, const F3& mx) {
        F3 d{mx.x - mn.x, mx.y - mn.y, mx.z - mn.z};
        if (d.x > d.y && d.x > d.z) return 0;
        if (d.y > d.z) return 1;
        return 2;
    }
    // bvh.cpp:326-332 followed by the int truncation and the 8 -> 7 clamp of :211-223
    static int bucket_of(const F3& cmin, const F3& cmax, const F3& p, int dim) {
        constexpr int kBuckets = 8;
        F3 r{p.x - cmin.x, p.y - cmin.y, p.z - cmin.z};
        if (cmax.x > cmin.x) r.x /= (cmax.x - cmin.x);
        if (cmax.y > cmin.y) r.y /= (cmax.y - cmin.y);
        if (cmax.z > cmin.z) r.z /= (cmax.z - cmin.z);
        int b = kBuckets * axis_of(r, dim);
        if (b == kBuckets) b = kBuckets - 1;
        return b;
    }

    int emit_leaf(int node, uint32_t start, uint32_t end, const F3& mn, const F3& mx) {
        BuildNode& nd = pool[node];
        nd.first = start;               // == orderedTriangles.size() at this point of the reference's recursion (bvh.cpp:150-152)
        nd.count = end - start;
        nd.child0 = nd.child1 = -1;
        nd.size = 1;
        nd.mn = mn;
        nd.mx = mx;
        return node;
    }

    void median_split(uint32_t start, uint32_t mid, uint32_t end, int dim) {
        PrimBox* base = prims.data();
        std::nth_element(base + start, base + mid, base + end,
                         [dim](const PrimBox& a, const PrimBox& b) { return axis_of(a.center, dim) < axis_of(b.center, dim); });
    }

    int build(uint32_t start, uint32_t end, int depth = 0) {
        const int node = next_node.fetch_add(1);
        F3 mn{FLT_MAX, FLT_MAX, FLT_MAX}, mx{-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (uint32_t i = start; i < end; i++) {
            mn = lo3(mn, prims[i].mn);
            mx = hi3(mx, prims[i].mx);
        }
        uint32_t n = end - start;
        if (n <= 3) return emit_leaf(node, start, end, mn, mx);

        F3 cmin{0.f, 0.f, 0.f}, cmax{0.f, 0.f, 0.f};          // pinned zero start (D1)
        for (uint32_t i = start; i < end; i++) {
            cmin = lo3(cmin, prims[i].center);
            cmax = hi3(cmax, prims[i].center);
        }
        int dim = widest_axis(cmin, cmax);
        if (axis_of(cmax, dim) == axis_of(cmin, dim)) return emit_leaf(node, start, end, mn, mx);

        uint32_t mid = (start + end) / 2;
        if (n <= 4) {
            median_split(start, mid, end, dim);
        } else {
            constexpr int kBuckets = 8;
            struct Bucket { int count = 0; F3 mn{0.f, 0.f, 0.f}, mx{0.f, 0.f, 0.f}; } buckets[kBuckets];   // zero start (D1)
            for (uint32_t i = start; i < end; i++) {
                int b = bucket_of(cmin, cmax, prims[i].center, dim);
                buckets[b].count++;
                buckets[b].mn = lo3(buckets[b].mn, prims[i].mn);
                buckets[b].mx = hi3(buckets[b].mx, prims[i].mx);
            }
            float cost[kBuckets - 1];
            for (int i = 0; i < kBuckets - 1; i++) {
                F3 amn{0.f, 0.f, 0.f}, amx{0.f, 0.f, 0.f}, bmn{0.f, 0.f, 0.f}, bmx{0.f, 0.f, 0.f};     // zero start (D1)
                int ca = 0, cb = 0;
                for (int j = 0; j <= i; j++) {
                    amn = lo3(amn, buckets[j].mn);
                    amx = hi3(amx, buckets[j].mx);
                    ca += buckets[j].count;
                }
                for (int j = i + 1; j < kBuckets; j++) {
                    bmn = lo3(bmn, buckets[j].mn);
                    bmx = hi3(bmx, buckets[j].mx);
                    cb += buckets[j].count;
                }
                cost[i] = 1.f + (ca * box_measure(amn, amx) + cb * box_measure(bmn, bmx)) / box_measure(mn, mx);
            }
            int split = (int)(std::min_element(cost, cost + (kBuckets - 1)) - cost);
            // n > 4 > maxTrisInNode, so the reference always splits here (bvh.cpp:252, D5).
            PrimBox* base = prims.data();
            PrimBox* pm = std::partition(base + start, base + end, [=](const PrimBox& t) {
                return bucket_of(cmin, cmax, t.center, dim) <= split;
            });
            mid = (uint32_t)(pm - base);
            if (mid == start || mid == end) {
                mid = (start + end) / 2;
                median_split(start, mid, end, dim);
            }
        }
        int lower, upper;
        if (depth < kForkLevels && n >= fork_min()) {
            Worker t([&, this] { lower = build(start, mid, depth + 1); });
            upper = build(mid, end, depth + 1);
            t.join();
        } else {
            lower = build(start, mid, depth + 1);      // built and emitted FIRST (D3) ...
            upper = build(mid, end, depth + 1);
        }
        BuildNode& nd = pool[node];
        nd.child0 = upper;                  // ... but stored second: child 0 is the upper partition
        nd.child1 = lower;
        nd.first = nd.count = 0;
        nd.size = 1 + pool[upper].size + pool[lower].size;
        nd.mx = hi3(pool[upper].mx, pool[lower].mx);
        nd.mn = lo3(pool[upper].mn, pool[lower].mn);
        return node;
    }

    // The reference's pre-order layout (bvh.cpp:284-306): a node, the whole subtree of child 0, then child 1, whose index
    // the node stores.  With the subtree sizes known every subtree's place is known, so large ones are written by their
    // own threads.  `out` is pre-sized by the caller.
    void flatten(int node, uint32_t my, RefNode* out, int depth = 0) const {
        for (;;) {
            const BuildNode& bn = pool[node];
            RefNode& o = out[my];
            std::memset(&o, 0, sizeof(RefNode));
            o.mn = bn.mn;
            o.mx = bn.mx;
            if (bn.count > 0) {
                o.offset = (int32_t)bn.first;
                o.nTris = (int16_t)bn.count;
                o.axis = 0;
                return;
            }
            const uint32_t second = my + 1 + pool[bn.child0].size;
            o.offset = (int32_t)second;
            if (depth < kForkLevels && bn.size >= fork_min()) {
                Worker t([&, this] { flatten(bn.child0, my + 1, out, depth + 1); });
                flatten(bn.child1, second, out, depth + 1);
                t.join();
                return;
            }
            flatten(bn.child0, my + 1, out, depth + 1);
            node = bn.child1;               // the second child: iterate instead of recursing
            my = second;
            depth++;
        }
    }
};

}  // namespace

void build_reference_bvh(RefTriangle* tris, const float* positions, uint32_t n, std::vector<RefNode>& nodes_out) {
    Builder b;
    b.src = tris;
    nodes_out.clear();
    if (n == 0) return;
    b.prims.resize(n);
    PrimBox* const pbs = b.prims.data();
    parallel_for(n, [=](uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; i++) {      // bvh.cpp:79-96 (boxes from the three vertex positions)
            const float* p = positions + 9 * (size_t)i;
            F3 a{p[0], p[1], p[2]}, c{p[3], p[4], p[5]}, d{p[6], p[7], p[8]};
            PrimBox pb;
            pb.tri = (int)i;
            pb.mn = {std::min(a.x, std::min(c.x, d.x)), std::min(a.y, std::min(c.y, d.y)), std::min(a.z, std::min(c.z, d.z))};
            pb.mx = {std::max(a.x, std::max(c.x, d.x)), std::max(a.y, std::max(c.y, d.y)), std::max(a.z, std::max(c.z, d.z))};
            pb.center = {(pb.mn.x + pb.mx.x) * 0.5f, (pb.mn.y + pb.mx.y) * 0.5f, (pb.mn.z + pb.mx.z) * 0.5f};
            pbs[i] = pb;
        }
    });
    b.pool = raw_array<BuildNode>(2 * (size_t)n);
    int root = b.build(0, n);
    nodes_out.resize((size_t)b.pool[root].size);
    b.flatten(root, 0, nodes_out.data());
    auto ordered = raw_array<RefTriangle>(n);
    RefTriangle* const ord = ordered.get();
    const PrimBox* const prims = b.prims.data();
    parallel_for(n, [=](uint32_t lo, uint32_t hi) { for (uint32_t i = lo; i < hi; i++) ord[i] = tris[prims[i].tri]; });
    parallel_for(n, [=](uint32_t lo, uint32_t hi) { std::memcpy(tris + lo, ord + lo, (size_t)(hi - lo) * sizeof(RefTriangle)); });
}

void build_light_list(RefTriangle* tris, uint32_t n, float* mat_floats, const int32_t* mat_indices, uint32_t n_mats,
                      std::vector<uint32_t>& lights, float& lights_area) {
    constexpr int kEmissive = 5;                 // refl_t EMISSIVE (material.h:8-15)
    lights.clear();
    float area = 0;
    for (uint32_t i = 0; i < n; i++) {           // scene.cpp:148-158
        uint32_t m = tris[i].materialId;
        if (m < n_mats && mat_indices[2 * m] == kEmissive) {
            tris[i].triId = i;
            lights.push_back(i);
            area += tris[i].area;                // Scene::getLightsSurfaceArea, same summation order
        }
    }
    lights_area = area;
    for (uint32_t m = 0; m < n_mats; m++)        // scene.cpp:161-164: ka *= total light area
        for (int k = 0; k < 3; k++) mat_floats[12 * m + 3 + k] = mat_floats[12 * m + 3 + k] * area;
}

// ---- wide re-flattening -----------------------------------------------------------------------
namespace {
inline float half_area(const RefNode& n) {
    float dx = n.mx.x - n.mn.x, dy = n.mx.y - n.mn.y, dz = n.mx.z - n.mn.z;
    return dx * dy + dy * dz + dz * dx;
}
}  // namespace

// ---- SAH re-clustering of the reference's leaves ----------------------------------------------------
// Which triangles a ray tests is decided by the reference's LEAF boxes alone; the interior
// structure above them only has to be conservative.  The reference builder's interior structure is poor (its
// "surface area" is a volume, its centroid bounds always include the origin: ~100 node visits per ray on
// monkeys), so the leaves are re-clustered here with a conventional binned surface-area heuristic.  The result
// is emitted in the reference's own pre-order LinearBVHNode format (leaf records copied verbatim, interior boxes
// = exact min/max unions) and then collapsed to 4-ary nodes like the reference tree would be.
namespace {
struct LeafRef { uint32_t node; F3 c; };

struct Recluster {
    // A node of the tree under construction.  Subtrees over disjoint leaf ranges are built by their own threads (nodes
    // come from a pre-sized pool through an atomic counter); place() then writes the pre-order layout - a node, its first
    // subtree, its second subtree - from the subtree sizes, large subtrees again on their own threads.  The tree does not
    // depend on the number of threads: every split is a function of its own range alone.
    struct Node { F3 mn, mx; uint32_t a, b, size, leaf; };      // a, b: children in the pool; leaf: reference node, or ~0u
    const RefNode* ref;
    std::vector<LeafRef> leaves;
    std::unique_ptr<Node[], FreeDeleter> pool;
    std::atomic<uint32_t> next_node{0};

    static float area(const F3& mn, const F3& mx) {
        float dx = mx.x - mn.x, dy = mx.y - mn.y, dz = mx.z - mn.z;
        return dx * dy + dy * dz + dz * dx;
    }
    uint32_t build(uint32_t lo, uint32_t hi, int depth = 0) {          // [lo, hi) of `leaves`; returns the pool index
        const uint32_t my = next_node.fetch_add(1);
        if (hi - lo == 1) {
            Node& nd = pool[my];
            nd.leaf = leaves[lo].node; nd.size = 1; nd.a = nd.b = 0;
            nd.mn = ref[nd.leaf].mn; nd.mx = ref[nd.leaf].mx;
            return my;
        }
        F3 mn{FLT_MAX, FLT_MAX, FLT_MAX}, mx{-FLT_MAX, -FLT_MAX, -FLT_MAX}, cmn = mn, cmx = mx;
        for (uint32_t i = lo; i < hi; i++) {
            const RefNode& l = ref[leaves[i].node];
            mn = lo3(mn, l.mn); mx = hi3(mx, l.mx);
            cmn = lo3(cmn, leaves[i].c); cmx = hi3(cmx, leaves[i].c);
        }
        constexpr int kBins = 16;
        int bestAxis = -1, bestSplit = 0;
        float bestCost = FLT_MAX;
        for (int ax = 0; ax < 3; ax++) {
            const float c0 = axis_of(cmn, ax), c1 = axis_of(cmx, ax);
            if (!(c1 > c0)) continue;
            struct Bin { F3 mn{FLT_MAX, FLT_MAX, FLT_MAX}, mx{-FLT_MAX, -FLT_MAX, -FLT_MAX}; uint32_t w = 0; } bins[kBins];
            const float scale = kBins / (c1 - c0);
            for (uint32_t i = lo; i < hi; i++) {
                int b = (int)((axis_of(leaves[i].c, ax) - c0) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                const RefNode& l = ref[leaves[i].node];
                bins[b].mn = lo3(bins[b].mn, l.mn); bins[b].mx = hi3(bins[b].mx, l.mx);
                bins[b].w += (uint32_t)(uint16_t)l.nTris;
            }
            float rightA[kBins]; uint32_t rightW[kBins];
            F3 rmn{FLT_MAX, FLT_MAX, FLT_MAX}, rmx{-FLT_MAX, -FLT_MAX, -FLT_MAX};
            uint32_t rw = 0;
            for (int b = kBins - 1; b > 0; b--) {
                rmn = lo3(rmn, bins[b].mn); rmx = hi3(rmx, bins[b].mx); rw += bins[b].w;
                rightA[b] = rw ? area(rmn, rmx) : 0.f; rightW[b] = rw;
            }
            F3 lmn{FLT_MAX, FLT_MAX, FLT_MAX}, lmx{-FLT_MAX, -FLT_MAX, -FLT_MAX};
            uint32_t lw = 0;
            for (int b = 0; b < kBins - 1; b++) {
                lmn = lo3(lmn, bins[b].mn); lmx = hi3(lmx, bins[b].mx); lw += bins[b].w;
                if (!lw || !rightW[b + 1]) continue;
                const float cost = area(lmn, lmx) * lw + rightA[b + 1] * rightW[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestSplit = b; }
            }
        }
        uint32_t mid;
        if (bestAxis >= 0) {
            const float c0 = axis_of(cmn, bestAxis), scale = kBins / (axis_of(cmx, bestAxis) - c0);
            LeafRef* m = std::partition(leaves.data() + lo, leaves.data() + hi, [&](const LeafRef& l) {
                int b = (int)((axis_of(l.c, bestAxis) - c0) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                return b <= bestSplit;
            });
            mid = (uint32_t)(m - leaves.data());
        } else {
            mid = lo;
        }
        if (mid == lo || mid == hi) mid = (lo + hi) / 2;      // coincident centroids: split by count
        uint32_t first, second;
        if (depth < kForkLevels && hi - lo >= fork_min()) {
            Worker t([&, this] { first = build(lo, mid, depth + 1); });
            second = build(mid, hi, depth + 1);
            t.join();
        } else {
            first = build(lo, mid, depth + 1);
            second = build(mid, hi, depth + 1);
        }
        Node& nd = pool[my];
        nd.mn = mn; nd.mx = mx;
        nd.a = first; nd.b = second; nd.leaf = 0xffffffffu;
        nd.size = 1 + pool[first].size + pool[second].size;
        return my;
    }
    void place(uint32_t k, uint32_t my, RefNode* out, int depth = 0) const {
        for (;;) {
            const Node& nd = pool[k];
            if (nd.leaf != 0xffffffffu) { out[my] = ref[nd.leaf]; return; }      // leaf records verbatim
            const uint32_t second = my + 1 + pool[nd.a].size;
            RefNode& o = out[my];
            std::memset(&o, 0, sizeof o);
            o.mn = nd.mn; o.mx = nd.mx;
            o.offset = (int32_t)second;
            o.nTris = 0; o.axis = 0;
            if (depth < kForkLevels && nd.size >= fork_min()) {
                Worker t([&, this] { place(nd.a, my + 1, out, depth + 1); });
                place(nd.b, second, out, depth + 1);
                t.join();
                return;
            }
            place(nd.a, my + 1, out, depth + 1);
            k = nd.b; my = second; depth++;
        }
    }
};
}  // namespace

void recluster_leaves(const RefNode* nodes, uint32_t n_nodes, std::vector<RefNode>& out) {
    Recluster rc;
    rc.ref = nodes;
    for (uint32_t k = 0; k < n_nodes; k++)
        if (nodes[k].nTris > 0)
            rc.leaves.push_back({k, {(nodes[k].mn.x + nodes[k].mx.x) * 0.5f, (nodes[k].mn.y + nodes[k].mx.y) * 0.5f,
                                     (nodes[k].mn.z + nodes[k].mx.z) * 0.5f}});
    out.clear();
    if (rc.leaves.empty()) return;
    rc.pool = raw_array<Recluster::Node>(2 * rc.leaves.size());
    const uint32_t root = rc.build(0, (uint32_t)rc.leaves.size());
    out.resize(rc.pool[root].size);
    rc.place(root, 0, out.data());
}

CullBound triangle_cull_bound(const RefTriangle& t) {
    // Forward error of Triangle::intersect's t = (e2 . (T x e1)) / (e1 . (d x e2)) in fp32 (eps = 2^-24), |d| <= 1:
    //   |dt| <= 5 eps |e1||e2| (|T| + |t|) / |det|,  and the reference only accepts |det| >= 1e-5 (geometry.cu:43),
    //   |T| <= |t| + L with L the triangle's longest edge.  Hence  t_computed >= t (1 - 2s) - s L,
    //   s = 5 eps |e1||e2| / 1e-5 ~= 0.03 |e1||e2|; we use 0.06 (2x safety) and give up beyond s = 0.25.
    auto len = [](const F3& v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); };
    const F3 e3{t.e2.x - t.e1.x, t.e2.y - t.e1.y, t.e2.z - t.e1.z};
    const float l1 = len(t.e1), l2 = len(t.e2), l3 = len(e3);
    const float s = 0.06f * l1 * l2, L = std::max(l1, std::max(l2, l3));
    if (!(s < 0.25f) || !std::isfinite(L)) return {0.f, INFINITY};
    return {1.f - 2.f * s - 1e-6f, s * L + 1e-6f * L};
}

uint32_t pack_cull_bound(CullBound b) {
    uint32_t a, bb;
    std::memcpy(&a, &b.A, 4);
    std::memcpy(&bb, &b.B, 4);
    a &= 0xffff0000u;                                           // A rounded down (positive)
    if (bb & 0xffffu) bb = (bb + 0x10000u) & 0xffff0000u;       // B rounded up
    return a | (bb >> 16);
}

WideStats flatten_wide(const RefNode* nodes, uint32_t n_nodes, const RefTriangle* tris, uint32_t n_tris, std::vector<WideNode>& out) {
    WideStats st;
    out.clear();
    if (n_nodes == 0) return st;
    // cull bound per binary node, bottom-up (children always have larger indices in the pre-order layout)
    RawVec<CullBound> cb(n_nodes);
    CullBound* const cbp = cb.data();
    parallel_for(n_nodes, [=](uint32_t lo_, uint32_t hi_) {            // the leaves (square roots per triangle): all threads
        for (uint32_t k = lo_; k < hi_; k++) {
            const RefNode& nd = nodes[k];
            if (nd.nTris <= 0) continue;
            CullBound b{1.f, 0.f};
            for (uint32_t i = 0; i < (uint32_t)(uint16_t)nd.nTris && (uint32_t)nd.offset + i < n_tris; i++) {
                const CullBound t = triangle_cull_bound(tris[(uint32_t)nd.offset + i]);
                b.A = std::min(b.A, t.A);
                b.B = std::max(b.B, t.B);
            }
            cbp[k] = b;
        }
    });
    for (uint32_t k = n_nodes; k-- > 0;) {
        const RefNode& nd = nodes[k];
        if (nd.nTris > 0) continue;
        const CullBound &l = cb[k + 1], &r = cb[(uint32_t)nd.offset];
        cb[k] = {std::min(l.A, r.A), std::max(l.B, r.B)};
    }
    struct Item { uint32_t bin; uint32_t wide; uint32_t depth; };
    std::vector<Item> todo;

    // A slot is either a binary node index or (for chunks of an oversized reference leaf) an explicit range.
    struct Slot { uint32_t bin; uint32_t first, count; bool chunk; };

    std::vector<CullBound> child_cb;             // 4 per wide node, unpacked
    auto new_wide = [&]() {
        child_cb.resize(child_cb.size() + 4, CullBound{1.f, 0.f});
        out.emplace_back();
        WideNode& w = out.back();
        std::memset(&w, 0, sizeof w);
        for (int c = 0; c < 4; c++) {
            w.child[c] = kEmptyChild;
            // an empty slot's box can never pass the slab test (and is skipped by the sentinel anyway)
            w.lox[c] = w.loy[c] = w.loz[c] = FLT_MAX;
            w.hix[c] = w.hiy[c] = w.hiz[c] = -FLT_MAX;
        }
        return (uint32_t)out.size() - 1;
    };

    out.reserve(n_nodes / 2 + 16);
    child_cb.reserve(4 * (size_t)(n_nodes / 2 + 16));
    todo.reserve(n_nodes / 2 + 16);
    uint32_t root = new_wide();
    todo.push_back({0, root, 1});
    std::vector<Slot> slots;
    // breadth-first: wide-node indices grow level by level, so "the first M nodes" are the top of the tree
    // (that prefix is what the traversal kernels stage in shared memory)
    for (size_t head = 0; head < todo.size(); head++) {
        Item it = todo[head];
        st.depth = std::max(st.depth, it.depth);
        slots.clear();
        const RefNode& top = nodes[it.bin];
        auto push_leaf_chunks = [&](uint32_t bin) {
            const RefNode& lf = nodes[bin];
            uint32_t first = (uint32_t)lf.offset, cnt = (uint32_t)(uint16_t)lf.nTris;
            for (uint32_t o = 0; o < cnt; o += kMaxLeafTris) slots.push_back({bin, first + o, std::min(kMaxLeafTris, cnt - o), true});
        };
        if (top.nTris > 0) {
            push_leaf_chunks(it.bin);
        } else {
            slots.push_back({it.bin + 1, 0, 0, false});
            slots.push_back({(uint32_t)top.offset, 0, 0, false});
        }
        // Greedy collapse: keep opening the interior slot with the largest surface area while there is room.
        while (slots.size() < 4) {
            int best = -1;
            float bestA = -1.f;
            for (size_t k = 0; k < slots.size(); k++) {
                if (slots[k].chunk || nodes[slots[k].bin].nTris > 0) continue;
                float a = half_area(nodes[slots[k].bin]);
                if (!(a <= bestA)) { bestA = a; best = (int)k; }
            }
            if (best < 0) break;
            uint32_t b = slots[best].bin;
            slots[best] = {b + 1, 0, 0, false};
            slots.push_back({(uint32_t)nodes[b].offset, 0, 0, false});
        }
        // Oversized leaves (> kMaxLeafTris, only via the cmpMax==cmpMin path of bvh.cpp:155-163) may
        // leave more than 4 slots: spill the tail into a chained wide node with the same boxes.
        uint32_t cur = it.wide;
        size_t k = 0;
        while (k < slots.size()) {
            size_t room = (slots.size() - k > 4) ? 3 : 4;
            for (size_t c = 0; c < room && k < slots.size(); c++, k++) {
                const Slot& s = slots[k];
                const RefNode& bn = nodes[s.bin];
                WideNode& w = out[cur];
                w.lox[c] = bn.mn.x; w.loy[c] = bn.mn.y; w.loz[c] = bn.mn.z;
                w.hix[c] = bn.mx.x; w.hiy[c] = bn.mx.y; w.hiz[c] = bn.mx.z;
                w.aux[c] = pack_cull_bound(cb[s.bin]);
                child_cb[4 * (size_t)cur + c] = cb[s.bin];
                if (s.chunk) {
                    w.child[c] = leaf_ref(s.first, s.count);
                    st.n_leaf_refs++;
                } else if (bn.nTris > 0 && (uint32_t)(uint16_t)bn.nTris <= kMaxLeafTris) {
                    w.child[c] = leaf_ref((uint32_t)bn.offset, (uint32_t)bn.nTris);
                    st.n_leaf_refs++;
                } else {
                    uint32_t child = new_wide();
                    out[cur].child[c] = (int32_t)child;
                    todo.push_back({s.bin, child, it.depth + 1});
                }
            }
            if (k < slots.size()) {          // chain: 4th slot points at a continuation node holding the rest
                const RefNode& bn = nodes[it.bin];
                uint32_t cont = new_wide();
                WideNode& w = out[cur];
                w.lox[3] = bn.mn.x; w.loy[3] = bn.mn.y; w.loz[3] = bn.mn.z;
                w.hix[3] = bn.mx.x; w.hiy[3] = bn.mx.y; w.hiz[3] = bn.mx.z;
                w.aux[3] = pack_cull_bound(cb[it.bin]);
                child_cb[4 * (size_t)cur + 3] = cb[it.bin];
                w.child[3] = (int32_t)cont;
                cur = cont;
            }
        }
    }
    // One cull bound per wide node (the loosest of its children): the kernel decodes it once per node instead of once
    // per child.  aux[0] (and [3]) carry it packed; aux[1], aux[2] the float terms of the dynamic pad.
    for (size_t n = 0; n < out.size(); n++) {
        WideNode& w = out[n];
        float A = 1.f, B = 0.f;
        bool any = false;
        for (int c = 0; c < 4; c++) {
            if (w.child[c] == kEmptyChild) continue;
            const CullBound& b = child_cb[4 * n + c];
            A = any ? std::min(A, b.A) : b.A;
            B = any ? std::max(B, b.B) : b.B;
            any = true;
        }
        if (!any) { A = 0.f; B = INFINITY; }
        const uint32_t packed = pack_cull_bound({A, B});
        float a1, b2;
        dynamic_pad_terms(A, B, a1, b2);
        w.aux[0] = w.aux[3] = packed;
        std::memcpy(&w.aux[1], &a1, 4);
        std::memcpy(&w.aux[2], &b2, 4);
    }
    st.n_wide = (uint32_t)out.size();
    return st;
}

// ---- tight triangle BVH ---------------------------------------------------------------------------------------------
namespace {
// A primitive of the tight tree: one triangle with its inflated box, or (group) all the triangles of a reference leaf that
// are better off behind the leaf's own exact box.
struct TightPrim { uint32_t first, count; F3 mn, mx, c; };     // members[first .. first + count)

struct TightBuilder {
    RawVec<TightPrim> prims;
    RawVec<uint32_t> members;                    // triangle indices, grouped per primitive
    std::vector<uint32_t> order;                 // tight slot -> original triangle index
    // Large subtrees are built on their own threads (disjoint primitive ranges): nodes come from a pre-sized pool
    // through an atomic counter, a leaf's slots are known from the weight in front of it, and place() writes the pool
    // out in pre-order afterwards from the subtree sizes (large subtrees again on their own threads).  The result does
    // not depend on the number of threads.
    struct Node { F3 mn, mx; uint32_t a, b, size, first, count; };   // count > 0: leaf over slots [first, first + count)
    std::unique_ptr<Node[], FreeDeleter> pool;
    std::atomic<uint32_t> next_node{0};
    uint32_t root = 0;
    // Largest leaf, in triangles.  8 where a triangle test is cheap next to a node visit (static pad); 4 in the
    // dynamic-pad trees, whose leaves are far less uniform in size (the warp pass lasts as long as its longest leaf):
    // measured monkeys 1189 -> 1226, 1M terrain 602 -> 606 Msamples/s, but cornellSphere (static) 2266 -> 2048
    // (profiles/r02n_tuning_sweep.jsonl).
    uint32_t max_leaf = 8;
    bool finite = false;                         // every primitive box is finite: the top levels may scan in chunks

    void run() {
        uint32_t total = 0;
        for (const TightPrim& p : prims) total += p.count;
        {
            std::atomic<bool> ok{true};
            const TightPrim* const pp = prims.data();
            parallel_for((uint32_t)prims.size(), [&ok, pp](uint32_t b, uint32_t e) {
                bool f = true;
                for (uint32_t i = b; i < e; i++) f = f && finite3(pp[i].mn) && finite3(pp[i].mx) && finite3(pp[i].c);
                if (!f) ok = false;
            });
            finite = ok;
        }
        pool = raw_array<Node>(2 * prims.size() + 2);
        order.assign(total, 0u);
        next_node = 0;
        root = build(0, (uint32_t)prims.size(), 0, 0);
    }
    uint32_t n_nodes() const { return pool[root].size; }
    // Pre-order layout (first child = index + 1, offset = second child) of the subtree of pool node k at out[my ...];
    // `out` is the final array: leaf offsets are shifted by slot_base (the triangles of the trees spliced in front).
    void place(uint32_t k, uint32_t my, RefNode* out, uint32_t slot_base, int depth = 0) const {
        for (;;) {
            const Node& nd = pool[k];
            RefNode& o = out[my];
            std::memset(&o, 0, sizeof o);
            o.mn = nd.mn; o.mx = nd.mx;
            if (nd.count > 0) {
                o.offset = (int32_t)(nd.first + slot_base);
                o.nTris = (int16_t)nd.count;
                return;
            }
            const uint32_t second = my + 1 + pool[nd.a].size;
            o.offset = (int32_t)second;
            if (depth < kForkLevels && nd.size >= fork_min()) {
                Worker t([&, this] { place(nd.a, my + 1, out, slot_base, depth + 1); });
                place(nd.b, second, out, slot_base, depth + 1);
                t.join();
                return;
            }
            place(nd.a, my + 1, out, slot_base, depth + 1);
            k = nd.b; my = second; depth++;
        }
    }

    static double area(const F3& mn, const F3& mx) {
        const double dx = std::max(0.0, (double)mx.x - mn.x), dy = std::max(0.0, (double)mx.y - mn.y), dz = std::max(0.0, (double)mx.z - mn.z);
        return dx * dy + dy * dz + dz * dx;
    }
    uint32_t build(uint32_t lo, uint32_t hi, uint32_t slot_base, int depth) {
        const uint32_t my = next_node.fetch_add(1);
        const uint32_t n = hi - lo;
        const bool in_chunks = finite && depth < kForkLevels && n >= scan_min();
        struct Bounds { F3 mn{FLT_MAX, FLT_MAX, FLT_MAX}, mx{-FLT_MAX, -FLT_MAX, -FLT_MAX}, cmn{FLT_MAX, FLT_MAX, FLT_MAX}, cmx{-FLT_MAX, -FLT_MAX, -FLT_MAX}; uint32_t weight = 0; };
        auto scan_bounds = [this](uint32_t a, uint32_t b) {
            Bounds r;
            for (uint32_t i = a; i < b; i++) {
                r.mn = lo3(r.mn, prims[i].mn); r.mx = hi3(r.mx, prims[i].mx);
                r.cmn = lo3(r.cmn, prims[i].c); r.cmx = hi3(r.cmx, prims[i].c);
                r.weight += prims[i].count;
            }
            return r;
        };
        Bounds all;
        if (in_chunks) {
            Bounds part[kScanChunks];
            chunked(lo, hi, [&](uint32_t c, uint32_t a, uint32_t b) { part[c] = scan_bounds(a, b); });
            for (const Bounds& r : part) {
                all.mn = lo3(all.mn, r.mn); all.mx = hi3(all.mx, r.mx);
                all.cmn = lo3(all.cmn, r.cmn); all.cmx = hi3(all.cmx, r.cmx);
                all.weight += r.weight;
            }
        } else {
            all = scan_bounds(lo, hi);
        }
        const F3 mn = all.mn, mx = all.mx, cmn = all.cmn, cmx = all.cmx;
        const uint32_t weight = all.weight;
        auto make_leaf = [&]() {
            Node& nd = pool[my];
            nd.mn = mn; nd.mx = mx;
            nd.a = nd.b = 0; nd.size = 1;
            nd.first = slot_base;
            nd.count = weight;
            uint32_t slot = slot_base;
            for (uint32_t i = lo; i < hi; i++)
                for (uint32_t m = 0; m < prims[i].count; m++) order[slot++] = members[prims[i].first + m];
            return my;
        };
        // SAH with the traversal cost measured on B200: one wide-node visit costs about as much as eight triangle
        // tests (profiles/r01s_*: a wide node covers ~3 binary nodes, hence 3 per binary node), so leaves of up to 8
        // triangles beat the reference's <= 3: cornellSphere 1649 -> 1950 Msamples/s, monkeys 775 -> 815, terrain 1M 324 -> 354.
#ifndef CRT_TIGHT_TRAV
#define CRT_TIGHT_TRAV kTightTraversalCost
#endif
        constexpr double kTrav = CRT_TIGHT_TRAV;          // tuning builds override it (profiles/r02n_*)
#ifdef CRT_TIGHT_MAXLEAF
        const uint32_t kMaxLeaf = CRT_TIGHT_MAXLEAF;
#else
        const uint32_t kMaxLeaf = max_leaf;
#endif
        const bool leaf_ok = weight <= 32767u;
        if (n == 1 || (leaf_ok && weight <= 2)) return make_leaf();
        constexpr int kBins = 16;
        // the bins of all three axes in ONE pass over the primitives (an axis without extent gets none)
        struct Bin { F3 mn{FLT_MAX, FLT_MAX, FLT_MAX}, mx{-FLT_MAX, -FLT_MAX, -FLT_MAX}; uint32_t w = 0; };
        struct Bins { Bin b[3][kBins]; };
        float c0s[3], scales[3];
        bool use[3];
        for (int ax = 0; ax < 3; ax++) {
            const float c0 = axis_of(cmn, ax), c1 = axis_of(cmx, ax);
            use[ax] = c1 > c0;
            c0s[ax] = c0;
            scales[ax] = use[ax] ? kBins / (c1 - c0) : 0.f;
        }
        auto scan_bins = [&](Bins& out_bins, uint32_t a, uint32_t b_end) {
            for (uint32_t i = a; i < b_end; i++) {
                const TightPrim& p = prims[i];
                const float cc[3] = {p.c.x, p.c.y, p.c.z};
                for (int ax = 0; ax < 3; ax++) {
                    if (!use[ax]) continue;
                    int b = (int)((cc[ax] - c0s[ax]) * scales[ax]);
                    b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                    Bin& bn = out_bins.b[ax][b];
                    bn.mn = lo3(bn.mn, p.mn); bn.mx = hi3(bn.mx, p.mx);
                    bn.w += p.count;
                }
            }
        };
        Bins binned;
        if (in_chunks) {
            std::unique_ptr<Bins[]> part(new Bins[kScanChunks]);
            chunked(lo, hi, [&](uint32_t c, uint32_t a, uint32_t b) { scan_bins(part[c], a, b); });
            for (uint32_t c = 0; c < kScanChunks; c++)
                for (int ax = 0; ax < 3; ax++)
                    for (int b = 0; b < kBins; b++) {
                        Bin& bn = binned.b[ax][b];
                        const Bin& r = part[c].b[ax][b];
                        bn.mn = lo3(bn.mn, r.mn); bn.mx = hi3(bn.mx, r.mx);
                        bn.w += r.w;
                    }
        } else {
            scan_bins(binned, lo, hi);
        }
        const auto& bins = binned.b;
        int bestAxis = -1, bestSplit = 0;
        double bestCost = DBL_MAX;
        for (int ax = 0; ax < 3; ax++) {
            if (!use[ax]) continue;
            const Bin* const bx = bins[ax];
            double rightA[kBins]; uint32_t rightW[kBins];
            F3 rmn{FLT_MAX, FLT_MAX, FLT_MAX}, rmx{-FLT_MAX, -FLT_MAX, -FLT_MAX};
            uint32_t rw = 0;
            for (int b = kBins - 1; b > 0; b--) {
                rmn = lo3(rmn, bx[b].mn); rmx = hi3(rmx, bx[b].mx); rw += bx[b].w;
                rightA[b] = rw ? area(rmn, rmx) : 0.0; rightW[b] = rw;
            }
            F3 lmn{FLT_MAX, FLT_MAX, FLT_MAX}, lmx{-FLT_MAX, -FLT_MAX, -FLT_MAX};
            uint32_t lw = 0;
            for (int b = 0; b < kBins - 1; b++) {
                lmn = lo3(lmn, bx[b].mn); lmx = hi3(lmx, bx[b].mx); lw += bx[b].w;
                if (!lw || !rightW[b + 1]) continue;
                const double cost = area(lmn, lmx) * lw + rightA[b + 1] * rightW[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestSplit = b; }
            }
        }
        // leaf when splitting does not pay and the leaf stays small
        if (leaf_ok && weight <= kMaxLeaf && (bestAxis < 0 || bestCost + kTrav * area(mn, mx) >= area(mn, mx) * (double)weight)) return make_leaf();
        uint32_t mid = lo;
        if (bestAxis >= 0) {
            const float c0 = c0s[bestAxis], scale = scales[bestAxis];
            TightPrim* m = std::partition(prims.data() + lo, prims.data() + hi, [&](const TightPrim& p) {
                int b = (int)((axis_of(p.c, bestAxis) - c0) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                return b <= bestSplit;
            });
            mid = (uint32_t)(m - prims.data());
        }
        if (mid == lo || mid == hi) {
            if (leaf_ok && weight <= kMaxLeaf) return make_leaf();
            mid = (lo + hi) / 2;
        }
        uint32_t left_weight = 0;
        for (uint32_t i = lo; i < mid; i++) left_weight += prims[i].count;
        uint32_t first, second;
        if (depth < kForkLevels && n >= fork_min()) {
            Worker t([&, this] { first = build(lo, mid, slot_base, depth + 1); });
            second = build(mid, hi, slot_base + left_weight, depth + 1);
            t.join();
        } else {
            first = build(lo, mid, slot_base, depth + 1);
            second = build(mid, hi, slot_base + left_weight, depth + 1);
        }
        Node& nd = pool[my];
        nd.mn = mn; nd.mx = mx;
        nd.a = first; nd.b = second;
        nd.first = nd.count = 0;
        nd.size = 1 + pool[first].size + pool[second].size;
        return my;
    }
};
inline float round_down(double v) { float f = (float)v; return (double)f > v ? std::nextafterf(f, -INFINITY) : f; }
inline float round_up(double v) { float f = (float)v; return (double)f < v ? std::nextafterf(f, INFINITY) : f; }

// Inflated box of one triangle for origins inside [region_lo, region_hi]; false when the fp32 test admits no useful bound.
// tight_mn / tight_mx: the box of the real triangle alone, rounded outwards (what the dynamic pad mode stores).
// pad_over_edge: the baked-in pad in units of the triangle's longest edge (what the choice of the pad mode looks at).
bool inflated_box(const RefTriangle& t, const F3& region_lo, const F3& region_hi, F3& mn_out, F3& mx_out, F3& tight_mn, F3& tight_mx,
                  double& pad_over_edge) {
    // The triangle the reference intersects is (v1, v1 + e1, v1 + e2) in REAL arithmetic on the stored fp32 values.
    const double p[3][3] = {{t.v1.x, t.v1.y, t.v1.z},
                            {(double)t.v1.x + t.e1.x, (double)t.v1.y + t.e1.y, (double)t.v1.z + t.e1.z},
                            {(double)t.v1.x + t.e2.x, (double)t.v1.y + t.e2.y, (double)t.v1.z + t.e2.z}};
    auto len = [](const F3& v) { return std::sqrt((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z); };
    const F3 e3{t.e2.x - t.e1.x, t.e2.y - t.e1.y, t.e2.z - t.e1.z};
    const double l1 = len(t.e1), l2 = len(t.e2), L = std::max(l1, std::max(l2, len(e3)));
    // |T| = |origin - v1| <= D for every origin inside the region
    double D2 = 0;
    const double lo[3] = {region_lo.x, region_lo.y, region_lo.z}, hi[3] = {region_hi.x, region_hi.y, region_hi.z};
    for (int k = 0; k < 3; k++) { const double m = std::max(std::fabs(p[0][k] - lo[k]), std::fabs(hi[k] - p[0][k])); D2 += m * m; }
    const double D = std::sqrt(D2);
    // Forward error of the accepted fp32 barycentrics (|det| >= 1e-5, |d| <= 1, eps = 2^-24):
    //   |du| |e1| <= (8 eps |T| + 7 eps |u| |e1|) |e1||e2| / 1e-5 <= 0.8 s |T| + 0.7 s L,  s = 10 eps |e1||e2| / 1e-5
    // (same s as triangle_cull_bound) and likewise for v: the REAL intersection point of the ray's line with the
    // triangle's plane lies within 1.6 s D + 1.4 s L of the triangle, measured IN the plane.  pad rounds that up to
    // 2 s (D + L); iso covers the rounding of the box test itself (3 eps D per plane) with a 50x margin.
    const double s = 0.06 * l1 * l2;
    const double pad = 2.0 * s * (D + L), iso = 1e-5 * D + 1e-3 * pad;
    if (!(s < 0.25) || !std::isfinite(pad) || !std::isfinite(iso)) return false;
    pad_over_edge = L > 0 ? pad / L : 0.0;
    // in-plane disc of radius pad: half extent pad * sqrt(1 - n_k^2) on axis k
    const double nx = (double)t.e1.y * t.e2.z - (double)t.e1.z * t.e2.y, ny = (double)t.e1.z * t.e2.x - (double)t.e1.x * t.e2.z,
                 nz = (double)t.e1.x * t.e2.y - (double)t.e1.y * t.e2.x;
    const double nl = std::sqrt(nx * nx + ny * ny + nz * nz);
    double ext[3] = {pad, pad, pad};
    if (nl > 1e-3 * l1 * l2 && nl > 0) {
        const double n[3] = {nx / nl, ny / nl, nz / nl};
        for (int k = 0; k < 3; k++) ext[k] = pad * std::sqrt(std::max(0.0, 1.0 - n[k] * n[k])) * 1.001;
    }
    double mn[3], mx[3], tmn[3], tmx[3];
    for (int k = 0; k < 3; k++) {
        tmn[k] = std::min(p[0][k], std::min(p[1][k], p[2][k]));
        tmx[k] = std::max(p[0][k], std::max(p[1][k], p[2][k]));
        mn[k] = tmn[k] - ext[k] - iso;
        mx[k] = tmx[k] + ext[k] + iso;
    }
    mn_out = {round_down(mn[0]), round_down(mn[1]), round_down(mn[2])};
    mx_out = {round_up(mx[0]), round_up(mx[1]), round_up(mx[2])};
    tight_mn = {round_down(tmn[0]), round_down(tmn[1]), round_down(tmn[2])};
    tight_mx = {round_up(tmx[0]), round_up(tmx[1]), round_up(tmx[2])};
    return true;
}
}  // namespace

void build_tight_bvh(const RefTriangle* tris, uint32_t n_tris, const RefNode* ref_nodes, uint32_t n_ref_nodes,
                     const F3& region_lo, const F3& region_hi, PadMode pad_mode, TightBvh& out) {
    out = TightBvh();
    if (n_tris == 0) return;
    out.boxes.assign(6 * (size_t)n_tris, 0.f);
    TightBuilder tb, ub;                    // bounded primitives / groups of triangles without a distance bound
    std::vector<uint32_t> leaf_of(n_tris, 0xffffffffu);
    auto set_box = [&](uint32_t i, const F3& mn, const F3& mx) {
        float* b = &out.boxes[6 * (size_t)i];
        b[0] = mn.x; b[1] = mn.y; b[2] = mn.z; b[3] = mx.x; b[4] = mx.y; b[5] = mx.z;
    };
    // the per-triangle bound is the expensive part (double precision, square roots): all host threads
    std::vector<F3> infl_mn(n_tris), infl_mx(n_tris);
    std::vector<uint8_t> infl_ok(n_tris);
    {
        std::vector<F3> tight_mn(n_tris), tight_mx(n_tris);
        std::vector<float> rel_pad(n_tris, 0.f);
        parallel_for(n_tris, [&](uint32_t b, uint32_t e) {
            for (uint32_t i = b; i < e; i++) {
                double rp = 0;
                infl_ok[i] = inflated_box(tris[i], region_lo, region_hi, infl_mn[i], infl_mx[i], tight_mn[i], tight_mx[i], rp) ? 1 : 0;
                rel_pad[i] = (float)std::min(rp, 1e6);
            }
        });
        if (pad_mode == kPadAuto) {
            // The baked-in pad is free at run time; the per-visit pad costs ~16 instructions per child and takes the
            // loosest (a1, b2) of a node's children, so one wall-sized triangle pads its whole node.  It pays when the
            // fixed pad is comparable to the triangles themselves.  Mean pad / longest edge and the measured outcome
            // (profiles/r01z_pad_ab.txt): cornellSphere 0.02 (1980 static vs 882 dynamic Msamples/s), cornell 0.23,
            // monkeys 0.44 (803 vs 932), terrain 999 700: 0.75 (343 vs 447), terrain 131 074: 2.1 (232 vs 320).
            // (Moving the large-s primitives into a subtree of their own, like the unbounded ones, was measured too:
            // every ray then walks two overlapping trees - monkeys 932 -> 712, cornellSphere 882 -> 915; not kept.)
            double sum = 0, cnt = 0;
            for (uint32_t i = 0; i < n_tris; i++)
                if (infl_ok[i]) { sum += rel_pad[i]; cnt += 1; }
            pad_mode = (cnt > 0 && sum / cnt > 0.3) ? kPadDynamic : kPadStatic;
        }
        out.dynamic_pad = pad_mode == kPadDynamic;
        if (out.dynamic_pad) { infl_mn.swap(tight_mn); infl_mx.swap(tight_mx); }
        tb.max_leaf = ub.max_leaf = out.dynamic_pad ? 4 : 8;
    }
    // One pass over the reference's leaves decides, per leaf, between individually boxed triangles and the leaf kept whole.
    // The leaves are independent, so contiguous ranges of them are processed on the host's threads into lists of their own
    // that are concatenated in leaf order afterwards: the same primitive lists as one sequential pass.
    struct LeafChunk {
        std::vector<TightPrim> tb_prims, ub_prims;           // `first` relative to this chunk's member lists
        std::vector<uint32_t> tb_members, ub_members;
        uint32_t n_split = 0, n_unbounded = 0;
    };
    const uint32_t n_chunks = n_ref_nodes < (1u << 15) ? 1u : host_threads();
    std::vector<LeafChunk> chunks(n_chunks);
    uint32_t claimed_twice = 0;                               // set (atomically) when two leaves claim one triangle
    uint32_t* const leaf_idx = leaf_of.data();
    auto leaf_pass = [&](LeafChunk& ch, uint32_t k_begin, uint32_t k_end) {
        auto add = [](std::vector<TightPrim>& prims, std::vector<uint32_t>& members, const uint32_t* m, size_t n, const F3& mn, const F3& mx) {
            TightPrim pr;
            pr.first = (uint32_t)members.size(); pr.count = (uint32_t)n;
            pr.mn = mn; pr.mx = mx;
            pr.c = {(mn.x + mx.x) * 0.5f, (mn.y + mx.y) * 0.5f, (mn.z + mx.z) * 0.5f};
            members.insert(members.end(), m, m + n);
            prims.push_back(pr);
        };
        std::vector<uint32_t> bounded, unbounded;
        std::vector<F3> bmn, bmx;
        for (uint32_t k = k_begin; k < k_end; k++) {
            const RefNode& lf = ref_nodes[k];
            if (lf.nTris <= 0) continue;
            // The leaf's own box in min/max order (the reference's slab test sorts each axis, renderer.cu:339-343); a box
            // that is not finite is replaced by the region: the gate in the kernel stays the exact reference box either way.
            F3 gmn = lo3(lf.mn, lf.mx), gmx = hi3(lf.mn, lf.mx);
            const float gb[6] = {gmn.x, gmn.y, gmn.z, gmx.x, gmx.y, gmx.z};
            bool finite = true;
            for (float f : gb) finite = finite && std::isfinite(f);
            if (!finite) { gmn = region_lo; gmx = region_hi; }
            bounded.clear(); unbounded.clear(); bmn.clear(); bmx.clear();
            double split_cost = 0;
            bool duplicate = false;
            for (uint32_t i = 0; i < (uint32_t)(uint16_t)lf.nTris && (uint32_t)lf.offset + i < n_tris; i++) {
                const uint32_t idx = (uint32_t)lf.offset + i;
                // a triangle claimed by two leaves (not a tree the reference builds) is reachable through either box: the
                // gate could not say which, so such input keeps the reference-leaf tree only
                if (__atomic_exchange_n(&leaf_idx[idx], k, __ATOMIC_RELAXED) != 0xffffffffu) { duplicate = true; break; }
                const F3 mn = infl_mn[idx], mx = infl_mx[idx];
                if (infl_ok[idx]) {
                    bounded.push_back(idx); bmn.push_back(mn); bmx.push_back(mx);
                    split_cost += TightBuilder::area(mn, mx);
                } else {
                    unbounded.push_back(idx);
                    ch.n_unbounded++;
                }
            }
            if (duplicate) { __atomic_store_n(&claimed_twice, 1u, __ATOMIC_RELAXED); return; }
            if (!unbounded.empty()) {
                add(ch.ub_prims, ch.ub_members, unbounded.data(), unbounded.size(), gmn, gmx);
                for (uint32_t i : unbounded) set_box(i, gmn, gmx);
            }
            if (bounded.empty()) continue;
            // Individually inflated triangles, or the whole leaf behind its exact box?  SAH: sum of the inflated areas (plus
            // the extra nodes and gate tests, hence the 0.7) against the leaf area times its triangle count.
            const double group_cost = TightBuilder::area(gmn, gmx) * (double)bounded.size();
            if (split_cost < 0.7 * group_cost) {
                for (size_t j = 0; j < bounded.size(); j++) {
                    add(ch.tb_prims, ch.tb_members, &bounded[j], 1, bmn[j], bmx[j]);
                    set_box(bounded[j], bmn[j], bmx[j]);
                }
                ch.n_split += (uint32_t)bounded.size();
            } else {
                add(ch.tb_prims, ch.tb_members, bounded.data(), bounded.size(), gmn, gmx);
                for (uint32_t i : bounded) set_box(i, gmn, gmx);
            }
        }
    };
    if (n_chunks == 1) {
        leaf_pass(chunks[0], 0, n_ref_nodes);
    } else {
        std::vector<std::unique_ptr<Worker>> th;
        for (uint32_t p = 0; p < n_chunks; p++) {
            const uint32_t kb = (uint32_t)((uint64_t)n_ref_nodes * p / n_chunks), ke = (uint32_t)((uint64_t)n_ref_nodes * (p + 1) / n_chunks);
            th.emplace_back(new Worker([&, p, kb, ke] { leaf_pass(chunks[p], kb, ke); }));
        }
        std::exception_ptr first;
        for (auto& t : th) {
            try { t->join(); } catch (...) { if (!first) first = std::current_exception(); }
        }
        if (first) std::rethrow_exception(first);
    }
    if (__atomic_load_n(&claimed_twice, __ATOMIC_RELAXED)) { out = TightBvh(); return; }
    {   // concatenate the chunks' lists in leaf order; every chunk copies its own part to its known place
        std::vector<size_t> tbp(n_chunks + 1, 0), tbm(n_chunks + 1, 0), ubp(n_chunks + 1, 0), ubm(n_chunks + 1, 0);
        for (uint32_t p = 0; p < n_chunks; p++) {
            const LeafChunk& ch = chunks[p];
            tbp[p + 1] = tbp[p] + ch.tb_prims.size(); tbm[p + 1] = tbm[p] + ch.tb_members.size();
            ubp[p + 1] = ubp[p] + ch.ub_prims.size(); ubm[p + 1] = ubm[p] + ch.ub_members.size();
            out.n_split += ch.n_split;
            out.n_unbounded += ch.n_unbounded;
        }
        tb.prims.resize(tbp[n_chunks]); tb.members.resize(tbm[n_chunks]);
        ub.prims.resize(ubp[n_chunks]); ub.members.resize(ubm[n_chunks]);
        auto splice = [](TightBuilder& bd, const std::vector<TightPrim>& prims, const std::vector<uint32_t>& members, size_t prim_at, size_t member_at) {
            TightPrim* dst = bd.prims.data() + prim_at;
            for (size_t i = 0; i < prims.size(); i++) { dst[i] = prims[i]; dst[i].first += (uint32_t)member_at; }
            if (!members.empty()) std::memcpy(bd.members.data() + member_at, members.data(), members.size() * sizeof(uint32_t));
        };
        auto splice_chunk = [&](uint32_t p) {
            splice(tb, chunks[p].tb_prims, chunks[p].tb_members, tbp[p], tbm[p]);
            splice(ub, chunks[p].ub_prims, chunks[p].ub_members, ubp[p], ubm[p]);
        };
        if (n_chunks == 1) {
            splice_chunk(0);
        } else {
            std::vector<std::unique_ptr<Worker>> th;
            for (uint32_t p = 0; p < n_chunks; p++) th.emplace_back(new Worker([&, p] { splice_chunk(p); }));
            for (auto& t : th) t->join();
        }
    }
    if (tb.prims.empty() && ub.prims.empty()) { out = TightBvh(); return; }
    // One tree: root = (subtree of the groups without a distance bound, subtree of everything else).  Keeping the former
    // apart keeps their (A = 0, B = inf) cull bound out of every other node's bound (flatten_wide takes the loosest).
    std::vector<RefNode>& nodes = out.nodes;
    std::vector<uint32_t> order;
    auto append = [&](TightBuilder& bd) {          // build bd's tree and splice it behind `nodes` / `order`
        bd.run();
        const uint32_t nbase = (uint32_t)nodes.size(), obase = (uint32_t)order.size();
        nodes.resize((size_t)nbase + bd.n_nodes());
        bd.place(bd.root, nbase, nodes.data(), obase);
        order.insert(order.end(), bd.order.begin(), bd.order.end());
        return nbase;
    };
    if (!tb.prims.empty() && !ub.prims.empty()) {
        nodes.emplace_back();
        append(ub);
        const uint32_t second = append(tb);
        RefNode& rt = nodes[0];
        std::memset(&rt, 0, sizeof rt);
        rt.mn = lo3(nodes[1].mn, nodes[second].mn); rt.mx = hi3(nodes[1].mx, nodes[second].mx);
        rt.offset = (int32_t)second;
    } else {
        append(tb.prims.empty() ? ub : tb);
    }
    const uint32_t n_used = (uint32_t)order.size();          // triangles in no reference leaf are never reached: left out
    out.records.resize(n_used);
    out.tris.resize(n_used);
    TightRecord* const recs = out.records.data();
    RefTriangle* const ttris = out.tris.data();
    const uint32_t* const ord = order.data();
    parallel_for(n_used, [=](uint32_t b, uint32_t e) {
        for (uint32_t slot = b; slot < e; slot++) {
            const uint32_t i = ord[slot];
            const RefTriangle& t = tris[i];
            TightRecord& r = recs[slot];
            std::memset(&r, 0, sizeof r);
            r.v1[0] = t.v1.x; r.v1[1] = t.v1.y; r.v1[2] = t.v1.z;
            r.e1[0] = t.e1.x; r.e1[1] = t.e1.y; r.e1[2] = t.e1.z;
            r.e2[0] = t.e2.x; r.e2[1] = t.e2.y; r.e2[2] = t.e2.z;
            r.materialId = t.materialId;
            r.origIdx = i;
            const RefNode& lf = ref_nodes[leaf_idx[i]];
            r.leafFirst = (uint32_t)lf.offset;
            r.gateLo[0] = lf.mn.x; r.gateLo[1] = lf.mn.y; r.gateLo[2] = lf.mn.z;
            r.gateHiX = lf.mx.x; r.gateHiY = lf.mx.y; r.gateHiZ = lf.mx.z;
            ttris[slot] = t;
        }
    });
}

SahCost sah_cost(const RefNode* nodes, uint32_t n_nodes, const F3& clip_lo, const F3& clip_hi) {
    SahCost c;
    if (n_nodes == 0) return c;
    auto area = [&](const F3& mn, const F3& mx) {
        const double dx = std::max(0.0, (double)std::min(mx.x, clip_hi.x) - std::max(mn.x, clip_lo.x)),
                     dy = std::max(0.0, (double)std::min(mx.y, clip_hi.y) - std::max(mn.y, clip_lo.y)),
                     dz = std::max(0.0, (double)std::min(mx.z, clip_hi.z) - std::max(mn.z, clip_lo.z));
        return dx * dy + dy * dz + dz * dx;
    };
    const double root = area(clip_lo, clip_hi);
    if (!(root > 0) || !std::isfinite(root)) return c;
    for (uint32_t k = 0; k < n_nodes; k++) {
        const double a = area(lo3(nodes[k].mn, nodes[k].mx), hi3(nodes[k].mn, nodes[k].mx)) / root;
        if (!std::isfinite(a)) continue;
        if (nodes[k].nTris > 0) c.tris += a * (double)(uint16_t)nodes[k].nTris;
        else c.nodes += a;
    }
    return c;
}

void make_tri_records(const RefTriangle* tris, uint32_t n, const RefNode* nodes, uint32_t n_nodes,
                      RawVec<TriIsect>& isect, RawVec<TriShade>& shade) {
    isect.resize(n);
    shade.resize(n);
    TriIsect* const isect_p = isect.data();
    TriShade* const shade_p = shade.data();
    parallel_for(n, [=](uint32_t lo_, uint32_t hi_) {
    for (uint32_t i = lo_; i < hi_; i++) {
        const RefTriangle& t = tris[i];
        TriIsect& a = isect_p[i];
        a.v1[0] = t.v1.x; a.v1[1] = t.v1.y; a.v1[2] = t.v1.z;
        a.e1[0] = t.e1.x; a.e1[1] = t.e1.y; a.e1[2] = t.e1.z;
        a.e2[0] = t.e2.x; a.e2[1] = t.e2.y; a.e2[2] = t.e2.z;
        a.materialId = t.materialId;
        a.index = i;
        a.leafFirst = 0xffffffffu;          // in no reference leaf (never reached by the reference traversal) until shown otherwise
        TriShade& s = shade_p[i];
        std::memset(&s, 0, sizeof s);
        s.n1[0] = t.n1.x; s.n1[1] = t.n1.y; s.n1[2] = t.n1.z;
        s.n2[0] = t.n2.x; s.n2[1] = t.n2.y; s.n2[2] = t.n2.z;
        s.n3[0] = t.n3.x; s.n3[1] = t.n3.y; s.n3[2] = t.n3.z;
        s.uv1[0] = t.uv1.x; s.uv1[1] = t.uv1.y;
        s.uv2[0] = t.uv2.x; s.uv2[1] = t.uv2.y;
        s.uv3[0] = t.uv3.x; s.uv3[1] = t.uv3.y;
        s.area = t.area;
        s.materialId = t.materialId;
    }
    });
    // The tie rule needs, per triangle, the firstTriOffset of the REFERENCE leaf holding it.
    for (uint32_t k = 0; k < n_nodes; k++) {
        if (nodes[k].nTris <= 0) continue;
        uint32_t first = (uint32_t)nodes[k].offset, cnt = (uint32_t)(uint16_t)nodes[k].nTris;
        for (uint32_t i = 0; i < cnt && first + i < n; i++) isect[first + i].leafFirst = first;
    }
}

}  // namespace crt
