// cudart_b200/csrc/host_scene.h - host-side scene preparation for the B200 path tracer.
//
//  * RefTriangle / RefNode / RefCamera: the reference's POD layouts (geometry.h:30-51, bvh.h:18-27,
//    camera.h:9-22), the currency of the C ABI.
//  * build_reference_bvh(): the reference's BVH (bvh.cpp:65-323) rebuilt bit for bit.
//  * flatten_wide(): that binary tree re-flattened into the 128-byte 4-ary node the kernels read.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <exception>
#include <memory>
#include <new>
#include <type_traits>
#include <thread>
#include <utility>
#include <vector>

namespace crt {

// A host thread whose body cannot take the process down: an exception thrown inside is carried to join() and rethrown
// on the joining thread (where the C ABI's guard turns it into an error code).  The destructor joins.
class Worker {
    std::thread t_;
    std::exception_ptr err_;
public:
    Worker() = default;
    template <class F>
    explicit Worker(F f) {
        t_ = std::thread([this, f]() mutable {
            try { f(); } catch (...) { err_ = std::current_exception(); }
        });
    }
    Worker(const Worker&) = delete;
    Worker& operator=(const Worker&) = delete;
    bool joinable() const { return t_.joinable(); }
    void join() {
        if (t_.joinable()) t_.join();
        if (err_) { std::exception_ptr e = err_; err_ = nullptr; std::rethrow_exception(e); }
    }
    ~Worker() { if (t_.joinable()) t_.join(); }
};

// Host threads the builders and loaders use for one job: the machine's, at most 16; CRT_HOST_THREADS overrides it (1 .. 64).
// Nothing they produce depends on this number (tests/test_bvh_build.py runs the derivation with 1, 3 and 16).
inline uint32_t host_threads() {
    static const uint32_t v = [] {
        const char* e = getenv("CRT_HOST_THREADS");
        const int forced = e ? atoi(e) : 0;
        return forced > 0 ? (uint32_t)std::min(forced, 64) : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    }();
    return v;
}

// fn(begin, end) over [0, n) in contiguous chunks on the host's threads (the chunks write disjoint outputs)
template <class F>
inline void parallel_for(uint32_t n, F fn) {
    const uint32_t hw = host_threads();
    const uint32_t parts = n < (1u << 15) ? 1u : hw;
    if (parts == 1) { fn(0u, n); return; }
    std::vector<std::unique_ptr<Worker>> th;
    for (uint32_t p = 0; p < parts; p++) {
        const uint32_t b = (uint32_t)((uint64_t)n * p / parts), e = (uint32_t)((uint64_t)n * (p + 1) / parts);
        th.emplace_back(new Worker([=, &fn] { fn(b, e); }));
    }
    std::exception_ptr first;
    for (auto& t : th) {
        try { t->join(); } catch (...) { if (!first) first = std::current_exception(); }
    }
    if (first) std::rethrow_exception(first);
}

// std::vector whose resize() leaves trivially constructible elements uninitialised: the large record arrays below are
// sized once and then filled completely by the host's threads - value-initialising 200 MB on one thread first (and
// taking its page faults there) cost more than filling them.
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new ((void*)p) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
template <class T> using RawVec = std::vector<T, NoInitAlloc<T>>;

struct F3 { float x, y, z; };
struct F2 { float x, y; };

struct RefTriangle {          // 128 B, __attribute__((aligned(128))) in the reference
    uint32_t triId, materialId;
    F3 v1, e1, e2, n1, n2, n3;
    F2 uv1, uv2, uv3;
    float area;
    uint8_t pad[20];
};
static_assert(sizeof(RefTriangle) == 128, "reference Triangle layout");

struct RefNode {              // 32 B LinearBVHNode
    F3 mn, mx;
    int32_t offset;           // leaf: first triangle; interior: index of the second child
    int16_t nTris, axis;
};
static_assert(sizeof(RefNode) == 32, "reference LinearBVHNode layout");

struct RefCamera {            // 72 B
    F3 eye, dir, up, right;
    float fov, aspect, apertureWidth, focusDistance;
    int32_t xpixels, ypixels;
};
static_assert(sizeof(RefCamera) == 72, "reference Camera layout");

// ---- device-facing layouts -------------------------------------------------------------------
// 4-ary node, 128 B = one L1/L2 line, read as eight 16-byte vectors.  Child boxes are the EXACT fp32
// boxes of the reference's binary nodes (the bit-exact leaf test needs them unrounded), stored SoA so
// one float4 holds the same plane of all four children.
struct WideNode {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    int32_t child[4];         // >= 0: index of a WideNode; < 0: leaf, see leaf_ref(); kEmptyChild: unused slot
    uint32_t aux[4];          // [0], [3]: packed CullBound (pack_cull_bound) of the node = loosest of its children's
                              // subtrees; [1], [2]: the float terms (a1, b2) of the dynamic pad (dynamic_pad_terms)
};
static_assert(sizeof(WideNode) == 128, "wide node is one 128-byte line");

constexpr int32_t kEmptyChild = 0x7fffffff;
constexpr uint32_t kMaxLeafTris = 32;           // per leaf reference; bigger reference leaves are chunked
constexpr uint32_t kMaxTris = 1u << 26;
// leaf reference: bit31 set | (count-1) << 26 | first triangle
inline int32_t leaf_ref(uint32_t first, uint32_t count) { return (int32_t)(0x80000000u | ((count - 1) << 26) | first); }

// Intersection record, 48 B = 3 x float4: (v1, triangle index) (e1, leafFirst) (e2, materialId) - the same first three
// vectors as a TightRecord, so the traversal reads either through one (base, stride) pair per ray
struct TriIsect { float v1[3]; uint32_t index; float e1[3]; uint32_t leafFirst; float e2[3]; uint32_t materialId; };
static_assert(sizeof(TriIsect) == 48, "intersection record");
// Shading record, 80 B = 5 x float4: (n1, area) (n2, materialId bits) (n3, 0) (uv1, uv2) (uv3, 0, 0)
struct TriShade {
    float n1[3]; float area;
    float n2[3]; uint32_t materialId;
    float n3[3]; uint32_t pad0;
    float uv1[2], uv2[2];
    float uv3[2]; uint32_t pad1[2];
};
static_assert(sizeof(TriShade) == 80, "shading record");

// Reference BVH, rebuilt.  tris is reordered in place into BVH order.  positions: 9 floats per
// triangle (the three OBJ vertex positions; the reference boxes them, not v1/e1/e2: bvh.cpp:79-96).
void build_reference_bvh(RefTriangle* tris, const float* positions, uint32_t n, std::vector<RefNode>& nodes_out);

// Scene::constructLightList.
void build_light_list(RefTriangle* tris, uint32_t n, float* mat_floats, const int32_t* mat_indices, uint32_t n_mats,
                      std::vector<uint32_t>& lights, float& lights_area);

// Re-cluster the reference's leaves under a binned-SAH interior structure (leaf records and boxes verbatim).
void recluster_leaves(const RefNode* nodes, uint32_t n_nodes, std::vector<RefNode>& out);

struct WideStats { uint32_t n_wide = 0, depth = 0, n_leaf_refs = 0; };
// Collapse the reference's binary tree into WideNodes (root = 0).
WideStats flatten_wide(const RefNode* nodes, uint32_t n_nodes, const RefTriangle* tris, uint32_t n_tris, std::vector<WideNode>& out);

// Distance-culling bound of a subtree (DESIGN.md "distance culling"): for every triangle below, a Moeller-Trumbore
// t computed in fp32 by the reference's operation sequence satisfies  t_computed >= t_true * A - B  for rays with
// |dir| <= 1, so a child whose box entry distance t0 gives t0 * A - B > best cannot hold the closest hit.
// A = 0, B = +inf marks a subtree that must never be culled (triangles too large for the bound to mean anything).
struct CullBound { float A, B; };
CullBound triangle_cull_bound(const RefTriangle& t);
// aux word of a WideNode child: bf16(A) rounded down in the high half, bf16(B) rounded up in the low half
uint32_t pack_cull_bound(CullBound b);

// ---- tight triangle BVH ("fast path" acceleration structure) ------------------------------------------------------
// The reference's leaves decide which triangles a ray may hit, but they are poor spatial clusters (on the 1M-triangle
// terrain a <= 3-triangle leaf spans 0.6 x 1.9 units of a 0.085-unit grid: 358 triangle tests per ray).  The tight BVH is
// built over the individual triangles instead, with every triangle box inflated far enough that any ray whose fp32
// Moeller-Trumbore test (geometry.cu:38-65) ACCEPTS the triangle is guaranteed to pass the box; a candidate found this way
// still has to pass the reference's own slab test of its reference leaf box (the "gate") before it counts.  Same accepted
// set as the reference traversal, far fewer candidates.
struct TightRecord {            // 80 B = 5 x float4
    float v1[3]; uint32_t origIdx;          // index in the reference's (BVH-ordered) triangle array = what hits report
    float e1[3]; uint32_t leafFirst;        // firstTriOffset of the reference leaf (tie key)
    float e2[3]; uint32_t materialId;
    float gateLo[3]; float gateHiX;         // the reference leaf's exact box
    float gateHiY, gateHiZ; uint32_t pad1[2];
};
static_assert(sizeof(TightRecord) == 80, "tight triangle record");
// region = box that contains every ray origin the bound is valid for (scene box united with the camera), see DESIGN.md
struct TightBvh {
    std::vector<RefNode> nodes;             // binary tree in the reference's pre-order layout over the tight order
    RawVec<TightRecord> records;            // triangles in tight order
    RawVec<RefTriangle> tris;               // same order, for the cull bounds of flatten_wide
    std::vector<float> boxes;               // per ORIGINAL triangle index, lo.xyz hi.xyz: the box a ray must pass before the
                                            // triangle is tested (its own inflated box, or its reference leaf's exact box)
    uint32_t n_unbounded = 0;               // triangles too large for a distance bound (kept behind their reference leaf's box)
    uint32_t n_split = 0;                   // triangles that got their own inflated box
    bool dynamic_pad = false;               // boxes are NOT inflated: the traversal pads them per visit (see below)
};
// Where the inflation of a triangle's box happens.  The bound on how far outside a triangle the fp32 test still accepts is
// pad = 2 s (|T| + L) with |T| the distance from the ray origin to the triangle.
//   kPadStatic : |T| <= D, the farthest point of the origin region - a fixed pad baked into the boxes.  Free at run time,
//                but on a large fine mesh D / (triangle size) is huge: 0.16 units of pad on the 0.085-unit grid of the
//                999 700-triangle terrain, i.e. every box is several triangles wide.
//   kPadDynamic: the boxes stay tight and the traversal kernel grows a child box by  (1 - A) |T|max + 2 B  when it visits
//                it, with (A, B) the node's cull bound (1 - A >= 2 s, B >= s L for every triangle below) and |T|max the
//                distance from the ray origin to the farthest corner of the child box: the pad follows the actual
//                distance (about 1e-3 of it on the terrain) at ~13 more instructions per child.
//   kPadAuto   : dynamic when the static inflation more than doubles the summed box area of the triangles.
enum PadMode { kPadAuto = 0, kPadStatic = 1, kPadDynamic = 2 };
// The two per-node constants of the dynamic pad, pad = a1 * |T|max + b2, from the (unpacked) cull bound of the node:
// 1 - A = 2 s + 1e-6 and B = s L (1 + 1e-6), so a1 >= 2 s (1 + 1e-3) + 2e-5 (the second term covers the rounding of the
// slab test itself, 3 eps |T|max per plane) and b2 >= 2 s L.  Stored as plain floats in aux[1], aux[2] of a WideNode:
// the bf16 halves of the packed cull bound would round 1 - A = 1e-3 up to 2..4e-3.
inline void dynamic_pad_terms(float A, float B, float& a1, float& b2) {
    a1 = 1.001f * (1.00002f - A);
    b2 = 2.002f * B;
}
void build_tight_bvh(const RefTriangle* tris, uint32_t n_tris, const RefNode* ref_nodes, uint32_t n_ref_nodes,
                     const F3& region_lo, const F3& region_hi, PadMode pad_mode, TightBvh& out);

// Surface-area-heuristic cost of a binary tree in the reference's pre-order layout: expected (interior nodes visited,
// triangles tested) by a random line that crosses the clip box (node boxes are clipped to it, so that two trees over
// the same scene are measured against the same set of lines).
struct SahCost { double nodes = 0, tris = 0; };
constexpr double kTightTraversalCost = 3.0;      // cost of visiting one binary node, in triangle tests (measured, see host_scene.cpp)
SahCost sah_cost(const RefNode* nodes, uint32_t n_nodes, const F3& clip_lo, const F3& clip_hi);

void make_tri_records(const RefTriangle* tris, uint32_t n, const RefNode* nodes, uint32_t n_nodes,
                      RawVec<TriIsect>& isect, RawVec<TriShade>& shade);

}  // namespace crt
