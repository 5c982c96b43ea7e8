// cudart_b200/csrc/crt_api.cu - the C ABI of include/crt_abi.h: context, scene upload, wave scheduling.
//
// A context owns (on one device): the re-flattened scene, the wavefront buffers sized for
// spp_per_wave x owned pixels paths, the compact float4 accumulator of the owned row tiles, and one
// stream.  crt_render() enqueues, per wave, raygen -> [extend, shade, shadow] x 6 -> accumulate; queue
// sizes never come back to the host, so a whole render is a fixed launch sequence with no host syncs.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/crt_abi.h"
#include "host_scene.h"
#include "kernels.cuh"

using namespace crt;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// No exception crosses the C boundary (crt_abi.h): every extern "C" body that allocates, spawns threads or calls into the
// host scene builders runs inside guarded().
template <class F>
int guarded(const char* what, F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return fail(CRT_ERR_NOMEM, "%s: out of host memory", what);
    } catch (const std::exception& e) {
        return fail(CRT_ERR_INVALID, "%s: %s", what, e.what());
    } catch (...) {
        return fail(CRT_ERR_INVALID, "%s: unknown exception", what);
    }
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(CRT_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    size_t cap = 0;
    cudaError_t alloc(size_t count) {                 // keeps the allocation when it is already large enough
        n = count;
        if (p && count <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = std::max<size_t>(count, 1);
        return cudaMalloc(&p, cap * sizeof(T));
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Two independent 64-bit multiply-xorshift hashes over 8-byte words, four lanes each for instruction-level parallelism
// (~8 bytes per cycle: 160 MB of triangles and nodes in ~10 ms).  Not cryptographic: it only answers "is this the
// geometry this context flattened last time?" for crt_update_scene.
struct Fingerprint {
    uint64_t a = 0, b = 0;
    bool operator==(const Fingerprint& o) const { return a == o.a && b == o.b; }
};
void fingerprint_mix_serial(Fingerprint& fp, const void* data, size_t bytes) {
    const unsigned char* p = (const unsigned char*)data;
    uint64_t ha[4] = {fp.a ^ bytes, 0x243F6A8885A308D3ull, 0x13198A2E03707344ull, 0xA4093822299F31D0ull};
    uint64_t hb[4] = {fp.b + bytes, 0x082EFA98EC4E6C89ull, 0x452821E638D01377ull, 0xBE5466CF34E90C6Cull};
    size_t i = 0;
    for (; i + 32 <= bytes; i += 32) {
        uint64_t w[4];
        memcpy(w, p + i, 32);
        for (int k = 0; k < 4; k++) {
            ha[k] = (ha[k] ^ w[k]) * 0x9E3779B97F4A7C15ull; ha[k] ^= ha[k] >> 29;
            hb[k] = (hb[k] + w[k]) * 0xC2B2AE3D27D4EB4Full; hb[k] ^= hb[k] >> 31;
        }
    }
    uint64_t tail[4] = {0, 0, 0, 0};
    memcpy(tail, p + i, bytes - i);
    for (int k = 0; k < 4; k++) {
        ha[k] = (ha[k] ^ tail[k]) * 0x9E3779B97F4A7C15ull; ha[k] ^= ha[k] >> 29;
        hb[k] = (hb[k] + tail[k]) * 0xC2B2AE3D27D4EB4Full; hb[k] ^= hb[k] >> 31;
    }
    fp.a = ((ha[0] * 3 + ha[1]) * 5 + ha[2]) * 7 + ha[3];
    fp.b = ((hb[0] * 11 + hb[1]) * 13 + hb[2]) * 17 + hb[3];
}
// Large arrays are hashed in fixed 4 MB chunks on the host's threads and the chunk hashes folded in order, so the value
// does not depend on the number of threads (one thread: 40 ms for the 155 MB of a 999 700-triangle scene, i.e. most of
// what crt_update_scene costs per frame once nothing has to be rebuilt).
void fingerprint_mix(Fingerprint& fp, const void* data, size_t bytes) {
    constexpr size_t kChunk = 4u << 20;
    if (bytes <= 2 * kChunk) { fingerprint_mix_serial(fp, data, bytes); return; }
    const size_t n_chunks = (bytes + kChunk - 1) / kChunk;
    std::vector<Fingerprint> part(n_chunks);
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (size_t c; (c = next.fetch_add(1)) < n_chunks;) {
            part[c].a = c;
            part[c].b = ~c;
            fingerprint_mix_serial(part[c], (const unsigned char*)data + c * kChunk, std::min(kChunk, bytes - c * kChunk));
        }
    };
    const size_t n_threads = std::min<size_t>(host_threads(), n_chunks);
    std::vector<std::unique_ptr<Worker>> th;
    for (size_t t = 1; t < n_threads; t++) th.emplace_back(new Worker(work));
    work();
    for (auto& t : th) t->join();
    fingerprint_mix_serial(fp, part.data(), n_chunks * sizeof(Fingerprint));
}

// What upload_scene derives from the caller's triangles and BVH nodes on the host.  Kept per context so that
// crt_update_scene with unchanged geometry (a new camera direction, new materials, the same frame again) re-uploads
// these arrays without rebuilding them: the 999 700-triangle terrain takes 0.4 s to derive and 15 ms to re-upload.
struct HostSceneCache {
    bool valid = false;
    Fingerprint fp;
    std::vector<WideNode> wide;             // tight tree (if any) followed by the reference-leaf tree
    RawVec<TriIsect> isect;                 // RawVec: sized once, then filled completely (host_scene.h)
    RawVec<TriShade> shade;
    RawVec<TriIsect> tight_isect;           // the tight tree's triangle records, hot part (what every test reads)
    RawVec<float4> tight_gate;              // ... and their reference-leaf boxes (read for would-be winners only)
    std::vector<float4> light_boxes;
    WideStats wide_stats{};
    uint32_t n_tight_wide = 0, tight_split = 0, n_front = 0;
    bool dynamic_pad = false;
    F3 region_lo{FLT_MAX, FLT_MAX, FLT_MAX}, region_hi{-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool coords_bounded = false;
    std::vector<const void*> pinned;        // cudaHostRegister'ed staging arrays (async copies at full PCIe rate)
    void unpin() { for (const void* q : pinned) cudaHostUnregister(const_cast<void*>(q)); pinned.clear(); }
    template <class V> void pin(const V& v) {
        if (v.size() * sizeof(v[0]) < (1u << 20)) return;                  // small arrays: not worth a registration
        if (cudaHostRegister((void*)v.data(), v.size() * sizeof(v[0]), cudaHostRegisterDefault) == cudaSuccess) pinned.push_back(v.data());
        else cudaGetLastError();                                           // pageable copy still works
    }
};

}  // namespace

// One wavefront's worth of path state.  A context has one or two (crt_options.waves_in_flight): with two, consecutive
// waves run on two streams so that the latency-bound shade launches of one wave can fill the issue slots and launch
// tails the traversal of the other leaves idle.
struct WaveSlot {
    DevBuf<float4> rq_o[2], rq_d[2], rq_m[2], hits, sq_o, sq_d, sq_c, color;
    DevBuf<uint32_t> rq_path[2], sq_path;
    DevBuf<uint2> rq_k[2];
    DevBuf<DCounters> counters;
    DWave wv{};
    LaunchCfg cfg{};
    cudaEvent_t accumulated = nullptr;       // recorded after this slot's k_accumulate: the next wave's accumulate waits for it
    void release() {
        for (int i = 0; i < 2; i++) { rq_o[i].release(); rq_d[i].release(); rq_path[i].release(); rq_m[i].release(); rq_k[i].release(); }
        hits.release(); sq_o.release(); sq_d.release(); sq_c.release(); color.release();
        sq_path.release(); counters.release();
        if (accumulated) cudaEventDestroy(accumulated);
        accumulated = nullptr;
    }
};

struct crt_ctx {
    int device = 0;
    bool own_stream = false;
    cudaStream_t stream = nullptr;           // the context's stream: everything a caller can observe is ordered on it
    // Wavefront batches run on the context's own work streams, one per wave slot.  A batch only has to follow the last
    // operation of this context that touched what it reads or writes (scene upload, reset, a read-out of the accumulator):
    // that point is `ev_epoch`, re-recorded on `stream` after every such operation.  `stream` itself only WAITS for a
    // call's batches (their last launch, the accumulation), so everything a caller can observe is still ordered on it -
    // but consecutive calls, and the consecutive batches inside one, overlap on the two work streams.
    cudaStream_t work[2] = {nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_epoch = nullptr, ev_order = nullptr;
    uint64_t epoch = 0, waited_epoch[2] = {~0ull, ~0ull};
    cudaEvent_t last_accum = nullptr;        // the accumulation of the most recent batch: the next one's accumulation follows it
    std::vector<cudaEvent_t> event_pool;     // timing events are recycled, never created per call
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;   // render spans not yet folded into render_ms
    LaunchCfg cfg{};
    uint32_t flags = 0, seed = 1234, spp_per_wave = 1;
    uint32_t samples = 0;                    // what Renderer::getSamplesRendered presents: samples asked for so far
    uint32_t launched = 0;                   // samples whose waves are already enqueued (samples - launched are deferred)
    uint32_t n_slots = 1, next_slot = 0;
    DScene sc{};
    DTiling tl{};
    WaveSlot slot[2];
    WideStats wide_stats{};
    HostSceneCache host_scene;
    uint64_t scene_rebuilds = 0, scene_uploads = 0, scene_upload_bytes = 0;
    // scene
    DevBuf<float4> wide, tri_isect, tri_shade, tri_tight, tri_gate, mats, light_boxes;
    uint32_t tight_split = 0, tight_nodes = 0;
    DevBuf<uint32_t> lights, tex_dims, tex_offsets;
    DevBuf<float> tex_data;
    DevBuf<DStats> stats;
    DevBuf<float4> accum;
    DevBuf<uint8_t> gamma_lut;
    // scratch for resolve / trace
    DevBuf<uchar4> rgba, rgba_local;
    DevBuf<float> f32;
    DevBuf<uint32_t> head;
    uint64_t launches = 0;
    double render_ms = 0;
    double stage_ms[CRT_STAGE_COUNT] = {0};

    cudaEvent_t take_event() {
        if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    void give_event(cudaEvent_t e) { if (e) event_pool.push_back(e); }
};

namespace {

thread_local uint64_t g_uploaded_bytes = 0;            // host -> device bytes of the upload_scene call in progress

template <typename T>
int upload(DevBuf<T>& b, const void* src, size_t count, cudaStream_t s) {
    CU(b.alloc(count));
    if (count) CU(cudaMemcpyAsync(b.p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    g_uploaded_bytes += count * sizeof(T);
    return CRT_OK;
}

void free_all(crt_ctx* c) {
    cudaSetDevice(c->device);
    c->host_scene.unpin();
    c->wide.release(); c->tri_isect.release(); c->tri_shade.release(); c->tri_tight.release(); c->tri_gate.release(); c->mats.release();
    c->lights.release(); c->light_boxes.release(); c->tex_dims.release(); c->tex_offsets.release(); c->tex_data.release();
    for (WaveSlot& w : c->slot) w.release();
    c->stats.release(); c->accum.release();
    c->gamma_lut.release(); c->rgba.release(); c->rgba_local.release(); c->f32.release(); c->head.release();
    for (auto& pr : c->pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    c->pending.clear();
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    c->event_pool.clear();
    for (cudaEvent_t* e : {&c->ev0, &c->ev1, &c->ev_epoch, &c->ev_order}) {
        if (*e) cudaEventDestroy(*e);
        *e = nullptr;
    }
    for (cudaStream_t& w : c->work) { if (w) cudaStreamDestroy(w); w = nullptr; }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = nullptr;
}

// Called after every operation on c->stream that later batches must follow (see crt_ctx::work).
int mark_epoch(crt_ctx* c) {
    CU(cudaEventRecord(c->ev_epoch, c->stream));
    c->epoch++;
    return CRT_OK;
}

int check_kernel(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CRT_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
    return CRT_OK;
}

}  // namespace

// Everything upload_scene derives from the caller's triangles and nodes on the host, into the context's HostSceneCache.
static int derive_host_scene(crt_ctx* c, const crt_scene_desc* s, const Fingerprint& fp) {
    const RefTriangle* tris = (const RefTriangle*)s->tris;
    const RefNode* nodes = (const RefNode*)s->nodes;
    HostSceneCache& hc = c->host_scene;
    hc.valid = false;
    hc.unpin();
    // The reference-leaf tree (re-cluster + 4-ary flatten), the shading records and the tight BVH below are independent
    // host jobs: the first two run on their own threads while this one builds the tight tree, and each of them forks
    // subtrees and fills its arrays on all threads (host_scene.cpp) (large scenes: 1.8 s -> 0.9 s -> 0.4 s
    // for 999 700 triangles).
    std::vector<WideNode>& wide = hc.wide;
    wide.clear();
    std::vector<RefNode> ref_tree;          // the binary tree behind `wide`, for the cost comparison below
    Worker ref_job([&] {
        std::vector<RefNode> clustered;
        if (!(c->flags & CRT_FLAG_REFERENCE_TOPOLOGY)) recluster_leaves(nodes, s->n_nodes, clustered);
        hc.wide_stats = clustered.empty() ? flatten_wide(nodes, s->n_nodes, tris, s->n_tris, wide)
                                          : flatten_wide(clustered.data(), (uint32_t)clustered.size(), tris, s->n_tris, wide);
        if (clustered.empty()) ref_tree.assign(nodes, nodes + s->n_nodes); else ref_tree.swap(clustered);
    });
    RawVec<TriIsect>& isect = hc.isect;
    RawVec<TriShade>& shade = hc.shade;
    Worker rec_job([&] { make_tri_records(tris, s->n_tris, nodes, s->n_nodes, isect, shade); });
    // ---- tight per-triangle BVH in front of the reference-leaf tree (host_scene.h) ----
    TightBvh tight;
    F3 region_lo{FLT_MAX, FLT_MAX, FLT_MAX}, region_hi{-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t n_tight_wide = 0;
    hc.tight_split = 0;
    if (!(c->flags & (CRT_FLAG_NO_TIGHT_BVH | CRT_FLAG_EXACT_VISIT))) {
        // region: every vertex and the camera, grown by 5 % of the diagonal (hit points on the scene's hull round outwards)
        auto grow = [&](double x, double y, double z) {
            region_lo = {std::min(region_lo.x, (float)x), std::min(region_lo.y, (float)y), std::min(region_lo.z, (float)z)};
            region_hi = {std::max(region_hi.x, (float)x), std::max(region_hi.y, (float)y), std::max(region_hi.z, (float)z)};
        };
        bool finite = true;
        for (uint32_t i = 0; i < s->n_tris; i++) {
            const RefTriangle& t = tris[i];
            grow(t.v1.x, t.v1.y, t.v1.z);
            grow((double)t.v1.x + t.e1.x, (double)t.v1.y + t.e1.y, (double)t.v1.z + t.e1.z);
            grow((double)t.v1.x + t.e2.x, (double)t.v1.y + t.e2.y, (double)t.v1.z + t.e2.z);
        }
        const RefCamera* cam = (const RefCamera*)s->camera;
        grow(cam->eye.x, cam->eye.y, cam->eye.z);
        const double dx = (double)region_hi.x - region_lo.x, dy = (double)region_hi.y - region_lo.y, dz = (double)region_hi.z - region_lo.z;
        const double m = 0.05 * std::sqrt(dx * dx + dy * dy + dz * dz) + std::fabs((double)cam->apertureWidth) + 1e-3;
        finite = std::isfinite(m) && m < 1e7;
        if (finite) {
            region_lo = {(float)(region_lo.x - m), (float)(region_lo.y - m), (float)(region_lo.z - m)};
            region_hi = {(float)(region_hi.x + m), (float)(region_hi.y + m), (float)(region_hi.z + m)};
            const PadMode pad_mode = (c->flags & CRT_FLAG_DYNAMIC_PAD) ? kPadDynamic : (c->flags & CRT_FLAG_STATIC_PAD) ? kPadStatic : kPadAuto;
            build_tight_bvh(tris, s->n_tris, nodes, s->n_nodes, region_lo, region_hi, pad_mode, tight);
            ref_job.join();
            // Per reference leaf the builder already chose between individually inflated triangles and the leaf behind its
            // own box, and it may merge several reference leaves into one tight leaf (the gate keeps that exact), which the
            // reference-leaf tree cannot.  Use it unless the surface-area-heuristic cost (expected node visits x their
            // measured weight + triangle tests of a random line) says the reference-leaf tree is cheaper.
            const SahCost tc = sah_cost(tight.nodes.data(), (uint32_t)tight.nodes.size(), region_lo, region_hi);
            const SahCost ref_cost = sah_cost(ref_tree.data(), (uint32_t)ref_tree.size(), region_lo, region_hi);
            const bool cheaper = kTightTraversalCost * tc.nodes + tc.tris < kTightTraversalCost * ref_cost.nodes + ref_cost.tris;
            const bool want = (c->flags & CRT_FLAG_TIGHT_BVH) || cheaper;
            if (!want || tight.nodes.empty()) tight = TightBvh();
        }
    }
    if (ref_job.joinable()) ref_job.join();
    if (3 * hc.wide_stats.depth + 4 > 64) return (fail(CRT_ERR_LIMIT, "scene: wide BVH depth %u exceeds the traversal stack", hc.wide_stats.depth));
    if (!tight.nodes.empty()) {
        std::vector<WideNode> tw;
        const WideStats ts = flatten_wide(tight.nodes.data(), (uint32_t)tight.nodes.size(), tight.tris.data(), s->n_tris, tw);
        if (3 * ts.depth + 4 > 64) {
            tight = TightBvh();
        } else {
            n_tight_wide = (uint32_t)tw.size();
            for (WideNode& w : wide)                                   // the reference-leaf tree moves behind the tight one
                for (int k = 0; k < 4; k++)
                    if (w.child[k] >= 0 && w.child[k] != kEmptyChild) w.child[k] += (int32_t)n_tight_wide;
            tw.insert(tw.end(), wide.begin(), wide.end());
            wide.swap(tw);
            hc.tight_split = tight.n_split;
        }
    }
    rec_job.join();
    hc.n_front = (uint32_t)(n_tight_wide ? n_tight_wide : wide.size());    // the tree the render path's rays walk
    hc.n_tight_wide = n_tight_wide;
    hc.dynamic_pad = n_tight_wide && tight.dynamic_pad;
    hc.tight_isect.clear();
    hc.tight_gate.clear();
    if (n_tight_wide) {
        // The builder's 80-byte records split by temperature: a leaf of eight 48-byte hot records spans three 128-byte
        // lines instead of five, and the 32-byte gates are only touched by candidates that would win.
        hc.tight_isect.resize(tight.records.size());
        hc.tight_gate.resize(2 * tight.records.size());
        const TightRecord* const recs = tight.records.data();
        TriIsect* const hot = hc.tight_isect.data();
        float4* const gate = hc.tight_gate.data();
        parallel_for((uint32_t)tight.records.size(), [=](uint32_t b, uint32_t e) {
            for (uint32_t i = b; i < e; i++) {
                const TightRecord& r = recs[i];
                TriIsect& h = hot[i];
                memcpy(h.v1, r.v1, 12); h.index = r.origIdx;
                memcpy(h.e1, r.e1, 12); h.leafFirst = r.leafFirst;
                memcpy(h.e2, r.e2, 12); h.materialId = r.materialId;
                gate[2 * (size_t)i] = make_float4(r.gateLo[0], r.gateLo[1], r.gateLo[2], 0.f);
                gate[2 * (size_t)i + 1] = make_float4(r.gateHiX, r.gateHiY, r.gateHiZ, 0.f);
            }
        });
    }
    hc.region_lo = region_lo; hc.region_hi = region_hi;
    {   // the reference leaf holding each light triangle
        hc.light_boxes.assign(2 * (size_t)s->n_lights, make_float4(0.f, 0.f, 0.f, 0.f));
        std::vector<uint32_t> leaf_of_first(s->n_tris, 0xffffffffu);
        for (uint32_t j = 0; j < s->n_nodes; j++)
            if (nodes[j].nTris > 0) leaf_of_first[(uint32_t)nodes[j].offset] = j;
        for (uint32_t k = 0; k < s->n_lights; k++) {
            const uint32_t lf = isect[s->light_indices[k]].leafFirst;
            const uint32_t j = lf < s->n_tris ? leaf_of_first[lf] : 0xffffffffu;
            if (j == 0xffffffffu) return (fail(CRT_ERR_INVALID, "scene: light %u is in no BVH leaf", k));
            hc.light_boxes[2 * k] = make_float4(nodes[j].mn.x, nodes[j].mn.y, nodes[j].mn.z, 0.f);
            hc.light_boxes[2 * k + 1] = make_float4(nodes[j].mx.x, nodes[j].mx.y, nodes[j].mx.z, 0.f);
        }
    }
    {
        bool bounded = true;
        for (uint32_t k = 0; k < s->n_nodes && bounded; k++) {
            const float v[6] = {nodes[k].mn.x, nodes[k].mn.y, nodes[k].mn.z, nodes[k].mx.x, nodes[k].mx.y, nodes[k].mx.z};
            for (float f : v) bounded = bounded && std::isfinite(f) && std::fabs(f) < 1e8f;
        }
        hc.coords_bounded = bounded;
    }
    hc.pin(hc.wide); hc.pin(hc.isect); hc.pin(hc.shade); hc.pin(hc.tight_isect); hc.pin(hc.tight_gate);
    hc.fp = fp;
    hc.valid = true;
    c->scene_rebuilds++;
    return CRT_OK;
}

// Validate, re-flatten and upload a scene into a context (crt_create and crt_update_scene).
static int upload_scene(crt_ctx* c, const crt_scene_desc* s, uint32_t W, uint32_t H) {
    // ---- scene: re-flatten on the host, upload ----
    const RefTriangle* tris = (const RefTriangle*)s->tris;
    const RefNode* nodes = (const RefNode*)s->nodes;
    for (uint32_t k = 0; k < s->n_nodes; k++) {
        const RefNode& nd = nodes[k];
        bool ok = nd.nTris > 0 ? (nd.offset >= 0 && (int64_t)nd.offset + (int64_t)nd.nTris <= (int64_t)s->n_tris)
                               : (nd.nTris == 0 && nd.offset > (int32_t)k && (uint32_t)nd.offset < s->n_nodes && k + 1 < s->n_nodes);
        if (!ok) return (fail(CRT_ERR_INVALID, "scene: BVH node %u is malformed", k));
    }
    if (s->n_lights == 0 || !s->light_indices || !(s->lights_area > 0.f))
        return fail(CRT_ERR_INVALID, "scene: no emissive triangles (the reference divides by a zero light area and indexes an empty "
                                     "light list in that case, SURVEY.md D13)");
    for (uint32_t k = 0; k < s->n_lights; k++)
        if (s->light_indices[k] >= s->n_tris) return (fail(CRT_ERR_INVALID, "scene: light index out of range"));
    if (s->n_mats == 0 || !s->mat_floats || !s->mat_indices)
        return fail(CRT_ERR_INVALID, "scene: no materials (every triangle indexes the material table, renderer.cu:154-157)");
    if (s->n_tex && (!s->tex_dims || !s->tex_offsets || (s->tex_texels && !s->tex_data)))
        return fail(CRT_ERR_INVALID, "scene: %u textures but null texture arrays", s->n_tex);
    // Fingerprint of everything the derived arrays depend on: geometry, BVH, lights, the flags that pick the trees, and
    // the camera position / aperture (the tight tree's region of valid ray origins contains the eye).
    const RefCamera* cam_in = (const RefCamera*)s->camera;
    Fingerprint fp;
    {
        const uint32_t hdr[6] = {s->n_tris, s->n_nodes, s->n_lights,
                                 c->flags & (CRT_FLAG_REFERENCE_TOPOLOGY | CRT_FLAG_NO_TIGHT_BVH | CRT_FLAG_EXACT_VISIT | CRT_FLAG_TIGHT_BVH |
                                            CRT_FLAG_DYNAMIC_PAD | CRT_FLAG_STATIC_PAD),
                                 (uint32_t)sizeof(RefTriangle), (uint32_t)sizeof(RefNode)};
        fingerprint_mix(fp, hdr, sizeof hdr);
        const float eye[4] = {cam_in->eye.x, cam_in->eye.y, cam_in->eye.z, cam_in->apertureWidth};
        fingerprint_mix(fp, eye, sizeof eye);
        fingerprint_mix(fp, s->light_indices, sizeof(uint32_t) * (size_t)s->n_lights);
        fingerprint_mix(fp, nodes, sizeof(RefNode) * (size_t)s->n_nodes);
        fingerprint_mix(fp, tris, sizeof(RefTriangle) * (size_t)s->n_tris);
    }
    HostSceneCache& hc = c->host_scene;
    if (!(hc.valid && hc.fp == fp)) {
        const int rc_derive = derive_host_scene(c, s, fp);
        if (rc_derive != CRT_OK) return rc_derive;
    }
    c->wide_stats = hc.wide_stats;
    c->tight_split = hc.tight_split;
    c->tight_nodes = hc.n_tight_wide;
    const uint32_t n_tight_wide = hc.n_tight_wide;
    for (uint32_t i = 0; i < s->n_tris; i++)
        if (tris[i].materialId >= s->n_mats) return (fail(CRT_ERR_INVALID, "scene: triangle %u has material %u of %u", i, tris[i].materialId, s->n_mats));
    std::vector<float4> mats(4 * (size_t)s->n_mats);
    for (uint32_t m = 0; m < s->n_mats; m++) {
        const float* f = s->mat_floats + 12 * (size_t)m;
        float bs, tx;
        memcpy(&bs, &s->mat_indices[2 * m], 4);
        memcpy(&tx, &s->mat_indices[2 * m + 1], 4);
        if (s->mat_indices[2 * m + 1] >= (int32_t)s->n_tex || s->mat_indices[2 * m + 1] < -1)
            return (fail(CRT_ERR_INVALID, "scene: material %u texture index %d out of range", m, s->mat_indices[2 * m + 1]));
        mats[4 * m + 0] = make_float4(f[0], f[1], f[2], bs);
        mats[4 * m + 1] = make_float4(f[3], f[4], f[5], tx);
        mats[4 * m + 2] = make_float4(f[6], f[7], f[8], 0.f);
        mats[4 * m + 3] = make_float4(f[9], f[10], f[11], 0.f);
    }
    int rc;
    g_uploaded_bytes = 0;
#define UP(buf, src, count) if ((rc = upload(c->buf, src, count, c->stream)) != CRT_OK) return rc
    UP(wide, hc.wide.data(), hc.wide.size() * 8);
    UP(tri_isect, hc.isect.data(), hc.isect.size() * 3);
    UP(tri_shade, hc.shade.data(), hc.shade.size() * 5);
    if (n_tight_wide) { UP(tri_tight, hc.tight_isect.data(), hc.tight_isect.size() * 3); UP(tri_gate, hc.tight_gate.data(), hc.tight_gate.size()); }
    else { c->tri_tight.release(); c->tri_gate.release(); }
    UP(mats, mats.data(), mats.size());
    UP(lights, s->light_indices, s->n_lights);
    UP(light_boxes, hc.light_boxes.data(), hc.light_boxes.size());
    UP(tex_dims, s->tex_dims, 2 * (size_t)s->n_tex);
    UP(tex_offsets, s->tex_offsets, s->n_tex ? s->n_tex + 1 : 0);
    UP(tex_data, s->tex_data, 3 * (size_t)s->tex_texels);
    uint8_t lut[256];
    {   // gammaCorrectPixel (renderer.cu:495-504) tabulated with the host's powf, i.e. the reference's own libm
        const float invGamma = 1.f / 2.2f;
        for (int v = 0; v < 256; v++) lut[v] = (unsigned char)(powf((float)v * 1.f / 255.f, invGamma) * 255.f);
    }
    UP(gamma_lut, lut, 256);
#undef UP
    c->scene_uploads++;
    c->scene_upload_bytes = g_uploaded_bytes;
    DScene& sc = c->sc;
    sc.wide = c->wide.p; sc.tri_isect = c->tri_isect.p; sc.tri_shade = c->tri_shade.p; sc.mats = c->mats.p;
    sc.lights = c->lights.p; sc.light_boxes = c->light_boxes.p; sc.tex_data = c->tex_data.p; sc.tex_dims = c->tex_dims.p; sc.tex_offsets = c->tex_offsets.p;
    // Shared-memory staging of the BVH top (breadth-first prefix of the wide nodes).  Measured on B200 (profiles/r01l_*):
    // when the whole tree fits, L1 wavefronts drop 50 -> 39 % and long-scoreboard stalls 1.9 -> 1.5 at equal speed;
    // when only the top fits (monkeys: 288 of 1058 nodes) the mixed smem/global fetch diverges and costs 3.7 %.
    // Policy: stage iff the whole tree fits, unless forced either way.
    sc.n_smem_nodes = 0;
    const size_t n_front = hc.n_front;                                     // the tree the render path's rays walk
    if (!(c->flags & CRT_FLAG_NO_SMEM_NODES) && (n_front <= kMaxSmemNodes || (c->flags & CRT_FLAG_SMEM_TOP_LEVELS)))
        sc.n_smem_nodes = std::min<uint32_t>((uint32_t)n_front, kMaxSmemNodes);
    // Tile-ordered camera rays pay where the tree is large (monkeys +2 %, 1M terrain +2 %, 131k terrain +2 %) and cost about
    // 1 % where the whole tree sits in shared memory and coherence buys nothing (the Cornell scenes):
    // profiles/r02af_raygen_tile_order_ab.jsonl, r02z_bench_all.jsonl.  CRT_RAYGEN_BLOCK overrides (A/B measurements).
    sc.raygen_block = n_front > kMaxSmemNodes ? 32u : 0u;
    if (const char* e = getenv("CRT_RAYGEN_BLOCK")) sc.raygen_block = (uint32_t)atoi(e) / 8u * 8u;
    sc.tri_tight = n_tight_wide ? c->tri_tight.p : nullptr;
    sc.tri_gate = n_tight_wide ? c->tri_gate.p : nullptr;
    sc.dyn_pad = hc.dynamic_pad ? 1 : 0;
    sc.ref_root = n_tight_wide;
    sc.region_lo[0] = hc.region_lo.x; sc.region_lo[1] = hc.region_lo.y; sc.region_lo[2] = hc.region_lo.z;
    sc.region_hi[0] = hc.region_hi.x; sc.region_hi[1] = hc.region_hi.y; sc.region_hi[2] = hc.region_hi.z;
    sc.n_wide = (uint32_t)hc.wide.size(); sc.n_tris = s->n_tris; sc.n_mats = s->n_mats; sc.n_lights = s->n_lights;
    sc.n_tex = s->n_tex; sc.tex_texels = s->tex_texels; sc.lights_area = s->lights_area;
    memcpy(&sc.cam, s->camera, sizeof(RefCamera));
    sc.cam.xpixels = (int32_t)W;
    sc.cam.ypixels = (int32_t)H;
    sc.coords_bounded = hc.coords_bounded ? 1 : 0;

    return CRT_OK;
}

extern "C" {

int crt_abi_version(void) { return CRT_ABI_VERSION; }
const char* crt_last_error(void) { return g_err.c_str(); }

static int create_impl(const crt_scene_desc* s, uint32_t W, uint32_t H, const crt_options* opt, crt_ctx*& c) {
    if (!W || !H) return fail(CRT_ERR_INVALID, "crt_create: null scene/out or empty image");
    if (!s->tris || !s->nodes || !s->n_tris || !s->n_nodes || !s->camera) return fail(CRT_ERR_INVALID, "crt_create: empty scene");
    if (s->n_mats > 128) return fail(CRT_ERR_LIMIT, "crt_create: %u materials > MAX_MATERIALS 128 (renderer.h:22)", s->n_mats);
    if (s->n_tris >= kMaxTris) return fail(CRT_ERR_LIMIT, "crt_create: %u triangles >= %u", s->n_tris, kMaxTris);
    if ((uint64_t)W * H >= (1ull << 31)) return fail(CRT_ERR_LIMIT, "crt_create: image too large");
    crt_options o;
    memset(&o, 0, sizeof o);
    o.device = -1;
    if (opt) memcpy(&o, opt, std::min<size_t>(opt->struct_size ? opt->struct_size : sizeof o, sizeof o));
    if (o.world == 0) o.world = 1;
    if (o.rank >= o.world) return fail(CRT_ERR_INVALID, "crt_create: rank %u >= world %u", o.rank, o.world);
    if (o.tile_rows == 0) o.tile_rows = 8;
    if (o.seed == 0) o.seed = 1234;
    if (const char* e = getenv("CRT_WAVES_IN_FLIGHT")) { if (!o.waves_in_flight) o.waves_in_flight = (uint32_t)atoi(e); }   // A/B measurements
    if (o.waves_in_flight > 2) return fail(CRT_ERR_INVALID, "crt_create: waves_in_flight %u (0, 1 or 2)", o.waves_in_flight);
    // default: two batches in flight (measured +1.5 ... +3 %, profiles/r02e_waves_in_flight_ab.jsonl).  Per-stage event
    // timing only means something when the stages do not overlap: CRT_FLAG_PROFILE runs one batch at a time.
    if (o.waves_in_flight == 0) o.waves_in_flight = 2;
    if (o.flags & CRT_FLAG_PROFILE) o.waves_in_flight = 1;

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(CRT_ERR_CUDA, "crt_create: no CUDA device (%s); this path has no CPU fallback", cudaGetErrorString(e));
    int dev = o.device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(CRT_ERR_INVALID, "crt_create: device %d of %d", dev, ndev);
    CU(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) return fail(CRT_ERR_CUDA, "crt_create: device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);

    c = new crt_ctx();
    c->device = dev;
    c->flags = o.flags;
    c->seed = o.seed;
    c->n_slots = o.waves_in_flight;
    if (o.stream) c->stream = (cudaStream_t)o.stream;
    else {
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    for (uint32_t si = 0; si < c->n_slots; si++) CU(cudaStreamCreateWithFlags(&c->work[si], cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    for (cudaEvent_t* ev : {&c->ev_epoch, &c->ev_order}) CU(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    if (configure_kernels() != 0) return fail(CRT_ERR_CUDA, "crt_create: cannot opt in to %u KB of shared memory per CTA", 56u);
    c->cfg.sms = prop.multiProcessorCount;
    c->cfg.stream = c->stream;
    c->cfg.trav_blocks = 0;
    if (const char* e = getenv("CRT_TRAV_BLOCKS")) c->cfg.trav_blocks = atoi(e);       // A/B measurements: traversal CTAs per SM

    { const int rc_scene = upload_scene(c, s, W, H); if (rc_scene != CRT_OK) return rc_scene; }

    // ---- pixel ownership ----
    DTiling& tl = c->tl;
    tl.W = W; tl.H = H; tl.rank = o.rank; tl.world = o.world; tl.tile_rows = o.tile_rows;
    const uint32_t n_tiles = (H + o.tile_rows - 1) / o.tile_rows;
    uint32_t rows = 0;
    for (uint32_t t = o.rank; t < n_tiles; t += o.world) rows += std::min(o.tile_rows, H - t * o.tile_rows);
    tl.local_rows = rows;
    tl.local_pixels = rows * W;
    tl.padded_pixels = ((n_tiles + o.world - 1) / o.world) * o.tile_rows * W;

    // ---- wave buffers ----
    uint32_t k = o.spp_per_wave;
    if (k == 0) {       // auto: about 16M paths per batch (3.4 GB of wave state each; measured +12 % over 4M in round 1, and
                        // +2 % for two batches of 16M over two of 8M: profiles/r02o_wave_size.jsonl), all batches together
                        // within a quarter of the free device memory, at least 1 spp per batch
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const uint64_t by_mem = std::max<uint64_t>(1u << 20, free_b / 4 / 208 / c->n_slots);
        const uint64_t target = std::min<uint64_t>(16ull << 20, by_mem);
        k = (uint32_t)std::max<uint64_t>(1, target / std::max<uint32_t>(tl.local_pixels, 1));
        k = std::min<uint32_t>(k, 64);
    }
    c->spp_per_wave = k;
    const size_t cap = std::max<size_t>((size_t)tl.local_pixels * k, 32);
    if (cap >= (1ull << 31)) return fail(CRT_ERR_LIMIT, "crt_create: %zu paths per wave", cap);
#define AL(buf, count) if (buf.alloc(count) != cudaSuccess) return fail(CRT_ERR_NOMEM, "crt_create: cudaMalloc(%s) failed: %s", #buf, cudaGetErrorString(cudaGetLastError()))
    AL(c->stats, 1); AL(c->head, 1);
    AL(c->accum, tl.padded_pixels);
    for (uint32_t si = 0; si < c->n_slots; si++) {
        WaveSlot& w = c->slot[si];
        for (int i = 0; i < 2; i++) { AL(w.rq_o[i], cap); AL(w.rq_d[i], cap); AL(w.rq_path[i], cap); AL(w.rq_m[i], cap); AL(w.rq_k[i], cap); }
        AL(w.hits, cap); AL(w.sq_o, cap); AL(w.sq_d, cap); AL(w.sq_c, cap); AL(w.sq_path, cap);
        AL(w.color, cap);
        AL(w.counters, 1);
        DWave& wv = w.wv;
        for (int i = 0; i < 2; i++) {
            wv.rq_o[i] = w.rq_o[i].p; wv.rq_d[i] = w.rq_d[i].p; wv.rq_path[i] = w.rq_path[i].p;
            wv.rq_m[i] = w.rq_m[i].p; wv.rq_k[i] = w.rq_k[i].p;
        }
        wv.hits = w.hits.p; wv.sq_o = w.sq_o.p; wv.sq_d = w.sq_d.p; wv.sq_c = w.sq_c.p; wv.sq_path = w.sq_path.p;
        wv.color = w.color.p;
        wv.counters = w.counters.p; wv.stats = c->stats.p; wv.capacity = (uint32_t)cap;
        w.cfg = c->cfg;
        w.cfg.stream = c->work[si];
        CU(cudaEventCreateWithFlags(&w.accumulated, cudaEventDisableTiming));
    }
#undef AL
    CU(cudaMemsetAsync(c->accum.p, 0, (size_t)tl.padded_pixels * sizeof(float4), c->stream));
    CU(cudaMemsetAsync(c->stats.p, 0, sizeof(DStats), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return mark_epoch(c);
}

int crt_create(const crt_scene_desc* s, uint32_t W, uint32_t H, const crt_options* opt, crt_ctx** out) {
    if (!s || !out) return fail(CRT_ERR_INVALID, "crt_create: null scene/out or empty image");
    crt_ctx* c = nullptr;
    const int rc = guarded("crt_create", [&] { return create_impl(s, W, H, opt, c); });
    if (rc != CRT_OK) {
        if (c) { free_all(c); delete c; }
        return rc;
    }
    *out = c;
    return CRT_OK;
}

void crt_destroy(crt_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (cudaStream_t w : c->work) if (w) cudaStreamSynchronize(w);
    cudaStreamSynchronize(c->stream);
    free_all(c);
    delete c;
}

// Enqueue the waves of every sample asked for so far (c->launched .. c->samples).  Consecutive waves alternate between the
// wave slots - across calls too - each on its slot's work stream; the per-pixel accumulation, the only thing two waves
// share, is chained in wave order by events, so the sums have the same bits whatever the number of slots; the context's
// stream waits for the accumulation of the call's last wave on every slot it used.
static int launch_pending(crt_ctx* c) {
    const uint32_t spp = c->samples - c->launched;
    if (spp == 0 || c->tl.local_pixels == 0) { c->launched = c->samples; return CRT_OK; }
    const bool exact = c->flags & CRT_FLAG_EXACT_VISIT, stats = c->flags & CRT_FLAG_STATS;
    const bool prof = c->flags & CRT_FLAG_PROFILE;            // implies one slot (crt_create)
    cudaEvent_t a = c->take_event(), b = c->take_event();
    CU(cudaEventRecord(a, c->stream));
    struct Span { int stage; cudaEvent_t e; };
    std::vector<Span> spans;        // CRT_FLAG_PROFILE: an event after every launch; stage -1 opens a wave
    auto mark = [&](int stage) {
        if (!prof) return;
        cudaEvent_t e = c->take_event();
        cudaEventRecord(e, c->work[0]);
        spans.push_back({stage, e});
    };
    uint32_t done = 0;
    bool used[2] = {false, false};
    while (done < spp) {
        const uint32_t k = std::min(c->spp_per_wave, spp - done);
        const uint32_t si = c->next_slot;
        WaveSlot& w = c->slot[si];
        const LaunchCfg& cfg = w.cfg;
        if (c->waited_epoch[si] != c->epoch) {
            CU(cudaStreamWaitEvent(cfg.stream, c->ev_epoch, 0));
            c->waited_epoch[si] = c->epoch;
        }
        used[si] = true;
        CU(cudaMemsetAsync(w.counters.p, 0, sizeof(DCounters), cfg.stream));
        mark(-1);
        launch_raygen(c->sc, c->tl, w.wv, c->launched + done, k, c->seed, cfg);
        mark(CRT_STAGE_RAYGEN);
        // traverse(E0); shade(0); traverse(E1 + S0); shade(1); ... traverse(E5 + S4); shade(5); traverse(S5)
        for (int bounce = 0; bounce < kMaxBounces; bounce++) {
            launch_traverse(c->sc, w.wv, bounce, bounce - 1, exact, stats, cfg);
            mark(CRT_STAGE_EXTEND);
            launch_shade(c->sc, c->tl, w.wv, bounce, cfg);
            mark(CRT_STAGE_SHADE);
        }
        launch_traverse(c->sc, w.wv, -1, kMaxBounces - 1, exact, stats, cfg);
        mark(CRT_STAGE_SHADOW);
        if (c->last_accum && c->last_accum != w.accumulated) CU(cudaStreamWaitEvent(cfg.stream, c->last_accum, 0));   // accumulate in wave order
        launch_accumulate(c->tl, w.wv, c->accum.p, k, cfg);
        CU(cudaEventRecord(w.accumulated, cfg.stream));
        c->last_accum = w.accumulated;
        mark(CRT_STAGE_ACCUM);
        c->launches += 3 + 2 * kMaxBounces;
        done += k;
        c->next_slot = (si + 1) % c->n_slots;
    }
    for (uint32_t si = 0; si < c->n_slots; si++)
        if (used[si]) CU(cudaStreamWaitEvent(c->stream, c->slot[si].accumulated, 0));   // accumulate is a wave's last launch
    CU(cudaEventRecord(b, c->stream));
    c->pending.emplace_back(a, b);
    c->launched = c->samples;
    int rc = check_kernel("render wave");
    if (rc) return rc;
    if (prof) {
        CU(cudaStreamSynchronize(c->work[0]));
        for (size_t i = 1; i < spans.size(); i++) {
            if (spans[i].stage < 0) continue;
            float ms = 0;
            cudaEventElapsedTime(&ms, spans[i - 1].e, spans[i].e);
            c->stage_ms[spans[i].stage] += ms;
        }
        for (auto& sp : spans) c->give_event(sp.e);
    }
    return CRT_OK;
}

int crt_render(crt_ctx* c, uint32_t spp) {
    if (!c) return fail(CRT_ERR_INVALID, "crt_render: null ctx");
    return guarded("crt_render", [&] {
        CU(cudaSetDevice(c->device));
        if ((uint64_t)c->samples + spp > 0xffffffffull) return fail(CRT_ERR_LIMIT, "crt_render: sample counter overflow");
        c->samples += spp;
        // The literal seam renders one sample per call (Renderer::renderOneSamplePerPixel x spp, terminal_launcher.cpp:12-14)
        // while the machine wants whole wavefront batches: calls that add up to less than one batch are queued here and
        // launched by the call that completes the batch (or by crt_flush / crt_sync / any read-out).  The presented counter
        // (crt_samples_rendered) advances with every call; sample indices, and therefore the image, do not depend on how
        // the samples were grouped into waves.
        if (c->samples - c->launched < c->spp_per_wave * c->n_slots) return (int)CRT_OK;
        return launch_pending(c);
    });
}

int crt_flush(crt_ctx* c) {
    if (!c) return fail(CRT_ERR_INVALID, "crt_flush: null ctx");
    return guarded("crt_flush", [&] {
        CU(cudaSetDevice(c->device));
        return launch_pending(c);
    });
}

int crt_sync(crt_ctx* c) {
    if (!c) return fail(CRT_ERR_INVALID, "crt_sync: null ctx");
    int rc = crt_flush(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    for (auto& pr : c->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) c->render_ms += ms;
        c->give_event(pr.first);
        c->give_event(pr.second);
    }
    c->pending.clear();
    return CRT_OK;
}

uint32_t crt_samples_rendered(const crt_ctx* c) { return c ? c->samples : 0; }
uint32_t crt_spp_per_wave(const crt_ctx* c) { return c ? c->spp_per_wave : 0; }

static int resolve_common(crt_ctx* c, const float4* src, bool gathered, uint8_t* rgba_dst, float* f32_dst, int f32_mode) {
    int rc = crt_flush(c);
    if (rc) return rc;
    const size_t n = (size_t)c->tl.W * c->tl.H;
    if (rgba_dst && !c->rgba.p) CU(c->rgba.alloc(n));
    if (f32_dst && !c->f32.p) CU(c->f32.alloc(3 * n));
    if (rgba_dst) CU(cudaMemsetAsync(c->rgba.p, 0, n * 4, c->stream));
    if (f32_dst) CU(cudaMemsetAsync(c->f32.p, 0, 3 * n * sizeof(float), c->stream));
    launch_resolve(c->tl, src, gathered, std::max(c->samples, 1u), c->gamma_lut.p, rgba_dst ? c->rgba.p : nullptr,
                   f32_dst ? c->f32.p : nullptr, f32_mode, c->cfg);
    c->launches++;
    rc = check_kernel("resolve");
    if (rc) return rc;
    rc = mark_epoch(c);
    if (rc) return rc;
    if (rgba_dst) CU(cudaMemcpyAsync(rgba_dst, c->rgba.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (f32_dst) CU(cudaMemcpyAsync(f32_dst, c->f32.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    return crt_sync(c);
}

int crt_resolve_rgba8(crt_ctx* c, uint8_t* dst) {
    if (!c || !dst) return fail(CRT_ERR_INVALID, "crt_resolve_rgba8: null argument");
    return resolve_common(c, c->accum.p, false, dst, nullptr, 0);
}
int crt_resolve_f32(crt_ctx* c, float* dst) {
    if (!c || !dst) return fail(CRT_ERR_INVALID, "crt_resolve_f32: null argument");
    return resolve_common(c, c->accum.p, false, nullptr, dst, 1);
}
int crt_read_accum(crt_ctx* c, float* dst) {
    if (!c || !dst) return fail(CRT_ERR_INVALID, "crt_read_accum: null argument");
    return resolve_common(c, c->accum.p, false, nullptr, dst, 2);
}
int crt_resolve_gathered(crt_ctx* c, const void* d_gathered, uint8_t* rgba8_dst, float* f32_dst) {
    if (!c || !d_gathered) return fail(CRT_ERR_INVALID, "crt_resolve_gathered: null argument");
    return resolve_common(c, (const float4*)d_gathered, true, rgba8_dst, f32_dst, 1);
}

// ---- multi-GPU read-out in display bytes (4 B per pixel instead of 16) -----------------------------------------------
int crt_resolve_local_rgba8_device(crt_ctx* c, void** d_ptr, size_t* bytes) {
    if (!c || !d_ptr) return fail(CRT_ERR_INVALID, "crt_resolve_local_rgba8_device: null argument");
    int rc = crt_flush(c);
    if (rc) return rc;
    CU(c->rgba_local.alloc(std::max<size_t>(c->tl.padded_pixels, 1)));
    launch_resolve_local(c->tl, c->accum.p, std::max(c->samples, 1u), c->gamma_lut.p, c->rgba_local.p, c->cfg);
    c->launches++;
    rc = check_kernel("resolve_local");
    if (rc) return rc;
    rc = mark_epoch(c);                          // later batches must not add to the accumulator while it is being read
    if (rc) return rc;
    *d_ptr = c->rgba_local.p;
    if (bytes) *bytes = (size_t)c->tl.padded_pixels * 4;
    return CRT_OK;
}

int crt_assemble_rgba8(crt_ctx* c, const void* d_gathered, uint8_t* dst) {
    if (!c || !d_gathered || !dst) return fail(CRT_ERR_INVALID, "crt_assemble_rgba8: null argument");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)c->tl.W * c->tl.H;
    if (!c->rgba.p) CU(c->rgba.alloc(n));
    launch_assemble(c->tl, (const uchar4*)d_gathered, c->rgba.p, c->cfg);
    c->launches++;
    int rc = check_kernel("assemble");
    if (rc) return rc;
    CU(cudaMemcpyAsync(dst, c->rgba.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    return crt_sync(c);
}

// ---- stream ordering against a caller's stream (collectives run there), without host synchronisation ----------------
int crt_stream_wait_ctx(crt_ctx* c, void* stream) {
    if (!c) return fail(CRT_ERR_INVALID, "crt_stream_wait_ctx: null ctx");
    int rc = crt_flush(c);
    if (rc) return rc;
    if ((cudaStream_t)stream == c->stream) return CRT_OK;
    CU(cudaEventRecord(c->ev_order, c->stream));
    CU(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_order, 0));
    return CRT_OK;
}
int crt_ctx_wait_stream(crt_ctx* c, void* stream) {
    if (!c) return fail(CRT_ERR_INVALID, "crt_ctx_wait_stream: null ctx");
    CU(cudaSetDevice(c->device));
    if ((cudaStream_t)stream == c->stream) return mark_epoch(c);
    CU(cudaEventRecord(c->ev_order, (cudaStream_t)stream));
    CU(cudaStreamWaitEvent(c->stream, c->ev_order, 0));
    return mark_epoch(c);                        // batches launched from now on follow what the caller's stream did
}

int crt_update_scene(crt_ctx* c, const crt_scene_desc* s) {
    if (!c || !s) return fail(CRT_ERR_INVALID, "crt_update_scene: null argument");
    if (!s->tris || !s->nodes || !s->n_tris || !s->n_nodes || !s->camera) return fail(CRT_ERR_INVALID, "crt_update_scene: empty scene");
    if (s->n_mats > 128) return fail(CRT_ERR_LIMIT, "crt_update_scene: %u materials > MAX_MATERIALS 128", s->n_mats);
    if (s->n_tris >= kMaxTris) return fail(CRT_ERR_LIMIT, "crt_update_scene: %u triangles >= %u", s->n_tris, kMaxTris);
    return guarded("crt_update_scene", [&] {
        CU(cudaSetDevice(c->device));
        c->samples = c->launched;                                // samples still queued on the host belong to the old scene: dropped
        CU(cudaStreamSynchronize(c->stream));                    // the previous frame may still be reading the old arrays
        int rc = upload_scene(c, s, c->tl.W, c->tl.H);
        if (rc) return rc;
        CU(cudaStreamSynchronize(c->stream));                    // host staging vectors die with this call
        return crt_reset(c);
    });
}

int crt_reset(crt_ctx* c) {
    if (!c) return fail(CRT_ERR_INVALID, "crt_reset: null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->accum.p, 0, (size_t)c->tl.padded_pixels * sizeof(float4), c->stream));
    c->samples = 0;
    c->launched = 0;
    return mark_epoch(c);
}

int crt_get_counters(crt_ctx* c, crt_counters* out) {
    if (!c || !out) return fail(CRT_ERR_INVALID, "crt_get_counters: null argument");
    int rc = crt_sync(c);
    if (rc) return rc;
    DStats st;
    CU(cudaMemcpy(&st, c->stats.p, sizeof st, cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof *out);
    out->samples = st.samples; out->extend_rays = st.extend_rays; out->shadow_rays = st.shadow_rays;
    out->node_visits = st.node_visits; out->tri_tests = st.tri_tests;
    out->kernel_launches = c->launches;
    out->render_ms = c->render_ms;
    for (int i = 0; i < CRT_STAGE_COUNT; i++) out->stage_ms[i] = c->stage_ms[i];
    return CRT_OK;
}

int crt_trace_closest_device(crt_ctx* c, const void* d_rays, size_t n, void* d_hits, uint32_t flags, float* kernel_ms) {
    if (!c || (n && (!d_rays || !d_hits))) return fail(CRT_ERR_INVALID, "crt_trace_closest_device: null argument");
    if (n >= (1ull << 31)) return fail(CRT_ERR_LIMIT, "crt_trace_closest_device: too many rays");
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->head.p, 0, sizeof(uint32_t), c->stream));
    CU(cudaEventRecord(c->ev0, c->stream));
    launch_trace(c->sc, (const float4*)d_rays, (uint32_t)n, (float4*)d_hits, c->head.p, c->stats.p,
                 flags & CRT_FLAG_EXACT_VISIT, (flags | c->flags) & CRT_FLAG_STATS, c->cfg);
    CU(cudaEventRecord(c->ev1, c->stream));
    c->launches++;
    int rc = check_kernel("trace");
    if (rc) return rc;
    if (kernel_ms) {
        CU(cudaEventSynchronize(c->ev1));
        CU(cudaEventElapsedTime(kernel_ms, c->ev0, c->ev1));
    }
    return CRT_OK;
}

int crt_trace_closest(crt_ctx* c, const crt_ray* rays, size_t n, crt_hit* hits, uint32_t flags) {
    if (!c || (n && (!rays || !hits))) return fail(CRT_ERR_INVALID, "crt_trace_closest: null argument");
    if (n == 0) return CRT_OK;
    CU(cudaSetDevice(c->device));
    DevBuf<float4> dr, dh;
    CU(dr.alloc(2 * n));
    if (dh.alloc(n) != cudaSuccess) { dr.release(); return fail(CRT_ERR_NOMEM, "crt_trace_closest: cudaMalloc failed"); }
    int rc = CRT_OK;
    if (cudaMemcpyAsync(dr.p, rays, n * sizeof(crt_ray), cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
        rc = fail(CRT_ERR_CUDA, "crt_trace_closest: H2D failed");
    if (!rc) rc = crt_trace_closest_device(c, dr.p, n, dh.p, flags, nullptr);
    if (!rc && cudaMemcpyAsync(hits, dh.p, n * sizeof(crt_hit), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
        rc = fail(CRT_ERR_CUDA, "crt_trace_closest: D2H failed");
    if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess)
        rc = fail(CRT_ERR_CUDA, "crt_trace_closest: %s", cudaGetErrorString(cudaGetLastError()));
    dr.release();
    dh.release();
    return rc;
}

int crt_accum_device_ptr(crt_ctx* c, void** d_ptr, size_t* padded_pixels) {
    if (!c || !d_ptr) return fail(CRT_ERR_INVALID, "crt_accum_device_ptr: null argument");
    int rc = crt_flush(c);                      // what the caller reads next must already be enqueued
    if (rc) return rc;
    *d_ptr = c->accum.p;
    if (padded_pixels) *padded_pixels = c->tl.padded_pixels;
    return CRT_OK;
}

// ---- host scene build ----------------------------------------------------------------------------
int crt_build_bvh(void* tris, const float* positions, uint32_t n_tris, void* nodes_out, uint32_t* n_nodes_out) {
    if (!tris || !positions || !nodes_out || !n_nodes_out) return fail(CRT_ERR_INVALID, "crt_build_bvh: null argument");
    return guarded("crt_build_bvh", [&] {
        std::vector<RefNode> nodes;
        build_reference_bvh((RefTriangle*)tris, positions, n_tris, nodes);
        memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(RefNode));
        *n_nodes_out = (uint32_t)nodes.size();
        return (int)CRT_OK;
    });
}

int crt_build_lights(void* tris, uint32_t n_tris, float* mat_floats, const int32_t* mat_indices, uint32_t n_mats,
                     uint32_t* light_indices_out, uint32_t* n_lights_out, float* lights_area_out) {
    if (!tris || !mat_floats || !mat_indices || !light_indices_out || !n_lights_out || !lights_area_out)
        return fail(CRT_ERR_INVALID, "crt_build_lights: null argument");
    return guarded("crt_build_lights", [&] {
        std::vector<uint32_t> lights;
        float area = 0;
        build_light_list((RefTriangle*)tris, n_tris, mat_floats, mat_indices, n_mats, lights, area);
        memcpy(light_indices_out, lights.data(), lights.size() * 4);
        *n_lights_out = (uint32_t)lights.size();
        *lights_area_out = area;
        return (int)CRT_OK;
    });
}

int crt_wide_bvh(const void* nodes, uint32_t n_nodes, void* wide_out, uint32_t cap, uint32_t* n_wide_out) {
    if (!nodes || !n_wide_out) return fail(CRT_ERR_INVALID, "crt_wide_bvh: null argument");
    return guarded("crt_wide_bvh", [&] {
        std::vector<WideNode> wide;
        std::vector<RefNode> clustered;
        recluster_leaves((const RefNode*)nodes, n_nodes, clustered);
        flatten_wide(clustered.data(), (uint32_t)clustered.size(), nullptr, 0, wide);
        *n_wide_out = (uint32_t)wide.size();
        if (wide_out) memcpy(wide_out, wide.data(), std::min<size_t>(cap, wide.size()) * sizeof(WideNode));
        return (int)CRT_OK;
    });
}


int crt_tight_bvh(const void* tris, uint32_t n_tris, const void* nodes, uint32_t n_nodes, const float* region, uint32_t pad_mode,
                  float* tri_boxes_out, float* tri_pad_out, uint32_t* dynamic_out,
                  uint32_t* n_split_out, uint32_t* n_tight_nodes_out, uint32_t* n_wide_out, double* sah_out) {
    if (!tris || !nodes || !region || pad_mode > 2) return fail(CRT_ERR_INVALID, "crt_tight_bvh: bad argument");
    return guarded("crt_tight_bvh", [&] {
    TightBvh tb;
    build_tight_bvh((const RefTriangle*)tris, n_tris, (const RefNode*)nodes, n_nodes, F3{region[0], region[1], region[2]},
                    F3{region[3], region[4], region[5]}, (PadMode)pad_mode, tb);
    if (tri_boxes_out && tb.boxes.size() == 6 * (size_t)n_tris) memcpy(tri_boxes_out, tb.boxes.data(), tb.boxes.size() * sizeof(float));
    if (dynamic_out) *dynamic_out = tb.dynamic_pad ? 1u : 0u;
    if (tri_pad_out)         // (a1, b2) per triangle from its own cull bound; a node's terms are never smaller
        for (uint32_t i = 0; i < n_tris; i++) {
            const CullBound b = triangle_cull_bound(((const RefTriangle*)tris)[i]);
            dynamic_pad_terms(b.A, b.B, tri_pad_out[2 * i], tri_pad_out[2 * i + 1]);
        }
    if (n_split_out) *n_split_out = tb.n_split;
    if (n_tight_nodes_out) *n_tight_nodes_out = (uint32_t)tb.nodes.size();
    if (sah_out) {      // (nodes, triangles) expected per random line: tight tree, then the re-clustered reference-leaf tree
        std::vector<RefNode> clustered;
        recluster_leaves((const RefNode*)nodes, n_nodes, clustered);
        const F3 rl{region[0], region[1], region[2]}, rh{region[3], region[4], region[5]};
        const SahCost a = sah_cost(tb.nodes.data(), (uint32_t)tb.nodes.size(), rl, rh), b = sah_cost(clustered.data(), (uint32_t)clustered.size(), rl, rh);
        sah_out[0] = a.nodes; sah_out[1] = a.tris; sah_out[2] = b.nodes; sah_out[3] = b.tris;
    }
    if (n_wide_out) {
        std::vector<WideNode> tw;
        if (!tb.nodes.empty()) flatten_wide(tb.nodes.data(), (uint32_t)tb.nodes.size(), tb.tris.data(), n_tris, tw);
        *n_wide_out = (uint32_t)tw.size();
    }
    return (int)CRT_OK;
    });
}

int crt_derived_fingerprint(const void* tris_, uint32_t n_tris, const void* nodes_, uint32_t n_nodes, const float* region, uint32_t pad_mode,
                            uint64_t out[2]) {
    if (!tris_ || !nodes_ || !region || !out || pad_mode > 2) return fail(CRT_ERR_INVALID, "crt_derived_fingerprint: bad argument");
    return guarded("crt_derived_fingerprint", [&] {
        const RefTriangle* tris = (const RefTriangle*)tris_;
        const RefNode* nodes = (const RefNode*)nodes_;
        Fingerprint fp;
        auto mix = [&](const void* p, size_t bytes) { const uint64_t n = bytes; fingerprint_mix(fp, &n, sizeof n); fingerprint_mix(fp, p, bytes); };
        std::vector<RefNode> clustered;
        recluster_leaves(nodes, n_nodes, clustered);
        mix(clustered.data(), clustered.size() * sizeof(RefNode));
        std::vector<WideNode> wide;
        flatten_wide(clustered.data(), (uint32_t)clustered.size(), tris, n_tris, wide);
        mix(wide.data(), wide.size() * sizeof(WideNode));
        RawVec<TriIsect> isect;
        RawVec<TriShade> shade;
        make_tri_records(tris, n_tris, nodes, n_nodes, isect, shade);
        mix(isect.data(), isect.size() * sizeof(TriIsect));
        mix(shade.data(), shade.size() * sizeof(TriShade));
        TightBvh tb;
        build_tight_bvh(tris, n_tris, nodes, n_nodes, F3{region[0], region[1], region[2]}, F3{region[3], region[4], region[5]}, (PadMode)pad_mode, tb);
        mix(tb.nodes.data(), tb.nodes.size() * sizeof(RefNode));
        mix(tb.records.data(), tb.records.size() * sizeof(TightRecord));
        mix(tb.boxes.data(), tb.boxes.size() * sizeof(float));
        wide.clear();
        if (!tb.nodes.empty()) flatten_wide(tb.nodes.data(), (uint32_t)tb.nodes.size(), tb.tris.data(), n_tris, wide);
        mix(wide.data(), wide.size() * sizeof(WideNode));
        out[0] = fp.a;
        out[1] = fp.b;
        return (int)CRT_OK;
    });
}

int crt_fingerprint(const void* data, size_t bytes, uint64_t out[2]) {
    if ((!data && bytes) || !out) return fail(CRT_ERR_INVALID, "crt_fingerprint: null argument");
    return guarded("crt_fingerprint", [&] {
        Fingerprint fp;
        fingerprint_mix(fp, data, bytes);
        out[0] = fp.a;
        out[1] = fp.b;
        return (int)CRT_OK;
    });
}

int crt_lane_stats(crt_ctx* c, uint64_t out[8]) {
    if (!c || !out) return fail(CRT_ERR_INVALID, "crt_lane_stats: null argument");
    int rc = crt_sync(c);
    if (rc) return rc;
    DStats st;
    CU(cudaMemcpy(&st, c->stats.p, sizeof st, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; k++) out[k] = st.lanes[k];
    return CRT_OK;
}

int crt_scene_info(const crt_ctx* c, uint64_t* host_rebuilds_out, uint64_t* uploads_out, uint64_t* last_upload_bytes_out) {
    if (!c) return fail(CRT_ERR_INVALID, "crt_scene_info: null ctx");
    if (host_rebuilds_out) *host_rebuilds_out = c->scene_rebuilds;
    if (uploads_out) *uploads_out = c->scene_uploads;
    if (last_upload_bytes_out) *last_upload_bytes_out = c->scene_upload_bytes;
    return CRT_OK;
}

int crt_tight_pad_mode(const crt_ctx* c) { return c && c->tight_nodes ? (c->sc.dyn_pad ? 2 : 1) : 0; }

int crt_tight_info(const crt_ctx* c, uint32_t* n_tight_wide_out, uint32_t* n_split_out) {
    if (!c) return fail(CRT_ERR_INVALID, "crt_tight_info: null ctx");
    if (n_tight_wide_out) *n_tight_wide_out = c->tight_nodes;
    if (n_split_out) *n_split_out = c->tight_split;
    return CRT_OK;
}

}  // extern "C"
