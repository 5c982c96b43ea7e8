/* include/crt_abi.h - C ABI of the B200-native path-tracing hot path (libcrt_b200.so).
 *
 * This is the drop-in boundary behind the reference's `Renderer` seam
 * (mattboran/CudaRT src/include/renderer.h:78-100): a host adapter with the reference's five-call
 * contract (renderOneSamplePerPixel / copyImageBytes / getImgBytesPointer / createMaterialsData /
 * getSamplesRendered) forwards to these functions; INTEGRATION.md shows that adapter.  Everything
 * crossing the boundary is plain pointers and sizes, in the reference's own POD layouts
 * (Triangle 128 B geometry.h:30-51, LinearBVHNode 32 B bvh.h:18-27, Camera 72 B camera.h:9-22,
 * Ray 32 B geometry.h:17-27, material tables as packed by sequential_renderer.cu:48-64), so the
 * reference's Scene getters (scene.h:26-49) can be passed through unchanged.
 *
 * There is no CPU fallback: every entry point that computes fails with CRT_ERR_CUDA when no sm_100
 * device is usable.  All functions return 0 on success, a CRT_ERR_* code otherwise and never throw
 * or exit (the reference prints and exit(1)s, cuda_error_check.h:15-21; an adapter may do that on top):
 * C++ exceptions of the host-side builders (bad_alloc, a thread that cannot be spawned) are caught at the
 * boundary, on worker threads too, and come back as CRT_ERR_NOMEM / CRT_ERR_INVALID.
 * A context is not re-entrant; different contexts may be driven from different threads.
 */
#ifndef CRT_ABI_H
#define CRT_ABI_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRT_ABI_VERSION 1

enum {
    CRT_OK = 0,
    CRT_ERR_INVALID = 1,   /* bad argument                                   */
    CRT_ERR_CUDA = 2,      /* CUDA runtime failure / no usable device        */
    CRT_ERR_NOMEM = 3,
    CRT_ERR_LIMIT = 4,     /* scene exceeds a documented limit               */
    CRT_ERR_IO = 5
};

/* Scene arrays exactly as the reference's Scene exposes them after loadTriangles / constructBvh /
 * constructLightList (main.cpp:133-139).  crt_create COPIES everything; the caller keeps ownership. */
typedef struct crt_scene_desc {
    const void* tris;             /* Triangle[n_tris], BVH order            (Scene::getTriPtr)            */
    uint32_t n_tris;
    const void* nodes;            /* LinearBVHNode[n_nodes], pre-order      (Scene::getBvhPtr)            */
    uint32_t n_nodes;
    const float* mat_floats;      /* n_mats x 4 x float3: kd, ka, ks, (ns, ni, diffuseCoefficient)       */
    const int32_t* mat_indices;   /* n_mats x int2: (bsdf refl_t, texKdIdx or -1)                         */
    uint32_t n_mats;              /* <= 128 (MAX_MATERIALS, renderer.h:22)                                */
    const uint32_t* light_indices;/* indices of EMISSIVE triangles          (Scene::getLightsIndicesPtr)  */
    uint32_t n_lights;
    float lights_area;            /* Scene::getLightsSurfaceArea                                          */
    const void* camera;           /* Camera, 72 B                           (Scene::getCameraPtr)         */
    const float* tex_data;        /* float3 texels, flattened + padded      (Scene::getTexturePtr)        */
    const uint32_t* tex_dims;     /* 2 x n_tex                                                            */
    const uint32_t* tex_offsets;  /* n_tex + 1                                                            */
    uint32_t n_tex;
    uint32_t tex_texels;          /* float3 count in tex_data                                             */
} crt_scene_desc;

/* Optional knobs; pass NULL for defaults. Zero-initialise, set struct_size = sizeof(crt_options). */
typedef struct crt_options {
    uint32_t struct_size;
    int32_t device;           /* CUDA device ordinal, -1 = current device                                  */
    uint32_t rank, world;     /* image-space sharding: this context renders the row tiles k with
                                 k % world == rank (world = 0 or 1: the whole image)                       */
    uint32_t tile_rows;       /* rows per tile (default 8)                                                 */
    uint32_t seed;            /* RNG seed (default 1234, the reference's curand seed)                      */
    uint32_t spp_per_wave;    /* samples per pixel traced per wavefront batch (0 = auto)                   */
    uint32_t flags;           /* CRT_FLAG_*                                                                */
    void* stream;             /* cudaStream_t to enqueue on (NULL: a private non-blocking stream)          */
    uint32_t waves_in_flight; /* wavefront batches in flight, 1 or 2 (0 = default: 2, or 1 with CRT_FLAG_PROFILE).
                                 Batches run on private work streams, consecutive ones (of one call or of
                                 consecutive calls) on alternating streams, and overlap; `stream` waits for
                                 every call's batches, so everything a caller can observe stays ordered on it,
                                 and the image is unchanged                                                   */
} crt_options;

#define CRT_FLAG_EXACT_VISIT 1u  /* no distance culling: visit exactly the leaves the reference visits     */
#define CRT_FLAG_PROFILE 2u      /* time every stage with CUDA events (crt_get_counters.stage_ms); one batch in flight */
#define CRT_FLAG_REFERENCE_TOPOLOGY 8u /* keep the reference's interior BVH structure (no SAH re-clustering of its leaves) */
#define CRT_FLAG_NO_SMEM_NODES 16u /* never stage BVH nodes in shared memory (default: staged when the whole wide tree fits) */
#define CRT_FLAG_SMEM_TOP_LEVELS 32u /* stage the top levels even when the tree does not fit entirely (A/B measurements) */
#define CRT_FLAG_TIGHT_BVH 64u   /* force the tight per-triangle BVH (default: used when few triangles are too large for it) */
#define CRT_FLAG_NO_TIGHT_BVH 128u /* never use the tight BVH: traverse the reference's leaves only            */
#define CRT_FLAG_DYNAMIC_PAD 256u /* tight BVH with un-inflated boxes, padded per node visit in proportion to the distance */
#define CRT_FLAG_STATIC_PAD 512u  /* tight BVH with the pad baked into its boxes (default: chosen per scene) */
#define CRT_FLAG_STATS 4u        /* count wide-node fetches and triangle tests (slows traversal a little)  */

typedef struct crt_ray { float ox, oy, oz, dx, dy, dz, tmin, tmax; } crt_ray;   /* == reference Ray       */
typedef struct crt_hit { float t, u, v; uint32_t tri; } crt_hit;                /* tri = CRT_MISS: t = tmax */
#define CRT_MISS 0xFFFFFFFFu

enum { CRT_STAGE_RAYGEN = 0, CRT_STAGE_EXTEND, CRT_STAGE_SHADE, CRT_STAGE_SHADOW, CRT_STAGE_ACCUM, CRT_STAGE_COUNT };

typedef struct crt_counters {
    uint64_t samples;         /* path samples started (one = one samplePixel of the reference)            */
    uint64_t extend_rays;     /* closest-hit queries from the bounce loop                                 */
    uint64_t shadow_rays;     /* closest-hit queries from next-event estimation                           */
    uint64_t node_visits;     /* wide nodes fetched (only with CRT_FLAG_STATS)                            */
    uint64_t tri_tests;       /* triangle tests (only with CRT_FLAG_STATS)                                */
    uint64_t kernel_launches; /* kernels launched by this context so far                                  */
    double render_ms;         /* device time spent in crt_render calls so far (CUDA events)               */
    double stage_ms[CRT_STAGE_COUNT];   /* per-stage device time (CRT_FLAG_PROFILE only)                  */
} crt_counters;

typedef struct crt_ctx crt_ctx;

int crt_abi_version(void);
const char* crt_last_error(void);          /* thread-local message of the last failure */

/* ---- renderer (replaces ParallelRenderer, parallel_renderer.cu:37-310) ---------------------- */
int crt_create(const crt_scene_desc* scene, uint32_t width, uint32_t height, const crt_options* opt, crt_ctx** out);
void crt_destroy(crt_ctx* ctx);
/* Trace `spp` more samples for every pixel this context owns and add them to its accumulator
 * (== `spp` calls of Renderer::renderOneSamplePerPixel). Asynchronous; ordered on the ctx stream.
 * The reference's launcher asks for ONE sample per call (terminal_launcher.cpp:12-14) while the machine wants
 * whole wavefront batches (spp_per_wave samples per pixel, waves_in_flight of them at a time): calls that add up to
 * less than that are queued on the host and launched by the call that completes it, by crt_flush / crt_sync, or by
 * any read-out.
 * crt_samples_rendered advances with every call (Renderer::samplesRendered, renderer.h:91,99); sample indices,
 * and with them the image, do not depend on how calls were grouped. */
int crt_render(crt_ctx* ctx, uint32_t spp);
/* Launch whatever crt_render queued (asynchronous).  crt_sync = crt_flush + wait for the ctx stream. */
int crt_flush(crt_ctx* ctx);
int crt_sync(crt_ctx* ctx);
uint32_t crt_samples_rendered(const crt_ctx* ctx);           /* Renderer::getSamplesRendered */
uint32_t crt_spp_per_wave(const crt_ctx* ctx);               /* samples per pixel in one wavefront batch of this context */
/* Renderer::copyImageBytes: W*H RGBA8 into host memory, row 0 = top, float3ToUchar4(sum/spp)
 * (renderer.h:149-156) then gammaCorrectPixel (renderer.cu:495-504), alpha 255.  When world > 1 the rows
 * owned by other ranks are written as zeros (use crt_resolve_gathered / crt_assemble_rgba8 for the whole image). */
int crt_resolve_rgba8(crt_ctx* ctx, uint8_t* dst);
/* W*H*3 floats: radiance sums divided by the samples rendered (no clamp, no gamma). */
int crt_resolve_f32(crt_ctx* ctx, float* dst);
/* Raw radiance SUMS, W*H*3 floats (what SequentialRenderer keeps in h_imgVectorPtr). */
int crt_read_accum(crt_ctx* ctx, float* dst);
int crt_reset(crt_ctx* ctx);                                  /* zero accumulator and sample count */
/* Replace the scene of a live context (next frame of an animation: new geometry / camera / materials), keeping the
 * image size, sharding and all device buffers; copies host -> device like crt_create and resets the accumulator.
 * The arrays derived on the host (4-ary trees, triangle records) are kept per context and only rebuilt when the
 * triangles, BVH nodes, light list or camera position differ from the previous call; they are uploaded every call. */
int crt_update_scene(crt_ctx* ctx, const crt_scene_desc* scene);
/* The 128-bit fingerprint crt_update_scene compares (host only, no device needed): introspection for tests. */
int crt_fingerprint(const void* data, size_t bytes, uint64_t out[2]);
/* How often this context rebuilt its derived host arrays, how often it uploaded a scene, and the host -> device bytes
 * of the last upload. */
int crt_scene_info(const crt_ctx* ctx, uint64_t* host_rebuilds_out, uint64_t* uploads_out, uint64_t* last_upload_bytes_out);
int crt_get_counters(crt_ctx* ctx, crt_counters* out);

/* ---- bit-exact test entry: closest hit of the reference traversal (renderer.cu:283-330) -------- */
/* rays/hits are HOST arrays. flags: 0 = the render path's traversal, CRT_FLAG_EXACT_VISIT = no culling. */
int crt_trace_closest(crt_ctx* ctx, const crt_ray* rays, size_t n, crt_hit* hits, uint32_t flags);
/* Same on DEVICE arrays (no copies), for timing the extend stage alone; kernel_ms may be NULL. */
int crt_trace_closest_device(crt_ctx* ctx, const void* d_rays, size_t n, void* d_hits, uint32_t flags, float* kernel_ms);

/* ---- multi-GPU plumbing (one context per GPU; the caller owns the collective) ---------------- */
/* The context's compact accumulator: float4 (r,g,b sums, unused) per owned pixel, tile-major.
 * `padded_pixels` is identical on every rank, so an all-gather of padded_pixels*16 bytes per rank
 * yields the buffer crt_resolve_gathered expects. */
int crt_accum_device_ptr(crt_ctx* ctx, void** d_ptr, size_t* padded_pixels);   /* launches queued samples first */
/* d_gathered: world x padded_pixels float4 on this context's device (rank-major).
 * Writes the full image; rgba8_dst and/or f32_dst may be NULL. */
int crt_resolve_gathered(crt_ctx* ctx, const void* d_gathered, uint8_t* rgba8_dst, float* f32_dst);
/* Read-out in display bytes (4 B per pixel instead of 16): every rank resolves its OWN pixels to RGBA8 on its
 * device - *d_ptr receives a context-owned buffer of *bytes = padded_pixels * 4 bytes, compact tile-major like the
 * accumulator, pad slots zero; asynchronous on the ctx stream - the caller collects the world buffers rank-major on one
 * rank (ncclSend / ncclRecv, torch.distributed.gather), and that rank's crt_assemble_rgba8 de-interleaves them into the
 * W*H RGBA8 host image (blocking, like crt_resolve_rgba8).  Byte-identical to crt_resolve_gathered's RGBA8. */
int crt_resolve_local_rgba8_device(crt_ctx* ctx, void** d_ptr, size_t* bytes);
int crt_assemble_rgba8(crt_ctx* ctx, const void* d_gathered, uint8_t* rgba8_dst);
/* Stream ordering for a caller that runs its collective on ANOTHER stream than the context's, without blocking the
 * host: crt_stream_wait_ctx makes `stream` wait for everything enqueued on the ctx stream so far (queued samples are
 * launched first); crt_ctx_wait_stream makes the ctx stream wait for everything enqueued on `stream` so far.
 * crt_resolve_gathered / crt_assemble_rgba8 run on the ctx stream: call crt_ctx_wait_stream after the collective. */
int crt_stream_wait_ctx(crt_ctx* ctx, void* stream);
int crt_ctx_wait_stream(crt_ctx* ctx, void* stream);

/* ---- host scene build (replaces scene.cpp:117-165 + bvh.cpp:65-323; no GPU needed) ------------- */
/* In: Triangle[n] in OBJ order with v1/e1/e2/normals/uvs/material ids/area filled (scene.cpp:93-134)
 * and positions[9*n], the three vertex positions of every triangle (the reference boxes those,
 * bvh.cpp:79-96; v1+e1 would not round to v2).
 * Out: tris reordered in place into BVH order, nodes_out (capacity 2*n) in the reference's pre-order
 * layout, *n_nodes_out.  Reproduces the zero-initialised reference tree bit for bit. */
int crt_build_bvh(void* tris, const float* positions, uint32_t n_tris, void* nodes_out, uint32_t* n_nodes_out);
/* Scene::constructLightList (scene.cpp:148-165): light_indices_out (capacity n_tris), lights area,
 * _triId rewrite for lights and ka *= area on mat_floats (in place). */
int crt_build_lights(void* tris, uint32_t n_tris, float* mat_floats, const int32_t* mat_indices, uint32_t n_mats,
                     uint32_t* light_indices_out, uint32_t* n_lights_out, float* lights_area_out);
/* Introspection of the wide (4-ary) flattening, for tests: writes up to cap 128-byte nodes. */
int crt_wide_bvh(const void* nodes, uint32_t n_nodes, void* wide_out, uint32_t cap, uint32_t* n_wide_out);
/* Introspection of the tight per-triangle BVH (no reference counterpart), for tests: region = lo.xyz, hi.xyz of
 * the box that holds every ray origin; pad_mode 0 = chosen per scene, 1 = pad baked into the boxes, 2 = tight boxes padded
 * by the traversal (*dynamic_out says which one was built); tri_boxes_out (6 floats per triangle, reference order) receives
 * the box a ray must pass before the triangle is tested (the triangle's own box, or its reference leaf's exact box when the
 * leaf was kept whole); tri_pad_out (2 floats per triangle) the terms (a1, b2) of the run-time pad a1 * |T|max + b2 the
 * kernel would apply to a node holding only this triangle; *n_split_out = triangles with their own box; binary / wide
 * node counts; sah_out = surface-area-heuristic (nodes, triangles) per random line for the tight tree and for the
 * re-clustered reference-leaf tree. */
int crt_tight_bvh(const void* tris, uint32_t n_tris, const void* nodes, uint32_t n_nodes, const float* region, uint32_t pad_mode,
                  float* tri_boxes_out, float* tri_pad_out, uint32_t* dynamic_out,
                  uint32_t* n_split_out, uint32_t* n_tight_nodes_out, uint32_t* n_wide_out, double* sah_out /* 4 doubles or null */);
/* Fingerprint (crt_fingerprint's hash) of every array crt_create derives on the host from these triangles and nodes: the
 * re-clustered reference-leaf tree and its 4-ary flattening, the intersection / shading records, the tight BVH (nodes,
 * records, boxes; region / pad_mode as in crt_tight_bvh) and its flattening.  Host only; the tests use it to pin that the
 * threaded builders produce the same bytes whatever the number of threads. */
int crt_derived_fingerprint(const void* tris, uint32_t n_tris, const void* nodes, uint32_t n_nodes, const float* region, uint32_t pad_mode,
                            uint64_t out[2]);
/* 0 = this context's render path does not walk the tight BVH, 1 = static pad, 2 = dynamic pad. */
int crt_tight_pad_mode(const crt_ctx* ctx);
/* Whether this context's render path walks the tight BVH: its wide-node count (0 = off) and individually boxed triangles. */
int crt_tight_info(const crt_ctx* ctx, uint32_t* n_tight_wide_out, uint32_t* n_split_out);
/* With CRT_FLAG_STATS: where the traversal's SIMT lanes go.  out[0..2] = warp-level iterations of the interior-node loop,
 * lanes that worked in them, lanes that had no ray; out[3..5] = the same for the triangle loop; out[6..7] = queue
 * refills and the rays they fetched. */
int crt_lane_stats(crt_ctx* ctx, uint64_t out[8]);

#ifdef __cplusplus
}
#endif
#endif
