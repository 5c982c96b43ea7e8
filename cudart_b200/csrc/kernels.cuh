// cudart_b200/csrc/kernels.cuh - device-side data model and kernel launchers of the wavefront path tracer.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "host_scene.h"

namespace crt {

// Read-only scene, resident in HBM (L2/L1 for the bundled scenes).
struct DScene {
    const float4* wide;        // n_wide x 8 float4  (WideNode)
    const float4* tri_isect;   // n_tris x 3 float4  (TriIsect)
    const float4* tri_shade;   // n_tris x 5 float4  (TriShade)
    const float4* tri_tight;   // n_tris x 3 float4  (v1|index, e1|leafFirst, e2|material in tight-BVH order); null when the tight BVH is off
    const float4* tri_gate;    // n_tris x 2 float4  (tight-BVH order): the exact box of the triangle's reference leaf, read only
                               // for a candidate that would become the closest hit
    const float4* mats;        // n_mats x 4 float4: (kd, bsdf) (ka, texKdIdx) (ks, 0) (ns, ni, diffuseCoefficient, 0)
    const uint32_t* lights;
    const float4* light_boxes; // per light: the reference leaf box holding it (lo.xyz, hi.xyz), for the shadow seed
    const float* tex_data;     // float3 texels
    const uint32_t* tex_dims;
    const uint32_t* tex_offsets;
    uint32_t n_wide, n_tris, n_mats, n_lights, n_tex, tex_texels;
    uint32_t n_smem_nodes;     // the first n_smem_nodes wide nodes (top levels, breadth-first) are staged in shared memory
    uint32_t raygen_block;     // camera rays are queued in 8 x 4-pixel tiles inside blocks of this edge (0: plain row order)
    float lights_area;
    int coords_bounded;        // every node/triangle coordinate is finite and below 1e8 in magnitude
    // Tight BVH (host_scene.h): its wide nodes are [0, ref_root), the reference-leaf tree starts at ref_root.  Rays that
    // qualify (culling allowed, origin inside the region the inflation was computed for) start at node 0.
    uint32_t ref_root;
    int dyn_pad;               // tight-tree boxes are not inflated: pad each child box per visit (host_scene.h kPadDynamic)
    float region_lo[3], region_hi[3];
    RefCamera cam;
};

// Which pixels this context owns: interleaved row tiles (tile k -> rank k % world).
struct DTiling {
    uint32_t W, H, rank, world, tile_rows;
    uint32_t local_rows;       // rows owned (valid ones; they are a prefix of the padded local rows)
    uint32_t local_pixels;     // local_rows * W
    uint32_t padded_pixels;    // identical on all ranks: ceil(n_tiles / world) * tile_rows * W
};

constexpr int kMaxBounces = 6;                 // renderer.cu:25,33

// Per-wave counters (device memory, zeroed once per wave).
struct DCounters {
    uint32_t qcount[kMaxBounces + 1];          // extend queue sizes per bounce
    uint32_t scount[kMaxBounces];              // shadow queue sizes per bounce
    uint32_t ehead[kMaxBounces];               // persistent-fetch heads
    uint32_t shead[kMaxBounces];
};
struct DStats {                                // running totals (never reset by a wave)
    unsigned long long extend_rays, shadow_rays, node_visits, tri_tests, samples;
    // CRT_FLAG_STATS: warp-level loop iterations and lane occupancy of the traversal (crt_lane_stats):
    // [0] node-loop iterations [1] lanes working in them [2] lanes idle (no ray) during them
    // [3] leaf-loop (triangle) iterations [4] lanes working [5] lanes idle   [6] refills [7] rays fetched by them
    unsigned long long lanes[8];
};

// Wavefront state for up to `capacity` paths in flight.  Everything is SoA at 16-byte granularity:
// consecutive lanes touch consecutive float4s.
struct DWave {
    float4* rq_o[2];           // extend queue (ping-pong): origin.xyz, tMin
    float4* rq_d[2];           //                           dir.xyz,    tMax
    uint32_t* rq_path[2];      //                           path id
    float4* rq_m[2];           //                           throughput.rgb, meta bits (dim | bounce<<8 | prevBsdf<<12)
    uint2* rq_k[2];            //                           RNG key (include/crt_rng.h)
    float4* hits;              // t, u, v, tri (bits) per extend-queue slot
    float4* sq_o;              // shadow queue: origin.xyz, tMin
    float4* sq_d;              //               dir.xyz, tMax
    float4* sq_c;              //               throughput.rgb after the diffuse bounce, light ordinal (bits)
    uint32_t* sq_path;
    float4* color;             // per path: radiance accumulated so far
    DCounters* counters;
    DStats* stats;
    uint32_t capacity;
};

struct LaunchCfg { int sms; cudaStream_t stream; int trav_blocks; /* traversal CTAs per SM, 0 = the compiled default */ };

constexpr uint32_t kMaxSmemNodes = 256;        // 256 x 144 B = 36 KB of nodes + 20 KB of stack (+ 1 KB reserved) per CTA, 4 CTAs per SM: 228 KB
int configure_kernels();                       // opt the traversal kernels into > 48 KB of dynamic shared memory

void launch_raygen(const DScene& sc, const DTiling& tl, const DWave& wv, uint32_t first_sample, uint32_t wave_spp,
                   uint32_t seed, const LaunchCfg& cfg);
// One persistent traversal launch for the extend queue of `ebounce` and the shadow queue of `sbounce` (-1: none).
void launch_traverse(const DScene& sc, const DWave& wv, int ebounce, int sbounce, bool exact, bool stats, const LaunchCfg& cfg);
void launch_shade(const DScene& sc, const DTiling& tl, const DWave& wv, int bounce, const LaunchCfg& cfg);
void launch_accumulate(const DTiling& tl, const DWave& wv, float4* accum, uint32_t wave_spp, const LaunchCfg& cfg);
// Stand-alone closest-hit over a flat ray array (crt_trace_closest): rays as 8 floats, hits as (t,u,v,tri).
void launch_trace(const DScene& sc, const float4* rays, uint32_t n, float4* hits, uint32_t* head, DStats* stats,
                  bool exact, bool with_stats, const LaunchCfg& cfg);
// Resolve: accum is rank-major (world x padded_pixels float4) when gathered, else this rank's compact buffer.
void launch_resolve(const DTiling& tl, const float4* accum, bool gathered, uint32_t spp, const uint8_t* gamma_lut,
                    uchar4* rgba_out, float* f32_out, int f32_mode /*0 none,1 mean,2 sum*/, const LaunchCfg& cfg);
// Multi-GPU read-out in display bytes: this rank's owned pixels as RGBA8 in its compact tile-major order
// (padded_pixels uchar4, identical size on every rank; pad slots zero) ...
void launch_resolve_local(const DTiling& tl, const float4* accum, uint32_t spp, const uint8_t* gamma_lut, uchar4* rgba_local,
                          const LaunchCfg& cfg);
// ... and, on the rank that collected them (rank-major: world x padded_pixels uchar4), the de-interleaved W x H image.
void launch_assemble(const DTiling& tl, const uchar4* gathered, uchar4* rgba_out, const LaunchCfg& cfg);
}  // namespace crt
