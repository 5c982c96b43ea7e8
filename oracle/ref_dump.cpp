// oracle/ref_dump.cpp - TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A small driver that links the UNMODIFIED reference translation units from
// /root/reference/src (compiled where they lie by oracle/Makefile, output into
// oracle/_ref/) and dumps what the parity tests need as golden data:
//   * the flattened scene arrays exactly as the reference's Scene produces them
//     (Triangle[] in BVH order, LinearBVHNode[], material tables, lights, camera, textures),
//   * the triangle array BEFORE constructBvh() (input of the product's own BVH builder),
//   * closest-hit records from the reference's own intersectBVH (renderer.cu:283-330)
//     for a ray set read from a file,
//   * float radiance sums from the reference's own samplePixel (renderer.cu:119-262) in the
//     loop order of SequentialRenderer::renderOneSamplePerPixel (sequential_renderer.cu:91-116),
//     single thread, srand(seed),
//   * the RGBA8 image through the reference's public Renderer API
//     (renderOneSamplePerPixel x spp, copyImageBytes; terminal_launcher.cpp:10-16).
//
// The reference keeps builder/material state in file statics (bvh.cpp:60-61, scene.cpp:24-25),
// so this program handles exactly ONE scene per process.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <chrono>

#include "scene.h"
#include "renderer.h"
#ifdef _OPENMP
#include <omp.h>
#endif

// Non-static functions of the reference's renderer.cu (renderer.cu:43-62), re-declared to link.
typedef void* dataPtr_t;
__host__ __device__ bool intersectBVH(dataPtr_t p_bvhData, Triangle* p_triangles, SurfaceInteraction& interaction, Ray& ray);

struct RayRec { float ox, oy, oz, dx, dy, dz, tmin, tmax; };   // 32 B, == reference Ray
struct HitRec { float t, u, v; uint32_t tri; };                // 16 B, tri = 0xFFFFFFFF on miss

static void put(FILE* f, const void* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(2); }
}

static const char* arg(int argc, char** argv, const char* key, const char* dflt = nullptr) {
    for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], key)) return argv[i + 1];
    return dflt;
}

int main(int argc, char** argv) {
    const char* obj = arg(argc, argv, "--obj");
    const char* cam = arg(argc, argv, "--camera");
    const char* tex = arg(argc, argv, "--textures");
    if (!obj || !cam || !tex) {
        fprintf(stderr, "usage: ref_dump --obj X.obj --camera X-camera.json --textures DIR/ -w W -h H\n"
                        "  [--scene-out F] [--pretris-out F] [--positions-out F] [--premats-out F] [--rays-in F --hits-out F]\n"
                        "  [--spp S --seed K --accum-out F] [--rgba-out F]\n");
        return 1;
    }
#ifdef _OPENMP
    omp_set_num_threads(1);   // rand() under the reference's omp-parallel loop is only reproducible single-threaded
#endif
    unsigned W = atoi(arg(argc, argv, "-w", "64")), H = atoi(arg(argc, argv, "-h", "64"));

    Scene scene;
    scene.loadObj(obj);
    scene.loadCamera(cam, W, H);
    scene.loadTextures(tex);
    scene.loadTriangles();

    static_assert(sizeof(Triangle) == 128, "reference Triangle is 128 B");
    static_assert(sizeof(LinearBVHNode) == 32, "reference LinearBVHNode is 32 B");
    static_assert(sizeof(Camera) == 72, "reference Camera is 72 B");
    static_assert(sizeof(Ray) == 32, "reference Ray is 32 B");

    uint32_t nTris = scene.getNumTriangles();
    if (const char* p = arg(argc, argv, "--pretris-out")) {
        FILE* f = fopen(p, "wb");
        put(f, scene.getTriPtr(), (size_t)nTris * sizeof(Triangle));
        fclose(f);
    }

    if (const char* p = arg(argc, argv, "--premats-out")) {
        // material floats BEFORE constructLightList scales ka by the total light area (scene.cpp:161-164)
        Material* m0 = scene.getMaterialsPtr();
        FILE* f = fopen(p, "wb");
        for (uint32_t i = 0; i < scene.getNumMaterials(); i++) {
            float3 aux = make_float3(m0[i].ns, m0[i].ni, m0[i].diffuseCoefficient);
            put(f, &m0[i].kd, 12); put(f, &m0[i].ka, 12); put(f, &m0[i].ks, 12); put(f, &aux, 12);
        }
        fclose(f);
    }

    auto tb0 = std::chrono::high_resolution_clock::now();
    scene.constructBvh();
    if (const char* p = arg(argc, argv, "--positions-out")) {
        // the three vertex positions per triangle that createTriangleBboxes boxes (bvh.cpp:79-96), OBJ order
        objl::Vertex* vtx = scene.getVertexPtr();
        uint* idx = scene.getVertexIndicesPtr();
        FILE* f = fopen(p, "wb");
        for (uint32_t i = 0; i < nTris; i++)
            for (int k = 0; k < 3; k++) put(f, &vtx[idx[3 * i + k]].Position, 12);
        fclose(f);
    }
    auto tb1 = std::chrono::high_resolution_clock::now();
    scene.constructLightList();
    fprintf(stderr, "ref_dump: bvh build %.3f s\n", std::chrono::duration<double>(tb1 - tb0).count());

    uint32_t nNodes = scene.getNumBvhNodes(), nMats = scene.getNumMaterials(), nLights = scene.getNumLights();
    uint32_t nTex = scene.getNumTextures();
    float lightsArea = scene.getLightsSurfaceArea();
    Material* mats = scene.getMaterialsPtr();
    // Same packing as SequentialRenderer::createMaterialsData (sequential_renderer.cu:48-64).
    std::vector<float3> matFloats(MAX_MATERIALS * MATERIALS_FLOAT_COMPONENTS);
    std::vector<int2> matIndices(MAX_MATERIALS);
    for (uint32_t i = 0; i < nMats; i++) {
        matFloats[4 * i + 0] = mats[i].kd;
        matFloats[4 * i + 1] = mats[i].ka;
        matFloats[4 * i + 2] = mats[i].ks;
        matFloats[4 * i + 3] = make_float3(mats[i].ns, mats[i].ni, mats[i].diffuseCoefficient);
        matIndices[i] = make_int2((int32_t)mats[i].bsdf, (int32_t)mats[i].texKdIdx);
    }
    pixels_t* texDims = scene.getTextureDimensionsPtr();
    pixels_t* texOffs = scene.getTextureOffsetsPtr();
    // texture_store.cpp:68-70: the flattened array is padded by (width of the last texture) zero texels.
    uint32_t texTexels = nTex ? scene.getTotalTexturePixels() + texDims[2 * (nTex - 1)] : 0;

    if (const char* p = arg(argc, argv, "--scene-out")) {
        FILE* f = fopen(p, "wb");
        uint32_t hdr[16] = {0};
        memcpy(&hdr[0], "CRTS", 4);
        hdr[1] = 1; hdr[2] = nTris; hdr[3] = nNodes; hdr[4] = nMats; hdr[5] = nLights; hdr[6] = nTex;
        hdr[7] = texTexels; hdr[8] = W; hdr[9] = H; memcpy(&hdr[10], &lightsArea, 4);
        put(f, hdr, sizeof hdr);
        put(f, scene.getCameraPtr(), sizeof(Camera));
        put(f, scene.getTriPtr(), (size_t)nTris * sizeof(Triangle));
        put(f, scene.getBvhPtr(), (size_t)nNodes * sizeof(LinearBVHNode));
        put(f, matFloats.data(), (size_t)nMats * 4 * sizeof(float3));
        put(f, matIndices.data(), (size_t)nMats * sizeof(int2));
        put(f, scene.getLightsIndicesPtr(), (size_t)nLights * 4);
        put(f, texDims, (size_t)nTex * 2 * 4);
        put(f, texOffs, (size_t)(nTex + 1) * 4);
        put(f, scene.getTexturePtr(), (size_t)texTexels * sizeof(float3));
        fclose(f);
        fprintf(stderr, "ref_dump: %u tris, %u nodes, %u materials, %u lights (area %g), %u textures\n",
                nTris, nNodes, nMats, nLights, lightsArea, nTex);
    }

    const char* raysIn = arg(argc, argv, "--rays-in");
    const char* hitsOut = arg(argc, argv, "--hits-out");
    if (raysIn && hitsOut) {
        FILE* f = fopen(raysIn, "rb");
        if (!f) { perror(raysIn); return 2; }
        fseek(f, 0, SEEK_END); size_t n = ftell(f) / sizeof(RayRec); fseek(f, 0, SEEK_SET);
        std::vector<RayRec> rays(n);
        if (fread(rays.data(), sizeof(RayRec), n, f) != n) { perror("fread"); return 2; }
        fclose(f);
        std::vector<HitRec> hits(n);
        for (size_t i = 0; i < n; i++) {
            const RayRec& r = rays[i];
            Ray ray(make_float3(r.ox, r.oy, r.oz), make_float3(r.dx, r.dy, r.dz));
            ray.tMin = r.tmin; ray.tMax = r.tmax;
            SurfaceInteraction si = SurfaceInteraction();
            bool hit = intersectBVH((dataPtr_t)scene.getBvhPtr(), scene.getTriPtr(), si, ray);
            hits[i].t = ray.tMax; hits[i].u = hit ? si.u : 0.f; hits[i].v = hit ? si.v : 0.f;
            hits[i].tri = hit ? si.hitTriIdx : 0xFFFFFFFFu;
        }
        f = fopen(hitsOut, "wb"); put(f, hits.data(), n * sizeof(HitRec)); fclose(f);
        fprintf(stderr, "ref_dump: traced %zu rays through the reference intersectBVH\n", n);
    }

    unsigned spp = atoi(arg(argc, argv, "--spp", "0"));
    unsigned seed = atoi(arg(argc, argv, "--seed", "1"));
    if (const char* p = arg(argc, argv, "--accum-out")) {
        // samplePixel called directly, same loop nest as sequential_renderer.cu:93-113, one thread.
        SceneData sd;
        memset(&sd, 0, sizeof sd);
        sd.p_triangles = scene.getTriPtr();
        sd.p_bvh = scene.getBvhPtr();
        sd.p_textureData = scene.getTexturePtr();
        sd.p_textureDimensions = texDims;
        sd.p_textureOffsets = texOffs;
        sd.numBVHNodes = nNodes;
        sd.numTextures = nTex;
        std::vector<float3> accum((size_t)W * H, make_float3(0, 0, 0));
        Camera camera = *scene.getCameraPtr();
        Sampler sampler;
        srand(seed);
        auto t0 = std::chrono::high_resolution_clock::now();
        for (unsigned s = 0; s < spp; s++)
            for (unsigned x = 0; x < W; x++)
                for (unsigned y = 0; y < H; y++) {
                    size_t idx = (size_t)y * W + x;
                    float3 c = samplePixel(x, y, camera, &sd, scene.getLightsIndicesPtr(), nLights, lightsArea,
                                           &sampler, matFloats.data(), matIndices.data());
                    accum[idx] = accum[idx] + c;
                }
        auto t1 = std::chrono::high_resolution_clock::now();
        fprintf(stderr, "ref_dump: samplePixel x %u spp x %ux%u in %.3f s\n", spp, W, H,
                std::chrono::duration<double>(t1 - t0).count());
        FILE* f = fopen(p, "wb"); put(f, accum.data(), accum.size() * sizeof(float3)); fclose(f);
    }
    if (const char* p = arg(argc, argv, "--rgba-out")) {
        // Through the reference's public Renderer seam, as TerminalLauncher::render drives it.
        SequentialRenderer renderer(&scene, W, H, spp);
        srand(seed);
        uchar4* img = renderer.getImgBytesPointer();
        for (unsigned s = 0; s < spp; s++) renderer.renderOneSamplePerPixel(img);
        renderer.copyImageBytes(img);
        FILE* f = fopen(p, "wb"); put(f, renderer.h_imgPtr, (size_t)W * H * 4); fclose(f);
    }
    return 0;
}
