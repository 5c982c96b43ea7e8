// integration/b200_renderer.cpp - see b200_renderer.h.  Two things live here:
//
//  1. B200Renderer, the documented binding: `p_renderer = new B200Renderer(&scene, width, height, samples);` in
//     main.cpp:163 is the one line a maintainer changes.
//  2. A link-time stand-in for the reference's ParallelRenderer (declared in renderer.h:102-127, defined in
//     renderer/parallel_renderer.cu): the same five member functions, implemented by delegating to a B200Renderer.
//     Linking THIS file instead of parallel_renderer.cu lets the reference's main.cpp - not one character changed -
//     construct "ParallelRenderer" (main.cpp:163) and drive libcrt_b200.so through TerminalLauncher::render
//     (terminal_launcher.cpp:10-16) and Launcher::saveToImage (launcher.cpp:13-21).  tests/test_dropin.py checks that
//     the PNG it writes holds exactly the pixels of the repo's own CLI.
//
// Errors: the reference prints and exit(1)s on any CUDA failure (cuda_error_check.h:15-21); so does this adapter, on top
// of the library's error codes.
#include "b200_renderer.h"

#include <cstdlib>
#include <cstring>
#include <iostream>

static void check(int rc) {
    if (rc != CRT_OK) {
        std::cerr << "libcrt_b200: " << crt_last_error() << std::endl;
        exit(1);
    }
}

B200Renderer::B200Renderer(Scene* s, pixels_t w, pixels_t h, uint spp) : Renderer(s, w, h, spp) {
    useCuda = true;
    createMaterialsData();
    crt_scene_desc d;
    memset(&d, 0, sizeof d);
    d.tris = s->getTriPtr();                    d.n_tris = (uint32_t)s->getNumTriangles();     // scene.h:27,33
    d.nodes = s->getBvhPtr();                   d.n_nodes = s->getNumBvhNodes();               // scene.h:30,38
    d.mat_floats = (const float*)matFloats;     d.mat_indices = (const int32_t*)matIndices;
    d.n_mats = s->getNumMaterials();                                                           // scene.h:31
    d.light_indices = s->getLightsIndicesPtr(); d.n_lights = (uint32_t)s->getNumLights();      // scene.h:28,35
    d.lights_area = s->getLightsSurfaceArea();                                                 // scene.h:32
    d.camera = s->getCameraPtr();                                                              // scene.h:41
    d.n_tex = s->getNumTextures();                                                             // scene.h:44
    if (d.n_tex) {
        d.tex_data = (const float*)s->getTexturePtr();                                         // scene.h:47
        d.tex_dims = s->getTextureDimensionsPtr();
        d.tex_offsets = s->getTextureOffsetsPtr();
        // the flattened store is padded by (width of the last texture) zero texels, texture_store.cpp:68-70
        d.tex_texels = s->getTotalTexturePixels() + d.tex_dims[2 * (d.n_tex - 1)];
    }
    check(crt_create(&d, w, h, nullptr, &ctx));
}

B200Renderer::~B200Renderer() { crt_destroy(ctx); }

void B200Renderer::createMaterialsData() {
    Material* p_materials = p_scene->getMaterialsPtr();
    const uint n = p_scene->getNumMaterials();
    for (uint i = 0; i < n && i < MAX_MATERIALS; i++) {
        matFloats[i * MATERIALS_FLOAT_COMPONENTS + KD_OFFSET] = p_materials[i].kd;
        matFloats[i * MATERIALS_FLOAT_COMPONENTS + KA_OFFSET] = p_materials[i].ka;
        matFloats[i * MATERIALS_FLOAT_COMPONENTS + KS_OFFSET] = p_materials[i].ks;
        matFloats[i * MATERIALS_FLOAT_COMPONENTS + AUX_OFFSET] = make_float3(p_materials[i].ns, p_materials[i].ni, p_materials[i].diffuseCoefficient);
        matIndices[i] = make_int2((int32_t)p_materials[i].bsdf, (int32_t)p_materials[i].texKdIdx);
    }
}

// One call per sample, as the reference's launchers make it.  The library queues single samples and launches whole
// wavefront batches; the counter the launcher reads advances by one per call all the same (renderer.h:91,99).
void B200Renderer::renderOneSamplePerPixel(uchar4*) {
    check(crt_render(ctx, 1));
    samplesRendered++;
}

void B200Renderer::resolveInto(uchar4* dst) { check(crt_resolve_rgba8(ctx, (uint8_t*)dst)); }
void B200Renderer::copyImageBytes(uchar4*) { resolveInto(h_imgPtr); }

// ---- link-time stand-in for parallel_renderer.cu ---------------------------------------------------------------------
// ParallelRenderer's private members (renderer.h:112-121) are device pointers of the renderer this file replaces; the
// stand-in keeps its B200Renderer in the first of them and leaves the others null.
static B200Renderer*& impl_of(float3*& slot) { return reinterpret_cast<B200Renderer*&>(slot); }

ParallelRenderer::ParallelRenderer(Scene* s, pixels_t w, pixels_t h, uint spp) : Renderer(s, w, h, spp) {
    useCuda = true;
    d_imgBytesPtr = nullptr;       // getImgBytesPointer() hands this to the launcher, which only passes it back to us
    d_sceneData = nullptr; d_triPtr = nullptr; d_lightsIndices = nullptr; d_cudaTexObjects = nullptr;
    d_camPtr = nullptr; d_curandStatePtr = nullptr; threadsPerBlock = 0; gridBlocks = 0;
    impl_of(d_imgVectorPtr) = new B200Renderer(s, w, h, spp);
}
ParallelRenderer::~ParallelRenderer() { delete impl_of(d_imgVectorPtr); }
void ParallelRenderer::renderOneSamplePerPixel(uchar4* p_img) {
    impl_of(d_imgVectorPtr)->renderOneSamplePerPixel(p_img);
    samplesRendered++;
}
void ParallelRenderer::copyImageBytes(uchar4*) { impl_of(d_imgVectorPtr)->resolveInto(h_imgPtr); }
void ParallelRenderer::createMaterialsData() {}
