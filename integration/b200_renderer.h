// integration/b200_renderer.h - the reference-side binding of libcrt_b200.so, COMPILED against the reference's own
// headers (src/include/renderer.h:78-100, scene.h:26-49): a third renderer next to SequentialRenderer and
// ParallelRenderer.  Everything it does is forward the reference's Scene getters to the C ABI (include/crt_abi.h).
// This is the class INTEGRATION.md describes; integration/Makefile builds it into integration/_build/CudaRT_dropin
// together with the reference's UNMODIFIED main.cpp, scene.cpp, bvh.cpp, launcher.cpp and terminal_launcher.cpp.
#pragma once
#include "renderer.h"          // the reference's (found through -I$(REF)/include)

#include "../include/crt_abi.h"

class B200Renderer : public Renderer {
public:
    B200Renderer(Scene* scene, pixels_t w, pixels_t h, uint spp);
    ~B200Renderer();
    void renderOneSamplePerPixel(uchar4* p_img) override;     // renderer.h:84: one more sample per pixel
    void copyImageBytes(uchar4* p_img) override;              // renderer.h:85: RGBA8, gamma-corrected, into h_imgPtr
    uchar4* getImgBytesPointer() override { return h_imgPtr; }
    void createMaterialsData() override;                      // the packing of sequential_renderer.cu:48-64
    // the same read-out into a caller's W*H buffer (the ParallelRenderer shim below owns its own h_imgPtr)
    void resolveInto(uchar4* dst);
private:
    crt_ctx* ctx = nullptr;
    float3 matFloats[MAX_MATERIALS * MATERIALS_FLOAT_COMPONENTS];
    int2 matIndices[MAX_MATERIALS];
};
