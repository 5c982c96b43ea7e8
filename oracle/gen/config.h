/* Mirrors the reference's src/include/config.h.in (version 0.1.0), headless build. */
#define CUDA_RT_VERSION 0
#define CUDA_RT_VERSION_MAJOR 1
#define CUDA_RT_VERSION_MINOR 0
#define SKIP_OPENGL
