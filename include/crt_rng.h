/* include/crt_rng.h - the counter-based uniform stream of the B200 path tracer.
 *
 * The reference draws uniforms from one shared glibc rand() on the host (renderer.cu:76) and from
 * 256 racily shared XORWOW states on the device (parallel_renderer.cu:52,70,328-330).
 * The new path replaces both by a stateless stream keyed by (seed, global pixel index, sample index)
 * and indexed by the draw number inside the path ("dimension"), so that
 *   - the image does not depend on how pixels are scheduled or sharded across GPUs, and
 *   - the CPU oracle (oracle/crt_oracle.c, ORC_RNG_COUNTER) can replay exactly the same paths.
 *
 * Key: 64 bits, two words.  For a fixed seed, pixel -> k0 and sample -> k1 are bijections of 32-bit words, so
 * no two (pixel, sample) pairs of a render share a key (BASELINE config 5 has 8.5e9 paths: more than a 32-bit
 * key space holds).  Draw `dim` of a path is  mix(k1 + mix(k0 ^ dim * phi)) : the dimension enters through a
 * hash, not as an offset added to the key, so the streams of two paths are never shifted copies of one another
 * (tests/test_host_logic.py checks 1e7 keys).  32-bit integer multiplies and shifts only.
 * Uniforms lie in [0,1) with 24 random bits (the reference's lie in [0,1] resp. (0,1]).
 *
 * Plain C99; usable from C, C++ and CUDA.
 */
#ifndef CRT_RNG_H
#define CRT_RNG_H
#include <stdint.h>

#if defined(__CUDACC__)
#define CRT_HD __host__ __device__ __forceinline__
#else
#define CRT_HD static inline
#endif

/* 32-bit finaliser ("lowbias32"): a bijection with full avalanche in two multiply/xorshift rounds. */
CRT_HD uint32_t crt_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

typedef struct crt_rng_key_t { uint32_t k0, k1; } crt_rng_key_t;

/* Per-path key. */
CRT_HD crt_rng_key_t crt_rng_key(uint32_t seed, uint32_t pixel, uint32_t sample) {
    const uint32_t s = crt_mix32(seed ^ 0x85ebca6bu);
    crt_rng_key_t k;
    k.k0 = crt_mix32(pixel ^ s);
    k.k1 = crt_mix32((sample + 0x7f4a7c15u) ^ ((s << 16) | (s >> 16)));
    return k;
}

/* dim-th draw of a path. */
CRT_HD uint32_t crt_rng_bits(crt_rng_key_t key, uint32_t dim) {
    return crt_mix32(key.k1 + crt_mix32(key.k0 ^ (dim * 0x9e3779b9u)));
}

CRT_HD float crt_rng_uniform(crt_rng_key_t key, uint32_t dim) {
    return (float)(crt_rng_bits(key, dim) >> 8) * (1.0f / 16777216.0f);
}

#endif
