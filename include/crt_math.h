/* include/crt_math.h - the ONE sine/cosine the B200 path tracer's shading uses, shared with its CPU checker.
 *
 * The reference calls libm's cosf/sinf on a uniform angle (renderer.cu:384, camera.cu:48-49).  glibc's and CUDA's
 * implementations differ in the last bit, which is enough to fork a path at the next Russian-roulette or Fresnel
 * decision.  The product therefore evaluates the function below - explicit fmaf/multiply steps only, every one a single
 * IEEE rounding on the host and on the device - and the oracle's replay mode (oracle/crt_oracle.c, ORC_RNG_COUNTER)
 * evaluates the same function, so that a same-seed replay is BIT-EXACT (tests/test_gpu_parity.py).  The oracle's
 * rand() mode keeps libm: that mode is pinned against the reference's own output.
 *
 * Valid for |a| <= 64 (the callers pass angles in [-pi, 2 pi)); maximum error 1.5 ulp (tests/test_host_logic.py).
 * Two-term Cody-Waite reduction by pi/2 with fused multiply-adds, then the classic degree-7 / degree-8 minimax
 * polynomials on [-pi/4, pi/4].
 *
 * Must be compiled without floating-point contraction (nvcc -fmad=false, gcc -ffp-contract=off): the products that are
 * not written as fmaf() have to stay products.
 */
#ifndef CRT_MATH_H
#define CRT_MATH_H
#include <math.h>

#if defined(__CUDACC__)
#define CRT_MATH_HD __host__ __device__ __forceinline__
#else
#define CRT_MATH_HD static inline
#endif

CRT_MATH_HD void crt_sincosf(float a, float* sn, float* cs) {
    const float q = rintf(a * 0.636619772367581343f);                 /* nearest multiple of pi/2 */
    float r = fmaf(q, -1.57079637050628662109375f, a);                /* pi/2 rounded to float ... */
    r = fmaf(q, 4.37113900018624283e-8f, r);                          /* ... and what the rounding dropped */
    const float z = r * r;
    const float ps = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    const float pc = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    const float s = fmaf(r * z, ps, r);
    const float c = fmaf(z * z, pc, fmaf(-0.5f, z, 1.0f));
    const int n = (int)q;
    const float ss = (n & 1) ? c : s, cc = (n & 1) ? s : c;
    *sn = (n & 2) ? -ss : ss;
    *cs = ((n + 1) & 2) ? -cc : cc;
}

#endif
