/* oracle/crt_oracle.c - TEST INFRASTRUCTURE ONLY (see crt_oracle.h).
 *
 * Plain-C restatement of the reference's hot path.  Every function names the reference lines it
 * follows; the arithmetic is written operation by operation in the reference's order and this file
 * must be compiled WITHOUT floating-point contraction or fast-math (oracle/Makefile:
 * -O2 -ffp-contract=off, no -march) so that each fp32 operation rounds once, as in the reference's
 * host instantiation.  Parity is pinned by tests/test_oracle_pinned.py (bit-exact against golden
 * data produced by the unmodified reference).
 */
#include "crt_oracle.h"
#include "../include/crt_rng.h"
#include "../include/crt_math.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- reference constants ------------------------------------------------------------------ */
#define EPSILON 0.00001f     /* geometry.h:9  */
#define EPSILON_2 0.000005f  /* geometry.h:10 */
#define MAX_BOUNCES 6        /* renderer.cu:25,33 */
static const float SKYBOX[3] = {0.0f, 0.025f, 0.05f}; /* renderer.cu:66-69 */
enum { DIFFUSE = 0, SPECULAR = 1, DIFFSPEC = 2, MICROFACET = 3, REFRACTIVE = 4, EMISSIVE = 5 }; /* material.h:8-15 */
#define KD_OFFSET 0
#define KA_OFFSET 1
#define KS_OFFSET 2
#define AUX_OFFSET 3
#define ORC_STACK 64         /* the reference uses int[16] and overflows on big scenes */

/* ---- reference POD layouts -------------------------------------------------- */
typedef struct { float x, y, z; } f3;
typedef struct { float x, y; } f2;
typedef struct {
    uint32_t triId, materialId;
    f3 v1, e1, e2, n1, n2, n3;
    f2 uv1, uv2, uv3;
    float area;
    uint8_t pad[128 - 108];
} Tri;
typedef struct { f3 mn, mx; int32_t offset; int16_t nTris, axis; } Node;
typedef struct { f3 eye, dir, up, right; float fov, aspect, apertureWidth, focusDistance; int32_t xpixels, ypixels; } Cam;
typedef struct { f3 origin, dir; float tMin, tMax; } Ray;
typedef struct { f3 position, normal, inputDirection, outputDirection; float pdf, u, v, t; uint32_t hitTriIdx; } SurfaceInteraction;

_Static_assert(sizeof(Tri) == 128, "Triangle");
_Static_assert(sizeof(Node) == 32, "LinearBVHNode");
_Static_assert(sizeof(Cam) == 72, "Camera");
_Static_assert(sizeof(Ray) == 32, "Ray");

/* ---- linalg.h:7-40 ------------------------------------------------------------------------ */
static inline f3 mk3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 mul3(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 mulf(f3 a, float b) { return mk3(a.x * b, a.y * b, a.z * b); }
static inline f3 divf(f3 a, float b) { return mk3(a.x / b, a.y / b, a.z / b); }
static inline f3 normalize3(f3 v) { float norm = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); return mk3(v.x / norm, v.y / norm, v.z / norm); }
static inline f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float maxComponent(f3 v) { float m = v.x > v.y ? v.x : v.y; return m > v.z ? m : v.z; }
static inline f2 add2(f2 a, f2 b) { f2 r = {a.x + b.x, a.y + b.y}; return r; }
static inline f2 mul2f(f2 a, float b) { f2 r = {a.x * b, a.y * b}; return r; }

/* ---- Sampler (renderer.cu:72-78) ----------------------------------------------------------- */
typedef struct { int mode; crt_rng_key_t key; uint32_t dim; } Sampler;
static inline float nextFloat(Sampler* s) {
    if (s->mode == ORC_RNG_RAND) return rand() / (float)RAND_MAX;
    return crt_rng_uniform(s->key, s->dim++);
}
float orc_counter_uniform(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t dim) {
    return crt_rng_uniform(crt_rng_key(seed, pixel, sample), dim);
}
uint32_t orc_counter_bits(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t dim) {
    return crt_rng_bits(crt_rng_key(seed, pixel, sample), dim);
}
/* cosf/sinf of the reference (renderer.cu:384, camera.cu:48-49).  The rand() mode is the reference as it stands (libm);
 * the replay mode evaluates the product's own function (include/crt_math.h) so that both sides round identically. */
static inline void sinCos(const Sampler* s, float a, float* sn, float* cs) {
    if (s->mode == ORC_RNG_RAND) { *cs = cosf(a); *sn = sinf(a); }
    else crt_sincosf(a, sn, cs);
}
void orc_sincos(float a, float* sn, float* cs) { crt_sincosf(a, sn, cs); }

/* ---- Triangle (geometry.cu:38-81) ----------------------------------------------------------- */
static float triIntersect(const Tri* tr, const Ray* r, float* o_u, float* o_v) {
    f3 P = cross3(r->dir, tr->e2);
    float det = dot3(tr->e1, P);
    if (fabsf(det) < EPSILON) return FLT_MAX;
    float inv_det = 1.f / det;
    f3 T = sub3(r->origin, tr->v1);
    float u = dot3(T, P) * inv_det;
    if (u < 0.f || u > 1.f) return FLT_MAX;
    f3 Q = cross3(T, tr->e1);
    float v = dot3(r->dir, Q) * inv_det;
    if (v < 0.f || u + v > 1.f) return FLT_MAX;
    float t = dot3(tr->e2, Q) * inv_det;
    if (t > EPSILON) { *o_u = u; *o_v = v; return t; }
    return FLT_MAX;
}
static f3 triNormal(const Tri* tr, float u, float v) {
    float w = 1.f - u - v;
    return normalize3(add3(add3(mulf(tr->n1, w), mulf(tr->n2, u)), mulf(tr->n3, v)));
}
static f3 triRandomPointOn(const Tri* tr, Sampler* s) {
    float u = nextFloat(s);
    float v = nextFloat(s);
    if (u + v >= 1.0f) { u = 1.0f - u; v = 1.0f - v; }
    return add3(add3(tr->v1, mulf(tr->e1, u)), mulf(tr->e2, v));
}

/* ---- rayIntersectsBox (renderer.cu:333-373): infinite-line slab test, tMin/tMax unused ------- */
static int rayIntersectsBox(const Ray* ray, const f3* mn, const f3* mx) {
    float t0 = -FLT_MAX, t1 = FLT_MAX;
    float invRayDir = 1.f / ray->dir.x;
    float tNear = (mn->x - ray->origin.x) * invRayDir;
    float tFar = (mx->x - ray->origin.x) * invRayDir;
    if (tNear > tFar) { float tmp = tNear; tNear = tFar; tFar = tmp; }
    t0 = tNear > t0 ? tNear : t0;
    t1 = tFar < t1 ? tFar : t1;
    if (t0 > t1) return 0;
    invRayDir = 1.f / ray->dir.y;
    tNear = (mn->y - ray->origin.y) * invRayDir;
    tFar = (mx->y - ray->origin.y) * invRayDir;
    if (tNear > tFar) { float tmp = tNear; tNear = tFar; tFar = tmp; }
    t0 = tNear > t0 ? tNear : t0;
    t1 = tFar < t1 ? tFar : t1;
    if (t0 > t1) return 0;
    invRayDir = 1.f / ray->dir.z;
    tNear = (mn->z - ray->origin.z) * invRayDir;
    tFar = (mx->z - ray->origin.z) * invRayDir;
    if (tNear > tFar) { float tmp = tNear; tNear = tFar; tFar = tmp; }
    t0 = tNear > t0 ? tNear : t0;
    t1 = tFar < t1 ? tFar : t1;
    if (t0 > t1) return 0;
    return 1;
}

/* ---- intersectTriangles (renderer.cu:264-280) ------------------------------------------------ */
static int intersectTriangles(const Tri* tris, int16_t numTriangles, uint32_t offset, SurfaceInteraction* si, Ray* ray,
                              orc_counters* c) {
    const float tInitial = ray->tMax;
    float t, u, v;
    const Tri* cur = tris;
    for (uint32_t i = 0; i < (uint32_t)numTriangles; i++) {
        t = triIntersect(cur, ray, &u, &v);
        if (c) c->tris_tested++;
        if (t < ray->tMax && t > ray->tMin) {
            ray->tMax = t;
            si->hitTriIdx = offset + i;
            si->u = u;
            si->v = v;
        }
        cur++;
    }
    return ray->tMax < tInitial;
}

/* ---- intersectBVH (renderer.cu:283-330), host variant; stack deepened from 16 to ORC_STACK ---- */
static int intersectBVH(const Node* bvh, const Tri* tris, SurfaceInteraction* si, Ray* ray, orc_counters* c) {
    int hit = 0;
    f3 invDir = mk3(1 / ray->dir.x, 1 / ray->dir.y, 1 / ray->dir.z);
    int dirIsNeg[3] = {invDir.x < 0, invDir.y < 0, invDir.z < 0};
    int toVisitOffset = 0, currentNodeIndex = 0;
    int stack[ORC_STACK];
    while (1) {
        const Node* node = &bvh[currentNodeIndex];
        f3 minBound = node->mn, maxBound = node->mx;
        int32_t offset = node->offset;
        int16_t numTriangles = node->nTris;
        int16_t axis = node->axis;
        if (c) c->nodes_visited++;
        if (rayIntersectsBox(ray, &minBound, &maxBound)) {
            if (numTriangles > 0) {
                if (intersectTriangles(&tris[offset], numTriangles, (uint32_t)offset, si, ray, c)) hit = 1;
                if (toVisitOffset == 0) break;
                currentNodeIndex = stack[--toVisitOffset];
            } else {
                if (dirIsNeg[axis]) {
                    stack[toVisitOffset++] = currentNodeIndex + 1;
                    currentNodeIndex = offset;
                } else {
                    stack[toVisitOffset++] = offset;
                    currentNodeIndex = currentNodeIndex + 1;
                }
                if (c && (uint32_t)toVisitOffset > c->max_stack) c->max_stack = (uint32_t)toVisitOffset;
                if (toVisitOffset >= ORC_STACK) abort();
            }
        } else {
            if (toVisitOffset == 0) break;
            currentNodeIndex = stack[--toVisitOffset];
        }
    }
    return hit;
}

int orc_trace_closest(const orc_scene* s, const orc_ray* rays, size_t n, orc_hit* hits, orc_counters* c) {
    const Node* bvh = (const Node*)s->nodes;
    const Tri* tris = (const Tri*)s->tris;
    /* counters are a single-thread affair; without them the rays are independent */
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 4096) if (c == NULL && n > 65536)
#endif
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Ray ray;
        memcpy(&ray, &rays[i], sizeof ray);
        SurfaceInteraction si;
        memset(&si, 0, sizeof si);
        int hit = intersectBVH(bvh, tris, &si, &ray, c);
        hits[i].t = ray.tMax;
        hits[i].u = hit ? si.u : 0.f;
        hits[i].v = hit ? si.v : 0.f;
        hits[i].tri = hit ? si.hitTriIdx : 0xFFFFFFFFu;
        if (c) c->extend_rays++;
    }
    return 0;
}

/* Order-independent evaluator: SURVEY.md 8(a)(i)-(ii) / Appendix C.  A triangle is tested iff its
 * leaf's OWN box passes the reference slab test; min t wins; on an exact t tie the reference keeps
 * the first triangle it visited, which is (dir.x < 0) the lowest index, else the lowest index inside
 * the leaf with the highest firstTriOffset. */
int orc_trace_closest_unordered(const orc_scene* s, const orc_ray* rays, size_t n, orc_hit* hits) {
    const Node* bvh = (const Node*)s->nodes;
    const Tri* tris = (const Tri*)s->tris;
    for (size_t r = 0; r < n; r++) {
        Ray ray;
        memcpy(&ray, &rays[r], sizeof ray);
        int dirneg = (1.f / ray.dir.x) < 0;
        float best = ray.tMax, bu = 0.f, bv = 0.f;
        uint32_t bestIdx = 0xFFFFFFFFu, bestLeaf = 0;
        for (uint32_t k = 0; k < s->n_nodes; k++) {
            const Node* nd = &bvh[k];
            if (nd->nTris <= 0) continue;
            if (!rayIntersectsBox(&ray, &nd->mn, &nd->mx)) continue;
            for (uint32_t i = 0; i < (uint32_t)nd->nTris; i++) {
                uint32_t idx = (uint32_t)nd->offset + i;
                float u, v;
                float t = triIntersect(&tris[idx], &ray, &u, &v);
                if (t == FLT_MAX || !(t > ray.tMin)) continue;
                int take = 0;
                if (t < best) take = 1;
                else if (t == best && bestIdx != 0xFFFFFFFFu) {
                    uint32_t leaf = (uint32_t)nd->offset;
                    take = dirneg ? (idx < bestIdx) : (leaf > bestLeaf || (leaf == bestLeaf && idx < bestIdx));
                }
                if (take) { best = t; bu = u; bv = v; bestIdx = idx; bestLeaf = (uint32_t)nd->offset; }
            }
        }
        hits[r].t = best; hits[r].u = bu; hits[r].v = bv; hits[r].tri = bestIdx;
    }
    return 0;
}

/* ---- camera (camera.cu:13-57) --------------------------------------------------------------- */
static void tentFilter(float* i, float* j, Sampler* s) {
    float r1, r2;
    r1 = 2.f * nextFloat(s);
    r2 = 2.f * nextFloat(s);
    if (r1 < 1.f) *i = sqrtf(r1) - 1.f; else *i = 1.f - sqrtf(2.f - r1);
    if (r2 < 1) *j = sqrtf(r2) - 1.f; else *j = 1.f - sqrtf(2.f - r2);
}
static Ray computeCameraRay(const Cam* cam, int i, int j, Sampler* s) {
    float dx, dy;
    tentFilter(&dx, &dy, s);
    /* `- 0.5` is a double literal in the reference: float quotient widened, subtracted, narrowed */
    float normalized_i = (float)((double)(((float)i + dx) / (float)cam->xpixels) - 0.5);
    float normalized_j = 0.5f - (((float)j + dy) / (float)cam->ypixels);
    f3 direction = cam->dir;
    float rightJitter = 1.0f, upJitter = 1.0f;
    f3 origin = cam->eye;
    direction = add3(direction, mulf(mulf(mulf(cam->right, cam->fov), cam->aspect), normalized_i));
    direction = add3(direction, mulf(mulf(cam->up, cam->fov), normalized_j));
    if (cam->apertureWidth > 0.0f) {
        /* `rand - 0.5f * 2 * M_PI`: (0.5f*2) is float 1.0f, times M_PI in double, float minus double */
        float r1 = (float)((double)nextFloat(s) - (double)(0.5f * 2) * M_PI);
        float r2 = nextFloat(s) - 0.5f;
        float sn, cs;
        sinCos(s, r1, &sn, &cs);
        rightJitter = cs * r2 * cam->apertureWidth;
        upJitter = sn * r2 * cam->apertureWidth;
        origin = add3(origin, add3(mulf(cam->right, rightJitter), mulf(cam->up, upJitter)));
        f3 focalPoint = add3(cam->eye, mulf(direction, cam->focusDistance));
        direction = sub3(focalPoint, origin);
    }
    Ray r;
    r.origin = origin;
    r.dir = normalize3(direction);
    r.tMin = EPSILON;
    r.tMax = FLT_MAX;
    return r;
}

/* ---- sampleTexture, host variant (renderer.cu:92-115) ----------------------------------------- */
static f3 texel(const orc_scene* s, const f3* base, uint32_t idx) {
    /* the reference indexes without bounds; guard the flattened+padded array instead of faulting */
    const f3* end = (const f3*)s->tex_data + s->tex_texels;
    const f3* p = base + idx;
    if (p >= end || p < (const f3*)s->tex_data) return mk3(0.f, 0.f, 0.f);
    return *p;
}
static f3 sampleTexture(const orc_scene* s, int texIdx, float u, float v) {
    const f3* p_texture = (const f3*)s->tex_data + s->tex_offsets[texIdx];
    uint32_t width = s->tex_dims[2 * texIdx], height = s->tex_dims[2 * texIdx + 1];
    float pixelCoordU = u * (float)width;
    float pixelCoordV = v * (float)height;
    uint32_t i = (uint32_t)truncf(pixelCoordU);
    uint32_t j = (uint32_t)truncf(pixelCoordV);
    float floorPixelCoordU = floorf(pixelCoordU), floorPixelCoordV = floorf(pixelCoordV);
    float ceilPixelCoordU = ceilf(pixelCoordU), ceilPixelCoordV = ceilf(pixelCoordV);
    f3 valA1 = mulf(texel(s, p_texture, j * width + i + 1), pixelCoordU - floorPixelCoordU);
    f3 valA2 = mulf(texel(s, p_texture, j * width + i), ceilPixelCoordU - pixelCoordU);
    f3 valB1 = mulf(texel(s, p_texture, (j + 1) * width + i + 1), pixelCoordU - floorPixelCoordU);
    f3 valB2 = mulf(texel(s, p_texture, (j + 1) * width + i), ceilPixelCoordU - pixelCoordU);
    f3 valC1 = mulf(add3(valA1, valA2), ceilPixelCoordV - pixelCoordV);
    f3 valC2 = mulf(add3(valB1, valB2), pixelCoordV - floorPixelCoordV);
    return add3(valC1, valC2);
}

/* ---- BSDFs (renderer.cu:375-412, 464-493) ------------------------------------------------------ */
static f3 sampleDiffuseBSDF(const orc_scene* s, SurfaceInteraction* si, const Tri* hitTri, f3 diffuseColor, int texIdx,
                            Sampler* smp) {
    float r1 = (float)(2 * M_PI * (double)nextFloat(smp));
    float r2 = nextFloat(smp);
    float r2sq = sqrtf(r2);
    f3 w = si->normal;
    f3 u = normalize3(cross3((fabsf(w.x) > 0.1f ? mk3(0.f, 1.f, 0.f) : mk3(1.f, 0.f, 0.f)), w));
    f3 v = cross3(w, u);
    float sn, cs;
    sinCos(smp, r1, &sn, &cs);
    si->inputDirection = normalize3(add3(add3(mulf(mulf(u, cs), r2sq), mulf(mulf(v, sn), r2sq)), mulf(w, sqrtf(1.f - r2))));
    si->pdf = 0.5f;
    float cosineWeight = dot3(si->inputDirection, si->normal);
    f3 kd = diffuseColor;
    if (texIdx != -1) {
        float bu = si->u, bv = si->v;
        float bw = 1.f - bu - bv;
        f2 uv = add2(add2(mul2f(hitTri->uv1, bw), mul2f(hitTri->uv2, bu)), mul2f(hitTri->uv3, bv));
        kd = sampleTexture(s, texIdx, uv.x, uv.y);
    }
    return mulf(kd, cosineWeight);
}
static f3 reflect3(f3 incedent, f3 normal) { return sub3(incedent, mulf(mulf(normal, dot3(incedent, normal)), 2.f)); }

typedef struct { float probReflection, probTransmission; } Fresnel;
static Fresnel getFresnelReflectance(const SurfaceInteraction* si, float ior, f3* transmittedDir) {
    Fresnel fr;
    f3 incedent = si->outputDirection;
    f3 normal = si->normal;
    float cosi = dot3(incedent, normal);
    float etai = 1, etat = ior;
    f3 n = normal;
    if (cosi < 0.0f) {
        cosi = -cosi;
    } else {
        float temp = etai; etai = etat; etat = temp;
        n = mulf(normal, -1.0f);
    }
    float eta = etai / etat;
    float k = 1 - eta * eta * (1 - cosi * cosi);
    if (k < 0) {
        fr.probReflection = 1.0f;
        fr.probTransmission = 0.0f;
    } else {
        *transmittedDir = add3(mulf(incedent, eta), mulf(n, eta * cosi - sqrtf(k)));
        float R0 = (etai - etat) * (etai - etat) / (etat + etai) * (etat + etai); /* sic (D18) */
        float c = 1.f - dot3(*transmittedDir, normal);
        float Re = R0 + (1.f - R0) * c * c * c * c * c;
        fr.probReflection = Re;
        fr.probTransmission = 1.f - Re;
    }
    return fr;
}

/* ---- path loop ---------------------------------------------------------------------------------- */
typedef struct {
    const orc_scene* s;
    orc_counters* c;
    orc_ray* log; size_t cap; size_t n;
} Ctx;
static void logRay(Ctx* cx, const Ray* r) {
    if (cx->log && cx->n < cx->cap) memcpy(&cx->log[cx->n], r, sizeof *r);
    cx->n++;
}

/* estimateDirectLighting (renderer.cu:414-445) */
static f3 estimateDirectLighting(Ctx* cx, const Tri* light, uint32_t lightIdx, f3 lightColor, float lightsSurfaceArea,
                                 const SurfaceInteraction* si, Sampler* smp) {
    if (si->hitTriIdx == lightIdx) return mk3(0.0f, 0.0f, 0.0f);
    f3 directLighting = mk3(0.0f, 0.0f, 0.0f);
    f3 rayOrigin = add3(si->position, mulf(si->normal, EPSILON));
    Ray ray;
    ray.origin = rayOrigin;
    ray.dir = normalize3(sub3(triRandomPointOn(light, smp), si->position));
    ray.tMin = EPSILON;
    ray.tMax = FLT_MAX;
    SurfaceInteraction li;
    memset(&li, 0, sizeof li);
    logRay(cx, &ray);
    if (cx->c) cx->c->shadow_rays++;
    int intersectsLight = intersectBVH((const Node*)cx->s->nodes, (const Tri*)cx->s->tris, &li, &ray, cx->c);
    if (intersectsLight && li.hitTriIdx == lightIdx) {
        float surfaceArea = light->area;
        float distanceSquared = ray.tMax * ray.tMax;
        float cosTheta = fabsf(dot3(triNormal(light, li.u, li.v), mulf(ray.dir, -1.0f)));
        float weightFactor = surfaceArea / (distanceSquared * lightsSurfaceArea) * cosTheta;
        directLighting = add3(directLighting, mulf(lightColor, weightFactor));
    }
    return directLighting;
}

/* samplePixel (renderer.cu:119-262) */
static f3 samplePixel(Ctx* cx, int x, int y, Sampler* smp) {
    const orc_scene* s = cx->s;
    const Tri* tris = (const Tri*)s->tris;
    const f3* matFloats = (const f3*)s->mat_floats;
    const int32_t* matIndices = s->mat_indices;
    const float lightsSurfaceArea = s->lights_area;
    const uint32_t numLights = s->n_lights;
    Ray ray = computeCameraRay((const Cam*)s->camera, x, y, smp);
    f3 color = mk3(0.f, 0.f, 0.f), mask = mk3(1.f, 1.f, 1.f);
    SurfaceInteraction si;
    memset(&si, 0, sizeof si);
    const Tri* hitTri = NULL;
    int currentBsdf = DIFFUSE, previousBsdf = DIFFUSE;
    uint32_t materialId;
    float oneOverLightsSA = 1.f / lightsSurfaceArea;
    if (cx->c) cx->c->samples++;
    for (unsigned bounces = 0; bounces < MAX_BOUNCES; bounces++) {
        logRay(cx, &ray);
        if (cx->c) cx->c->extend_rays++;
        if (!intersectBVH((const Node*)s->nodes, tris, &si, &ray, cx->c)) {
            color = add3(color, mul3(mask, mk3(SKYBOX[0], SKYBOX[1], SKYBOX[2])));
            break;
        }
        hitTri = tris + si.hitTriIdx;
        materialId = hitTri->materialId;
        if (bounces == 0 || previousBsdf == SPECULAR || previousBsdf == REFRACTIVE)
            color = add3(color, mulf(mul3(mask, matFloats[materialId * 4 + KA_OFFSET]), oneOverLightsSA));

        si.normal = triNormal(hitTri, si.u, si.v);
        si.position = add3(add3(ray.origin, mulf(ray.dir, ray.tMax)), mulf(si.normal, EPSILON_2));
        si.outputDirection = normalize3(ray.dir);

        currentBsdf = matIndices[2 * materialId];
        if (currentBsdf == DIFFSPEC) {
            float p = matFloats[materialId * 4 + AUX_OFFSET].z;
            if (nextFloat(smp) < p) {
                currentBsdf = DIFFUSE;
                mask = mulf(mask, 1.0f / p);
            } else {
                currentBsdf = SPECULAR;
                mask = mulf(mask, 1.0f / (1.0f - p));
            }
        }
        if (currentBsdf == REFRACTIVE) {
            f3 transmittedDir = mk3(0.f, 0.f, 0.f);
            float ni = matFloats[materialId * 4 + AUX_OFFSET].y;
            Fresnel fresnel = getFresnelReflectance(&si, ni, &transmittedDir);
            if (fresnel.probReflection == 1.0f) {
                currentBsdf = SPECULAR;
            } else {
                float unknownMagicNumber = 0.25f;
                float P = unknownMagicNumber + .5f * fresnel.probReflection;
                float RP = fresnel.probReflection / P;
                float TP = fresnel.probTransmission / (1.f - P);
                if (nextFloat(smp) > unknownMagicNumber) {
                    si.inputDirection = transmittedDir;
                    mask = mulf(mul3(mask, matFloats[materialId * 4 + KS_OFFSET]), TP);
                } else {
                    mask = mulf(mask, RP);
                    currentBsdf = SPECULAR;
                }
            }
        }
        if (currentBsdf == DIFFUSE) {
            int texIdx = matIndices[2 * materialId + 1];
            f3 diffuseColor = matFloats[materialId * 4 + KD_OFFSET];
            f3 diffuseSample = sampleDiffuseBSDF(s, &si, hitTri, diffuseColor, texIdx, smp);
            mask = divf(mul3(mask, diffuseSample), si.pdf);
            float randomNumber = nextFloat(smp) * ((float)numLights - .00001f);
            uint32_t selectedLightIdx = s->light_indices[(uint32_t)truncf(randomNumber)];
            const Tri* light = tris + selectedLightIdx;
            f3 lightColor = matFloats[light->materialId * 4 + KA_OFFSET];
            f3 directLighting = estimateDirectLighting(cx, light, selectedLightIdx, lightColor, lightsSurfaceArea, &si, smp);
            color = add3(color, mul3(mask, directLighting));
        }
        if (currentBsdf == SPECULAR) {
            f3 specularColor = matFloats[materialId * 4 + KS_OFFSET];
            si.inputDirection = reflect3(si.outputDirection, si.normal);
            si.pdf = 1.0f;
            mask = divf(mul3(mask, specularColor), si.pdf);
        }
        previousBsdf = currentBsdf;
        ray.origin = add3(si.position, mulf(si.inputDirection, EPSILON));
        ray.dir = si.inputDirection;
        ray.tMin = EPSILON;
        ray.tMax = FLT_MAX;
        if (bounces >= 3) {
            float p = maxComponent(mask);
            if (nextFloat(smp) > p) break;
            mask = divf(mask, p);
        }
    }
    return color;
}

int orc_render_accum(const orc_scene* s, uint32_t W, uint32_t H, uint32_t first_sample, uint32_t spp, int rng_mode,
                     uint32_t seed, int threads, uint32_t row0, uint32_t row1, float* accum, orc_counters* c,
                     orc_ray* ray_log, size_t ray_log_cap, size_t* ray_log_n) {
    f3* acc = (f3*)accum;
    if (row1 > H) row1 = H;
    if (rng_mode == ORC_RNG_RAND || threads <= 1 || ray_log) {
        Ctx cx = {s, c, ray_log, ray_log_cap, 0};
        Sampler smp = {rng_mode, {0, 0}, 0};
        for (uint32_t sm = 0; sm < spp; sm++)
            for (uint32_t x = 0; x < W; x++)
                for (uint32_t y = row0; y < row1; y++) {
                    size_t idx = (size_t)y * W + x;
                    smp.key = crt_rng_key(seed, (uint32_t)idx, first_sample + sm);
                    smp.dim = 0;
                    f3 col = samplePixel(&cx, (int)x, (int)y, &smp);
                    acc[idx] = add3(acc[idx], col);
                }
        if (ray_log_n) *ray_log_n = cx.n;
        return 0;
    }
    orc_counters total;
    memset(&total, 0, sizeof total);
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
    {
        orc_counters local;
        memset(&local, 0, sizeof local);
        Ctx cx = {s, c ? &local : NULL, NULL, 0, 0};
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int64_t y = row0; y < (int64_t)row1; y++)
            for (uint32_t x = 0; x < W; x++) {
                size_t idx = (size_t)y * W + x;
                f3 sum = acc[idx];
                for (uint32_t sm = 0; sm < spp; sm++) {
                    Sampler smp = {rng_mode, crt_rng_key(seed, (uint32_t)idx, first_sample + sm), 0};
                    sum = add3(sum, samplePixel(&cx, (int)x, (int)y, &smp));
                }
                acc[idx] = sum;
            }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            total.samples += local.samples; total.extend_rays += local.extend_rays; total.shadow_rays += local.shadow_rays;
            total.nodes_visited += local.nodes_visited; total.tris_tested += local.tris_tested;
            if (local.max_stack > total.max_stack) total.max_stack = local.max_stack;
        }
    }
    if (c) {
        c->samples += total.samples; c->extend_rays += total.extend_rays; c->shadow_rays += total.shadow_rays;
        c->nodes_visited += total.nodes_visited; c->tris_tested += total.tris_tested;
        if (total.max_stack > c->max_stack) c->max_stack = total.max_stack;
    }
    if (ray_log_n) *ray_log_n = 0;
    return 0;
}

/* float3ToUchar4 (renderer.h:149-156) on accum/spp (sequential_renderer.cu:111), then
 * gammaCorrectPixel (renderer.cu:495-504) as copyImageBytes applies it (sequential_renderer.cu:118-125). */
void orc_resolve_rgba8(const float* accum, uint32_t n_pixels, uint32_t spp, uint8_t* rgba) {
    const f3* acc = (const f3*)accum;
    float denom = (float)(spp - 1) + 1;
    float invGamma = 1.f / 2.2f;
    for (uint32_t i = 0; i < n_pixels; i++) {
        f3 v = divf(acc[i], denom);
        uint8_t px[3];
        px[0] = (uint8_t)((v.x > 1.0f ? 1.0f : v.x) * (255.f));
        px[1] = (uint8_t)((v.y > 1.0f ? 1.0f : v.y) * (255.f));
        px[2] = (uint8_t)((v.z > 1.0f ? 1.0f : v.z) * (255.f));
        for (int k = 0; k < 3; k++) {
            float fp = powf((float)px[k] * 1.f / 255.f, invGamma);
            rgba[4 * i + k] = (uint8_t)(fp * 255.f);
        }
        rgba[4 * i + 3] = 255u;
    }
}
