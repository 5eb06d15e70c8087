// svcm_device.cuh — device-side building blocks of the VCM path (sm_100a).
//
// Everything the wavefront kernels call per path vertex: vector math, the per-path random
// stream, ray/scene intersection, the BSDF, the light models and the camera. These are
// written from scratch for one warp lane per path with the whole scene in the kernel's
// constant bank (a __grid_constant__ DevScene parameter); they FOLLOW the reference's
// arithmetic (file:line cited at each function) so results match the CPU renderer, but
// none of this is shared with, or calls into, oracle/.
#ifndef SVCM_DEVICE_CUH_
#define SVCM_DEVICE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include "svcm.h"
#include "svcm_bvh.h"

#define SVCM_HD __host__ __device__ __forceinline__
#define SVCM_D  __device__ __forceinline__

namespace svcm {

// reference src/math.hxx:29-30, src/utils.hxx:32-33, src/bsdf.hxx:56
#define SVCM_PI       3.14159265358979f
#define SVCM_INV_PI   (1.f / SVCM_PI)
#define SVCM_EPS_COS  1e-6f
#define SVCM_EPS_RAY  1e-3f
#define SVCM_EPS_PHONG 1e-3f

// ---------------------------------------------------------------------------------------
// float3 helpers (reference src/math.hxx:88-165: all operators component-wise)
// ---------------------------------------------------------------------------------------
struct f3 { float x, y, z; };

SVCM_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
SVCM_HD f3 mk3(float a) { return mk3(a, a, a); }
SVCM_HD f3 ld3(const float *p) { return mk3(p[0], p[1], p[2]); }
SVCM_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
SVCM_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
SVCM_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
SVCM_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
SVCM_HD f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
SVCM_HD f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
SVCM_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
SVCM_HD float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SVCM_HD f3 cross(f3 a, f3 b)
{
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
SVCM_HD float lensqr(f3 a) { return dot(a, a); }
SVCM_HD bool  iszero(f3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
SVCM_HD float maxcomp(f3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
SVCM_HD f3 normalize(f3 a) { return a / sqrtf(dot(a, a)); }            // src/math.hxx:156-161
SVCM_HD float sqr(float a) { return a * a; }
// a + b*s with the product and the sum rounded separately (no FMA contraction). Used for
// the two places where the reference's double rounding is load-bearing: the ray-origin
// offset and the hit point (reference src/vertexcm.hxx:333,340). Hit points on the
// axis-aligned Cornell walls sit exactly ON the light-vertex bounding box that
// HashGrid::Process tests queries against (src/hashgrid.hxx:116-122); a fused multiply-add
// changes their last-ulp scatter and with it the fraction of wall queries that test rejects
// (measured: 1 % of all merge queries, a 1.5 % darker BPM image).
SVCM_D f3 mulAddUnfused(f3 a, f3 b, float s)
{
    return mk3(__fadd_rn(a.x, __fmul_rn(b.x, s)), __fadd_rn(a.y, __fmul_rn(b.y, s)),
               __fadd_rn(a.z, __fmul_rn(b.z, s)));
}
SVCM_HD float luminance(f3 c)                                          // src/utils.hxx:36-41
{
    return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z;
}

// Frame, reference src/frame.hxx:32-78
struct Frame3
{
    f3 x, y, z;
    SVCM_HD void setFromZ(f3 n)
    {
        const f3 tz = z = normalize(n);
        const f3 tx = (fabsf(tz.x) > 0.99f) ? mk3(0, 1, 0) : mk3(1, 0, 0);
        y = normalize(cross(tz, tx));
        x = cross(y, tz);
    }
    SVCM_HD f3 toWorld(f3 a) const { return x * a.x + y * a.y + z * a.z; }
    SVCM_HD f3 toLocal(f3 a) const { return mk3(dot(a, x), dot(a, y), dot(a, z)); }
};

// ---------------------------------------------------------------------------------------
// per-path random stream (DESIGN.md "random numbers"): Philox2x32-7, counter = (draw index,
// path index), key = seed + 0x85EBCA6B * (2*iteration + phase); float = (top 24 bits | 1) / 2^24, i.e. 2^23 equally spaced values strictly inside (0, 1).
// Replaces the reference's single sequential mt19937_64 stream (src/rng.hxx:41-86), which a
// parallel renderer cannot reproduce; the draw ORDER inside a path is the reference's.
// ---------------------------------------------------------------------------------------
SVCM_HD uint32_t rngKey(int32_t seed, int32_t iteration, int32_t phase)
{
    return uint32_t(seed) + 0x85EBCA6Bu * (2u * uint32_t(iteration) + uint32_t(phase));
}

SVCM_HD uint32_t philox2x32_7(uint32_t c0, uint32_t c1, uint32_t key)
{
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for(int round = 0; round < 7; round++)
    {
#ifdef __CUDA_ARCH__
        const uint32_t hi = __umulhi(0xD256D193u, c0);
#else
        const uint32_t hi = uint32_t((uint64_t(0xD256D193u) * c0) >> 32);
#endif
        const uint32_t lo = 0xD256D193u * c0;
        c0 = hi ^ key ^ c1;
        c1 = lo;
        key += 0x9E3779B9u;
    }
    return c0;
}

struct PathRng
{
    uint32_t key, path, draw;
    SVCM_HD float next()
    {
        // odd multiples of 2^-24: strictly inside (0, 1), every value exact in float32
        return float((philox2x32_7(draw++, path, key) >> 8) | 1u) * (1.0f / 16777216.0f);
    }
};

// ---------------------------------------------------------------------------------------
// the scene as the kernels see it: one POD block passed as a __grid_constant__ parameter
// ---------------------------------------------------------------------------------------
struct DevPrim  { float v[12]; int32_t type; int32_t matID; };
// a primitive of a scene too large for the constant bank: one 64-byte record in global memory,
// in the hierarchy's slot order; `orig` = its index in the caller's list (svcm_bvh.h)
struct alignas(16) DevPrimG { float v[12]; int32_t type, matID, orig, pad; };
struct DevLight
{
    int32_t type;
    float p0[3], e1[3], e2[3];
    float fx[3], fy[3], fz[3];
    float intensity[3];          // background: already multiplied by mScale
    float invArea;
};
struct DevScene
{
    int32_t numPrims, numMats, numLights, background;
    float   camPos[3], camFwd[3];
    float   resX, resY;
    float   rasterToWorld[16], worldToRaster[16];
    float   imagePlaneDist;
    float   sphereCenter[3], sphereRadius, sphereInvRadiusSqr;
    int32_t matToLight[SVCM_MAX_MATERIALS];
    SvcmMaterial mats[SVCM_MAX_MATERIALS];
    DevLight lights[SVCM_MAX_LIGHTS];
    DevPrim  prims[SVCM_MAX_PRIMITIVES];
    // The constant-bank list cut into contiguous runs (list order kept). A run whose bit is set in
    // runCull consists of triangles inside a small padded box (lo / hi): a ray that misses the box
    // skips the run, like a leaf of the hierarchy below and under the same rules (svcm_bvh.h).
    // Walls, spheres and anything large sit in runs that are always tested.
    int32_t  numRuns;
    uint32_t runCull;
    BvhNode  runs[SVCM_MAX_RUNS];
    // numPrims > SVCM_MAX_PRIMITIVES: the list lives in global memory under a hierarchy instead
    const BvhNode  *bvh;         // NULL for the constant-bank list above
    const DevPrimG *bvhPrims;
};

struct Hit { float dist; int32_t matID; int32_t prim; f3 normal; };

// Triangle::Intersect, reference src/geometry.hxx:125-156
SVCM_HD bool hitTriangle(const float *v, f3 org, f3 dir, float tmax, float &oDist)
{
    const f3 ao = ld3(v) - org;
    const f3 n = ld3(v + 9);
    const float num = dot(n, ao), den = dot(n, dir);
    // The reference evaluates the three edge tests first and the plane distance num / den last; a hit
    // needs both, so the order is free. The distance is the cheap half and, for a segment between two
    // points inside a closed room, almost always out of range: decide that first, WITHOUT the division
    // and only where the IEEE quotient is certain to fail 0 < distance < tmax —
    //   * num and den of opposite sign, or num == 0, or den == 0: the quotient is negative, a zero, NaN
    //     or an infinity;
    //   * |num| > tmax |den| by more than 1e-5 relative (the quotient and the product round by 6e-8
    //     each; products in the subnormal range, where that no longer holds, are left to the division).
    // Everything else runs the reference's test unchanged, so the result is the same bit for bit.
    const float sn = (den < 0.f) ? -num : num;
    const float q = tmax * fabsf(den);
    if(!(sn > 0.f) || (q >= 1e-30f && sn > q * 1.00001f)) return false;
    const f3 bo = ld3(v + 3) - org;
    const f3 co = ld3(v + 6) - org;
    const float v0d = dot(cross(co, bo), dir);
    const float v1d = dot(cross(bo, ao), dir);
    const float v2d = dot(cross(ao, co), dir);
    if(((v0d < 0.f) && (v1d < 0.f) && (v2d < 0.f)) || ((v0d >= 0.f) && (v1d >= 0.f) && (v2d >= 0.f)))
    {
        const float distance = num / den;
        if((distance > 0.f) && (distance < tmax)) { oDist = distance; return true; }
    }
    return false;
}

// dot product evaluated like the reference's Dot (src/math.hxx:129-130): ((0 + x) + y) + z
// with every product and sum rounded on its own (no FMA contraction)
SVCM_D float dotUnfused(f3 a, f3 b)
{
    return __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(a.x, b.x)), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}

// Sphere::Intersect, reference src/geometry.hxx:198-237: A, B, C and the discriminant in
// float, the roots in double (the reference's fix for grazing-ray epsilon leaks). The float
// part is evaluated unfused, in the reference's order: B*B - 4*A*C cancels catastrophically
// for grazing rays, so a contracted evaluation would move those hits by far more than an ulp
// (there are at most two spheres per scene, the extra instructions are noise).
SVCM_D bool hitSphere(const float *v, f3 org, f3 dir, float tmax, float &oDist)
{
    const f3 to = org - ld3(v);
    const float radius = v[3];
    const float A = dotUnfused(dir, dir);
    const float B = __fmul_rn(2.f, dotUnfused(dir, to));
    const float C = __fsub_rn(dotUnfused(to, to), __fmul_rn(radius, radius));
    const float discF = __fsub_rn(__fmul_rn(B, B), __fmul_rn(__fmul_rn(4.f, A), C));
    if(discF < 0.f) return false;
    const double discSqrt = sqrt(double(discF));
    const double q = (B < 0.f) ? ((-double(B) - discSqrt) / 2.0) : ((-double(B) + discSqrt) / 2.0);
    double t0 = q / double(A);
    double t1 = double(C) / q;
    if(t0 > t1) { const double t = t0; t0 = t1; t1 = t; }
    if(t0 > 0.0 && t0 < double(tmax))      { oDist = float(t0); return true; }
    else if(t1 > 0.0 && t1 < double(tmax)) { oDist = float(t1); return true; }
    return false;
}

// The same two queries for a scene under a hierarchy (svcm_bvh.h).
SVCM_D bool bvhSceneIntersect(const BvhNode *nodes, const DevPrimG *prims, f3 org, f3 dir, Hit &hit)
{
    float best = 1e36f;
    int32_t bestOrig = 0x7FFFFFFF, bestSlot = -1;
    bvhClosest(nodes, org.x, org.y, org.z, dir.x, dir.y, dir.z, best, bestOrig, bestSlot,
               [&](int32_t slot, float &d, int32_t &orig) -> bool
               {
                   const DevPrimG p = prims[slot];
                   orig = p.orig;
                   return (p.type == SVCM_PRIM_TRIANGLE) ? hitTriangle(p.v, org, dir, 1e36f, d)
                                                         : hitSphere(p.v, org, dir, 1e36f, d);
               });
    if(bestSlot < 0) return false;
    const DevPrimG p = prims[bestSlot];
    hit.dist = best; hit.prim = bestOrig; hit.matID = p.matID;
    if(p.type == SVCM_PRIM_TRIANGLE)
        hit.normal = ld3(p.v + 9);
    else
    {
        const f3 n = mulAddUnfused(org - ld3(p.v), dir, best);
        hit.normal = n / sqrtf(dotUnfused(n, n));
    }
    return true;
}
SVCM_D bool bvhSceneOccluded(const BvhNode *nodes, const DevPrimG *prims, f3 org, f3 dir, float tmax)
{
    return bvhAny(nodes, org.x, org.y, org.z, dir.x, dir.y, dir.z, tmax,
                  [&](int32_t slot) -> bool
                  {
                      const DevPrimG p = prims[slot];
                      float d;
                      return (p.type == SVCM_PRIM_TRIANGLE) ? hitTriangle(p.v, org, dir, tmax, d)
                                                            : hitSphere(p.v, org, dir, tmax, d);
                  });
}

// Scene::Intersect -> GeometryList::Intersect, reference src/scene.hxx:53-70 and
// src/geometry.hxx:65-78: closest hit over the whole list, list order kept (strict <).
// The primitive index is warp-uniform, so every load below is a constant-bank broadcast.
// kBvh is a compile-time property of the kernel (the launchers pick the instantiation from
// sc.bvh): kernels of small scenes carry no trace of the hierarchy code.
template<bool kBvh>
SVCM_D bool sceneIntersect(const DevScene &sc, f3 org, f3 dir, Hit &hit)
{
    if(kBvh) return bvhSceneIntersect(sc.bvh, sc.bvhPrims, org, dir, hit);
    float best = 1e36f;
    int prim = -1;
    for(int r = 0; r < sc.numRuns; r++)
    {
        if((sc.runCull >> r) & 1u)
        {
            float tEnter;           // the box against [0, best]: a run behind the closest hit so far is skipped too
            if(!bvhSlab(sc.runs[r], org.x, org.y, org.z, __frcp_rn(dir.x), __frcp_rn(dir.y), __frcp_rn(dir.z), best, tEnter))
                continue;
        }
        const int end = sc.runs[r].first + sc.runs[r].count;
        for(int i = sc.runs[r].first; i < end; i++)
        {
            float d;
            const bool h = (sc.prims[i].type == SVCM_PRIM_TRIANGLE)
                ? hitTriangle(sc.prims[i].v, org, dir, best, d)
                : hitSphere(sc.prims[i].v, org, dir, best, d);
            if(h) { best = d; prim = i; }
        }
    }
    if(prim < 0) return false;
    hit.dist  = best;
    hit.prim  = prim;
    hit.matID = sc.prims[prim].matID;
    if(sc.prims[prim].type == SVCM_PRIM_TRIANGLE)
        hit.normal = ld3(sc.prims[prim].v + 9);
    else
    {
        // Normalize(transformedOrigin + Vec3f(resT) * aRay.dir), src/geometry.hxx:235 (unfused)
        const f3 n = mulAddUnfused(org - ld3(sc.prims[prim].v), dir, best);
        hit.normal = n / sqrtf(dotUnfused(n, n));
    }
    return true;
}

// Scene::Occluded -> GeometryList::IntersectP, reference src/scene.hxx:72-85,
// src/geometry.hxx:80-91: any hit in (0, tMax - 2 EPS_RAY) from point + dir*EPS_RAY
template<bool kBvh>
SVCM_D bool sceneOccluded(const DevScene &sc, f3 point, f3 dir, float tMax)
{
    const f3 org = point + dir * SVCM_EPS_RAY;
    const float tmax = tMax - 2.f * SVCM_EPS_RAY;
    if(kBvh) return bvhSceneOccluded(sc.bvh, sc.bvhPrims, org, dir, tmax);
    for(int r = 0; r < sc.numRuns; r++)
    {
        if((sc.runCull >> r) & 1u)
        {
            float tEnter;
            if(!bvhSlab(sc.runs[r], org.x, org.y, org.z, __frcp_rn(dir.x), __frcp_rn(dir.y), __frcp_rn(dir.z), tmax, tEnter))
                continue;
        }
        const int end = sc.runs[r].first + sc.runs[r].count;
        for(int i = sc.runs[r].first; i < end; i++)
        {
            float d;
            const bool h = (sc.prims[i].type == SVCM_PRIM_TRIANGLE)
                ? hitTriangle(sc.prims[i].v, org, dir, tmax, d)
                : hitSphere(sc.prims[i].v, org, dir, tmax, d);
            if(h) return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------
// sampling helpers, reference src/utils.hxx
// ---------------------------------------------------------------------------------------
SVCM_D float fresnelDielectric(float cosInc, float ior)               // src/utils.hxx:43-74
{
    if(ior < 0.f) return 1.f;
    float eta;
    if(cosInc < 0.f) { cosInc = -cosInc; eta = ior; }
    else             { eta = 1.f / ior; }
    const float sinTrans2 = sqr(eta) * (1.f - sqr(cosInc));
    const float cosTrans  = sqrtf(fmaxf(0.f, 1.f - sinTrans2));
    const float term1 = eta * cosTrans;
    const float rPar  = (cosInc - term1) / (cosInc + term1);
    const float term2 = eta * cosInc;
    const float rPerp = (term2 - cosTrans) / (term2 + cosTrans);
    return 0.5f * (sqr(rPar) + sqr(rPerp));
}
SVCM_D f3 reflectLocal(f3 v) { return mk3(-v.x, -v.y, v.z); }          // src/utils.hxx:77-80

SVCM_D f3 samplePowerCosHemisphereW(float sx, float sy, float power)   // src/utils.hxx:85-103
{
    const float term1 = 2.f * SVCM_PI * sx;
    const float term2 = powf(sy, 1.f / (power + 1.f));
    const float term3 = sqrtf(1.f - term2 * term2);
    float s, c; sincosf(term1, &s, &c);
    return mk3(c * term3, s * term3, term2);
}
SVCM_D float powerCosHemispherePdfW(f3 normal, f3 dir, float power)    // src/utils.hxx:105-113
{
    const float cosTheta = fmaxf(0.f, dot(normal, dir));
    return (power + 1.f) * powf(cosTheta, power) * (SVCM_INV_PI * 0.5f);
}
SVCM_D void sampleConcentricDisc(float sx, float sy, float &ox, float &oy) // src/utils.hxx:119-163
{
    float phi, r;
    const float a = 2.f * sx - 1.f;
    const float b = 2.f * sy - 1.f;
    if(a > -b)
    {
        if(a > b) { r = a;  phi = (SVCM_PI / 4.f) * (b / a); }
        else      { r = b;  phi = (SVCM_PI / 4.f) * (2.f - (a / b)); }
    }
    else
    {
        if(a < b) { r = -a; phi = (SVCM_PI / 4.f) * (4.f + (b / a)); }
        else      { r = -b; phi = (b != 0.f) ? (SVCM_PI / 4.f) * (6.f - (a / b)) : 0.f; }
    }
    float s, c; sincosf(phi, &s, &c);
    ox = r * c;
    oy = r * s;
}
SVCM_D f3 sampleCosHemisphereW(float sx, float sy, float &oPdfW)       // src/utils.hxx:174-191
{
    const float term1 = 2.f * SVCM_PI * sx;
    const float term2 = sqrtf(1.f - sy);
    float s, c; sincosf(term1, &s, &c);
    const f3 ret = mk3(c * term2, s * term2, sqrtf(sy));
    oPdfW = ret.z * SVCM_INV_PI;
    return ret;
}
SVCM_D f3 sampleUniformSphereW(float sx, float sy, float &oPdf)        // src/utils.hxx:212-231
{
    const float term1 = 2.f * SVCM_PI * sx;
    const float term2 = 2.f * sqrtf(sy - sy * sy);
    float s, c; sincosf(term1, &s, &c);
    oPdf = SVCM_INV_PI * 0.25f;
    return mk3(c * term2, s * term2, 1.f - 2.f * sy);
}
SVCM_D float pdfWtoA(float pdfW, float dist, float cosThere)           // src/utils.hxx:245-251
{
    return pdfW * fabsf(cosThere) / sqr(dist);
}

// ---------------------------------------------------------------------------------------
// BSDF<FixIsLight>, reference src/bsdf.hxx:58-575
// ---------------------------------------------------------------------------------------
enum { EV_DIFFUSE = 1, EV_PHONG = 2, EV_REFLECT = 4, EV_REFRACT = 8, EV_SPECULAR = 12 };

struct Bsdf
{
    Frame3  frame;
    f3      localDirFix;
    int32_t matID;          // < 0: invalid
    float   diffProb, phongProb, reflProb, refrProb;
    float   contProb, reflectCoeff;
    SVCM_D bool  isValid() const { return matID >= 0; }
    SVCM_D bool  isDelta() const { return diffProb == 0.f && phongProb == 0.f; }
    SVCM_D float cosThetaFix() const { return localDirFix.z; }
    SVCM_D f3    worldDirFix() const { return frame.toWorld(localDirFix); }
};

// GetComponentProbabilities, reference src/bsdf.hxx:528-566
SVCM_D void bsdfProbabilities(const SvcmMaterial &m, Bsdf &b)
{
    b.reflectCoeff = fresnelDielectric(b.localDirFix.z, m.ior);
    const f3 diffuse = ld3(m.diffuse), phong = ld3(m.phong), mirror = ld3(m.mirror);
    const float aD = luminance(diffuse);
    const float aP = luminance(phong);
    const float aR = b.reflectCoeff * luminance(mirror);
    const float aT = (1.f - b.reflectCoeff) * (m.ior > 0.f ? 1.f : 0.f);
    const float total = aD + aP + aR + aT;
    if(total < 1e-9f)
    {
        b.diffProb = b.phongProb = b.reflProb = b.refrProb = 0.f;
        b.contProb = 0.f;
    }
    else
    {
        b.diffProb  = aD / total;
        b.phongProb = aP / total;
        b.reflProb  = aR / total;
        b.refrProb  = aT / total;
        const float c = maxcomp(diffuse + phong + b.reflectCoeff * mirror) + (1.f - b.reflectCoeff);
        b.contProb = fminf(1.f, fmaxf(0.f, c));
    }
}

// BSDF::Setup from an already-built frame and local fixed direction (used when a stored
// vertex is re-hydrated: frame and localDirFix are what Setup computed at the hit)
SVCM_D void bsdfFromLocal(const DevScene &sc, int32_t matID, Bsdf &b)
{
    if(fabsf(b.localDirFix.z) < SVCM_EPS_COS) { b.matID = -1; return; }
    bsdfProbabilities(sc.mats[matID], b);
    b.matID = matID;
}

// BSDF::Setup, reference src/bsdf.hxx:95-117
SVCM_D void bsdfSetup(const DevScene &sc, f3 rayDir, f3 normal, int32_t matID, Bsdf &b)
{
    b.frame.setFromZ(normal);
    b.localDirFix = b.frame.toLocal(-rayDir);
    b.diffProb = b.phongProb = b.reflProb = b.refrProb = 0.f;
    b.contProb = 0.f; b.reflectCoeff = 0.f;
    bsdfFromLocal(sc, matID, b);
}

// EvaluateDiffuse + EvaluatePhong with both pdfs, reference src/bsdf.hxx:393-450
SVCM_D f3 bsdfEvaluateLocal(const SvcmMaterial &m, const Bsdf &b, f3 localDirGen,
                            float &oDirectPdfW, float &oReversePdfW)
{
    f3 result = mk3(0.f);
    oDirectPdfW = 0.f; oReversePdfW = 0.f;
    if(b.localDirFix.z < SVCM_EPS_COS || localDirGen.z < SVCM_EPS_COS) return result;
    if(b.diffProb != 0.f)
    {
        oDirectPdfW  += b.diffProb * fmaxf(0.f, localDirGen.z * SVCM_INV_PI);
        oReversePdfW += b.diffProb * fmaxf(0.f, b.localDirFix.z * SVCM_INV_PI);
        result = result + ld3(m.diffuse) * SVCM_INV_PI;
    }
    if(b.phongProb != 0.f)
    {
        const f3 refl = reflectLocal(b.localDirFix);
        const float dotRWi = dot(refl, localDirGen);
        if(dotRWi > SVCM_EPS_PHONG)
        {
            // pow(dot, n) feeds both the pdf and the lobe; one evaluation serves both
            const float pw = powf(dotRWi, m.phongExponent);
            const float pdfW = b.phongProb * ((m.phongExponent + 1.f) * pw * (SVCM_INV_PI * 0.5f));
            oDirectPdfW  += pdfW;
            oReversePdfW += pdfW;
            const f3 rho = ld3(m.phong) * (m.phongExponent + 2.f) * 0.5f * SVCM_INV_PI;
            result = result + rho * pw;
        }
    }
    return result;
}

// BSDF::Evaluate, reference src/bsdf.hxx:128-153. Returns zero when the directions are on
// opposite sides; oCosThetaGen = |cos| of the generated direction.
SVCM_D f3 bsdfEvaluate(const DevScene &sc, const Bsdf &b, f3 worldDirGen, float &oCosThetaGen,
                       float &oDirectPdfW, float &oReversePdfW)
{
    oDirectPdfW = 0.f; oReversePdfW = 0.f; oCosThetaGen = 0.f;
    const f3 localDirGen = b.frame.toLocal(worldDirGen);
    if(localDirGen.z * b.localDirFix.z < 0.f) return mk3(0.f);
    oCosThetaGen = fabsf(localDirGen.z);
    return bsdfEvaluateLocal(sc.mats[b.matID], b, localDirGen, oDirectPdfW, oReversePdfW);
}

// BSDF::Pdf(aEvalRevPdf = true), reference src/bsdf.hxx:161-180 with PdfDiffuse / PdfPhong
// (src/bsdf.hxx:456-502) — note: unlike Evaluate* these have no EPS_COSINE rejection
SVCM_D float bsdfReversePdf(const DevScene &sc, const Bsdf &b, f3 worldDirGen)
{
    const f3 localDirGen = b.frame.toLocal(worldDirGen);
    if(localDirGen.z * b.localDirFix.z < 0.f) return 0.f;
    const SvcmMaterial &m = sc.mats[b.matID];
    float rev = 0.f;
    if(b.diffProb != 0.f) rev += b.diffProb * fmaxf(0.f, b.localDirFix.z * SVCM_INV_PI);
    if(b.phongProb != 0.f)
    {
        const f3 refl = reflectLocal(b.localDirFix);
        const float dotRWi = dot(refl, localDirGen);
        if(dotRWi > SVCM_EPS_PHONG)
            rev += powerCosHemispherePdfW(refl, localDirGen, m.phongExponent) * b.phongProb;
    }
    return rev;
}

// BSDF::Sample and the four Sample* helpers, reference src/bsdf.hxx:191-387.
// Returns zero (sample to be discarded) exactly where the reference does.
template<bool kFixIsLight>
SVCM_D f3 bsdfSample(const DevScene &sc, const Bsdf &b, float r0, float r1, float r2,
                     f3 &oWorldDirGen, float &oPdfW, float &oCosThetaGen, uint32_t &oEvent)
{
    uint32_t ev;
    if(r2 < b.diffProb)                               ev = EV_DIFFUSE;
    else if(r2 < b.diffProb + b.phongProb)            ev = EV_PHONG;
    else if(r2 < b.diffProb + b.phongProb + b.reflProb) ev = EV_REFLECT;
    else                                              ev = EV_REFRACT;
    oEvent = ev;

    const SvcmMaterial &m = sc.mats[b.matID];
    oPdfW = 0.f;
    f3 result = mk3(0.f);
    f3 localDirGen = mk3(0.f);

    if(ev == EV_DIFFUSE)
    {
        if(b.localDirFix.z < SVCM_EPS_COS) return mk3(0.f);               // SampleDiffuse :274-289
        float unweighted;
        localDirGen = sampleCosHemisphereW(r0, r1, unweighted);
        oPdfW += unweighted * b.diffProb;
        result = ld3(m.diffuse) * SVCM_INV_PI;
        if(iszero(result)) return mk3(0.f);
        if(b.phongProb != 0.f && localDirGen.z >= SVCM_EPS_COS)           // + EvaluatePhong
        {
            const f3 refl = reflectLocal(b.localDirFix);
            const float dotRWi = dot(refl, localDirGen);
            if(dotRWi > SVCM_EPS_PHONG)
            {
                const float pw = powf(dotRWi, m.phongExponent);
                oPdfW += b.phongProb * ((m.phongExponent + 1.f) * pw * (SVCM_INV_PI * 0.5f));
                result = result + (ld3(m.phong) * (m.phongExponent + 2.f) * 0.5f * SVCM_INV_PI) * pw;
            }
        }
    }
    else if(ev == EV_PHONG)
    {
        localDirGen = samplePowerCosHemisphereW(r0, r1, m.phongExponent);   // SamplePhong :291-318
        const f3 refl = reflectLocal(b.localDirFix);
        {
            Frame3 fr; fr.setFromZ(refl);
            localDirGen = fr.toWorld(localDirGen);
        }
        const float dotRWi = dot(refl, localDirGen);
        if(dotRWi <= SVCM_EPS_PHONG) return mk3(0.f);
        const float pw = powf(dotRWi, m.phongExponent);
        // PdfPhong (no cosine rejection), src/bsdf.hxx:476-502; dotRWi > 0 here, so the
        // max(0, .) of PowerCosHemispherePdfW is the identity and pw serves pdf and lobe
        oPdfW += ((m.phongExponent + 1.f) * pw * (SVCM_INV_PI * 0.5f)) * b.phongProb;
        result = (ld3(m.phong) * (m.phongExponent + 2.f) * 0.5f * SVCM_INV_PI) * pw;
        if(iszero(result)) return mk3(0.f);
        if(b.diffProb != 0.f && b.localDirFix.z >= SVCM_EPS_COS && localDirGen.z >= SVCM_EPS_COS)
        {
            oPdfW += b.diffProb * fmaxf(0.f, localDirGen.z * SVCM_INV_PI);  // + EvaluateDiffuse
            result = result + ld3(m.diffuse) * SVCM_INV_PI;
        }
    }
    else if(ev == EV_REFLECT)
    {
        localDirGen = reflectLocal(b.localDirFix);                          // SampleReflect :320-333
        oPdfW += b.reflProb;
        result = (b.reflectCoeff * ld3(m.mirror)) / fabsf(localDirGen.z);
        if(iszero(result)) return mk3(0.f);
    }
    else
    {
        if(m.ior < 0.f) return mk3(0.f);                                    // SampleRefract :335-387
        float cosI = b.localDirFix.z, cosT, eta;
        if(cosI < 0.f) { eta = m.ior; cosI = -cosI; cosT = 1.f; }
        else           { eta = 1.f / m.ior; cosT = -1.f; }
        const float sinI2 = 1.f - cosI * cosI;
        const float sinT2 = sqr(eta) * sinI2;
        if(!(sinT2 < 1.f)) return mk3(0.f);                                 // total internal reflection
        cosT *= sqrtf(fmaxf(0.f, 1.f - sinT2));
        localDirGen = mk3(-eta * b.localDirFix.x, -eta * b.localDirFix.y, cosT);
        oPdfW += b.refrProb;
        const float refractCoeff = 1.f - b.reflectCoeff;
        // only camera paths carry the eta^2 radiance scaling (adjoint BSDF), :378-381
        result = kFixIsLight ? mk3(refractCoeff / fabsf(cosT))
                             : mk3(refractCoeff * sqr(eta) / fabsf(cosT));
        if(iszero(result)) return mk3(0.f);
    }

    oCosThetaGen = fabsf(localDirGen.z);
    if(oCosThetaGen < SVCM_EPS_COS) return mk3(0.f);
    oWorldDirGen = b.frame.toWorld(localDirGen);
    return result;
}

// ---------------------------------------------------------------------------------------
// lights, reference src/lights.hxx
// ---------------------------------------------------------------------------------------
SVCM_D bool lightIsFinite(int32_t type) { return type == SVCM_LIGHT_AREA || type == SVCM_LIGHT_POINT; }
SVCM_D bool lightIsDelta(int32_t type)  { return type == SVCM_LIGHT_DIRECTIONAL || type == SVCM_LIGHT_POINT; }

// AbstractLight::Illuminate: src/lights.hxx:131-171 (area), :244-265 (directional),
// :329-353 (point), :410-440 (background)
SVCM_D f3 lightIlluminate(const DevScene &sc, const DevLight &l, f3 receivingPos, float rx, float ry,
                          f3 &oDirToLight, float &oDistance, float &oDirectPdfW,
                          float &oEmissionPdfW, float &oCosAtLight)
{
    const f3 intensity = ld3(l.intensity);
    if(l.type == SVCM_LIGHT_AREA)
    {
        const float term = sqrtf(rx);                                  // SampleUniformTriangle
        const float u = 1.f - term, v = ry * term;
        const f3 lightPoint = ld3(l.p0) + ld3(l.e1) * u + ld3(l.e2) * v;
        oDirToLight = lightPoint - receivingPos;
        const float distSqr = lensqr(oDirToLight);
        oDistance   = sqrtf(distSqr);
        oDirToLight = oDirToLight / oDistance;
        const float cosNormalDir = dot(ld3(l.fz), -oDirToLight);
        if(cosNormalDir < SVCM_EPS_COS) return mk3(0.f);
        oDirectPdfW   = l.invArea * distSqr / cosNormalDir;
        oCosAtLight   = cosNormalDir;
        oEmissionPdfW = l.invArea * cosNormalDir * SVCM_INV_PI;
        return intensity;
    }
    if(l.type == SVCM_LIGHT_DIRECTIONAL)
    {
        oDirToLight   = -ld3(l.fz);
        oDistance     = 1e36f;
        oDirectPdfW   = 1.f;
        oCosAtLight   = 1.f;
        oEmissionPdfW = SVCM_INV_PI * sc.sphereInvRadiusSqr;
        return intensity;
    }
    if(l.type == SVCM_LIGHT_POINT)
    {
        oDirToLight = ld3(l.p0) - receivingPos;
        const float distSqr = lensqr(oDirToLight);
        oDirectPdfW   = distSqr;
        oDistance     = sqrtf(distSqr);
        oDirToLight   = oDirToLight / oDistance;
        oCosAtLight   = 1.f;
        oEmissionPdfW = SVCM_INV_PI * 0.25f;
        return intensity;
    }
    oDirToLight   = sampleUniformSphereW(rx, ry, oDirectPdfW);           // background
    oDistance     = 1e36f;
    oEmissionPdfW = oDirectPdfW * SVCM_INV_PI * sc.sphereInvRadiusSqr;
    oCosAtLight   = 1.f;
    return intensity;
}

// AbstractLight::Emit: src/lights.hxx:173-201 (area), :267-292 (directional),
// :355-377 (point), :442-480 (background)
SVCM_D f3 lightEmit(const DevScene &sc, const DevLight &l, float dirX, float dirY, float posX,
                    float posY, f3 &oPosition, f3 &oDirection, float &oEmissionPdfW,
                    float &oDirectPdfA, float &oCosThetaLight)
{
    const f3 intensity = ld3(l.intensity);
    const f3 center = ld3(sc.sphereCenter);
    if(l.type == SVCM_LIGHT_AREA)
    {
        const float term = sqrtf(posX);
        const float u = 1.f - term, v = posY * term;
        oPosition = ld3(l.p0) + ld3(l.e1) * u + ld3(l.e2) * v;
        f3 localDirOut = sampleCosHemisphereW(dirX, dirY, oEmissionPdfW);
        oEmissionPdfW *= l.invArea;
        localDirOut.z = fmaxf(localDirOut.z, SVCM_EPS_COS);
        oDirection     = ld3(l.fx) * localDirOut.x + ld3(l.fy) * localDirOut.y + ld3(l.fz) * localDirOut.z;
        oDirectPdfA    = l.invArea;
        oCosThetaLight = localDirOut.z;
        return intensity * localDirOut.z;
    }
    if(l.type == SVCM_LIGHT_DIRECTIONAL)
    {
        float dx, dy;
        sampleConcentricDisc(posX, posY, dx, dy);
        oPosition  = center + sc.sphereRadius * (-ld3(l.fz) + ld3(l.fx) * dx + ld3(l.fy) * dy);
        oDirection = ld3(l.fz);
        oEmissionPdfW  = SVCM_INV_PI * sc.sphereInvRadiusSqr;
        oDirectPdfA    = 1.f;
        oCosThetaLight = 1.f;
        return intensity;
    }
    if(l.type == SVCM_LIGHT_POINT)
    {
        oPosition  = ld3(l.p0);
        oDirection = sampleUniformSphereW(dirX, dirY, oEmissionPdfW);
        oDirectPdfA    = 1.f;
        oCosThetaLight = 1.f;
        return intensity;
    }
    float directPdf;                                                    // background
    oDirection = sampleUniformSphereW(dirX, dirY, directPdf);
    float dx, dy;
    sampleConcentricDisc(posX, posY, dx, dy);
    Frame3 fr; fr.setFromZ(oDirection);
    oPosition = center + sc.sphereRadius * (-oDirection + fr.x * dx + fr.y * dy);
    oEmissionPdfW  = directPdf * SVCM_INV_PI * sc.sphereInvRadiusSqr;
    oDirectPdfA    = directPdf;
    oCosThetaLight = 1.f;
    return intensity;
}

// AbstractLight::GetRadiance: src/lights.hxx:203-223 (area), :482-502 (background); the two
// delta lights return zero (:294-303, :379-387)
SVCM_D f3 lightRadiance(const DevScene &sc, const DevLight &l, f3 rayDir, float &oDirectPdfA,
                        float &oEmissionPdfW)
{
    oDirectPdfA = 0.f; oEmissionPdfW = 0.f;
    if(l.type == SVCM_LIGHT_AREA)
    {
        const float cosOutL = fmaxf(0.f, dot(ld3(l.fz), -rayDir));
        if(cosOutL == 0.f) return mk3(0.f);
        oDirectPdfA   = l.invArea;
        oEmissionPdfW = (cosOutL * SVCM_INV_PI) * l.invArea;
        return ld3(l.intensity);
    }
    if(l.type == SVCM_LIGHT_BACKGROUND)
    {
        const float directPdf = SVCM_INV_PI * 0.25f;
        oDirectPdfA   = directPdf;
        oEmissionPdfW = directPdf * (SVCM_INV_PI * sc.sphereInvRadiusSqr);
        return ld3(l.intensity);
    }
    return mk3(0.f);
}

// Mat4f::TransformPoint, reference src/math.hxx:202-223 (element (r,c) at m[r + 4c])
SVCM_D f3 transformPoint(const float *m, f3 v)
{
    const float w = m[15] + m[3] * v.x + m[7] * v.y + m[11] * v.z;
    const float invW = 1.f / w;
    return mk3((m[12] + v.x * m[0] + v.y * m[4] + v.z * m[8])  * invW,
               (m[13] + v.x * m[1] + v.y * m[5] + v.z * m[9])  * invW,
               (m[14] + v.x * m[2] + v.y * m[6] + v.z * m[10]) * invW);
}

} // namespace svcm
#endif // SVCM_DEVICE_CUH_
