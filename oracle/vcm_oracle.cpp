// =========================================================================================
// TEST INFRASTRUCTURE — the parity oracle. NOT part of the product.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load this library; libsvcm.so never links, loads or calls it.
//
// What it is: a serial CPU restatement, in plain C++ (g++ -O3 -ffp-contract=off, no
// fast-math), of the algorithm on SmallVCM's hot path — VertexCM::RunIteration and
// everything it calls. Every function cites the reference file:line it follows. It works
// on the FLAT scene of include/svcm.h (the same bytes the CUDA library receives), not on
// the reference's classes.
//
// Pinning (tests/test_oracle.py, run where oracle/_ref exists):
//   * rngMode 0 (one sequential std::mt19937_64 stream, src/rng.hxx:41-86) reproduces the
//     UNMODIFIED reference's framebuffer BIT FOR BIT (BASELINE.md §3 known answers:
//     scene 1 vcm 512x512 TotalLuminance 69769.8, ...);
//   * rngMode 1 (per-path counter stream, the one the GPU uses) reproduces, bit for bit,
//     the reference arithmetic compiled with that generator swapped in
//     (oracle/_ref/libsvcm_ref_ctr.so): images, light vertices, hash grid, merge sums.
// Bit-exactness needs the same association order as the reference in every float
// expression; where that dictates an odd-looking form (0 + a*b, (a*b)*c ...) it is kept.
// =========================================================================================
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <random>
#include <algorithm>
#include <omp.h>

#include "svcm.h"

#define ORACLE_API extern "C" __attribute__((visibility("default")))

namespace {

typedef unsigned int uint;

// src/math.hxx:29-30
static const float kPi    = 3.14159265358979f;
static const float kInvPi = 1.f / kPi;
// src/utils.hxx:32-33, src/bsdf.hxx:56
static const float kEpsCosine = 1e-6f;
static const float kEpsRay    = 1e-3f;
static const float kEpsPhong  = 1e-3f;

// ---------------------------------------------------------------------------------------
// vectors: src/math.hxx:88-165. All binary operators are component-wise; a scalar operand
// is broadcast first (Vec3x(T a) constructor), which for * and / equals the scalar form.
// ---------------------------------------------------------------------------------------
struct V3
{
    float x, y, z;
    V3() {}
    explicit V3(float a) : x(a), y(a), z(a) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
};

inline V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(const V3 &a, const V3 &b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(const V3 &a, const V3 &b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3 operator*(const V3 &a, float s)     { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, const V3 &a)     { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(const V3 &a, float s)     { return V3(a.x / s, a.y / s, a.z / s); }
inline V3 operator-(const V3 &a)              { return V3(-a.x, -a.y, -a.z); }

// Dot accumulates from zero: ((0 + x) + y) + z  (src/math.hxx:129-130)
inline float Dot(const V3 &a, const V3 &b)
{
    float res = 0.f;
    res += a.x * b.x;
    res += a.y * b.y;
    res += a.z * b.z;
    return res;
}
inline float LenSqr(const V3 &a) { return Dot(a, a); }
inline float Length(const V3 &a) { return std::sqrt(LenSqr(a)); }
inline bool  IsZero(const V3 &a) { return a.x == 0 && a.y == 0 && a.z == 0; }
inline float MaxComp(const V3 &a) { return std::max(std::max(a.x, a.y), a.z); }

// src/math.hxx:145-154
inline V3 Cross(const V3 &a, const V3 &b)
{
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// src/math.hxx:156-161
inline V3 Normalize(const V3 &a)
{
    const float len = std::sqrt(Dot(a, a));
    return a / len;
}
inline float Sqr(float a) { return a * a; }
inline V3 Load3(const float *p) { return V3(p[0], p[1], p[2]); }
inline void Store3(float *p, const V3 &v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// src/utils.hxx:36-41
inline float Luminance(const V3 &c)
{
    return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z;
}

// ---------------------------------------------------------------------------------------
// Frame: src/frame.hxx:32-78
// ---------------------------------------------------------------------------------------
struct Frame
{
    V3 mX, mY, mZ;
    void SetFromZ(const V3 &z)
    {
        const V3 tmpZ = mZ = Normalize(z);
        const V3 tmpX = (std::abs(tmpZ.x) > 0.99f) ? V3(0, 1, 0) : V3(1, 0, 0);
        mY = Normalize(Cross(tmpZ, tmpX));
        mX = Cross(mY, tmpZ);
    }
    V3 ToWorld(const V3 &a) const { return mX * a.x + mY * a.y + mZ * a.z; }
    V3 ToLocal(const V3 &a) const { return V3(Dot(a, mX), Dot(a, mY), Dot(a, mZ)); }
};

// ---------------------------------------------------------------------------------------
// random numbers
// ---------------------------------------------------------------------------------------
struct Rng
{
    int mMode;  // 0: the reference's own generator; 1: per-path counter stream
    // mode 0 — src/rng.hxx:41-86
    std::mt19937_64 mEngine;
    std::uniform_real_distribution<float> mDistFloat;
    // mode 1 — Philox2x32-7, counter (draw, path), key seed + 0x85EBCA6B*(2*iteration+phase);
    // restated identically in oracle/ref_glue/ref_counter_rng.cpp and the product
    uint mSeed, mKey, mPath, mDraw;

    Rng(int aSeed, int aMode) : mMode(aMode), mEngine(aSeed), mSeed(uint(aSeed)),
        mKey(uint(aSeed)), mPath(0), mDraw(0) {}

    void SetPath(int aIteration, int aPhase, int aPathIdx)
    {
        mKey  = mSeed + 0x85EBCA6Bu * (2u * uint(aIteration) + uint(aPhase));
        mPath = uint(aPathIdx);
        mDraw = 0;
    }

    static uint Philox(uint c0, uint c1, uint key)
    {
        for(int round = 0; round < 7; round++)
        {
            const uint64_t prod = uint64_t(0xD256D193u) * c0;
            const uint hi = uint(prod >> 32), lo = uint(prod);
            c0 = hi ^ key ^ c1;
            c1 = lo;
            key += 0x9E3779B9u;
        }
        return c0;
    }

    float GetFloat()
    {
        if(mMode == 0) return mDistFloat(mEngine);
        return float((Philox(mDraw++, mPath, mKey) >> 8) | 1u) * (1.0f / 16777216.0f);
    }
};

// ---------------------------------------------------------------------------------------
// scene (flat copy of SvcmSceneDesc)
// ---------------------------------------------------------------------------------------
struct Scene
{
    std::vector<SvcmPrimitive> prims;
    std::vector<SvcmMaterial>  mats;
    std::vector<SvcmLight>     lights;
    std::vector<int>           mat2light;
    int                        background;
    SvcmCamera                 camera;
    SvcmSceneSphere            sphere;
};

struct Ray   { V3 org, dir; float tmin; };
struct Isect { float dist; int matID; int lightID; V3 normal; };

// Triangle::Intersect, src/geometry.hxx:125-156
inline bool IntersectTriangle(const SvcmPrimitive &t, const Ray &ray, Isect &res)
{
    const V3 p0 = Load3(t.v), p1 = Load3(t.v + 3), p2 = Load3(t.v + 6), n = Load3(t.v + 9);
    const V3 ao = p0 - ray.org;
    const V3 bo = p1 - ray.org;
    const V3 co = p2 - ray.org;

    const V3 v0 = Cross(co, bo);
    const V3 v1 = Cross(bo, ao);
    const V3 v2 = Cross(ao, co);

    const float v0d = Dot(v0, ray.dir);
    const float v1d = Dot(v1, ray.dir);
    const float v2d = Dot(v2, ray.dir);

    if(((v0d < 0.f)  && (v1d < 0.f)  && (v2d < 0.f)) ||
       ((v0d >= 0.f) && (v1d >= 0.f) && (v2d >= 0.f)))
    {
        const float distance = Dot(n, ao) / Dot(n, ray.dir);
        if((distance > ray.tmin) && (distance < res.dist))
        {
            res.normal = n;
            res.matID  = t.matID;
            res.dist   = distance;
            return true;
        }
    }
    return false;
}

// Sphere::Intersect, src/geometry.hxx:198-237 — float A,B,C and discriminant, then doubles
inline bool IntersectSphere(const SvcmPrimitive &s, const Ray &ray, Isect &res)
{
    const V3 center = Load3(s.v);
    const float radius = s.v[3];
    const V3 to = ray.org - center;

    const float A = Dot(ray.dir, ray.dir);
    const float B = 2 * Dot(ray.dir, to);
    const float C = Dot(to, to) - (radius * radius);

    const double disc = B * B - 4 * A * C;   // evaluated in float, widened afterwards
    if(disc < 0) return false;

    const double discSqrt = std::sqrt(disc);
    const double q = (B < 0) ? ((-B - discSqrt) / 2.f) : ((-B + discSqrt) / 2.f);

    double t0 = q / A;
    double t1 = C / q;
    if(t0 > t1) std::swap(t0, t1);

    float resT;
    if(t0 > ray.tmin && t0 < res.dist)      resT = float(t0);
    else if(t1 > ray.tmin && t1 < res.dist) resT = float(t1);
    else return false;

    res.dist   = resT;
    res.matID  = s.matID;
    res.normal = Normalize(to + V3(resT) * ray.dir);
    return true;
}

inline bool IntersectPrim(const SvcmPrimitive &p, const Ray &ray, Isect &res)
{
    return p.type == SVCM_PRIM_TRIANGLE ? IntersectTriangle(p, ray, res)
                                        : IntersectSphere(p, ray, res);
}

// Scene::Intersect -> GeometryList::Intersect, src/scene.hxx:53-70, src/geometry.hxx:65-78.
// oPrim (optional) = list index of the closest hit.
inline bool SceneIntersect(const Scene &sc, const Ray &ray, Isect &res, int *oPrim = NULL)
{
    bool any = false;
    for(size_t i = 0; i < sc.prims.size(); i++)
    {
        if(IntersectPrim(sc.prims[i], ray, res))
        {
            any = true;
            if(oPrim) *oPrim = int(i);
        }
    }
    if(any)
    {
        res.lightID = -1;
        if(res.matID >= 0 && res.matID < (int)sc.mat2light.size())
            res.lightID = sc.mat2light[res.matID];
    }
    return any;
}

// Scene::Occluded -> GeometryList::IntersectP, src/scene.hxx:72-85, src/geometry.hxx:80-91
inline bool SceneOccluded(const Scene &sc, const V3 &point, const V3 &dir, float tMax)
{
    Ray ray;
    ray.org  = point + dir * kEpsRay;
    ray.dir  = dir;
    ray.tmin = 0;
    Isect isect;
    isect.dist = tMax - 2 * kEpsRay;
    for(size_t i = 0; i < sc.prims.size(); i++)
        if(IntersectPrim(sc.prims[i], ray, isect))
            return true;
    return false;
}

// ---------------------------------------------------------------------------------------
// sampling helpers, src/utils.hxx
// ---------------------------------------------------------------------------------------
// src/utils.hxx:43-74
inline float FresnelDielectric(float cosInc, float ior)
{
    if(ior < 0) return 1.f;
    float etaIncOverEtaTrans;
    if(cosInc < 0.f) { cosInc = -cosInc; etaIncOverEtaTrans = ior; }
    else             { etaIncOverEtaTrans = 1.f / ior; }

    const float sinTrans2 = Sqr(etaIncOverEtaTrans) * (1.f - Sqr(cosInc));
    const float cosTrans  = std::sqrt(std::max(0.f, 1.f - sinTrans2));

    const float term1 = etaIncOverEtaTrans * cosTrans;
    const float rParallel = (cosInc - term1) / (cosInc + term1);
    const float term2 = etaIncOverEtaTrans * cosInc;
    const float rPerpendicular = (term2 - cosTrans) / (term2 + cosTrans);
    return 0.5f * (Sqr(rParallel) + Sqr(rPerpendicular));
}

// src/utils.hxx:77-80
inline V3 ReflectLocal(const V3 &v) { return V3(-v.x, -v.y, v.z); }

// src/utils.hxx:85-103
inline V3 SamplePowerCosHemisphereW(float sx, float sy, float power)
{
    const float term1 = 2.f * kPi * sx;
    const float term2 = std::pow(sy, 1.f / (power + 1.f));
    const float term3 = std::sqrt(1.f - term2 * term2);
    return V3(std::cos(term1) * term3, std::sin(term1) * term3, term2);
}

// src/utils.hxx:105-113
inline float PowerCosHemispherePdfW(const V3 &normal, const V3 &dir, float power)
{
    const float cosTheta = std::max(0.f, Dot(normal, dir));
    return (power + 1.f) * std::pow(cosTheta, power) * (kInvPi * 0.5f);
}

// src/utils.hxx:119-163
inline void SampleConcentricDisc(float sx, float sy, float &ox, float &oy)
{
    float phi, r;
    float a = 2 * sx - 1;
    float b = 2 * sy - 1;
    if(a > -b)
    {
        if(a > b) { r = a;  phi = (kPi / 4.f) * (b / a); }
        else      { r = b;  phi = (kPi / 4.f) * (2.f - (a / b)); }
    }
    else
    {
        if(a < b) { r = -a; phi = (kPi / 4.f) * (4.f + (b / a)); }
        else
        {
            r = -b;
            if(b != 0) phi = (kPi / 4.f) * (6.f - (a / b));
            else       phi = 0;
        }
    }
    ox = r * std::cos(phi);
    oy = r * std::sin(phi);
}
inline float ConcentricDiscPdfA() { return kInvPi; }               // src/utils.hxx:165-168

// src/utils.hxx:174-191
inline V3 SampleCosHemisphereW(float sx, float sy, float *oPdfW)
{
    const float term1 = 2.f * kPi * sx;
    const float term2 = std::sqrt(1.f - sy);
    const V3 ret(std::cos(term1) * term2, std::sin(term1) * term2, std::sqrt(sy));
    if(oPdfW) *oPdfW = ret.z * kInvPi;
    return ret;
}
// src/utils.hxx:193-198
inline float CosHemispherePdfW(const V3 &normal, const V3 &dir)
{
    return std::max(0.f, Dot(normal, dir)) * kInvPi;
}
// src/utils.hxx:202-207
inline void SampleUniformTriangle(float sx, float sy, float &u, float &v)
{
    const float term = std::sqrt(sx);
    u = 1.f - term;
    v = sy * term;
}
// src/utils.hxx:212-231
inline V3 SampleUniformSphereW(float sx, float sy, float *oPdf)
{
    const float term1 = 2.f * kPi * sx;
    const float term2 = 2.f * std::sqrt(sy - sy * sy);
    const V3 ret(std::cos(term1) * term2, std::sin(term1) * term2, 1.f - 2.f * sy);
    if(oPdf) *oPdf = kInvPi * 0.25f;
    return ret;
}
inline float UniformSpherePdfW() { return kInvPi * 0.25f; }       // src/utils.hxx:233-237
// src/utils.hxx:245-251
inline float PdfWtoA(float pdfW, float dist, float cosThere)
{
    return pdfW * std::abs(cosThere) / Sqr(dist);
}

// ---------------------------------------------------------------------------------------
// BSDF<FixIsLight>, src/bsdf.hxx:58-575. fixIsLight is a run-time flag here; it only
// matters in SampleRefract (src/bsdf.hxx:378-381).
// ---------------------------------------------------------------------------------------
enum { kDiffuse = 1, kPhong = 2, kReflect = 4, kRefract = 8, kSpecular = 12 };

struct Bsdf
{
    int   matID;        // < 0: invalid
    Frame frame;
    V3    localDirFix;
    bool  isDelta;
    float diffProb, phongProb, reflProb, refrProb;
    float contProb;
    float reflectCoeff;
    bool  fixIsLight;

    float CosThetaFix() const { return localDirFix.z; }
    V3    WorldDirFix() const { return frame.ToWorld(localDirFix); }
};

// GetComponentProbabilities, src/bsdf.hxx:528-566
inline void ComponentProbabilities(const SvcmMaterial &m, Bsdf &b)
{
    b.reflectCoeff = FresnelDielectric(b.localDirFix.z, m.ior);

    const V3 diffuse = Load3(m.diffuse), phong = Load3(m.phong), mirror = Load3(m.mirror);
    const float albedoDiffuse = Luminance(diffuse);
    const float albedoPhong   = Luminance(phong);
    const float albedoReflect = b.reflectCoeff         * Luminance(mirror);
    const float albedoRefract = (1.f - b.reflectCoeff) * (m.ior > 0.f ? 1.f : 0.f);

    const float totalAlbedo = albedoDiffuse + albedoPhong + albedoReflect + albedoRefract;

    if(totalAlbedo < 1e-9f)
    {
        b.diffProb = b.phongProb = b.reflProb = b.refrProb = 0.f;
        b.contProb = 0.f;
    }
    else
    {
        b.diffProb  = albedoDiffuse / totalAlbedo;
        b.phongProb = albedoPhong   / totalAlbedo;
        b.reflProb  = albedoReflect / totalAlbedo;
        b.refrProb  = albedoRefract / totalAlbedo;
        b.contProb  = MaxComp(diffuse + phong + b.reflectCoeff * mirror) + (1.f - b.reflectCoeff);
        b.contProb  = std::min(1.f, std::max(0.f, b.contProb));
    }
}

// BSDF::Setup, src/bsdf.hxx:95-117
inline void BsdfSetup(const Scene &sc, const V3 &rayDir, const Isect &isect, bool fixIsLight, Bsdf &b)
{
    b.matID = -1;
    b.fixIsLight = fixIsLight;
    b.isDelta = false;
    b.diffProb = b.phongProb = b.reflProb = b.refrProb = 0.f;
    b.contProb = 0.f;
    b.reflectCoeff = 0.f;
    b.frame.SetFromZ(isect.normal);
    b.localDirFix = b.frame.ToLocal(-rayDir);

    if(std::abs(b.localDirFix.z) < kEpsCosine) return;

    ComponentProbabilities(sc.mats[isect.matID], b);
    b.isDelta = (b.diffProb == 0) && (b.phongProb == 0);
    b.matID = isect.matID;
}

// EvaluateDiffuse, src/bsdf.hxx:393-412
inline V3 EvaluateDiffuse(const Bsdf &b, const SvcmMaterial &m, const V3 &localDirGen,
                          float *oDirectPdfW, float *oReversePdfW)
{
    if(b.diffProb == 0) return V3(0);
    if(b.localDirFix.z < kEpsCosine || localDirGen.z < kEpsCosine) return V3(0);
    if(oDirectPdfW)  *oDirectPdfW  += b.diffProb * std::max(0.f, localDirGen.z * kInvPi);
    if(oReversePdfW) *oReversePdfW += b.diffProb * std::max(0.f, b.localDirFix.z * kInvPi);
    return Load3(m.diffuse) * kInvPi;
}

// EvaluatePhong, src/bsdf.hxx:414-450
inline V3 EvaluatePhong(const Bsdf &b, const SvcmMaterial &m, const V3 &localDirGen,
                        float *oDirectPdfW, float *oReversePdfW)
{
    if(b.phongProb == 0) return V3(0);
    if(b.localDirFix.z < kEpsCosine || localDirGen.z < kEpsCosine) return V3(0);

    const V3 reflLocalDirIn = ReflectLocal(b.localDirFix);
    const float dot_R_Wi = Dot(reflLocalDirIn, localDirGen);
    if(dot_R_Wi <= kEpsPhong) return V3(0.f);

    if(oDirectPdfW || oReversePdfW)
    {
        const float pdfW = b.phongProb *
            PowerCosHemispherePdfW(reflLocalDirIn, localDirGen, m.phongExponent);
        if(oDirectPdfW)  *oDirectPdfW  += pdfW;
        if(oReversePdfW) *oReversePdfW += pdfW;
    }

    const V3 rho = Load3(m.phong) * (m.phongExponent + 2.f) * 0.5f * kInvPi;
    return rho * std::pow(dot_R_Wi, m.phongExponent);
}

// PdfDiffuse / PdfPhong, src/bsdf.hxx:456-502 (note: PdfPhong multiplies in the other order)
inline void PdfDiffuse(const Bsdf &b, const V3 &localDirGen, float *oDirectPdfW, float *oReversePdfW)
{
    if(b.diffProb == 0) return;
    if(oDirectPdfW)  *oDirectPdfW  += b.diffProb * std::max(0.f, localDirGen.z * kInvPi);
    if(oReversePdfW) *oReversePdfW += b.diffProb * std::max(0.f, b.localDirFix.z * kInvPi);
}
inline void PdfPhong(const Bsdf &b, const SvcmMaterial &m, const V3 &localDirGen,
                     float *oDirectPdfW, float *oReversePdfW)
{
    if(b.phongProb == 0) return;
    const V3 reflLocalDirIn = ReflectLocal(b.localDirFix);
    const float dot_R_Wi = Dot(reflLocalDirIn, localDirGen);
    if(dot_R_Wi <= kEpsPhong) return;
    if(oDirectPdfW || oReversePdfW)
    {
        const float pdfW = PowerCosHemispherePdfW(reflLocalDirIn, localDirGen, m.phongExponent) *
            b.phongProb;
        if(oDirectPdfW)  *oDirectPdfW  += pdfW;
        if(oReversePdfW) *oReversePdfW += pdfW;
    }
}

// BSDF::Evaluate, src/bsdf.hxx:128-153
inline V3 BsdfEvaluate(const Scene &sc, const Bsdf &b, const V3 &worldDirGen, float &oCosThetaGen,
                       float *oDirectPdfW, float *oReversePdfW)
{
    V3 result(0);
    if(oDirectPdfW)  *oDirectPdfW = 0;
    if(oReversePdfW) *oReversePdfW = 0;

    const V3 localDirGen = b.frame.ToLocal(worldDirGen);
    if(localDirGen.z * b.localDirFix.z < 0) return result;

    oCosThetaGen = std::abs(localDirGen.z);
    const SvcmMaterial &m = sc.mats[b.matID];
    result = result + EvaluateDiffuse(b, m, localDirGen, oDirectPdfW, oReversePdfW);
    result = result + EvaluatePhong(b, m, localDirGen, oDirectPdfW, oReversePdfW);
    return result;
}

// BSDF::Pdf, src/bsdf.hxx:161-180
inline float BsdfPdf(const Scene &sc, const Bsdf &b, const V3 &worldDirGen, bool evalRevPdf)
{
    const V3 localDirGen = b.frame.ToLocal(worldDirGen);
    if(localDirGen.z * b.localDirFix.z < 0) return 0;
    const SvcmMaterial &m = sc.mats[b.matID];
    float directPdfW = 0, reversePdfW = 0;
    PdfDiffuse(b, localDirGen, &directPdfW, &reversePdfW);
    PdfPhong(b, m, localDirGen, &directPdfW, &reversePdfW);
    return evalRevPdf ? reversePdfW : directPdfW;
}

// BSDF::Sample and its four Sample* helpers, src/bsdf.hxx:191-387
inline V3 BsdfSample(const Scene &sc, const Bsdf &b, const V3 &rnd, V3 &oWorldDirGen,
                     float &oPdfW, float &oCosThetaGen, uint *oSampledEvent)
{
    uint sampledEvent;
    if(rnd.z < b.diffProb)                                   sampledEvent = kDiffuse;
    else if(rnd.z < b.diffProb + b.phongProb)                sampledEvent = kPhong;
    else if(rnd.z < b.diffProb + b.phongProb + b.reflProb)   sampledEvent = kReflect;
    else                                                     sampledEvent = kRefract;
    if(oSampledEvent) *oSampledEvent = sampledEvent;

    const SvcmMaterial &m = sc.mats[b.matID];
    oPdfW = 0;
    V3 result(0);
    V3 localDirGen(0);

    if(sampledEvent == kDiffuse)
    {
        // SampleDiffuse, src/bsdf.hxx:274-289
        if(b.localDirFix.z < kEpsCosine) return V3(0);
        float unweightedPdfW;
        localDirGen = SampleCosHemisphereW(rnd.x, rnd.y, &unweightedPdfW);
        oPdfW += unweightedPdfW * b.diffProb;
        result = result + Load3(m.diffuse) * kInvPi;
        if(IsZero(result)) return V3(0);
        result = result + EvaluatePhong(b, m, localDirGen, &oPdfW, NULL);
    }
    else if(sampledEvent == kPhong)
    {
        // SamplePhong, src/bsdf.hxx:291-318
        localDirGen = SamplePowerCosHemisphereW(rnd.x, rnd.y, m.phongExponent);
        const V3 reflLocalDirFixed = ReflectLocal(b.localDirFix);
        {
            Frame frame;
            frame.SetFromZ(reflLocalDirFixed);
            localDirGen = frame.ToWorld(localDirGen);
        }
        const float dot_R_Wi = Dot(reflLocalDirFixed, localDirGen);
        if(dot_R_Wi <= kEpsPhong) return V3(0);
        PdfPhong(b, m, localDirGen, &oPdfW, NULL);
        const V3 rho = Load3(m.phong) * (m.phongExponent + 2.f) * 0.5f * kInvPi;
        result = result + rho * std::pow(dot_R_Wi, m.phongExponent);
        if(IsZero(result)) return V3(0);
        result = result + EvaluateDiffuse(b, m, localDirGen, &oPdfW, NULL);
    }
    else if(sampledEvent == kReflect)
    {
        // SampleReflect, src/bsdf.hxx:320-333
        localDirGen = ReflectLocal(b.localDirFix);
        oPdfW += b.reflProb;
        result = result + b.reflectCoeff * Load3(m.mirror) / std::abs(localDirGen.z);
        if(IsZero(result)) return V3(0);
    }
    else
    {
        // SampleRefract, src/bsdf.hxx:335-387
        if(m.ior < 0) return V3(0);
        float cosI = b.localDirFix.z;
        float cosT;
        float etaIncOverEtaTrans;
        if(cosI < 0.f) { etaIncOverEtaTrans = m.ior; cosI = -cosI; cosT = 1.f; }
        else           { etaIncOverEtaTrans = 1.f / m.ior; cosT = -1.f; }

        const float sinI2 = 1.f - cosI * cosI;
        const float sinT2 = Sqr(etaIncOverEtaTrans) * sinI2;
        if(sinT2 < 1.f)
        {
            cosT *= std::sqrt(std::max(0.f, 1.f - sinT2));
            localDirGen = V3(-etaIncOverEtaTrans * b.localDirFix.x,
                             -etaIncOverEtaTrans * b.localDirFix.y, cosT);
            oPdfW += b.refrProb;
            const float refractCoeff = 1.f - b.reflectCoeff;
            if(!b.fixIsLight)
                result = result + V3(refractCoeff * Sqr(etaIncOverEtaTrans) / std::abs(cosT));
            else
                result = result + V3(refractCoeff / std::abs(cosT));
        }
        if(IsZero(result)) return V3(0);
    }

    oCosThetaGen = std::abs(localDirGen.z);
    if(oCosThetaGen < kEpsCosine) return V3(0.f);

    oWorldDirGen = b.frame.ToWorld(localDirGen);
    return result;
}

// ---------------------------------------------------------------------------------------
// lights, src/lights.hxx
// ---------------------------------------------------------------------------------------
inline bool LightIsFinite(const SvcmLight &l) { return l.type == SVCM_LIGHT_AREA || l.type == SVCM_LIGHT_POINT; }
inline bool LightIsDelta(const SvcmLight &l)  { return l.type == SVCM_LIGHT_DIRECTIONAL || l.type == SVCM_LIGHT_POINT; }

// AbstractLight::Illuminate for the four classes: src/lights.hxx:131-171 (area),
// :244-265 (directional), :329-353 (point), :410-440 (background)
inline V3 LightIlluminate(const SvcmLight &l, const SvcmSceneSphere &ss, const V3 &receivingPos,
                          float rx, float ry, V3 &oDirToLight, float &oDistance,
                          float &oDirectPdfW, float &oEmissionPdfW, float &oCosAtLight)
{
    const V3 intensity = Load3(l.intensity);
    switch(l.type)
    {
    case SVCM_LIGHT_AREA:
    {
        float u, v;
        SampleUniformTriangle(rx, ry, u, v);
        const V3 lightPoint = Load3(l.p0) + Load3(l.e1) * u + Load3(l.e2) * v;
        oDirToLight = lightPoint - receivingPos;
        const float distSqr = LenSqr(oDirToLight);
        oDistance   = std::sqrt(distSqr);
        oDirToLight = oDirToLight / oDistance;
        const float cosNormalDir = Dot(Load3(l.frameZ), -oDirToLight);
        if(cosNormalDir < kEpsCosine) return V3(0.f);
        oDirectPdfW   = l.invArea * distSqr / cosNormalDir;
        oCosAtLight   = cosNormalDir;
        oEmissionPdfW = l.invArea * cosNormalDir * kInvPi;
        return intensity;
    }
    case SVCM_LIGHT_DIRECTIONAL:
        oDirToLight   = -Load3(l.frameZ);
        oDistance     = 1e36f;
        oDirectPdfW   = 1.f;
        oCosAtLight   = 1.f;
        oEmissionPdfW = ConcentricDiscPdfA() * ss.invRadiusSqr;
        return intensity;
    case SVCM_LIGHT_POINT:
    {
        oDirToLight = Load3(l.p0) - receivingPos;
        const float distSqr = LenSqr(oDirToLight);
        oDirectPdfW   = distSqr;
        oDistance     = std::sqrt(distSqr);
        oDirToLight   = oDirToLight / oDistance;
        oCosAtLight   = 1.f;
        oEmissionPdfW = UniformSpherePdfW();
        return intensity;
    }
    default: // background
    {
        oDirToLight = SampleUniformSphereW(rx, ry, &oDirectPdfW);
        const V3 radiance = intensity * l.scale;
        oDistance     = 1e36f;
        oEmissionPdfW = oDirectPdfW * ConcentricDiscPdfA() * ss.invRadiusSqr;
        oCosAtLight   = 1.f;
        return radiance;
    }
    }
}

// AbstractLight::Emit: src/lights.hxx:173-201 (area), :267-292 (directional),
// :355-377 (point), :442-480 (background)
inline V3 LightEmit(const SvcmLight &l, const SvcmSceneSphere &ss, float dirX, float dirY,
                    float posX, float posY, V3 &oPosition, V3 &oDirection,
                    float &oEmissionPdfW, float &oDirectPdfA, float &oCosThetaLight)
{
    const V3 intensity = Load3(l.intensity);
    const V3 center = Load3(ss.center);
    switch(l.type)
    {
    case SVCM_LIGHT_AREA:
    {
        float u, v;
        SampleUniformTriangle(posX, posY, u, v);
        oPosition = Load3(l.p0) + Load3(l.e1) * u + Load3(l.e2) * v;
        V3 localDirOut = SampleCosHemisphereW(dirX, dirY, &oEmissionPdfW);
        oEmissionPdfW *= l.invArea;
        localDirOut.z = std::max(localDirOut.z, kEpsCosine);
        Frame f; f.mX = Load3(l.frameX); f.mY = Load3(l.frameY); f.mZ = Load3(l.frameZ);
        oDirection     = f.ToWorld(localDirOut);
        oDirectPdfA    = l.invArea;
        oCosThetaLight = localDirOut.z;
        return intensity * localDirOut.z;
    }
    case SVCM_LIGHT_DIRECTIONAL:
    {
        float dx, dy;
        SampleConcentricDisc(posX, posY, dx, dy);
        const V3 fx = Load3(l.frameX), fy = Load3(l.frameY), fz = Load3(l.frameZ);
        oPosition = center + ss.radius * (-fz + fx * dx + fy * dy);
        oDirection = fz;
        oEmissionPdfW  = ConcentricDiscPdfA() * ss.invRadiusSqr;
        oDirectPdfA    = 1.f;
        oCosThetaLight = 1.f;
        return intensity;
    }
    case SVCM_LIGHT_POINT:
        oPosition  = Load3(l.p0);
        oDirection = SampleUniformSphereW(dirX, dirY, &oEmissionPdfW);
        oDirectPdfA    = 1.f;
        oCosThetaLight = 1.f;
        return intensity;
    default: // background
    {
        float directPdf;
        oDirection = SampleUniformSphereW(dirX, dirY, &directPdf);
        const V3 radiance = intensity * l.scale;
        float dx, dy;
        SampleConcentricDisc(posX, posY, dx, dy);
        Frame frame;
        frame.SetFromZ(oDirection);
        oPosition = center + ss.radius * (-oDirection + frame.mX * dx + frame.mY * dy);
        oEmissionPdfW  = directPdf * ConcentricDiscPdfA() * ss.invRadiusSqr;
        oDirectPdfA    = directPdf;
        oCosThetaLight = 1.f;
        return radiance;
    }
    }
}

// AbstractLight::GetRadiance: src/lights.hxx:203-223 (area), :294-303 / :379-387 (delta
// lights: zero), :482-502 (background)
inline V3 LightGetRadiance(const SvcmLight &l, const SvcmSceneSphere &ss, const V3 &rayDir,
                           float &oDirectPdfA, float &oEmissionPdfW)
{
    switch(l.type)
    {
    case SVCM_LIGHT_AREA:
    {
        const V3 n = Load3(l.frameZ);
        const float cosOutL = std::max(0.f, Dot(n, -rayDir));
        if(cosOutL == 0) return V3(0);
        oDirectPdfA   = l.invArea;
        oEmissionPdfW = CosHemispherePdfW(n, -rayDir);
        oEmissionPdfW *= l.invArea;
        return Load3(l.intensity);
    }
    case SVCM_LIGHT_BACKGROUND:
    {
        const float directPdf = UniformSpherePdfW();
        const V3 radiance = Load3(l.intensity) * l.scale;
        const float positionPdf = ConcentricDiscPdfA() * ss.invRadiusSqr;
        oDirectPdfA   = directPdf;
        oEmissionPdfW = directPdf * positionPdf;
        return radiance;
    }
    default:
        return V3(0);
    }
}

// ---------------------------------------------------------------------------------------
// camera, src/camera.hxx + Mat4f::TransformPoint src/math.hxx:202-223
// ---------------------------------------------------------------------------------------
inline V3 TransformPoint(const float *m, const V3 &v)
{
    // element (r,c) at m[r + 4*c]
    const float in[3] = { v.x, v.y, v.z };
    float w = m[3 + 4 * 3];
    for(int c = 0; c < 3; c++) w += m[3 + 4 * c] * in[c];
    const float invW = 1.f / w;
    float res[3];
    for(int r = 0; r < 3; r++)
    {
        res[r] = m[r + 4 * 3];
        for(int c = 0; c < 3; c++) res[r] += in[c] * m[r + 4 * c];
        res[r] *= invW;
    }
    return V3(res[0], res[1], res[2]);
}

// ---------------------------------------------------------------------------------------
// hash grid, src/hashgrid.hxx
// ---------------------------------------------------------------------------------------
struct LightVertex   // PathVertex<true>, src/vertexcm.hxx:79-98
{
    V3    hitpoint;
    V3    throughput;
    uint  pathLength;
    Bsdf  bsdf;
    float dVCM, dVC, dVM;
};

struct HashGrid
{
    V3 bboxMin, bboxMax;
    std::vector<int> indices;
    std::vector<int> cellEnds;
    float radius, radiusSqr, cellSize, invCellSize;

    // GetCellIndex(Vec3i), src/hashgrid.hxx:179-187
    int CellIndex(int cx, int cy, int cz) const
    {
        const uint x = uint(cx), y = uint(cy), z = uint(cz);
        return int(((x * 73856093) ^ (y * 19349663) ^ (z * 83492791)) % uint(cellEnds.size()));
    }
    // GetCellIndex(Vec3f), src/hashgrid.hxx:189-202
    int CellIndex(const V3 &p) const
    {
        const V3 distMin = p - bboxMin;
        const V3 coordF(std::floor(invCellSize * distMin.x),
                        std::floor(invCellSize * distMin.y),
                        std::floor(invCellSize * distMin.z));
        return CellIndex(int(coordF.x), int(coordF.y), int(coordF.z));
    }
    // GetCellRange, src/hashgrid.hxx:173-177
    void CellRange(int cell, int &begin, int &end) const
    {
        if(cell == 0) { begin = 0; end = cellEnds[0]; return; }
        begin = cellEnds[cell - 1]; end = cellEnds[cell];
    }

    // Reserve + Build, src/hashgrid.hxx:35-107
    void Build(const std::vector<LightVertex> &particles, float aRadius, int numCells)
    {
        cellEnds.resize(numCells);
        radius      = aRadius;
        radiusSqr   = Sqr(radius);
        cellSize    = radius * 2.f;
        invCellSize = 1.f / cellSize;

        bboxMin = V3( 1e36f);
        bboxMax = V3(-1e36f);
        for(size_t i = 0; i < particles.size(); i++)
        {
            const V3 &pos = particles[i].hitpoint;
            bboxMax.x = std::max(bboxMax.x, pos.x); bboxMin.x = std::min(bboxMin.x, pos.x);
            bboxMax.y = std::max(bboxMax.y, pos.y); bboxMin.y = std::min(bboxMin.y, pos.y);
            bboxMax.z = std::max(bboxMax.z, pos.z); bboxMin.z = std::min(bboxMin.z, pos.z);
        }

        indices.resize(particles.size());
        std::fill(cellEnds.begin(), cellEnds.end(), 0);

        for(size_t i = 0; i < particles.size(); i++)
            cellEnds[CellIndex(particles[i].hitpoint)]++;

        int sum = 0;
        for(size_t i = 0; i < cellEnds.size(); i++)
        {
            const int temp = cellEnds[i];
            cellEnds[i] = sum;
            sum += temp;
        }

        for(size_t i = 0; i < particles.size(); i++)
        {
            const int targetIdx = cellEnds[CellIndex(particles[i].hitpoint)]++;
            indices[targetIdx] = int(i);
        }
    }
};

// ---------------------------------------------------------------------------------------
// the integrator, src/vertexcm.hxx
// ---------------------------------------------------------------------------------------
struct SubPathState   // src/vertexcm.hxx:64-76
{
    V3    origin, direction, throughput;
    uint  pathLength;
    bool  isFiniteLight;
    bool  specularPath;
    float dVCM, dVC, dVM;
};

struct Counters
{
    long long lightRays, cameraRays, shadowRays;
    long long mergeQueries, mergeCandidates, mergeAccepted;
    long long connections;
};

struct Oracle
{
    Scene scene;
    Rng   rng;
    bool  useVM, useVC, lightTraceOnly, ppm;
    float radiusAlpha, baseRadius;
    float misVmWeightFactor, misVcWeightFactor;
    float screenPixelCount, lightSubPathCount, vmNormalization;
    uint  maxPathLength, minPathLength;
    int   iterations;
    int   resX, resY;
    std::vector<V3>          framebuffer;     // mFramebuffer.mColor
    std::vector<LightVertex> lightVertices;
    std::vector<int>         pathEnds;
    HashGrid                 grid;
    float                    lastRadius;
    Counters                 cnt;
    // Path-parallel iterations (RunIterationParallel, per-path random streams only): a worker is a
    // copy of the renderer that traces a contiguous block of path indices into PRIVATE light
    // vertices / framebuffer / counters; in the camera phase it reads the master's complete light
    // vertices, path ends and grid through these views. A serial renderer views itself.
    const std::vector<LightVertex> *lvView;
    const std::vector<int>         *peView;
    const HashGrid                 *gridView;
    int                             pathBase;     // first path index of a worker's block (pathEnds is block-local)
    std::vector<Oracle*>            workers;

    Oracle(int seed, int rngMode) : rng(seed, rngMode), lvView(&lightVertices), peView(&pathEnds),
        gridView(&grid), pathBase(0) {}
    ~Oracle() { for(size_t i = 0; i < workers.size(); i++) delete workers[i]; }

    float Mis(float pdf) const { return pdf; }   // balance heuristic, src/vertexcm.hxx:553-557

    const SvcmLight &LightPtr(int idx) const     // Scene::GetLightPtr, src/scene.hxx:98-102
    {
        idx = std::min<int>(idx, int(scene.lights.size()) - 1);
        return scene.lights[idx];
    }

    // Framebuffer::AddColor, src/framebuffer.hxx:43-57
    void AddColor(float sx, float sy, const V3 &color)
    {
        if(sx < 0 || sx >= scene.camera.resolution[0]) return;
        if(sy < 0 || sy >= scene.camera.resolution[1]) return;
        const int x = int(sx), y = int(sy);
        framebuffer[x + y * resX] = framebuffer[x + y * resX] + color;
    }

    // RangeQuery::Process, src/vertexcm.hxx:130-169
    void MergeOne(const LightVertex &lv, const Bsdf &cameraBsdf, const SubPathState &cameraState,
                  V3 &contrib) const
    {
        if((lv.pathLength + cameraState.pathLength > maxPathLength) ||
           (lv.pathLength + cameraState.pathLength < minPathLength))
            return;

        const V3 lightDirection = lv.bsdf.WorldDirFix();
        float cosCamera, cameraBsdfDirPdfW, cameraBsdfRevPdfW;
        const V3 cameraBsdfFactor = BsdfEvaluate(scene, cameraBsdf, lightDirection, cosCamera,
                                                 &cameraBsdfDirPdfW, &cameraBsdfRevPdfW);
        if(IsZero(cameraBsdfFactor)) return;

        cameraBsdfDirPdfW *= cameraBsdf.contProb;
        cameraBsdfRevPdfW *= lv.bsdf.contProb;

        const float wLight  = lv.dVCM * misVcWeightFactor + lv.dVM * Mis(cameraBsdfDirPdfW);
        const float wCamera = cameraState.dVCM * misVcWeightFactor +
                              cameraState.dVM * Mis(cameraBsdfRevPdfW);
        const float misWeight = ppm ? 1.f : 1.f / (wLight + 1.f + wCamera);

        contrib = contrib + misWeight * cameraBsdfFactor * lv.throughput;
    }

    // HashGrid::Process, src/hashgrid.hxx:110-169, with the RangeQuery inlined
    // oQueries counts the queries that pass the bbox test (the N_q of SURVEY §8d)
    V3 RangeQuery(const V3 &queryPos, const Bsdf &cameraBsdf, const SubPathState &cameraState,
                  long long *oQueries, long long *oCandidates, long long *oAccepted) const
    {
        V3 contrib(0);
        const HashGrid &grid = *gridView;
        const std::vector<LightVertex> &lightVertices = *lvView;
        const V3 distMin = queryPos - grid.bboxMin;
        const V3 distMax = grid.bboxMax - queryPos;
        if(distMin.x < 0.f || distMax.x < 0.f) return contrib;
        if(distMin.y < 0.f || distMax.y < 0.f) return contrib;
        if(distMin.z < 0.f || distMax.z < 0.f) return contrib;
        if(oQueries) (*oQueries)++;

        const V3 cellPt = grid.invCellSize * distMin;
        const V3 coordF(std::floor(cellPt.x), std::floor(cellPt.y), std::floor(cellPt.z));
        const int px = int(coordF.x), py = int(coordF.y), pz = int(coordF.z);
        const V3 fractCoord = cellPt - coordF;
        const int pxo = px + (fractCoord.x < 0.5f ? -1 : +1);
        const int pyo = py + (fractCoord.y < 0.5f ? -1 : +1);
        const int pzo = pz + (fractCoord.z < 0.5f ? -1 : +1);

        for(int j = 0; j < 8; j++)
        {
            const int cx = (j & 4) ? pxo : px;
            const int cy = (j & 2) ? pyo : py;
            const int cz = (j & 1) ? pzo : pz;
            int begin, end;
            grid.CellRange(grid.CellIndex(cx, cy, cz), begin, end);
            for(; begin < end; begin++)
            {
                const LightVertex &lv = lightVertices[grid.indices[begin]];
                const float distSqr = LenSqr(queryPos - lv.hitpoint);
                if(oCandidates) (*oCandidates)++;
                if(distSqr <= grid.radiusSqr)
                {
                    if(oAccepted) (*oAccepted)++;
                    MergeOne(lv, cameraBsdf, cameraState, contrib);
                }
            }
        }
        return contrib;
    }

    // GenerateLightSample, src/vertexcm.hxx:816-858
    void GenerateLightSample(SubPathState &st)
    {
        const int   lightCount    = int(scene.lights.size());
        const float lightPickProb = 1.f / lightCount;

        const int   lightID = int(rng.GetFloat() * lightCount);
        const float dirX = rng.GetFloat(), dirY = rng.GetFloat();
        const float posX = rng.GetFloat(), posY = rng.GetFloat();

        const SvcmLight &light = LightPtr(lightID);

        float emissionPdfW, directPdfA, cosLight;
        st.throughput = LightEmit(light, scene.sphere, dirX, dirY, posX, posY,
                                  st.origin, st.direction, emissionPdfW, directPdfA, cosLight);

        emissionPdfW *= lightPickProb;
        directPdfA   *= lightPickProb;

        st.throughput    = st.throughput / emissionPdfW;
        st.pathLength    = 1;
        st.isFiniteLight = LightIsFinite(light);
        st.specularPath  = false; // never read on light paths (src/vertexcm.hxx:836-838)

        st.dVCM = Mis(directPdfA / emissionPdfW);
        if(!LightIsDelta(light))
        {
            const float usedCosLight = LightIsFinite(light) ? cosLight : 1.f;
            st.dVC = Mis(usedCosLight / emissionPdfW);
        }
        else
        {
            st.dVC = 0.f;
        }
        st.dVM = st.dVC * misVcWeightFactor;
    }

    // GenerateCameraSample, src/vertexcm.hxx:564-606 + Camera::GenerateRay src/camera.hxx:108-117
    void GenerateCameraSample(int pixelIndex, SubPathState &st, float &sampleX, float &sampleY)
    {
        const SvcmCamera &cam = scene.camera;
        const int x = pixelIndex % resX;
        const int y = pixelIndex / resX;
        const float jx = rng.GetFloat(), jy = rng.GetFloat();
        sampleX = float(x) + jx;
        sampleY = float(y) + jy;

        const V3 camPos = Load3(cam.position), camFwd = Load3(cam.forward);
        const V3 worldRaster = TransformPoint(cam.rasterToWorld, V3(sampleX, sampleY, 0));
        const V3 dir = Normalize(worldRaster - camPos);

        const float cosAtCamera = Dot(camFwd, dir);
        const float imagePointToCameraDist = cam.imagePlaneDist / cosAtCamera;
        const float imageToSolidAngleFactor = Sqr(imagePointToCameraDist) / cosAtCamera;
        const float cameraPdfW = imageToSolidAngleFactor;

        st.origin       = camPos;
        st.direction    = dir;
        st.throughput   = V3(1);
        st.pathLength   = 1;
        st.specularPath = true;
        st.isFiniteLight = false;
        st.dVCM = Mis(lightSubPathCount / cameraPdfW);
        st.dVC  = 0;
        st.dVM  = 0;
    }

    // GetLightRadiance, src/vertexcm.hxx:617-658
    V3 GetLightRadiance(const SvcmLight &light, const SubPathState &st, const V3 &rayDir) const
    {
        const int   lightCount    = int(scene.lights.size());
        const float lightPickProb = 1.f / lightCount;

        float directPdfA = 0, emissionPdfW = 0;
        const V3 radiance = LightGetRadiance(light, scene.sphere, rayDir, directPdfA, emissionPdfW);
        if(IsZero(radiance)) return V3(0);
        if(st.pathLength == 1) return radiance;
        if(useVM && !useVC) return st.specularPath ? radiance : V3(0);

        directPdfA   *= lightPickProb;
        emissionPdfW *= lightPickProb;
        const float wCamera = Mis(directPdfA) * st.dVCM + Mis(emissionPdfW) * st.dVC;
        const float misWeight = 1.f / (1.f + wCamera);
        return misWeight * radiance;
    }

    // DirectIllumination, src/vertexcm.hxx:663-738
    V3 DirectIllumination(const SubPathState &st, const V3 &hitpoint, const Bsdf &bsdf)
    {
        const int   lightCount    = int(scene.lights.size());
        const float lightPickProb = 1.f / lightCount;

        const int   lightID = int(rng.GetFloat() * lightCount);
        const float rx = rng.GetFloat(), ry = rng.GetFloat();
        const SvcmLight &light = LightPtr(lightID);

        V3 directionToLight(0);
        float distance, directPdfW, emissionPdfW, cosAtLight;
        const V3 radiance = LightIlluminate(light, scene.sphere, hitpoint, rx, ry,
                                            directionToLight, distance, directPdfW,
                                            emissionPdfW, cosAtLight);
        if(IsZero(radiance)) return V3(0);

        float bsdfDirPdfW, bsdfRevPdfW, cosToLight;
        const V3 bsdfFactor = BsdfEvaluate(scene, bsdf, directionToLight, cosToLight,
                                           &bsdfDirPdfW, &bsdfRevPdfW);
        if(IsZero(bsdfFactor)) return V3(0);

        const float continuationProbability = bsdf.contProb;
        bsdfDirPdfW *= LightIsDelta(light) ? 0.f : continuationProbability;
        bsdfRevPdfW *= continuationProbability;

        const float wLight = Mis(bsdfDirPdfW / (lightPickProb * directPdfW));
        const float wCamera = Mis(emissionPdfW * cosToLight / (directPdfW * cosAtLight)) * (
            misVmWeightFactor + st.dVCM + st.dVC * Mis(bsdfRevPdfW));
        const float misWeight = 1.f / (wLight + 1.f + wCamera);

        const V3 contrib = (misWeight * cosToLight / (lightPickProb * directPdfW)) *
            (radiance * bsdfFactor);

        if(IsZero(contrib)) return V3(0);
        cnt.shadowRays++;
        if(SceneOccluded(scene, hitpoint, directionToLight, distance)) return V3(0);
        return contrib;
    }

    // ConnectVertices, src/vertexcm.hxx:743-809
    V3 ConnectVertices(const LightVertex &lv, const Bsdf &cameraBsdf, const V3 &cameraHitpoint,
                       const SubPathState &st)
    {
        V3 direction = lv.hitpoint - cameraHitpoint;
        const float dist2 = LenSqr(direction);
        float distance = std::sqrt(dist2);
        direction = direction / distance;

        float cosCamera, cameraBsdfDirPdfW, cameraBsdfRevPdfW;
        const V3 cameraBsdfFactor = BsdfEvaluate(scene, cameraBsdf, direction, cosCamera,
                                                 &cameraBsdfDirPdfW, &cameraBsdfRevPdfW);
        if(IsZero(cameraBsdfFactor)) return V3(0);

        const float cameraCont = cameraBsdf.contProb;
        cameraBsdfDirPdfW *= cameraCont;
        cameraBsdfRevPdfW *= cameraCont;

        float cosLight, lightBsdfDirPdfW, lightBsdfRevPdfW;
        const V3 lightBsdfFactor = BsdfEvaluate(scene, lv.bsdf, -direction, cosLight,
                                                &lightBsdfDirPdfW, &lightBsdfRevPdfW);
        if(IsZero(lightBsdfFactor)) return V3(0);

        const float lightCont = lv.bsdf.contProb;
        lightBsdfDirPdfW *= lightCont;
        lightBsdfRevPdfW *= lightCont;

        const float geometryTerm = cosLight * cosCamera / dist2;
        if(geometryTerm < 0) return V3(0);

        const float cameraBsdfDirPdfA = PdfWtoA(cameraBsdfDirPdfW, distance, cosLight);
        const float lightBsdfDirPdfA  = PdfWtoA(lightBsdfDirPdfW,  distance, cosCamera);

        const float wLight = Mis(cameraBsdfDirPdfA) * (
            misVmWeightFactor + lv.dVCM + lv.dVC * Mis(lightBsdfRevPdfW));
        const float wCamera = Mis(lightBsdfDirPdfA) * (
            misVmWeightFactor + st.dVCM + st.dVC * Mis(cameraBsdfRevPdfW));
        const float misWeight = 1.f / (wLight + 1.f + wCamera);

        const V3 contrib = (misWeight * geometryTerm) * cameraBsdfFactor * lightBsdfFactor;

        if(IsZero(contrib)) return V3(0);
        cnt.shadowRays++;
        if(SceneOccluded(scene, cameraHitpoint, direction, distance)) return V3(0);
        return contrib;
    }

    // ConnectToCamera, src/vertexcm.hxx:862-933
    void ConnectToCamera(const SubPathState &st, const V3 &hitpoint, const Bsdf &bsdf)
    {
        const SvcmCamera &cam = scene.camera;
        const V3 camPos = Load3(cam.position), camFwd = Load3(cam.forward);
        V3 directionToCamera = camPos - hitpoint;

        if(Dot(camFwd, -directionToCamera) <= 0.f) return;

        // Camera::WorldToRaster / CheckRaster, src/camera.hxx:95-105
        const V3 raster = TransformPoint(cam.worldToRaster, hitpoint);
        if(!(raster.x >= 0 && raster.y >= 0 &&
             raster.x < cam.resolution[0] && raster.y < cam.resolution[1]))
            return;

        const float distEye2 = LenSqr(directionToCamera);
        const float distance = std::sqrt(distEye2);
        directionToCamera = directionToCamera / distance;

        float cosToCamera, bsdfDirPdfW, bsdfRevPdfW;
        const V3 bsdfFactor = BsdfEvaluate(scene, bsdf, directionToCamera, cosToCamera,
                                           &bsdfDirPdfW, &bsdfRevPdfW);
        if(IsZero(bsdfFactor)) return;

        bsdfRevPdfW *= bsdf.contProb;

        const float cosAtCamera = Dot(camFwd, -directionToCamera);
        const float imagePointToCameraDist = cam.imagePlaneDist / cosAtCamera;
        const float imageToSolidAngleFactor = Sqr(imagePointToCameraDist) / cosAtCamera;
        const float imageToSurfaceFactor = imageToSolidAngleFactor * std::abs(cosToCamera) / Sqr(distance);
        const float cameraPdfA = imageToSurfaceFactor;

        const float wLight = Mis(cameraPdfA / lightSubPathCount) * (
            misVmWeightFactor + st.dVCM + st.dVC * Mis(bsdfRevPdfW));
        const float misWeight = lightTraceOnly ? 1.f : (1.f / (wLight + 1.f));
        const float surfaceToImageFactor = 1.f / imageToSurfaceFactor;

        const V3 contrib = misWeight * st.throughput * bsdfFactor /
            (lightSubPathCount * surfaceToImageFactor);

        if(!IsZero(contrib))
        {
            cnt.shadowRays++;
            if(SceneOccluded(scene, hitpoint, directionToCamera, distance)) return;
            AddColor(raster.x, raster.y, contrib);
        }
    }

    // SampleScattering, src/vertexcm.hxx:938-1006
    bool SampleScattering(const Bsdf &bsdf, const V3 &hitPoint, SubPathState &st)
    {
        const float r0 = rng.GetFloat(), r1 = rng.GetFloat(), r2 = rng.GetFloat();
        float bsdfDirPdfW, cosThetaOut;
        uint sampledEvent;
        const V3 bsdfFactor = BsdfSample(scene, bsdf, V3(r0, r1, r2), st.direction,
                                         bsdfDirPdfW, cosThetaOut, &sampledEvent);
        if(IsZero(bsdfFactor)) return false;

        float bsdfRevPdfW = bsdfDirPdfW;
        if((sampledEvent & kSpecular) == 0)
            bsdfRevPdfW = BsdfPdf(scene, bsdf, st.direction, true);

        const float contProb = bsdf.contProb;
        if(rng.GetFloat() > contProb) return false;

        bsdfDirPdfW *= contProb;
        bsdfRevPdfW *= contProb;

        if(sampledEvent & kSpecular)
        {
            st.dVCM = 0.f;
            st.dVC *= Mis(cosThetaOut);
            st.dVM *= Mis(cosThetaOut);
            st.specularPath = st.specularPath && true;
        }
        else
        {
            st.dVC = Mis(cosThetaOut / bsdfDirPdfW) * (
                st.dVC * Mis(bsdfRevPdfW) + st.dVCM + misVmWeightFactor);
            st.dVM = Mis(cosThetaOut / bsdfDirPdfW) * (
                st.dVM * Mis(bsdfRevPdfW) + st.dVCM * misVcWeightFactor + 1.f);
            st.dVCM = Mis(1.f / bsdfDirPdfW);
            st.specularPath = false;
        }

        st.origin = hitPoint;
        st.throughput = st.throughput * (bsdfFactor * (cosThetaOut / bsdfDirPdfW));
        return true;
    }

    // the per-iteration constants, src/vertexcm.hxx:288-308
    float SetupIteration(int aIteration)
    {
        const int pathCount = resX * resY;
        screenPixelCount  = float(pathCount);
        lightSubPathCount = float(pathCount);

        float radius = baseRadius;
        radius /= std::pow(float(aIteration + 1), 0.5f * (1 - radiusAlpha));
        radius = std::max(radius, 1e-7f);
        const float radiusSqr = Sqr(radius);

        vmNormalization = 1.f / (radiusSqr * kPi * lightSubPathCount);
        const float etaVCM = (kPi * radiusSqr) * lightSubPathCount;
        misVmWeightFactor = useVM ? Mis(etaVCM)       : 0.f;
        misVcWeightFactor = useVC ? Mis(1.f / etaVCM) : 0.f;
        lastRadius = radius;
        return radius;
    }

    // one light sub-path, src/vertexcm.hxx:321-396 (loop body)
    void TraceLightPath(int aIteration, int pathIdx)
    {
        rng.SetPath(aIteration, 0, pathIdx);
        SubPathState st;
        GenerateLightSample(st);

        for(;; ++st.pathLength)
        {
            Ray ray; ray.org = st.origin + st.direction * kEpsRay; ray.dir = st.direction; ray.tmin = 0;
            Isect isect; isect.dist = 1e36f;
            cnt.lightRays++;
            if(!SceneIntersect(scene, ray, isect)) break;

            const V3 hitPoint = ray.org + ray.dir * isect.dist;
            isect.dist += kEpsRay;

            Bsdf bsdf;
            BsdfSetup(scene, ray.dir, isect, true, bsdf);
            if(bsdf.matID < 0) break;

            if(st.pathLength > 1 || st.isFiniteLight)
                st.dVCM *= Mis(Sqr(isect.dist));
            st.dVCM /= Mis(std::abs(bsdf.CosThetaFix()));
            st.dVC  /= Mis(std::abs(bsdf.CosThetaFix()));
            st.dVM  /= Mis(std::abs(bsdf.CosThetaFix()));

            if(!bsdf.isDelta && (useVC || useVM))
            {
                LightVertex lv;
                lv.hitpoint   = hitPoint;
                lv.throughput = st.throughput;
                lv.pathLength = st.pathLength;
                lv.bsdf       = bsdf;
                lv.dVCM = st.dVCM; lv.dVC = st.dVC; lv.dVM = st.dVM;
                lightVertices.push_back(lv);
            }

            if(!bsdf.isDelta && (useVC || lightTraceOnly))
            {
                if(st.pathLength + 1 >= minPathLength)
                    ConnectToCamera(st, hitPoint, bsdf);
            }

            if(st.pathLength + 2 > maxPathLength) break;
            if(!SampleScattering(bsdf, hitPoint, st)) break;
        }
        pathEnds[pathIdx - pathBase] = int(lightVertices.size());
    }

    // one camera sub-path, src/vertexcm.hxx:415-545 (loop body)
    void TraceCameraPath(int aIteration, int pathIdx)
    {
        rng.SetPath(aIteration, 1, pathIdx);
        SubPathState st;
        float sampleX, sampleY;
        GenerateCameraSample(pathIdx, st, sampleX, sampleY);
        V3 color(0);

        for(;; ++st.pathLength)
        {
            Ray ray; ray.org = st.origin + st.direction * kEpsRay; ray.dir = st.direction; ray.tmin = 0;
            Isect isect; isect.dist = 1e36f;
            cnt.cameraRays++;

            if(!SceneIntersect(scene, ray, isect))
            {
                if(scene.background >= 0)
                {
                    if(st.pathLength >= minPathLength)
                        color = color + st.throughput *
                            GetLightRadiance(scene.lights[scene.background], st, ray.dir);
                }
                break;
            }

            const V3 hitPoint = ray.org + ray.dir * isect.dist;
            isect.dist += kEpsRay;

            Bsdf bsdf;
            BsdfSetup(scene, ray.dir, isect, false, bsdf);
            if(bsdf.matID < 0) break;

            st.dVCM *= Mis(Sqr(isect.dist));
            st.dVCM /= Mis(std::abs(bsdf.CosThetaFix()));
            st.dVC  /= Mis(std::abs(bsdf.CosThetaFix()));
            st.dVM  /= Mis(std::abs(bsdf.CosThetaFix()));

            if(isect.lightID >= 0)
            {
                const SvcmLight &light = LightPtr(isect.lightID);
                if(st.pathLength >= minPathLength)
                    color = color + st.throughput * GetLightRadiance(light, st, ray.dir);
                break;
            }

            if(st.pathLength >= maxPathLength) break;

            if(!bsdf.isDelta && useVC)
            {
                if(st.pathLength + 1 >= minPathLength)
                    color = color + st.throughput * DirectIllumination(st, hitPoint, bsdf);
            }

            if(!bsdf.isDelta && useVC)
            {
                const std::vector<int> &pathEnds = *peView;
                const std::vector<LightVertex> &lightVertices = *lvView;
                const int begin = (pathIdx == 0) ? 0 : pathEnds[pathIdx - 1];
                const int end   = pathEnds[pathIdx];
                for(int i = begin; i < end; i++)
                {
                    const LightVertex &lv = lightVertices[i];
                    if(lv.pathLength + 1 + st.pathLength < minPathLength) continue;
                    if(lv.pathLength + 1 + st.pathLength > maxPathLength) break;
                    cnt.connections++;
                    color = color + st.throughput * lv.throughput *
                        ConnectVertices(lv, bsdf, hitPoint, st);
                }
            }

            if(!bsdf.isDelta && useVM)
            {
                const V3 contrib = RangeQuery(hitPoint, bsdf, st, &cnt.mergeQueries,
                                              &cnt.mergeCandidates, &cnt.mergeAccepted);
                color = color + st.throughput * vmNormalization * contrib;
                if(ppm) break;
            }

            if(!SampleScattering(bsdf, hitPoint, st)) break;
        }

        AddColor(sampleX, sampleY, color);
    }

    // ---------------------------------------------------------------------------------
    // PathTracer::RunIteration, src/pathtracer.hxx:45-215 (next-event estimation + BSDF
    // sampling combined with the balance heuristic) — SURVEY §8(f) rank 1
    // ---------------------------------------------------------------------------------
    static float Mis2(float samplePdf, float otherPdf) { return samplePdf / (samplePdf + otherPdf); }  // :228-233

    void TracePathTracerPixel(int aIteration, int pixID)
    {
        rng.SetPath(aIteration, 1, pixID);
        const int   lightCount    = int(scene.lights.size());
        const float lightPickProb = 1.f / lightCount;
        const SvcmCamera &cam = scene.camera;

        const int x = pixID % resX;
        const int y = pixID / resX;
        const float jx = rng.GetFloat(), jy = rng.GetFloat();
        const float sampleX = float(x) + jx, sampleY = float(y) + jy;

        // Camera::GenerateRay, src/camera.hxx:108-117 (no origin offset on the first segment)
        const V3 camPos = Load3(cam.position);
        Ray ray;
        ray.org = camPos;
        ray.dir = Normalize(TransformPoint(cam.rasterToWorld, V3(sampleX, sampleY, 0)) - camPos);
        ray.tmin = 0;
        Isect isect; isect.dist = 1e36f;

        V3 pathWeight(1.f);
        V3 color(0.f);
        uint  pathLength   = 1;
        bool  lastSpecular = true;
        float lastPdfW     = 1;

        for(;; ++pathLength)
        {
            cnt.cameraRays++;
            if(!SceneIntersect(scene, ray, isect))
            {
                if(pathLength < minPathLength) break;
                if(scene.background < 0) break;
                float directPdfW = 0, emissionPdfW = 0;
                const V3 contrib = LightGetRadiance(scene.lights[scene.background], scene.sphere, ray.dir,
                                                    directPdfW, emissionPdfW);
                if(IsZero(contrib)) break;
                float misWeight = 1.f;
                if(pathLength > 1 && !lastSpecular)
                    misWeight = Mis2(lastPdfW, directPdfW * lightPickProb);
                color = color + pathWeight * misWeight * contrib;
                break;
            }

            const V3 hitPoint = ray.org + ray.dir * isect.dist;
            isect.dist += kEpsRay;

            Bsdf bsdf;
            BsdfSetup(scene, ray.dir, isect, false, bsdf);
            if(bsdf.matID < 0) break;

            if(isect.lightID >= 0)
            {
                if(pathLength < minPathLength) break;
                const SvcmLight &light = LightPtr(isect.lightID);
                float directPdfA = 0, emissionPdfW = 0;
                const V3 contrib = LightGetRadiance(light, scene.sphere, ray.dir, directPdfA, emissionPdfW);
                if(IsZero(contrib)) break;
                float misWeight = 1.f;
                if(pathLength > 1 && !lastSpecular)
                {
                    // PdfAtoW, src/utils.hxx:253-259
                    const float directPdfW = directPdfA * Sqr(isect.dist) / std::abs(bsdf.CosThetaFix());
                    misWeight = Mis2(lastPdfW, directPdfW * lightPickProb);
                }
                color = color + pathWeight * misWeight * contrib;
                break;
            }

            if(pathLength >= maxPathLength) break;
            if(bsdf.contProb == 0) break;

            // next event estimation
            if(!bsdf.isDelta && pathLength + 1 >= minPathLength)
            {
                const int lightID = int(rng.GetFloat() * lightCount);
                const SvcmLight &light = LightPtr(lightID);
                const float rx = rng.GetFloat(), ry = rng.GetFloat();
                V3 directionToLight(0);
                float distance, directPdfW, emissionPdfW, cosAtLight;
                const V3 radiance = LightIlluminate(light, scene.sphere, hitPoint, rx, ry, directionToLight,
                                                    distance, directPdfW, emissionPdfW, cosAtLight);
                if(!IsZero(radiance))
                {
                    float bsdfPdfW, cosThetaOut;
                    const V3 factor = BsdfEvaluate(scene, bsdf, directionToLight, cosThetaOut, &bsdfPdfW, NULL);
                    if(!IsZero(factor))
                    {
                        float weight = 1.f;
                        if(!LightIsDelta(light))
                        {
                            const float contProb = bsdf.contProb;
                            bsdfPdfW *= contProb;
                            weight = Mis2(directPdfW * lightPickProb, bsdfPdfW);
                        }
                        const V3 contrib = (weight * cosThetaOut / (lightPickProb * directPdfW)) * (radiance * factor);
                        cnt.shadowRays++;
                        if(!SceneOccluded(scene, hitPoint, directionToLight, distance))
                            color = color + pathWeight * contrib;
                    }
                }
            }

            // continue random walk
            {
                const float r0 = rng.GetFloat(), r1 = rng.GetFloat(), r2 = rng.GetFloat();
                float pdf, cosThetaOut;
                uint  sampledEvent;
                const V3 factor = BsdfSample(scene, bsdf, V3(r0, r1, r2), ray.dir, pdf, cosThetaOut, &sampledEvent);
                if(IsZero(factor)) break;

                const float contProb = bsdf.contProb;
                lastSpecular = (sampledEvent & kSpecular) != 0;
                lastPdfW     = pdf * contProb;
                if(contProb < 1.f)
                {
                    if(rng.GetFloat() > contProb) break;
                    pdf *= contProb;
                }
                pathWeight = pathWeight * (factor * (cosThetaOut / pdf));
                ray.org    = hitPoint + kEpsRay * ray.dir;
                ray.tmin   = 0.f;
                isect.dist = 1e36f;
            }
        }
        AddColor(sampleX, sampleY, color);
    }

    // EyeLight::RunIteration, src/eyelight.hxx:47-78 (N.L preview; iteration 1 is unjittered)
    void TraceEyeLightPixel(int aIteration, int pixID)
    {
        rng.SetPath(aIteration, 1, pixID);
        const SvcmCamera &cam = scene.camera;
        const int x = pixID % resX;
        const int y = pixID / resX;
        float jx = 0.5f, jy = 0.5f;
        if(aIteration != 1) { jx = rng.GetFloat(); jy = rng.GetFloat(); }
        const float sampleX = float(x) + jx, sampleY = float(y) + jy;
        const V3 camPos = Load3(cam.position);
        Ray ray;
        ray.org = camPos;
        ray.dir = Normalize(TransformPoint(cam.rasterToWorld, V3(sampleX, sampleY, 0)) - camPos);
        ray.tmin = 0;
        Isect isect; isect.dist = 1e36f;
        cnt.cameraRays++;
        if(SceneIntersect(scene, ray, isect))
        {
            const float dotLN = Dot(isect.normal, -ray.dir);
            if(dotLN > 0) AddColor(sampleX, sampleY, V3(dotLN));
            else          AddColor(sampleX, sampleY, V3(-dotLN, 0, 0));
        }
    }

    int algorithm;   // SVCM_ALG_*

    // VertexCM::RunIteration, src/vertexcm.hxx:284-548
    void RunIteration(int aIteration)
    {
        if(algorithm == SVCM_ALG_PATH_TRACE || algorithm == SVCM_ALG_EYE_LIGHT)
        {
            memset(&cnt, 0, sizeof(cnt));
            lightVertices.clear();
            for(int pixID = 0; pixID < resX * resY; pixID++)
            {
                if(algorithm == SVCM_ALG_PATH_TRACE) TracePathTracerPixel(aIteration, pixID);
                else                                 TraceEyeLightPixel(aIteration, pixID);
            }
            iterations++;
            return;
        }

        const int pathCount = resX * resY;
        const float radius = SetupIteration(aIteration);
        memset(&cnt, 0, sizeof(cnt));

        pathEnds.assign(pathCount, 0);
        lightVertices.reserve(pathCount);
        lightVertices.clear();

        for(int pathIdx = 0; pathIdx < pathCount; pathIdx++)
            TraceLightPath(aIteration, pathIdx);

        if(useVM)
            grid.Build(lightVertices, radius, pathCount);

        for(int pathIdx = 0; (pathIdx < pathCount) && (!lightTraceOnly); ++pathIdx)
            TraceCameraPath(aIteration, pathIdx);

        iterations++;
    }

    // The same iteration with the path loops of src/vertexcm.hxx:321-396 and :415-545 split into
    // `threads` contiguous blocks. Only meaningful with per-path random streams (rng mode 1), where
    // a path's result does not depend on which paths ran before it. Light vertices and path ends
    // come out in path order exactly as in RunIteration; framebuffer contributions are summed
    // per block and the blocks added in order, so the image differs from RunIteration's only in
    // the order of float additions.
    void RunIterationParallel(int aIteration, int threads)
    {
        const int pathCount = resX * resY;
        threads = std::max(1, std::min(threads, std::max(1, pathCount / 64)));
        if(threads == 1 || rng.mMode != 1) { RunIteration(aIteration); return; }
        const bool walkOnly = algorithm == SVCM_ALG_PATH_TRACE || algorithm == SVCM_ALG_EYE_LIGHT;
        const float radius = walkOnly ? 0.f : SetupIteration(aIteration);
        memset(&cnt, 0, sizeof(cnt));
        while(int(workers.size()) < threads)
        {
            Oracle *w = new Oracle(int(rng.mSeed), rng.mMode);
            w->scene = scene;
            workers.push_back(w);
        }
        for(int t = 0; t < threads; t++)
        {
            Oracle *w = workers[t];
            w->useVM = useVM; w->useVC = useVC; w->lightTraceOnly = lightTraceOnly; w->ppm = ppm;
            w->algorithm = algorithm;
            w->radiusAlpha = radiusAlpha; w->baseRadius = baseRadius;
            w->misVmWeightFactor = misVmWeightFactor; w->misVcWeightFactor = misVcWeightFactor;
            w->screenPixelCount = screenPixelCount; w->lightSubPathCount = lightSubPathCount;
            w->vmNormalization = vmNormalization;
            w->maxPathLength = maxPathLength; w->minPathLength = minPathLength;
            w->resX = resX; w->resY = resY; w->lastRadius = lastRadius;
            w->framebuffer.assign(framebuffer.size(), V3(0));
            memset(&w->cnt, 0, sizeof(w->cnt));
            w->pathBase = int((long long)pathCount * t / threads);
            w->lightVertices.clear();
            w->pathEnds.assign(size_t((long long)pathCount * (t + 1) / threads - w->pathBase), 0);
            w->lvView = &lightVertices; w->peView = &pathEnds; w->gridView = &grid;
        }
        omp_set_num_threads(threads);
        if(!walkOnly)
        {
#pragma omp parallel for schedule(static, 1)
            for(int t = 0; t < threads; t++)
            {
                Oracle *w = workers[t];
                const int lo = w->pathBase, hi = lo + int(w->pathEnds.size());
                for(int pathIdx = lo; pathIdx < hi; pathIdx++) w->TraceLightPath(aIteration, pathIdx);
            }
            pathEnds.assign(pathCount, 0);
            lightVertices.clear();
            for(int t = 0; t < threads; t++)
            {
                Oracle *w = workers[t];
                const int offset = int(lightVertices.size());
                lightVertices.insert(lightVertices.end(), w->lightVertices.begin(), w->lightVertices.end());
                for(size_t k = 0; k < w->pathEnds.size(); k++) pathEnds[size_t(w->pathBase) + k] = offset + w->pathEnds[k];
                w->lightVertices.clear(); w->lightVertices.shrink_to_fit();
            }
            if(useVM) grid.Build(lightVertices, radius, pathCount);
        }
        else lightVertices.clear();
#pragma omp parallel for schedule(static, 1)
        for(int t = 0; t < threads; t++)
        {
            Oracle *w = workers[t];
            const int lo = w->pathBase, hi = lo + int(w->pathEnds.size());
            for(int pathIdx = lo; pathIdx < hi && !lightTraceOnly; pathIdx++)
            {
                if(algorithm == SVCM_ALG_PATH_TRACE)     w->TracePathTracerPixel(aIteration, pathIdx);
                else if(algorithm == SVCM_ALG_EYE_LIGHT) w->TraceEyeLightPixel(aIteration, pathIdx);
                else                                     w->TraceCameraPath(aIteration, pathIdx);
            }
        }
        for(int t = 0; t < threads; t++)
        {
            Oracle *w = workers[t];
            for(size_t k = 0; k < framebuffer.size(); k++) framebuffer[k] = framebuffer[k] + w->framebuffer[k];
            cnt.lightRays += w->cnt.lightRays; cnt.cameraRays += w->cnt.cameraRays; cnt.shadowRays += w->cnt.shadowRays;
            cnt.mergeQueries += w->cnt.mergeQueries; cnt.mergeCandidates += w->cnt.mergeCandidates;
            cnt.mergeAccepted += w->cnt.mergeAccepted; cnt.connections += w->cnt.connections;
        }
        iterations++;
    }
};

void CopyScene(const SvcmSceneDesc *d, Scene &s)
{
    s.prims.assign(d->primitives, d->primitives + d->numPrimitives);
    s.mats.assign(d->materials, d->materials + d->numMaterials);
    s.lights.assign(d->lights, d->lights + d->numLights);
    s.mat2light.assign(d->materialToLight, d->materialToLight + d->numMaterials);
    s.background = d->backgroundLight;
    s.camera = d->camera;
    s.sphere = d->sceneSphere;
}

// a camera-side BSDF from the flat query record (what BSDF<false>(ray, isect, scene) builds
// at src/vertexcm.hxx:452 from ray.dir = -dirFix and isect.normal / matID)
void QueryBsdf(const Scene &sc, const float *normal, const float *dirFix, int matID,
               bool fixIsLight, Bsdf &b)
{
    Isect isect; isect.dist = 1.f; isect.matID = matID; isect.lightID = -1; isect.normal = Load3(normal);
    BsdfSetup(sc, -Load3(dirFix), isect, fixIsLight, b);
}

} // namespace

// =========================================================================================
// C interface (mirrors oracle/ref_glue/ref_counter_rng.cpp so tests can run either)
// =========================================================================================

// VertexCM::VertexCM, src/vertexcm.hxx:208-282. rngMode 0: reference stream, 1: per-path.
ORACLE_API void *oracle_create(const SvcmSceneDesc *scene, int algorithm, float radiusFactor,
                               float radiusAlpha, int seed, int rngMode,
                               unsigned maxPathLength, unsigned minPathLength)
{
    Oracle *o = new Oracle(seed, rngMode);
    CopyScene(scene, o->scene);
    o->lightTraceOnly = o->useVC = o->useVM = o->ppm = false;
    switch(algorithm)
    {
    case SVCM_ALG_LIGHT_TRACE: o->lightTraceOnly = true; break;
    case SVCM_ALG_PPM:         o->ppm = true; o->useVM = true; break;
    case SVCM_ALG_BPM:         o->useVM = true; break;
    case SVCM_ALG_BPT:         o->useVC = true; break;
    case SVCM_ALG_VCM:         o->useVC = true; o->useVM = true; break;
    case SVCM_ALG_PATH_TRACE:  break;
    case SVCM_ALG_EYE_LIGHT:   break;
    default: delete o; return NULL;
    }
    o->algorithm = algorithm;
    if(o->ppm)
    {
        // src/vertexcm.hxx:246-278: mixed specular / non-specular materials => BPM
        for(size_t i = 0; i < o->scene.mats.size(); i++)
        {
            const SvcmMaterial &m = o->scene.mats[i];
            const bool hasNonSpecular = MaxComp(Load3(m.diffuse)) > 0 || MaxComp(Load3(m.phong)) > 0;
            const bool hasSpecular    = MaxComp(Load3(m.mirror)) > 0 || m.ior > 0;
            if(hasNonSpecular && hasSpecular) { o->ppm = false; break; }
        }
    }
    o->baseRadius  = radiusFactor * o->scene.sphere.radius;
    o->radiusAlpha = radiusAlpha;
    o->maxPathLength = maxPathLength;
    o->minPathLength = minPathLength;
    o->iterations = 0;
    o->resX = int(o->scene.camera.resolution[0]);
    o->resY = int(o->scene.camera.resolution[1]);
    o->framebuffer.assign(size_t(o->resX) * o->resY, V3(0));
    o->misVmWeightFactor = o->misVcWeightFactor = 0;
    o->vmNormalization = 0;
    o->lastRadius = 0;
    o->grid.radius = o->grid.radiusSqr = o->grid.cellSize = o->grid.invCellSize = 0;
    o->grid.bboxMin = V3(1e36f); o->grid.bboxMax = V3(-1e36f);
    memset(&o->cnt, 0, sizeof(o->cnt));
    return o;
}

ORACLE_API void oracle_destroy(void *h) { delete (Oracle*)h; }

ORACLE_API void oracle_run_iteration(void *h, int iteration) { ((Oracle*)h)->RunIteration(iteration); }

// Path-parallel iteration (per-path random streams only; falls back to the serial loop otherwise).
// threads <= 0: all hardware threads.
ORACLE_API void oracle_run_iteration_parallel(void *h, int iteration, int threads)
{
    ((Oracle*)h)->RunIterationParallel(iteration, threads > 0 ? threads : omp_get_num_procs());
}

ORACLE_API int oracle_accum(void *h, float *outRgb)
{
    Oracle *o = (Oracle*)h;
    memcpy(outRgb, o->framebuffer.data(), sizeof(float) * 3 * o->framebuffer.size());
    return o->iterations;
}

ORACLE_API long long oracle_num_light_vertices(void *h) { return (long long)((Oracle*)h)->lightVertices.size(); }

ORACLE_API void oracle_light_vertices(void *h, SvcmMergeVertex *out, float *dVC, int *matID,
                                      float *normal, int *pathEnds)
{
    Oracle *o = (Oracle*)h;
    for(size_t i = 0; i < o->lightVertices.size(); i++)
    {
        const LightVertex &lv = o->lightVertices[i];
        if(out)
        {
            Store3(out[i].position, lv.hitpoint);
            Store3(out[i].throughput, lv.throughput);
            Store3(out[i].dirFix, lv.bsdf.WorldDirFix());
            out[i].pathLength = lv.pathLength;
            out[i].dVCM = lv.dVCM; out[i].dVM = lv.dVM;
            out[i].contProb = lv.bsdf.contProb;
        }
        if(dVC)    dVC[i] = lv.dVC;
        if(matID)  matID[i] = lv.bsdf.matID;
        if(normal) Store3(normal + 3 * i, lv.bsdf.frame.mZ);
    }
    if(pathEnds) memcpy(pathEnds, o->pathEnds.data(), sizeof(int) * o->pathEnds.size());
}

ORACLE_API void oracle_iteration_constants(void *h, int iteration, float *radius,
                                           float *vmNormalization, float *misVmWeightFactor,
                                           float *misVcWeightFactor)
{
    Oracle *o = (Oracle*)h;
    *radius            = o->SetupIteration(iteration);
    *vmNormalization   = o->vmNormalization;
    *misVmWeightFactor = o->misVmWeightFactor;
    *misVcWeightFactor = o->misVcWeightFactor;
}

ORACLE_API int oracle_grid(void *h, float bbox[6], int *cellEnds, int *indices)
{
    Oracle *o = (Oracle*)h;
    if(bbox) { Store3(bbox, o->grid.bboxMin); Store3(bbox + 3, o->grid.bboxMax); }
    if(cellEnds) memcpy(cellEnds, o->grid.cellEnds.data(), sizeof(int) * o->grid.cellEnds.size());
    if(indices)  memcpy(indices,  o->grid.indices.data(),  sizeof(int) * o->grid.indices.size());
    return (int)o->grid.cellEnds.size();
}

// Replaces the handle's light vertices by an arbitrary set (merge fields only) and rebuilds
// the grid with the given radius / cell count: lets kernel-level tests and benchmarks run
// HashGrid::Build / Process on any input. Also fixes the constants Process reads.
ORACLE_API void oracle_set_light_vertices(void *h, const SvcmMergeVertex *verts, long long count,
                                          float radius, int numCells, float misVcWeightFactor,
                                          int ppm, unsigned maxPathLength, unsigned minPathLength)
{
    Oracle *o = (Oracle*)h;
    o->lightVertices.resize(size_t(count));
    for(long long i = 0; i < count; i++)
    {
        LightVertex &lv = o->lightVertices[size_t(i)];
        lv.hitpoint   = Load3(verts[i].position);
        lv.throughput = Load3(verts[i].throughput);
        lv.pathLength = verts[i].pathLength;
        lv.dVCM = verts[i].dVCM; lv.dVC = 0; lv.dVM = verts[i].dVM;
        // only WorldDirFix() and ContinuationProb() of the light BSDF are read by the merge
        // (src/vertexcm.hxx:138,153): an identity frame makes WorldDirFix() return dirFix
        lv.bsdf.matID = 0;
        lv.bsdf.frame.mX = V3(1, 0, 0); lv.bsdf.frame.mY = V3(0, 1, 0); lv.bsdf.frame.mZ = V3(0, 0, 1);
        lv.bsdf.localDirFix = Load3(verts[i].dirFix);
        lv.bsdf.contProb = verts[i].contProb;
        lv.bsdf.isDelta = false;
        lv.bsdf.fixIsLight = true;
    }
    o->grid.Build(o->lightVertices, radius, numCells);
    o->misVcWeightFactor = misVcWeightFactor;
    o->ppm = ppm != 0;
    o->maxPathLength = maxPathLength;
    o->minPathLength = minPathLength;
}

// HashGrid::Process + RangeQuery::Process for caller-made camera vertices; counts (optional)
// = 2 ints per query: candidates, accepted.
ORACLE_API void oracle_merge_queries(void *h, const SvcmMergeQuery *queries, long long count,
                                     float *contrib, int *counts)
{
    Oracle *o = (Oracle*)h;
    for(long long i = 0; i < count; i++)
    {
        const SvcmMergeQuery &q = queries[i];
        Bsdf bsdf;
        QueryBsdf(o->scene, q.normal, q.dirFix, q.matID, false, bsdf);
        SubPathState st;
        st.origin = V3(0); st.direction = -Load3(q.dirFix); st.throughput = V3(1);
        st.pathLength = q.pathLength; st.isFiniteLight = false; st.specularPath = false;
        st.dVCM = q.dVCM; st.dVC = 0; st.dVM = q.dVM;
        V3 result(0);
        long long cand = 0, acc = 0;
        if(bsdf.matID >= 0 && !bsdf.isDelta)
            result = o->RangeQuery(Load3(q.position), bsdf, st, NULL, &cand, &acc);
        Store3(contrib + 3 * i, result);
        if(counts) { counts[2 * i] = int(cand); counts[2 * i + 1] = int(acc); }
    }
}

ORACLE_API void oracle_intersect(void *h, const float *rays, long long count, int *hit,
                                 float *dist, float *normal, int *matID, int *lightID)
{
    Oracle *o = (Oracle*)h;
    for(long long i = 0; i < count; i++)
    {
        Ray ray; ray.org = Load3(rays + 6 * i); ray.dir = Load3(rays + 6 * i + 3); ray.tmin = 0;
        Isect isect; isect.dist = 1e36f; isect.matID = -1; isect.lightID = -1; isect.normal = V3(0);
        int prim = -1;
        const bool found = SceneIntersect(o->scene, ray, isect, &prim);
        hit[i]  = found ? prim : -1;
        dist[i] = found ? isect.dist : 1e36f;
        Store3(normal + 3 * i, found ? isect.normal : V3(0));
        matID[i]   = found ? isect.matID : -1;
        lightID[i] = found ? isect.lightID : -1;
    }
}

ORACLE_API void oracle_occluded(void *h, const float *segments, long long count, int *occluded)
{
    Oracle *o = (Oracle*)h;
    for(long long i = 0; i < count; i++)
        occluded[i] = SceneOccluded(o->scene, Load3(segments + 7 * i), Load3(segments + 7 * i + 3),
                                    segments[7 * i + 6]) ? 1 : 0;
}

ORACLE_API void oracle_bsdf_evaluate(void *h, const float *in, long long count, int fixIsLight,
                                     float *out)
{
    Oracle *o = (Oracle*)h;
    for(long long i = 0; i < count; i++)
    {
        const float *p = in + 10 * i; float *r = out + 8 * i;
        int matID; memcpy(&matID, p + 9, 4);
        Bsdf b; QueryBsdf(o->scene, p, p + 3, matID, fixIsLight != 0, b);
        for(int k = 0; k < 8; k++) r[k] = 0;
        if(b.matID < 0) continue;
        float cosGen = 0, dirPdf = 0, revPdf = 0;
        const V3 f = BsdfEvaluate(o->scene, b, Load3(p + 6), cosGen, &dirPdf, &revPdf);
        Store3(r, f);
        r[3] = IsZero(f) ? 0.f : cosGen;
        r[4] = dirPdf; r[5] = revPdf; r[6] = b.contProb;
        r[7] = float(1 | (b.isDelta ? 2 : 0));
    }
}

ORACLE_API void oracle_bsdf_sample(void *h, const float *in, long long count, int fixIsLight,
                                   float *out, unsigned *events)
{
    Oracle *o = (Oracle*)h;
    for(long long i = 0; i < count; i++)
    {
        const float *p = in + 10 * i; float *r = out + 8 * i;
        int matID; memcpy(&matID, p + 9, 4);
        Bsdf b; QueryBsdf(o->scene, p, p + 3, matID, fixIsLight != 0, b);
        for(int k = 0; k < 8; k++) r[k] = 0;
        events[i] = 0;
        if(b.matID < 0) continue;
        V3 dirGen(0); float pdfW = 0, cosGen = 0; uint ev = 0;
        const V3 f = BsdfSample(o->scene, b, Load3(p + 6), dirGen, pdfW, cosGen, &ev);
        if(IsZero(f)) continue;
        Store3(r, f); Store3(r + 3, dirGen); r[6] = pdfW; r[7] = cosGen;
        events[i] = ev;
    }
}

// The light classes and the camera on random inputs (layouts: include/svcm.h, svcm_light_*_batch
// and svcm_camera_ray_batch). op 0 Emit, 1 Illuminate, 2 GetRadiance.
ORACLE_API void oracle_light(void *h, int op, int lightIndex, const float *in, long long count, float *out)
{
    Oracle *o = (Oracle*)h;
    const SvcmLight &l = o->scene.lights[size_t(lightIndex)];
    for(long long i = 0; i < count; i++)
    {
        const float *p = in + 8 * i; float *r = out + 12 * i;
        for(int k = 0; k < 12; k++) r[k] = 0;
        if(op == 0)
        {
            V3 pos(0), dir(0); float emissionPdfW = 0, directPdfA = 0, cosLight = 0;
            const V3 e = LightEmit(l, o->scene.sphere, p[0], p[1], p[2], p[3], pos, dir, emissionPdfW, directPdfA, cosLight);
            Store3(r, e); Store3(r + 3, pos); Store3(r + 6, dir);
            r[9] = emissionPdfW; r[10] = directPdfA; r[11] = cosLight;
        }
        else if(op == 1)
        {
            V3 dir(0); float dist = 0, directPdfW = 0, emissionPdfW = 0, cosAtLight = 0;
            const V3 e = LightIlluminate(l, o->scene.sphere, Load3(p), p[3], p[4], dir, dist, directPdfW,
                                         emissionPdfW, cosAtLight);
            Store3(r, e); Store3(r + 3, dir); r[6] = dist;
            if(!IsZero(e)) { r[7] = directPdfW; r[8] = emissionPdfW; r[9] = cosAtLight; }
        }
        else
        {
            float directPdfA = 0, emissionPdfW = 0;
            const V3 e = LightGetRadiance(l, o->scene.sphere, Load3(p), directPdfA, emissionPdfW);
            Store3(r, e); r[3] = directPdfA; r[4] = emissionPdfW;
        }
    }
}

ORACLE_API void oracle_camera(void *h, const float *in, long long count, float *out)
{
    Oracle *o = (Oracle*)h;
    const SvcmCamera &cam = o->scene.camera;
    for(long long i = 0; i < count; i++)
    {
        const float *p = in + 5 * i; float *r = out + 6 * i;
        const V3 camPos = Load3(cam.position);
        const V3 dir = Normalize(TransformPoint(cam.rasterToWorld, V3(p[0], p[1], 0)) - camPos);   // src/camera.hxx:108-117
        const V3 raster = TransformPoint(cam.worldToRaster, Load3(p + 2));                          // :95-99
        Store3(r, dir); r[3] = raster.x; r[4] = raster.y;
        r[5] = (raster.x >= 0 && raster.y >= 0 && raster.x < cam.resolution[0] && raster.y < cam.resolution[1]) ? 1.f : 0.f;
    }
}

ORACLE_API void oracle_rng_stream(int seed, int iteration, int phase, unsigned firstPath,
                                  int paths, int draws, float *out)
{
    Rng rng(seed, 1);
    for(int p = 0; p < paths; p++)
    {
        rng.SetPath(iteration, phase, int(firstPath + unsigned(p)));
        for(int k = 0; k < draws; k++)
            out[size_t(p) * draws + k] = rng.GetFloat();
    }
}

// counters of the last iteration: lightRays, cameraRays, shadowRays, mergeQueries,
// mergeCandidates, mergeAccepted, connections
ORACLE_API void oracle_counters(void *h, long long out[7])
{
    const Counters &c = ((Oracle*)h)->cnt;
    out[0] = c.lightRays; out[1] = c.cameraRays; out[2] = c.shadowRays;
    out[3] = c.mergeQueries; out[4] = c.mergeCandidates; out[5] = c.mergeAccepted;
    out[6] = c.connections;
}

// Framebuffer::TotalLuminance, src/framebuffer.hxx:89-102 (float accumulation, row-major)
ORACLE_API float oracle_total_luminance(const float *rgb, int width, int height)
{
    float lum = 0;
    for(int y = 0; y < height; y++)
        for(int x = 0; x < width; x++)
            lum += Luminance(Load3(rgb + 3 * (size_t(x) + size_t(y) * width)));
    return lum;
}

// The x86 float -> unsigned char conversion the reference's `byte(float)` casts compile to
// (cvttss2si to a 32-bit integer, low byte kept; NaN and out-of-range give the "integer
// indefinite" 0x80000000, i.e. byte 0). Written out so the device code can restate it.
static inline unsigned char ByteOfFloat(float f)
{
    if(!(std::fabs(f) < 2147483648.f)) return 0;
    return (unsigned char)(unsigned(int(f)) & 0xFFu);
}

// Pixel payload of Framebuffer::SaveBMP (src/framebuffer.hxx:196-215) applied to the image
// AbstractRenderer::GetFramebuffer returns (src/renderer.hxx:49-55: accumulated colour times
// 1.f / iterations): rows bottom-up, B G R bytes, no row padding, each byte
// byte(min(255, max(0, pow(c, 1/gamma) * 255))) with std::min/std::max's NaN behaviour.
ORACLE_API void oracle_bmp_payload(const float *accum, int width, int height, int iterations,
                                   float gamma, unsigned char *out)
{
    const float scale = iterations > 0 ? 1.f / iterations : 1.f;
    const float invGamma = 1.f / gamma;
    for(int y = 0; y < height; y++)
        for(int x = 0; x < width; x++)
        {
            const float *src = accum + 3 * (size_t(x) + size_t(height - y - 1) * width);
            for(int c = 0; c < 3; c++)
            {
                const float value = iterations > 0 ? src[2 - c] * scale : src[2 - c];
                const float g = std::pow(value, invGamma) * 255.f;
                *out++ = ByteOfFloat(std::min(255.f, std::max(0.f, g)));
            }
        }
}

// Pixel payload of Framebuffer::SaveHDR (src/framebuffer.hxx:229-249): rows top-down, shared-
// exponent R G B E bytes. Whether the unqualified frexp there binds to the float or the double
// overload does not matter: a double quotient of two 24-bit significands rounds to the same float
// as the float quotient (53 >= 2*24+2), which a 2e8-sample sweep confirmed.
ORACLE_API void oracle_hdr_payload(const float *accum, int width, int height, int iterations,
                                   unsigned char *out)
{
    const float scale = iterations > 0 ? 1.f / iterations : 1.f;
    for(size_t i = 0; i < size_t(width) * size_t(height); i++, out += 4)
    {
        float rgb[3];
        for(int c = 0; c < 3; c++) rgb[c] = iterations > 0 ? accum[3 * i + c] * scale : accum[3 * i + c];
        out[0] = out[1] = out[2] = out[3] = 0;
        float v = std::max(rgb[0], std::max(rgb[1], rgb[2]));
        if(v >= 1e-32f)
        {
            int e;
            v = float(::frexp(double(v), &e) * 256.f / v);
            out[0] = ByteOfFloat(rgb[0] * v);
            out[1] = ByteOfFloat(rgb[1] * v);
            out[2] = ByteOfFloat(rgb[2] * v);
            out[3] = (unsigned char)(unsigned(e + 128) & 0xFFu);
        }
    }
}

// The reference's OpenMP driver loop, src/smallvcm.cxx:52-151, over oracle renderers: one
// private renderer per thread with seed baseSeed + thread, iterations distributed by
// `omp parallel for`, framebuffers averaged. Used as the "port" CPU baseline when
// oracle/_ref is not available. Returns wall seconds; outRgb may be NULL.
ORACLE_API double oracle_render(const SvcmSceneDesc *scene, int algorithm, int iterations,
                                int threads, int baseSeed, int rngMode, unsigned maxPathLength,
                                unsigned minPathLength, float radiusFactor, float radiusAlpha,
                                float *outRgb)
{
    if(threads < 1) threads = 1;
    std::vector<void*> renderers(threads);
    for(int i = 0; i < threads; i++)
        renderers[i] = oracle_create(scene, algorithm, radiusFactor, radiusAlpha, baseSeed + i,
                                     rngMode, maxPathLength, minPathLength);

    const double t0 = omp_get_wtime();
    omp_set_num_threads(threads);
#pragma omp parallel for
    for(int iter = 0; iter < iterations; iter++)
        ((Oracle*)renderers[omp_get_thread_num()])->RunIteration(iter);
    const double t1 = omp_get_wtime();

    // GetFramebuffer of every used renderer (sum / mIterations), then the average over the
    // used renderers: src/renderer.hxx:49-55, src/smallvcm.cxx:116-142
    if(outRgb)
    {
        const size_t n = ((Oracle*)renderers[0])->framebuffer.size();
        std::vector<V3> sum(n, V3(0));
        int used = 0;
        for(int i = 0; i < threads; i++)
        {
            Oracle *o = (Oracle*)renderers[i];
            if(o->iterations == 0) continue;
            const float scale = 1.f / o->iterations;
            for(size_t k = 0; k < n; k++)
            {
                const V3 c = o->framebuffer[k] * V3(scale);
                sum[k] = used == 0 ? c : sum[k] + c;
            }
            used++;
        }
        const float inv = 1.f / used;
        for(size_t k = 0; k < n; k++) sum[k] = sum[k] * V3(inv);
        memcpy(outRgb, sum.data(), sizeof(float) * 3 * n);
    }
    for(int i = 0; i < threads; i++) oracle_destroy(renderers[i]);
    return t1 - t0;
}
