// TEST INFRASTRUCTURE — not part of the product.
//
// The REFERENCE ARITHMETIC with a per-path counter-based random stream: every reference
// header is compiled in place from /root/reference/src; only (1) rng.hxx is replaced
// through its include guard by the `Rng` below and (2) a throw-away copy of vertexcm.hxx
// carries two `mRng.SetPath(...)` hook lines (oracle/ref_glue/patch_vertexcm.sh).
// Built into oracle/_ref/libsvcm_ref_ctr.so.
//
// Why: the reference draws from ONE sequential mt19937_64 stream across all paths
// (src/vertexcm.hxx:1030), which no parallel renderer can reproduce; keyed per
// (seed, iteration, phase, path) the same arithmetic becomes reproducible lane by lane
// (SURVEY §8c). This library is what pins oracle/vcm_oracle.cpp (tests/test_oracle.py):
// both must agree BIT FOR BIT on images, light vertices, grids and merge contributions.
#include <vector>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <time.h>
#include <algorithm>
#include <map>
#include <string>

#include "math.hxx"

// ---------------------------------------------------------------------------------------
// replacement for src/rng.hxx (same five-method interface, src/rng.hxx:48-78)
// Generator: Philox2x32-7 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"),
// counter = (draw index, path index), key = seed + 0x85EBCA6B * (2*iteration + phase);
// float = (top 24 bits | 1) / 2^24: strictly inside (0,1).  The same definition is restated in
// oracle/vcm_oracle.cpp and in the product (smallvcm_b200/csrc/svcm_device.cuh).
// ---------------------------------------------------------------------------------------
#define __RNG_HXX__
class Rng
{
public:
    Rng(int aSeed = 1234) : mSeed(uint(aSeed)), mKey(uint(aSeed)), mPath(0), mDraw(0) {}

    void SetPath(int aIteration, int aPhase, int aPathIdx)
    {
        mKey  = mSeed + 0x85EBCA6Bu * (2u * uint(aIteration) + uint(aPhase));
        mPath = uint(aPathIdx);
        mDraw = 0;
    }

    uint GetUint()
    {
        uint c0 = mDraw++, c1 = mPath, key = mKey;
        for(int round = 0; round < 7; round++)
        {
            const unsigned long long prod = 0xD256D193ull * c0;
            const uint hi = uint(prod >> 32), lo = uint(prod);
            c0 = hi ^ key ^ c1;
            c1 = lo;
            key += 0x9E3779B9u;
        }
        return c0;
    }

    int   GetInt()   { return int(GetUint() >> 1); }
    float GetFloat() { return float((GetUint() >> 8) | 1u) * (1.0f / 16777216.0f); }

    Vec2f GetVec2f()
    {
        float a = GetFloat();
        float b = GetFloat();
        return Vec2f(a, b);
    }

    Vec3f GetVec3f()
    {
        float a = GetFloat();
        float b = GetFloat();
        float c = GetFloat();
        return Vec3f(a, b, c);
    }

private:
    uint mSeed, mKey, mPath, mDraw;
};

#include "ray.hxx"
#include "geometry.hxx"
#include "camera.hxx"
#include "framebuffer.hxx"
#include "scene.hxx"
#include <eyelight.hxx>   // the patched copies: first -I directory, angle brackets skip "."
#include <pathtracer.hxx>
#include "bsdf.hxx"
#include <vertexcm.hxx>

#define main smallvcm_reference_main
#include "smallvcm.cxx"
#undef main

#include "svcm.h"

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

struct Handle
{
    Scene            *scene;
    VertexCM         *renderer;   // NULL for the two unidirectional integrators
    AbstractRenderer *any;        // whatever CreateRenderer would have built
    int               width, height;
};

Scene *MakeScene(int sceneId, int width, int height)
{
    if(sceneId < 0 || sceneId >= SizeOfArray(g_SceneConfigs)) return NULL;
    Scene *scene = new Scene;
    scene->LoadCornellBox(Vec2i(width, height), g_SceneConfigs[sceneId]);
    scene->BuildSceneSphere();
    return scene;
}

Vec3f Get3(const float *p) { return Vec3f(p[0], p[1], p[2]); }
void  Put3(float *p, const Vec3f &v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

} // namespace

// One private renderer, exactly what CreateRenderer makes for the five VertexCM algorithms
// (src/config.hxx:124-138). `algorithm` is VertexCM::AlgorithmType (0 lt .. 4 vcm).
REF_API void *refctr_create(int sceneId, int algorithm, int width, int height, int seed,
                            float radiusFactor, float radiusAlpha,
                            unsigned maxPathLength, unsigned minPathLength)
{
    Scene *scene = MakeScene(sceneId, width, height);
    if(!scene) return NULL;
    Handle *h = new Handle;
    h->scene    = scene;
    h->width    = width;
    h->height   = height;
    h->renderer = NULL;
    if(algorithm == SVCM_ALG_PATH_TRACE)     h->any = new PathTracer(*scene, seed);   // src/config.hxx:121-122
    else if(algorithm == SVCM_ALG_EYE_LIGHT) h->any = new EyeLight(*scene, seed);     // src/config.hxx:119-120
    else
    {
        h->renderer = new VertexCM(*scene, VertexCM::AlgorithmType(algorithm),
                                   radiusFactor, radiusAlpha, seed);
        h->any = h->renderer;
    }
    h->any->mMaxPathLength = maxPathLength;   // after construction, like src/smallvcm.cxx:70-71
    h->any->mMinPathLength = minPathLength;
    return h;
}

REF_API void refctr_destroy(void *handle)
{
    Handle *h = (Handle*)handle;
    if(!h) return;
    delete h->any;
    delete h->scene;
    delete h;
}

REF_API void refctr_run_iteration(void *handle, int iteration)
{
    ((Handle*)handle)->any->RunIteration(iteration);
}

// The ACCUMULATED framebuffer (mFramebuffer, not divided by mIterations)
REF_API int refctr_accum(void *handle, float *outRgb)
{
    Handle *h = (Handle*)handle;
    memcpy(outRgb, &h->any->mFramebuffer.mColor[0],
           sizeof(float) * 3 * size_t(h->width) * size_t(h->height));
    return h->any->mIterations;
}

REF_API long long refctr_num_light_vertices(void *handle)
{
    Handle *h = (Handle*)handle;
    return h->renderer ? (long long)h->renderer->mLightVertices.size() : 0;
}

// mLightVertices of the last iteration in the 52-byte merge form + dVC, the shading
// normal frame and local direction needed by connections; pathEnds = mPathEnds.
REF_API void refctr_light_vertices(void *handle, SvcmMergeVertex *out, float *dVC,
                                   int *matID, float *normal, int *pathEnds)
{
    Handle *h = (Handle*)handle;
    const std::vector<VertexCM::LightVertex> &lv = h->renderer->mLightVertices;
    for(size_t i = 0; i < lv.size(); i++)
    {
        if(out)
        {
            Put3(out[i].position,   lv[i].mHitpoint);
            Put3(out[i].throughput, lv[i].mThroughput);
            Put3(out[i].dirFix,     lv[i].mBsdf.WorldDirFix());
            out[i].pathLength = lv[i].mPathLength;
            out[i].dVCM       = lv[i].dVCM;
            out[i].dVM        = lv[i].dVM;
            out[i].contProb   = lv[i].mBsdf.ContinuationProb();
        }
        if(dVC)    dVC[i]   = lv[i].dVC;
        if(matID)  matID[i] = lv[i].mBsdf.mMaterialID;
        if(normal) Put3(normal + 3 * i, lv[i].mBsdf.mFrame.mZ);
    }
    if(pathEnds)
        memcpy(pathEnds, &h->renderer->mPathEnds[0], sizeof(int) * h->renderer->mPathEnds.size());
}

// Iteration constants of the last RunIteration: radius is recomputed the reference's way
REF_API void refctr_iteration_constants(void *handle, int iteration, float *radius,
                                        float *vmNormalization, float *misVmWeightFactor,
                                        float *misVcWeightFactor)
{
    Handle *h = (Handle*)handle;
    float r = h->renderer->mBaseRadius;
    r /= std::pow(float(iteration + 1), 0.5f * (1 - h->renderer->mRadiusAlpha));
    r = std::max(r, 1e-7f);
    *radius            = r;
    *vmNormalization   = h->renderer->mVmNormalization;
    *misVmWeightFactor = h->renderer->mMisVmWeightFactor;
    *misVcWeightFactor = h->renderer->mMisVcWeightFactor;
}

// HashGrid members after the last Build (src/hashgrid.hxx:199-208)
REF_API int refctr_grid(void *handle, float bbox[6], int *cellEnds, int *indices)
{
    Handle *h = (Handle*)handle;
    const HashGrid &g = h->renderer->mHashGrid;
    if(bbox) { Put3(bbox, g.mBBoxMin); Put3(bbox + 3, g.mBBoxMax); }
    if(cellEnds) memcpy(cellEnds, &g.mCellEnds[0], sizeof(int) * g.mCellEnds.size());
    if(indices)  memcpy(indices,  &g.mIndices[0],  sizeof(int) * g.mIndices.size());
    return (int)g.mCellEnds.size();
}

// HashGrid::Process with the reference's own RangeQuery (src/vertexcm.hxx:109-178) for
// caller-made camera vertices, against the light vertices / grid of the last iteration.
REF_API void refctr_merge_queries(void *handle, const SvcmMergeQuery *queries, long long count,
                                  float *contrib)
{
    Handle *h = (Handle*)handle;
    VertexCM &vcm = *h->renderer;
    for(long long i = 0; i < count; i++)
    {
        const SvcmMergeQuery &q = queries[i];
        Ray ray(Vec3f(0), -Get3(q.dirFix), 0);
        Isect isect(1.f);
        isect.matID   = q.matID;
        isect.lightID = -1;
        isect.normal  = Get3(q.normal);
        VertexCM::CameraBSDF bsdf(ray, isect, *h->scene);

        VertexCM::SubPathState state;
        state.mOrigin = Vec3f(0); state.mDirection = ray.dir; state.mThroughput = Vec3f(1);
        state.mPathLength = q.pathLength; state.mIsFiniteLight = 0; state.mSpecularPath = 0;
        state.dVCM = q.dVCM; state.dVC = 0; state.dVM = q.dVM;

        const Vec3f pos = Get3(q.position);
        Vec3f result(0);
        if(bsdf.IsValid() && !bsdf.IsDelta())
        {
            VertexCM::RangeQuery query(vcm, pos, bsdf, state);
            vcm.mHashGrid.Process(vcm.mLightVertices, query);
            result = query.GetContrib();
        }
        Put3(contrib + 3 * i, result);
    }
}

// Scene::Intersect (src/scene.hxx:53-70): rays = 6 floats (org, dir); dist starts at 1e36
REF_API void refctr_intersect(void *handle, const float *rays, long long count, int *hit,
                              float *dist, float *normal, int *matID, int *lightID)
{
    Handle *h = (Handle*)handle;
    const GeometryList *list = dynamic_cast<const GeometryList*>(h->scene->mGeometry);
    for(long long i = 0; i < count; i++)
    {
        Ray ray(Get3(rays + 6 * i), Get3(rays + 6 * i + 3), 0);
        Isect isect(1e36f);
        isect.matID = -1; isect.lightID = -1; isect.normal = Vec3f(0);
        const bool found = h->scene->Intersect(ray, isect);
        int prim = -1;
        if(found)
        {
            // which list entry produced the hit: re-run the closest-hit loop
            Isect probe(1e36f);
            for(int k = 0; k < (int)list->mGeometry.size(); k++)
                if(list->mGeometry[k]->Intersect(ray, probe)) prim = k;
        }
        hit[i] = prim;
        dist[i] = found ? isect.dist : 1e36f;
        Put3(normal + 3 * i, found ? isect.normal : Vec3f(0));
        matID[i]   = found ? isect.matID : -1;
        lightID[i] = found ? isect.lightID : -1;
    }
}

// Scene::Occluded (src/scene.hxx:72-85): segments = 7 floats (point, dir, tMax)
REF_API void refctr_occluded(void *handle, const float *segments, long long count, int *occluded)
{
    Handle *h = (Handle*)handle;
    for(long long i = 0; i < count; i++)
        occluded[i] = h->scene->Occluded(Get3(segments + 7 * i), Get3(segments + 7 * i + 3),
                                         segments[7 * i + 6]) ? 1 : 0;
}

namespace {
template<bool tLight>
void EvaluateOne(const Scene &scene, const float *in, float *out)
{
    int matID; memcpy(&matID, in + 9, 4);
    Ray ray(Vec3f(0), -Get3(in + 3), 0);
    Isect isect(1.f);
    isect.matID = matID; isect.lightID = -1; isect.normal = Get3(in);
    BSDF<tLight> bsdf(ray, isect, scene);
    for(int k = 0; k < 8; k++) out[k] = 0;
    if(!bsdf.IsValid()) return;
    float cosGen = 0, dirPdf = 0, revPdf = 0;
    const Vec3f f = bsdf.Evaluate(scene, Get3(in + 6), cosGen, &dirPdf, &revPdf);
    Put3(out, f);
    out[3] = f.IsZero() ? 0.f : cosGen;
    out[4] = dirPdf; out[5] = revPdf;
    out[6] = bsdf.ContinuationProb();
    out[7] = float(1 | (bsdf.IsDelta() ? 2 : 0));
}

template<bool tLight>
void SampleOne(const Scene &scene, const float *in, float *out, unsigned *event)
{
    int matID; memcpy(&matID, in + 9, 4);
    Ray ray(Vec3f(0), -Get3(in + 3), 0);
    Isect isect(1.f);
    isect.matID = matID; isect.lightID = -1; isect.normal = Get3(in);
    BSDF<tLight> bsdf(ray, isect, scene);
    for(int k = 0; k < 8; k++) out[k] = 0;
    *event = 0;
    if(!bsdf.IsValid()) return;
    Vec3f dirGen(0); float pdfW = 0, cosGen = 0; uint ev = 0;
    const Vec3f f = bsdf.Sample(scene, Get3(in + 6), dirGen, pdfW, cosGen, &ev);
    if(f.IsZero()) return;
    Put3(out, f); Put3(out + 3, dirGen); out[6] = pdfW; out[7] = cosGen;
    *event = ev;
}
} // namespace

// BSDF<>::Setup + Evaluate, layout of svcm_bsdf_evaluate_batch (include/svcm.h)
REF_API void refctr_bsdf_evaluate(void *handle, const float *in, long long count,
                                  int fixIsLight, float *out)
{
    Handle *h = (Handle*)handle;
    for(long long i = 0; i < count; i++)
    {
        if(fixIsLight) EvaluateOne<true >(*h->scene, in + 10 * i, out + 8 * i);
        else           EvaluateOne<false>(*h->scene, in + 10 * i, out + 8 * i);
    }
}

// BSDF<>::Setup + Sample, layout of svcm_bsdf_sample_batch (include/svcm.h)
REF_API void refctr_bsdf_sample(void *handle, const float *in, long long count,
                                int fixIsLight, float *out, unsigned *events)
{
    Handle *h = (Handle*)handle;
    for(long long i = 0; i < count; i++)
    {
        if(fixIsLight) SampleOne<true >(*h->scene, in + 10 * i, out + 8 * i, events + i);
        else           SampleOne<false>(*h->scene, in + 10 * i, out + 8 * i, events + i);
    }
}

// The reference's light classes and camera on the same inputs (layouts: include/svcm.h,
// svcm_light_*_batch / svcm_camera_ray_batch). op 0 Emit, 1 Illuminate, 2 GetRadiance.
REF_API void refctr_light(void *handle, int op, int lightIndex, const float *in, long long count, float *out)
{
    Handle *h = (Handle*)handle;
    const AbstractLight *l = h->scene->GetLightPtr(lightIndex);
    const SceneSphere &ss = h->scene->mSceneSphere;
    for(long long i = 0; i < count; i++)
    {
        const float *p = in + 8 * i; float *r = out + 12 * i;
        for(int k = 0; k < 12; k++) r[k] = 0;
        if(op == 0)
        {
            Vec3f pos(0), dir(0); float emissionPdfW = 0, directPdfA = 0, cosLight = 0;
            const Vec3f e = l->Emit(ss, Vec2f(p[0], p[1]), Vec2f(p[2], p[3]), pos, dir, emissionPdfW, &directPdfA, &cosLight);
            Put3(r, e); Put3(r + 3, pos); Put3(r + 6, dir);
            r[9] = emissionPdfW; r[10] = directPdfA; r[11] = cosLight;
        }
        else if(op == 1)
        {
            Vec3f dir(0); float dist = 0, directPdfW = 0, emissionPdfW = 0, cosAtLight = 0;
            const Vec3f e = l->Illuminate(ss, Get3(p), Vec2f(p[3], p[4]), dir, dist, directPdfW, &emissionPdfW, &cosAtLight);
            Put3(r, e); Put3(r + 3, dir); r[6] = dist;
            if(!e.IsZero()) { r[7] = directPdfW; r[8] = emissionPdfW; r[9] = cosAtLight; }
        }
        else
        {
            float directPdfA = 0, emissionPdfW = 0;
            const Vec3f e = l->GetRadiance(ss, Get3(p), Get3(p + 3), &directPdfA, &emissionPdfW);
            Put3(r, e); r[3] = directPdfA; r[4] = emissionPdfW;
        }
    }
}

REF_API void refctr_camera(void *handle, const float *in, long long count, float *out)
{
    Handle *h = (Handle*)handle;
    const Camera &cam = h->scene->mCamera;
    for(long long i = 0; i < count; i++)
    {
        const float *p = in + 5 * i; float *r = out + 6 * i;
        const Ray ray = cam.GenerateRay(Vec2f(p[0], p[1]));
        const Vec2f raster = cam.WorldToRaster(Get3(p + 2));
        Put3(r, ray.dir); r[3] = raster.x; r[4] = raster.y;
        r[5] = cam.CheckRaster(raster) ? 1.f : 0.f;
    }
}

// the replaced generator, for pinning the three restatements of it against each other
REF_API void refctr_rng_stream(int seed, int iteration, int phase, unsigned firstPath,
                               int paths, int draws, float *out)
{
    Rng rng(seed);
    for(int p = 0; p < paths; p++)
    {
        rng.SetPath(iteration, phase, int(firstPath + unsigned(p)));
        for(int k = 0; k < draws; k++)
            out[size_t(p) * draws + k] = rng.GetFloat();
    }
}
