// svcm_kernels.cu — the wavefront kernels of one VCM iteration (sm_100a).
//
// One iteration of VertexCM::RunIteration (reference src/vertexcm.hxx:284-548) becomes
//   light pass : k_lightBounce  x (maxLen-1)   one warp lane per live light sub-path
//   grid build : k_gridBBox, k_gridKeys, 3-kernel exclusive scan, k_gridScatter
//   camera pass: k_cameraBounce x maxLen       one warp lane per live camera sub-path
//   merge      : k_merge                        G lanes cooperate on one range query
// Sub-path state lives in HBM as four 16-byte SoA streams (svcm_internal.h) and is
// compacted between bounces with warp-aggregated appends, so every bounce kernel reads and
// writes full coalesced 128-bit lanes. The scene (<= 64 primitives) rides in the constant
// bank as a __grid_constant__ parameter: the trace kernels are instruction-bound, only the
// vertex store, the grid build and the merge move real HBM traffic.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
#include "svcm_internal.h"
#include "svcm_kutil.cuh"

// Resident blocks per SM the register allocation of the hot kernels is held to. Unset: ptxas' own
// choice for a 256-thread block (A/B builds: -DSVCM_<KERNEL>_MINBLOCKS=n, tools/build_variant.sh).
#define SVCM_CAMERA_MINBLOCKS_DEFAULT 3
#ifndef SVCM_CAMERA_MINBLOCKS
#define SVCM_CAMERA_MINBLOCKS SVCM_CAMERA_MINBLOCKS_DEFAULT
#endif
#ifndef SVCM_LIGHT_MINBLOCKS
#define SVCM_LIGHT_MINBLOCKS 3       // like the camera walk: three resident blocks beat two with every value in a register (0.76 vs 0.83 ms)
#endif
#define SVCM_LB_LIGHT __launch_bounds__(SVCM_BLOCK, SVCM_LIGHT_MINBLOCKS)
#ifndef SVCM_CONNECT_MINBLOCKS
#define SVCM_CONNECT_MINBLOCKS 4      // 64 registers, 24 spill bytes: 1.40 -> 1.33 ms for the camera pass alone at 1024^2 (latency-bound gathers)
#endif
#ifdef SVCM_CONNECT_MINBLOCKS
#define SVCM_LB_CONNECT __launch_bounds__(SVCM_BLOCK, SVCM_CONNECT_MINBLOCKS)
#else
#define SVCM_LB_CONNECT __launch_bounds__(SVCM_BLOCK)
#endif
#ifdef SVCM_SHADOW_MINBLOCKS
#define SVCM_LB_SHADOW __launch_bounds__(SVCM_BLOCK, SVCM_SHADOW_MINBLOCKS)
#else
#define SVCM_LB_SHADOW __launch_bounds__(SVCM_BLOCK)
#endif
#ifdef SVCM_MERGE_MINBLOCKS
#define SVCM_LB_MERGE __launch_bounds__(SVCM_BLOCK, SVCM_MERGE_MINBLOCKS)
#else
#define SVCM_LB_MERGE __launch_bounds__(SVCM_BLOCK)
#endif

namespace svcm {

// Queue one visibility test: "add `contrib` to `pixel` unless point -> point + dir*tMax is
// blocked" (Scene::Occluded, reference src/scene.hxx:72-85). If the queue is full the ray is
// traced on the spot — slower, never wrong.
template<bool kBvh>
SVCM_D void emitShadowRay(const DevScene &sc, const ShadowQueue &q, uint32_t which, float *accum,
                          f3 point, f3 dir, float tMax, uint32_t pixel, f3 contrib)
{
    const uint32_t slot = activeAppend(q.counts + which);
    if(slot < q.capacity)
    {
        q.r0[slot] = mk4(point, tMax);
        q.r1[slot] = mk4(dir, __uint_as_float(pixel));
        q.r2[slot] = mk4(contrib, 0.f);
    }
    else if(!sceneOccluded<kBvh>(sc, point, dir, tMax)) addPixel(accum, pixel, contrib);
}

// local (per-rank) path index <-> global path index: rows are dealt round-robin to ranks
SVCM_D uint32_t localToGlobalPath(const IterConst &ic, uint32_t l)
{
    if(ic.worldSize == 1) return l;
    const uint32_t ly = l / uint32_t(ic.resX), x = l - ly * uint32_t(ic.resX);
    return (ly * uint32_t(ic.worldSize) + uint32_t(ic.rank)) * uint32_t(ic.resX) + x;
}
SVCM_D uint32_t globalToLocalPath(const IterConst &ic, uint32_t p)
{
    if(ic.worldSize == 1) return p;
    const uint32_t y = p / uint32_t(ic.resX), x = p - y * uint32_t(ic.resX);
    return (y / uint32_t(ic.worldSize)) * uint32_t(ic.resX) + x;
}

// ------------------------------------------------------------------------------------------
// sub-path state in registers
// ------------------------------------------------------------------------------------------
struct SubPath
{
    f3 origin, direction, throughput;
    float dVCM, dVC, dVM;
    uint32_t path;      // global path index (= pixel index of the pair, = RNG stream id)
    uint32_t length;    // mPathLength
    bool finiteLight;   // mIsFiniteLight
    bool specular;      // mSpecularPath
    uint32_t extra;     // light: number of stored vertices; camera: target pixel
    PathRng rng;
};

SVCM_D void loadSubPath(const PathQueue &q, uint32_t i, uint32_t key, SubPath &s)
{
    const float4 a = q.s0[i], b = q.s1[i], c = q.s2[i];
    const uint4  d = q.s3[i];
    s.origin = xyz(a); s.dVCM = a.w;
    s.direction = xyz(b); s.dVC = b.w;
    s.throughput = xyz(c); s.dVM = c.w;
    s.path = d.x;
    s.length = d.y & PATH_LEN_MASK;
    s.finiteLight = (d.y & PATH_FLAG_FINITE) != 0u;
    s.specular = (d.y & PATH_FLAG_SPECULAR) != 0u;
    s.rng.key = key; s.rng.path = d.x; s.rng.draw = d.z;
    s.extra = d.w;
}
SVCM_D void storeSubPath(const PathQueue &q, uint32_t i, const SubPath &s)
{
    q.s0[i] = mk4(s.origin, s.dVCM);
    q.s1[i] = mk4(s.direction, s.dVC);
    q.s2[i] = mk4(s.throughput, s.dVM);
    q.s3[i] = make_uint4(s.path, s.length | (s.finiteLight ? PATH_FLAG_FINITE : 0u) |
                         (s.specular ? PATH_FLAG_SPECULAR : 0u), s.rng.draw, s.extra);
}

// SampleScattering, reference src/vertexcm.hxx:938-1006. Returns false to terminate.
template<bool kLight>
SVCM_D bool sampleScattering(const DevScene &sc, const IterConst &ic, const Bsdf &bsdf, f3 hitPoint,
                             SubPath &s)
{
    const float r0 = s.rng.next(), r1 = s.rng.next(), r2 = s.rng.next();
    float bsdfDirPdfW, cosThetaOut;
    uint32_t ev;
    f3 dirOut = s.direction;
    const f3 bsdfFactor = bsdfSample<kLight>(sc, bsdf, r0, r1, r2, dirOut, bsdfDirPdfW, cosThetaOut, ev);
    if(iszero(bsdfFactor)) return false;

    float bsdfRevPdfW = bsdfDirPdfW;
    if((ev & EV_SPECULAR) == 0u) bsdfRevPdfW = bsdfReversePdf(sc, bsdf, dirOut);

    const float contProb = bsdf.contProb;          // Russian roulette
    if(s.rng.next() > contProb) return false;

    bsdfDirPdfW *= contProb;
    bsdfRevPdfW *= contProb;

    if(ev & EV_SPECULAR)
    {
        s.dVCM = 0.f;                              // [tech. rep. (53)-(55)]
        s.dVC *= cosThetaOut;
        s.dVM *= cosThetaOut;
    }
    else
    {
        const float f = cosThetaOut / bsdfDirPdfW; // [tech. rep. (34)-(36)]
        s.dVC = f * (s.dVC * bsdfRevPdfW + s.dVCM + ic.misVmWeightFactor);
        s.dVM = f * (s.dVM * bsdfRevPdfW + s.dVCM * ic.misVcWeightFactor + 1.f);
        s.dVCM = 1.f / bsdfDirPdfW;
        s.specular = false;
    }
    s.origin = hitPoint;
    s.direction = dirOut;
    s.throughput = s.throughput * (bsdfFactor * (cosThetaOut / bsdfDirPdfW));
    return true;
}

// ------------------------------------------------------------------------------------------
// light pass
// ------------------------------------------------------------------------------------------

// GenerateLightSample, reference src/vertexcm.hxx:816-858
SVCM_D void generateLightSample(const DevScene &sc, const IterConst &ic, SubPath &s)
{
    const int   lightCount = sc.numLights;
    const float lightPickProb = 1.f / float(lightCount);
    int lightID = int(s.rng.next() * float(lightCount));
    const float dirX = s.rng.next(), dirY = s.rng.next();
    const float posX = s.rng.next(), posY = s.rng.next();
    lightID = min(lightID, lightCount - 1);                          // Scene::GetLightPtr clamp
    const DevLight &light = sc.lights[lightID];

    float emissionPdfW, directPdfA, cosLight;
    s.throughput = lightEmit(sc, light, dirX, dirY, posX, posY, s.origin, s.direction,
                             emissionPdfW, directPdfA, cosLight);
    emissionPdfW *= lightPickProb;
    directPdfA   *= lightPickProb;

    s.throughput  = s.throughput / emissionPdfW;
    s.length      = 1u;
    s.finiteLight = lightIsFinite(light.type);
    s.specular    = false;
    s.dVCM = directPdfA / emissionPdfW;
    if(!lightIsDelta(light.type))
    {
        const float usedCosLight = lightIsFinite(light.type) ? cosLight : 1.f;
        s.dVC = usedCosLight / emissionPdfW;
    }
    else s.dVC = 0.f;
    s.dVM = s.dVC * ic.misVcWeightFactor;
    s.extra = 0u;
}

// ConnectToCamera, reference src/vertexcm.hxx:862-933: splat the light vertex onto the
// image through the pinhole; float atomics because the target pixel is arbitrary.
template<bool kBvh>
SVCM_D void connectToCamera(const DevScene &sc, const IterConst &ic, const SubPath &s, f3 hitPoint,
                            const Bsdf &bsdf, const TraceArgs &a)
{
    const f3 camPos = ld3(sc.camPos), camFwd = ld3(sc.camFwd);
    f3 dirToCamera = camPos - hitPoint;
    if(dot(camFwd, -dirToCamera) <= 0.f) return;

    const f3 raster = transformPoint(sc.worldToRaster, hitPoint);      // Camera::WorldToRaster
    if(!(raster.x >= 0.f && raster.y >= 0.f && raster.x < sc.resX && raster.y < sc.resY)) return;

    const float distEye2 = lensqr(dirToCamera);
    const float distance = sqrtf(distEye2);
    dirToCamera = dirToCamera / distance;

    float cosToCamera, bsdfDirPdfW, bsdfRevPdfW;
    const f3 bsdfFactor = bsdfEvaluate(sc, bsdf, dirToCamera, cosToCamera, bsdfDirPdfW, bsdfRevPdfW);
    if(iszero(bsdfFactor)) return;
    bsdfRevPdfW *= bsdf.contProb;

    const float cosAtCamera = dot(camFwd, -dirToCamera);
    const float imagePointToCameraDist = sc.imagePlaneDist / cosAtCamera;
    const float imageToSolidAngleFactor = sqr(imagePointToCameraDist) / cosAtCamera;
    const float imageToSurfaceFactor = imageToSolidAngleFactor * fabsf(cosToCamera) / sqr(distance);
    const float cameraPdfA = imageToSurfaceFactor;

    const float wLight = (cameraPdfA / ic.lightSubPathCount) *
        (ic.misVmWeightFactor + s.dVCM + s.dVC * bsdfRevPdfW);            // [tech. rep. (46)]
    const float misWeight = ic.lightTraceOnly ? 1.f : (1.f / (wLight + 1.f));
    const float surfaceToImageFactor = 1.f / imageToSurfaceFactor;

    const f3 contrib = (misWeight * s.throughput * bsdfFactor) /
        (ic.lightSubPathCount * surfaceToImageFactor);
    if(iszero(contrib)) return;

    // Framebuffer::AddColor target, reference src/framebuffer.hxx:43-57; visibility deferred
    const uint32_t pixel = uint32_t(int(raster.x)) + uint32_t(int(raster.y)) * uint32_t(ic.resX);
    emitShadowRay<kBvh>(sc, a.shadow, 0u, a.accum, hitPoint, dirToCamera, distance, pixel, contrib);
}

// One bounce of every live light sub-path: reference src/vertexcm.hxx:328-393 (loop body)
template<bool kBvh>
__global__ void SVCM_LB_LIGHT
k_lightBounce(const __grid_constant__ DevScene sc, const __grid_constant__ IterConst ic,
              const __grid_constant__ TraceArgs a, const uint32_t bounce)
{
    const uint32_t count = (bounce == 1u) ? ic.localPaths : a.queueCounts[bounce];
    uint32_t nRays = 0u;

    for(uint32_t i0 = blockIdx.x * blockDim.x; i0 < count; i0 += gridDim.x * blockDim.x)
    {
        const uint32_t i = i0 + threadIdx.x;
        bool alive = i < count;
        SubPath s;
        if(alive)
        {
            if(bounce == 1u)
            {
                s.path = localToGlobalPath(ic, i);
                s.rng.key = ic.keyLight; s.rng.path = s.path; s.rng.draw = 0u;
                generateLightSample(sc, ic, s);
            }
            else loadSubPath(a.in, i, ic.keyLight, s);
        }

        Hit hit;
        Bsdf bsdf;
        f3 hitPoint = mk3(0.f);
        bool storeVertex = false;
        if(alive)
        {
            // origin offset instead of tmin; dist extended after the hit point is known
            const f3 org = mulAddUnfused(s.origin, s.direction, SVCM_EPS_RAY);
            nRays++;
            alive = sceneIntersect<kBvh>(sc, org, s.direction, hit);
            if(alive)
            {
                hitPoint = mulAddUnfused(org, s.direction, hit.dist);
                hit.dist += SVCM_EPS_RAY;
                bsdfSetup(sc, s.direction, hit.normal, hit.matID, bsdf);
                alive = bsdf.isValid();
            }
            if(alive)
            {
                // MIS quantities at the vertex [tech. rep. (31)-(33) / (34)-(36)]
                if(s.length > 1u || s.finiteLight) s.dVCM *= sqr(hit.dist);
                const float invCos = fabsf(bsdf.cosThetaFix());
                s.dVCM /= invCos; s.dVC /= invCos; s.dVM /= invCos;
                storeVertex = !bsdf.isDelta() && (ic.useVC || ic.useVM);
            }
        }

        // store the light vertex (PathVertex<true>) unless purely specular
        const uint32_t vidx = warpAppend64(a.counters + CNT_LIGHT_VERTS, storeVertex);
        if(storeVertex)
        {
            if(vidx < a.verts.capacity)
            {
                const float4 posLen = mk4(hitPoint, __uint_as_float(s.length));
                const float4 thrVcm = mk4(s.throughput, s.dVCM);
                if(ic.useVM)
                {
                    a.verts.posLen[vidx]   = posLen;
                    a.verts.thrVcm[vidx]   = thrVcm;
                    a.verts.dirVm[vidx]    = mk4(bsdf.worldDirFix(), s.dVM);
                    a.verts.contProb[vidx] = bsdf.contProb;
                }
                if(ic.useVC)
                {
                    // path-major connection record: 64 contiguous bytes, two full sectors
                    const uint32_t l = globalToLocalPath(ic, s.path);
                    float4 *rec = a.verts.conn + (size_t(l) * a.verts.connSlots + s.extra) * 4u;
                    rec[0] = posLen;
                    rec[1] = thrVcm;
                    rec[2] = mk4(bsdf.frame.z, s.dVC);
                    rec[3] = mk4(bsdf.localDirFix, __int_as_float(bsdf.matID));
                    s.extra++;
                    a.pathVertCount[l] = uint8_t(s.extra);
                }
            }
            else atomicAdd(a.counters + CNT_OVERFLOW, 1ull);
        }

        if(alive)
        {
            if(!bsdf.isDelta() && (ic.useVC || ic.lightTraceOnly) && (s.length + 1u >= ic.minLen))
                connectToCamera<kBvh>(sc, ic, s, hitPoint, bsdf, a);

            // terminate if the path would become too long after scattering
            if(s.length + 2u > ic.maxLen) alive = false;
            else alive = sampleScattering<true>(sc, ic, bsdf, hitPoint, s);
            s.length++;
        }

        const uint32_t slot = warpAppend(a.queueCounts + bounce + 1u, alive);
        if(alive) storeSubPath(a.out, slot, s);
    }

    warpCount(a.counters + CNT_LIGHT_RAYS, nRays);
}

// ------------------------------------------------------------------------------------------
// camera pass
// ------------------------------------------------------------------------------------------

// GenerateCameraSample, reference src/vertexcm.hxx:564-606 + Camera::GenerateRay
// (src/camera.hxx:108-117). Returns false if the jittered sample rounds outside the image
// (Framebuffer::AddColor would drop the whole path's colour, src/framebuffer.hxx:47-51).
SVCM_D bool generateCameraSample(const DevScene &sc, const IterConst &ic, SubPath &s)
{
    const uint32_t x = s.path % uint32_t(ic.resX), y = s.path / uint32_t(ic.resX);
    const float jx = s.rng.next(), jy = s.rng.next();
    const float sx = float(x) + jx, sy = float(y) + jy;

    const f3 camPos = ld3(sc.camPos);
    const f3 worldRaster = transformPoint(sc.rasterToWorld, mk3(sx, sy, 0.f));
    const f3 dir = normalize(worldRaster - camPos);

    const float cosAtCamera = dot(ld3(sc.camFwd), dir);
    const float imagePointToCameraDist = sc.imagePlaneDist / cosAtCamera;
    const float cameraPdfW = sqr(imagePointToCameraDist) / cosAtCamera;

    s.origin = camPos; s.direction = dir; s.throughput = mk3(1.f);
    s.length = 1u; s.specular = true; s.finiteLight = false;
    s.dVCM = ic.lightSubPathCount / cameraPdfW;
    s.dVC = 0.f; s.dVM = 0.f;

    if(sx < 0.f || sx >= sc.resX || sy < 0.f || sy >= sc.resY) return false;
    s.extra = uint32_t(int(sx)) + uint32_t(int(sy)) * uint32_t(ic.resX);
    return true;
}

// GetLightRadiance, reference src/vertexcm.hxx:617-658
SVCM_D f3 getLightRadiance(const DevScene &sc, const IterConst &ic, const DevLight &light,
                           const SubPath &s, f3 rayDir)
{
    const float lightPickProb = 1.f / float(sc.numLights);
    float directPdfA, emissionPdfW;
    const f3 radiance = lightRadiance(sc, light, rayDir, directPdfA, emissionPdfW);
    if(iszero(radiance)) return mk3(0.f);
    if(s.length == 1u) return radiance;                       // seen directly: no weighting
    if(ic.useVM && !ic.useVC) return s.specular ? radiance : mk3(0.f);
    directPdfA   *= lightPickProb;
    emissionPdfW *= lightPickProb;
    const float wCamera = directPdfA * s.dVCM + emissionPdfW * s.dVC;   // [tech. rep. (43)]
    return (1.f / (1.f + wCamera)) * radiance;
}

// DirectIllumination, reference src/vertexcm.hxx:663-738
// (the shadow test of :734 is deferred: the caller queues oDirToLight / oDistance)
SVCM_D f3 directIllumination(const DevScene &sc, const IterConst &ic, SubPath &s, f3 hitPoint,
                             const Bsdf &bsdf, f3 &oDirToLight, float &oDistance)
{
    const int   lightCount = sc.numLights;
    const float lightPickProb = 1.f / float(lightCount);
    int lightID = int(s.rng.next() * float(lightCount));
    const float rx = s.rng.next(), ry = s.rng.next();
    lightID = min(lightID, lightCount - 1);
    const DevLight &light = sc.lights[lightID];

    f3 dirToLight = mk3(0.f);
    float distance = 0.f, directPdfW, emissionPdfW, cosAtLight;
    const f3 radiance = lightIlluminate(sc, light, hitPoint, rx, ry, dirToLight, distance,
                                        directPdfW, emissionPdfW, cosAtLight);
    oDirToLight = dirToLight; oDistance = distance;
    if(iszero(radiance)) return mk3(0.f);

    float bsdfDirPdfW, bsdfRevPdfW, cosToLight;
    const f3 bsdfFactor = bsdfEvaluate(sc, bsdf, dirToLight, cosToLight, bsdfDirPdfW, bsdfRevPdfW);
    if(iszero(bsdfFactor)) return mk3(0.f);

    const float contProb = bsdf.contProb;
    bsdfDirPdfW *= lightIsDelta(light.type) ? 0.f : contProb;
    bsdfRevPdfW *= contProb;

    const float wLight  = bsdfDirPdfW / (lightPickProb * directPdfW);                 // (44)
    const float wCamera = (emissionPdfW * cosToLight / (directPdfW * cosAtLight)) *
        (ic.misVmWeightFactor + s.dVCM + s.dVC * bsdfRevPdfW);                         // (45)
    const float misWeight = 1.f / (wLight + 1.f + wCamera);

    return (misWeight * cosToLight / (lightPickProb * directPdfW)) * (radiance * bsdfFactor);
}

// ConnectVertices, reference src/vertexcm.hxx:743-809; the light vertex's BSDF is rebuilt
// from its stored normal / local direction / material (what BSDF<true>::Setup computed).
// (the shadow test of :805 is deferred: the caller queues oDirection / oDistance)
SVCM_D f3 connectVertices(const DevScene &sc, const IterConst &ic, f3 lvPos, const Bsdf &lvBsdf,
                          float lvdVCM, float lvdVC, const Bsdf &camBsdf, f3 camHit,
                          const SubPath &s, f3 &oDirection, float &oDistance)
{
    f3 direction = lvPos - camHit;
    const float dist2 = lensqr(direction);
    const float distance = sqrtf(dist2);
    direction = direction / distance;
    oDirection = direction; oDistance = distance;

    float cosCamera, camDirPdfW, camRevPdfW;
    const f3 camFactor = bsdfEvaluate(sc, camBsdf, direction, cosCamera, camDirPdfW, camRevPdfW);
    if(iszero(camFactor)) return mk3(0.f);
    camDirPdfW *= camBsdf.contProb;
    camRevPdfW *= camBsdf.contProb;

    float cosLight, lightDirPdfW, lightRevPdfW;
    const f3 lightFactor = bsdfEvaluate(sc, lvBsdf, -direction, cosLight, lightDirPdfW, lightRevPdfW);
    if(iszero(lightFactor)) return mk3(0.f);
    lightDirPdfW *= lvBsdf.contProb;
    lightRevPdfW *= lvBsdf.contProb;

    const float geometryTerm = cosLight * cosCamera / dist2;
    if(geometryTerm < 0.f) return mk3(0.f);

    const float camDirPdfA   = pdfWtoA(camDirPdfW, distance, cosLight);
    const float lightDirPdfA = pdfWtoA(lightDirPdfW, distance, cosCamera);

    const float wLight  = camDirPdfA * (ic.misVmWeightFactor + lvdVCM + lvdVC * lightRevPdfW);   // (40)
    const float wCamera = lightDirPdfA * (ic.misVmWeightFactor + s.dVCM + s.dVC * camRevPdfW);    // (41)
    const float misWeight = 1.f / (wLight + 1.f + wCamera);

    return ((misWeight * geometryTerm) * camFactor) * lightFactor;
}

// One bounce of every live camera sub-path: reference src/vertexcm.hxx:423-542 (loop body).
// Merging is deferred: a non-delta vertex inside the light-vertex bbox appends a query
// record that k_merge resolves after the whole camera pass (merging draws no random
// numbers and does not steer the walk, so the split changes nothing but summation order).
template<bool kBvh>
// (three resident blocks per SM: 80 registers with a few spill stores beat 86 registers and two
// blocks — measured 1.73 -> 1.63 ms for the camera pass at 1024^2)
__global__ void __launch_bounds__(SVCM_BLOCK, SVCM_CAMERA_MINBLOCKS)
k_cameraBounce(const __grid_constant__ DevScene sc, const __grid_constant__ IterConst ic,
               const __grid_constant__ TraceArgs a, const uint32_t bounce)
{
    const uint32_t count = (bounce == 1u) ? ic.localPaths : a.queueCounts[bounce];
    uint32_t nRays = 0u;

    for(uint32_t i0 = blockIdx.x * blockDim.x; i0 < count; i0 += gridDim.x * blockDim.x)
    {
        const uint32_t i = i0 + threadIdx.x;
        bool alive = i < count;
        SubPath s;
        if(alive)
        {
            if(bounce == 1u)
            {
                s.path = localToGlobalPath(ic, i);
                s.rng.key = ic.keyCamera; s.rng.path = s.path; s.rng.draw = 0u;
                alive = generateCameraSample(sc, ic, s);
            }
            else loadSubPath(a.in, i, ic.keyCamera, s);
        }

        f3 color = mk3(0.f);
        Hit hit;
        Bsdf bsdf;
        f3 hitPoint = mk3(0.f);

        if(alive)
        {
            const f3 org = mulAddUnfused(s.origin, s.direction, SVCM_EPS_RAY);
            nRays++;
            if(!sceneIntersect<kBvh>(sc, org, s.direction, hit))
            {
                // environment: BEFORE the MIS update (src/vertexcm.hxx:434-447)
                if(sc.background >= 0 && s.length >= ic.minLen)
                    color = color + s.throughput *
                        getLightRadiance(sc, ic, sc.lights[sc.background], s, s.direction);
                alive = false;
            }
            else
            {
                hitPoint = mulAddUnfused(org, s.direction, hit.dist);
                hit.dist += SVCM_EPS_RAY;
                bsdfSetup(sc, s.direction, hit.normal, hit.matID, bsdf);
                alive = bsdf.isValid();
            }
        }

        if(alive)
        {
            s.dVCM *= sqr(hit.dist);
            const float invCos = fabsf(bsdf.cosThetaFix());
            s.dVCM /= invCos; s.dVC /= invCos; s.dVM /= invCos;

            const int lightID = sc.matToLight[hit.matID];
            if(lightID >= 0)
            {
                // hit an emitter: AFTER the MIS update; lights do not reflect (:468-479)
                if(s.length >= ic.minLen)
                    color = color + s.throughput *
                        getLightRadiance(sc, ic, sc.lights[min(lightID, sc.numLights - 1)], s, s.direction);
                alive = false;
            }
            else if(s.length >= ic.maxLen) alive = false;
        }

        // a non-delta vertex is recorded once; vertex connection (:498-526) and merging (:530-538)
        // consume the record later (k_connect, k_merge*) — neither draws random numbers nor
        // steers the walk, so deferring them changes nothing but the order of summation
        bool record = false, mergeable = false;
        uint32_t lpath = 0u;
        if(alive && !bsdf.isDelta())
        {
            if(ic.useVC)
            {
                // vertex connection to a light source: evaluate here, visibility deferred
                if(s.length + 1u >= ic.minLen)
                {
                    f3 dir; float dist;
                    const f3 c = s.throughput * directIllumination(sc, ic, s, hitPoint, bsdf, dir, dist);
                    if(!iszero(c)) emitShadowRay<kBvh>(sc, a.shadow, 1u, a.accum, hitPoint, dir, dist, s.extra, c);
                }
                lpath = globalToLocalPath(ic, s.path);
            }
            // Every non-delta vertex is a merge candidate; HashGrid::Process' early-out for queries
            // outside the particle bbox (src/hashgrid.hxx:116-122) is applied when the queries are
            // keyed for the merge — the camera pass then needs nothing from the grid build (nor
            // from the all-gather before it) and runs concurrently with both.
            mergeable = ic.useVM != 0;
            // vertex connection pairs this vertex with the stored vertices of its own light sub-path
            // (k_makePairs, once both passes are through): possible while a connected path of
            // length + 1 + (>= 1) still fits the limit
            record = mergeable || (ic.useVC && s.length + 2u <= ic.maxLen);
        }

        const uint32_t qidx = warpAppend64(a.counters + CNT_CAM_VERTS, record);
        if(record)
        {
            if(qidx < a.queries.capacity)
            {
                a.queries.q0[qidx] = mk4(hitPoint, __uint_as_float(s.extra));
                a.queries.q1[qidx] = mk4(hit.normal, __uint_as_float(s.length | (uint32_t(hit.matID) << 16) |
                                                                     (mergeable ? CAMV_MERGEABLE : 0u)));
                a.queries.q2[qidx] = mk4(bsdf.localDirFix, s.dVCM);
                a.queries.q3[qidx] = mk4(s.throughput, s.dVM);
                if(ic.useVC) a.queries.q4[qidx] = make_float4(s.dVC, __uint_as_float(lpath), 0.f, 0.f);
            }
            else atomicAdd(a.counters + CNT_OVERFLOW, 1ull);   // dropped: an error at the next synchronisation
        }

        if(alive)
        {
            // PPM merges only at the first non-specular vertex from the camera (:537)
            if(ic.ppm && ic.useVM && !bsdf.isDelta()) alive = false;
            else alive = sampleScattering<false>(sc, ic, bsdf, hitPoint, s);
            s.length++;
        }

        if(!iszero(color)) addPixel(a.accum, s.extra, color);

        const uint32_t slot = warpAppend(a.queueCounts + bounce + 1u, alive);
        if(alive) storeSubPath(a.out, slot, s);
    }

    warpCount(a.counters + CNT_CAMERA_RAYS, nRays);
}

// The pair list of vertex connection (reference src/vertexcm.hxx:498-526: every camera vertex
// against the stored vertices of ITS light sub-path): one lane per recorded camera vertex reserves
// a contiguous block of (vertex, k) pairs. Runs once both passes are through — the camera pass no
// longer reads anything the light pass writes and the two run concurrently.
__global__ void __launch_bounds__(SVCM_BLOCK)
k_makePairs(const __grid_constant__ IterConst ic, const __grid_constant__ TraceArgs a)
{
    const uint32_t count = uint32_t(min(a.counters[CNT_CAM_VERTS], (unsigned long long)a.queries.capacity));
    for(uint32_t i0 = blockIdx.x * blockDim.x; i0 < count; i0 += gridDim.x * blockDim.x)
    {
        const uint32_t i = i0 + threadIdx.x;
        uint32_t nv = 0u;
        if(i < count)
        {
            const uint32_t camLen = __float_as_uint(a.queries.q1[i].w) & 0xFFFFu;
            const uint32_t lpath = __float_as_uint(a.queries.q4[i].y);
            // the k-th stored light vertex has path length >= k+1, so only the first
            // maxLen-1-length of them can pass the window of src/vertexcm.hxx:519-521
            nv = (camLen + 1u < ic.maxLen) ? min(uint32_t(a.pathVertCount[lpath]), ic.maxLen - 1u - camLen) : 0u;
        }
        uint32_t inc = nv;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(FULL_MASK, inc, o);
            if(laneId() >= uint32_t(o)) inc += t;
        }
        const uint32_t total = __shfl_sync(FULL_MASK, inc, 31);
        unsigned long long base = 0ull;
        if(total && laneId() == 31u) base = atomicAdd(a.counters + CNT_PAIRS, (unsigned long long)total);
        base = __shfl_sync(FULL_MASK, base, 31) + inc - nv;
        for(uint32_t k = 0; k < nv; k++)
        {
            if(base + k < (unsigned long long)a.pairs.capacity) a.pairs.pairs[base + k] = make_uint2(i, k);
            else atomicAdd(a.counters + CNT_OVERFLOW, 1ull);
        }
    }
}

// Vertex connection (reference src/vertexcm.hxx:498-526 + ConnectVertices :743-809), one lane
// per (camera vertex, light vertex) pair: both BSDFs are rebuilt from the stored records, the
// connection is evaluated with its MIS weight and, if it carries energy, its visibility
// segment is queued for k_shadow.
template<bool kBvh>
__global__ void SVCM_LB_CONNECT
k_connect(const __grid_constant__ DevScene sc, const __grid_constant__ IterConst ic,
          const __grid_constant__ TraceArgs a)
{
    const uint32_t count = uint32_t(min(a.counters[CNT_PAIRS], (unsigned long long)a.pairs.capacity));
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        const uint2 pr = a.pairs.pairs[i];
        const float4 q0 = a.queries.q0[pr.x], q1 = a.queries.q1[pr.x], q2 = a.queries.q2[pr.x],
                     q3 = a.queries.q3[pr.x], q4 = a.queries.q4[pr.x];
        const uint32_t bits = __float_as_uint(q1.w);
        const uint32_t camLen = bits & 0xFFFFu;
        const uint32_t lpath = __float_as_uint(q4.y);

        // the k-th stored vertex of the camera vertex's own light sub-path: the lanes of one camera
        // vertex read consecutive 64-byte records
        const float4 *rec = a.verts.conn + (size_t(lpath) * a.verts.connSlots + pr.y) * 4u;
        const float4 pl = rec[0];
        const uint32_t lvLen = __float_as_uint(pl.w);
        // path-length window, src/vertexcm.hxx:512-521
        if(lvLen + 1u + camLen < ic.minLen || lvLen + 1u + camLen > ic.maxLen) continue;
        const float4 tv = rec[1];
        const float4 nv4 = rec[2];
        const float4 lm = rec[3];

        Bsdf camBsdf;                      // BSDF<false>::Setup at the camera vertex
        camBsdf.frame.setFromZ(xyz(q1));
        camBsdf.localDirFix = xyz(q2);
        camBsdf.diffProb = camBsdf.phongProb = camBsdf.reflProb = camBsdf.refrProb = 0.f;
        bsdfFromLocal(sc, int32_t((bits >> 16) & 0xFFu), camBsdf);
        Bsdf lvBsdf;                       // BSDF<true>::Setup at the light vertex
        lvBsdf.frame.setFromZ(xyz(nv4));
        lvBsdf.localDirFix = xyz(lm);
        lvBsdf.diffProb = lvBsdf.phongProb = lvBsdf.reflProb = lvBsdf.refrProb = 0.f;
        bsdfFromLocal(sc, __float_as_int(lm.w), lvBsdf);

        SubPath s;                         // the camera state ConnectVertices reads
        s.dVCM = q2.w; s.dVC = q4.x;
        f3 dir; float dist;
        const f3 c = (xyz(q3) * xyz(tv)) *
            connectVertices(sc, ic, xyz(pl), lvBsdf, tv.w, nv4.w, camBsdf, xyz(q0), s, dir, dist);
        if(!iszero(c)) emitShadowRay<kBvh>(sc, a.shadow, 1u, a.accum, xyz(q0), dir, dist, __float_as_uint(q0.w), c);
    }
}

// The deferred Scene::Occluded tests of one pass (reference src/scene.hxx:72-85), one lane per
// queued segment, then Framebuffer::AddColor for the unoccluded ones. Every lane runs the same
// primitive loop: this is where the bounce kernels' divergent any-hit loops went.
template<bool kBvh>
__global__ void SVCM_LB_SHADOW
k_shadow(const __grid_constant__ DevScene sc, const __grid_constant__ TraceArgs a, const uint32_t bounce)
{
    const uint32_t queued = a.shadow.counts[bounce];
    const uint32_t count = min(queued, a.shadow.capacity);
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        const float4 r0 = a.shadow.r0[i], r1 = a.shadow.r1[i];
        if(!sceneOccluded<kBvh>(sc, xyz(r0), xyz(r1), r0.w))
            addPixel(a.accum, __float_as_uint(r1.w), xyz(a.shadow.r2[i]));
    }
    if(blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.counters + CNT_SHADOW_RAYS, (unsigned long long)queued);
}

// ------------------------------------------------------------------------------------------
// the reference's unidirectional integrators on the same device functions
// ------------------------------------------------------------------------------------------

// One bounce of PathTracer::RunIteration's walk (reference src/pathtracer.hxx:68-208): emission
// (environment / area light) with the balance heuristic against next-event estimation, the
// next-event sample itself (visibility deferred to k_shadow) and BSDF sampling with Russian
// roulette. Sub-path slots: s0 = ray origin (already offset) | lastPdfW, s1 = direction | -,
// s2 = pathWeight | -, s3 = pixel index | length + flags (PATH_FLAG_SPECULAR = lastSpecular) |
// rng draw | target pixel.
template<bool kBvh>
__global__ void __launch_bounds__(SVCM_BLOCK)
k_pathTraceBounce(const __grid_constant__ DevScene sc, const __grid_constant__ IterConst ic,
                  const __grid_constant__ TraceArgs a, const uint32_t bounce)
{
    const uint32_t count = (bounce == 1u) ? ic.localPaths : a.queueCounts[bounce];
    const float lightPickProb = 1.f / float(sc.numLights);
    uint32_t nRays = 0u;

    for(uint32_t i0 = blockIdx.x * blockDim.x; i0 < count; i0 += gridDim.x * blockDim.x)
    {
        const uint32_t i = i0 + threadIdx.x;
        bool alive = i < count;
        SubPath s;                       // dVCM doubles as lastPdfW, specular as lastSpecular
        if(alive)
        {
            if(bounce == 1u)
            {
                s.path = localToGlobalPath(ic, i);
                s.rng.key = ic.keyCamera; s.rng.path = s.path; s.rng.draw = 0u;
                // jittered pixel sample + Camera::GenerateRay (:52-57, src/camera.hxx:108-117);
                // the first segment starts AT the camera (no epsilon offset, unlike VertexCM)
                const uint32_t x = s.path % uint32_t(ic.resX), y = s.path / uint32_t(ic.resX);
                const float jx = s.rng.next(), jy = s.rng.next();
                const float sx = float(x) + jx, sy = float(y) + jy;
                const f3 camPos = ld3(sc.camPos);
                s.origin = camPos;
                s.direction = normalize(transformPoint(sc.rasterToWorld, mk3(sx, sy, 0.f)) - camPos);
                s.throughput = mk3(1.f);
                s.length = 1u; s.specular = true; s.finiteLight = false;
                s.dVCM = 1.f; s.dVC = 0.f; s.dVM = 0.f;
                // Framebuffer::AddColor drops samples that round outside the image
                alive = !(sx < 0.f || sx >= sc.resX || sy < 0.f || sy >= sc.resY);
                s.extra = alive ? uint32_t(int(sx)) + uint32_t(int(sy)) * uint32_t(ic.resX) : 0u;
            }
            else loadSubPath(a.in, i, ic.keyCamera, s);
        }

        f3 color = mk3(0.f);
        if(alive)
        {
            Hit hit;
            nRays++;
            if(!sceneIntersect<kBvh>(sc, s.origin, s.direction, hit))
            {
                // environment (:72-93)
                alive = false;
                if(s.length >= ic.minLen && sc.background >= 0)
                {
                    float directPdfW, emissionPdfW;
                    const f3 contrib = lightRadiance(sc, sc.lights[sc.background], s.direction, directPdfW, emissionPdfW);
                    if(!iszero(contrib))
                    {
                        float misWeight = 1.f;
                        if(s.length > 1u && !s.specular)
                            misWeight = s.dVCM / (s.dVCM + directPdfW * lightPickProb);      // Mis2 (:228-233)
                        color = (s.throughput * misWeight) * contrib;
                    }
                }
            }
            else
            {
                const f3 hitPoint = mulAddUnfused(s.origin, s.direction, hit.dist);
                hit.dist += SVCM_EPS_RAY;
                Bsdf bsdf;
                bsdfSetup(sc, s.direction, hit.normal, hit.matID, bsdf);
                alive = bsdf.isValid();
                const int lightID = alive ? sc.matToLight[hit.matID] : -1;
                if(alive && lightID >= 0)
                {
                    // directly hit an emitter; lights do not reflect (:103-125)
                    alive = false;
                    if(s.length >= ic.minLen)
                    {
                        float directPdfA, emissionPdfW;
                        const f3 contrib = lightRadiance(sc, sc.lights[min(lightID, sc.numLights - 1)], s.direction,
                                                         directPdfA, emissionPdfW);
                        if(!iszero(contrib))
                        {
                            float misWeight = 1.f;
                            if(s.length > 1u && !s.specular)
                            {
                                // PdfAtoW, src/utils.hxx:253-259
                                const float directPdfW = directPdfA * sqr(hit.dist) / fabsf(bsdf.cosThetaFix());
                                misWeight = s.dVCM / (s.dVCM + directPdfW * lightPickProb);
                            }
                            color = (s.throughput * misWeight) * contrib;
                        }
                    }
                }
                if(alive && (s.length >= ic.maxLen || bsdf.contProb == 0.f)) alive = false;   // :127-131

                if(alive && !bsdf.isDelta() && s.length + 1u >= ic.minLen)
                {
                    // next event estimation (:134-169); Scene::Occluded deferred to k_shadow
                    int lid = int(s.rng.next() * float(sc.numLights));
                    lid = min(lid, sc.numLights - 1);
                    const DevLight &light = sc.lights[lid];
                    const float rx = s.rng.next(), ry = s.rng.next();
                    f3 dirToLight = mk3(0.f);
                    float distance = 0.f, directPdfW, emissionPdfW, cosAtLight;
                    const f3 radiance = lightIlluminate(sc, light, hitPoint, rx, ry, dirToLight, distance,
                                                        directPdfW, emissionPdfW, cosAtLight);
                    if(!iszero(radiance))
                    {
                        float bsdfPdfW, bsdfRevPdfW, cosThetaOut;
                        const f3 factor = bsdfEvaluate(sc, bsdf, dirToLight, cosThetaOut, bsdfPdfW, bsdfRevPdfW);
                        if(!iszero(factor))
                        {
                            float weight = 1.f;
                            if(!lightIsDelta(light.type))
                            {
                                bsdfPdfW *= bsdf.contProb;
                                const float lp = directPdfW * lightPickProb;
                                weight = lp / (lp + bsdfPdfW);
                            }
                            const f3 contrib = s.throughput *
                                ((weight * cosThetaOut / (lightPickProb * directPdfW)) * (radiance * factor));
                            if(!iszero(contrib))
                                emitShadowRay<kBvh>(sc, a.shadow, 1u, a.accum, hitPoint, dirToLight, distance, s.extra, contrib);
                        }
                    }
                }

                if(alive)
                {
                    // continue the random walk (:172-206)
                    const float r0 = s.rng.next(), r1 = s.rng.next(), r2 = s.rng.next();
                    float pdf, cosThetaOut;
                    uint32_t ev;
                    f3 dirOut = s.direction;
                    const f3 factor = bsdfSample<false>(sc, bsdf, r0, r1, r2, dirOut, pdf, cosThetaOut, ev);
                    if(iszero(factor)) alive = false;
                    else
                    {
                        const float contProb = bsdf.contProb;
                        s.specular = (ev & EV_SPECULAR) != 0u;
                        s.dVCM = pdf * contProb;                          // lastPdfW
                        if(contProb < 1.f)
                        {
                            if(s.rng.next() > contProb) alive = false;
                            pdf *= contProb;
                        }
                        if(alive)
                        {
                            s.throughput = s.throughput * (factor * (cosThetaOut / pdf));
                            s.direction = dirOut;
                            s.origin = mulAddUnfused(hitPoint, dirOut, SVCM_EPS_RAY);
                            s.length++;
                        }
                    }
                }
            }
        }

        if(!iszero(color)) addPixel(a.accum, s.extra, color);
        const uint32_t slot = warpAppend(a.queueCounts + bounce + 1u, alive);
        if(alive) storeSubPath(a.out, slot, s);
    }
    warpCount(a.counters + CNT_CAMERA_RAYS, nRays);
}

// EyeLight::RunIteration, reference src/eyelight.hxx:47-78: N.L preview, back faces in red; the
// iteration with index 1 samples pixel centres
template<bool kBvh>
__global__ void __launch_bounds__(SVCM_BLOCK)
k_eyeLight(const __grid_constant__ DevScene sc, const __grid_constant__ IterConst ic,
           const __grid_constant__ TraceArgs a, const uint32_t iteration)
{
    uint32_t nRays = 0u;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ic.localPaths; i += gridDim.x * blockDim.x)
    {
        const uint32_t path = localToGlobalPath(ic, i);
        PathRng rng; rng.key = ic.keyCamera; rng.path = path; rng.draw = 0u;
        const uint32_t x = path % uint32_t(ic.resX), y = path / uint32_t(ic.resX);
        float jx = 0.5f, jy = 0.5f;
        if(iteration != 1u) { jx = rng.next(); jy = rng.next(); }
        const float sx = float(x) + jx, sy = float(y) + jy;
        if(sx < 0.f || sx >= sc.resX || sy < 0.f || sy >= sc.resY) continue;
        const f3 camPos = ld3(sc.camPos);
        const f3 dir = normalize(transformPoint(sc.rasterToWorld, mk3(sx, sy, 0.f)) - camPos);
        Hit hit;
        nRays++;
        if(sceneIntersect<kBvh>(sc, camPos, dir, hit))
        {
            const float dotLN = dot(hit.normal, -dir);
            const uint32_t pixel = uint32_t(int(sx)) + uint32_t(int(sy)) * uint32_t(ic.resX);
            addPixel(a.accum, pixel, dotLN > 0.f ? mk3(dotLN) : mk3(-dotLN, 0.f, 0.f));
        }
    }
    warpCount(a.counters + CNT_CAMERA_RAYS, nRays);
}

// ------------------------------------------------------------------------------------------
// hash grid build (HashGrid::Build, reference src/hashgrid.hxx:41-107)
// ------------------------------------------------------------------------------------------

// (bounding box of the particles, src/hashgrid.hxx:50-61: k_cellBBox in svcm_grid.cu, launchParticleBBox)

// cell key of every particle + per-cell histogram (reference src/hashgrid.hxx:67-71,189-202)
__global__ void __launch_bounds__(SVCM_BLOCK)
k_gridKeys(const __grid_constant__ MergeSet set, const int32_t *bboxBits, float invCellSize,
           uint32_t numCells, uint32_t *cellKey, uint32_t *cellCount)
{
    const f3 bbMin = mk3(orderedToFloat(bboxBits[0]), orderedToFloat(bboxBits[1]), orderedToFloat(bboxBits[2]));
    const uint32_t total = mergeSetTotal(set);
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    {
        uint32_t key = 0xFFFFFFFFu;
        const MergeSlot ms = mergeSetSlot(set, i);
        if(mergeSetValid(set, ms))
        {
            const float4 p = *mergeSetStream(set, set.posLen, ms);
            const f3 distMin = xyz(p) - bbMin;
            key = cellHash(int32_t(floorf(invCellSize * distMin.x)), int32_t(floorf(invCellSize * distMin.y)),
                           int32_t(floorf(invCellSize * distMin.z)), numCells);
            atomicAdd(cellCount + key, 1u);
        }
        cellKey[i] = key;
    }
}

// exclusive prefix sum over the histogram (reference src/hashgrid.hxx:75-81), three kernels:
// per-tile scan + tile totals, scan of the totals, add-back.
#define SCAN_ITEMS 8
#define SCAN_TILE (SVCM_BLOCK * SCAN_ITEMS)

SVCM_D uint32_t blockExclusiveScan(uint32_t v, uint32_t *smem, uint32_t &total)
{
    // exclusive scan of one value per thread across the block; total = block sum
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(FULL_MASK, inc, o);
        if(lane >= uint32_t(o)) inc += t;
    }
    if(lane == 31u) smem[warp] = inc;
    __syncthreads();
    if(warp == 0u)
    {
        uint32_t w = (lane < (blockDim.x >> 5)) ? smem[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(FULL_MASK, winc, o);
            if(lane >= uint32_t(o)) winc += t;
        }
        smem[lane] = winc - w;            // exclusive warp offsets
        if(lane == 31u) smem[32] = winc;  // block total
    }
    __syncthreads();
    total = smem[32];
    const uint32_t res = smem[warp] + inc - v;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SVCM_BLOCK)
k_scanTiles(uint32_t *data, uint32_t n, uint32_t *tileSums)
{
    __shared__ uint32_t smem[33];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0u;
#pragma unroll
    for(int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? data[base + k] : 0u; sum += v[k]; }
    uint32_t total;
    uint32_t run = blockExclusiveScan(sum, smem, total);
#pragma unroll
    for(int k = 0; k < SCAN_ITEMS; k++) { if(base + k < n) data[base + k] = run; run += v[k]; }
    if(threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SVCM_BLOCK)
k_scanSums(uint32_t *tileSums, uint32_t numTiles)
{
    __shared__ uint32_t smem[33];
    uint32_t carry = 0u;
    for(uint32_t start = 0; start < numTiles; start += SCAN_TILE)
    {
        const uint32_t base = start + threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS], sum = 0u;
#pragma unroll
        for(int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < numTiles) ? tileSums[base + k] : 0u; sum += v[k]; }
        uint32_t total;
        uint32_t run = carry + blockExclusiveScan(sum, smem, total);
#pragma unroll
        for(int k = 0; k < SCAN_ITEMS; k++) { if(base + k < numTiles) tileSums[base + k] = run; run += v[k]; }
        carry += total;
    }
}

__global__ void __launch_bounds__(SVCM_BLOCK)
k_scanAdd(uint32_t *data, uint32_t n, const uint32_t *tileSums)
{
    const uint32_t add = tileSums[blockIdx.x];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for(int k = 0; k < SCAN_ITEMS; k++) if(base + k < n) data[base + k] += add;
}

// scatter into cell order (reference src/hashgrid.hxx:83-88) — but the RECORDS move, not
// just indices, so a range query later streams contiguous 16-byte positions instead of
// gathering 120-byte AoS vertices through mIndices. Like the reference, the cursor IS the
// cell-start array: afterwards cellCount[c] is the END of cell c (mCellEnds).
__global__ void __launch_bounds__(SVCM_BLOCK)
k_gridScatter(const __grid_constant__ MergeSet set, const __grid_constant__ GridBuffers g)
{
    // four records per thread per trip: the cursor atomics return values (300+ cycles each),
    // so they are issued back to back before any dependent store
    const uint32_t total = mergeSetTotal(set);
    for(uint32_t i0 = blockIdx.x * blockDim.x * 4u + threadIdx.x; i0 < total; i0 += gridDim.x * blockDim.x * 4u)
    {
        uint32_t key[4], dst[4];
#pragma unroll
        for(int k = 0; k < 4; k++)
        {
            const uint32_t i = i0 + uint32_t(k) * blockDim.x;
            key[k] = (i < total) ? g.cellKey[i] : 0xFFFFFFFFu;
        }
#pragma unroll
        for(int k = 0; k < 4; k++) dst[k] = (key[k] != 0xFFFFFFFFu) ? atomicAdd(g.cellCount + key[k], 1u) : 0u;
#pragma unroll
        for(int k = 0; k < 4; k++)
        {
            if(key[k] == 0xFFFFFFFFu) continue;
            const uint32_t i = i0 + uint32_t(k) * blockDim.x;
            const MergeSlot ms = mergeSetSlot(set, i);
            g.sPosLen[dst[k]]   = *mergeSetStream(set, set.posLen, ms);
            g.sThrVcm[dst[k]]   = *mergeSetStream(set, set.thrVcm, ms);
            g.sDirVm[dst[k]]    = *mergeSetStream(set, set.dirVm, ms);
            g.sContProb[dst[k]] = *mergeSetStream(set, set.contProb, ms);
            g.sIndex[dst[k]]    = int32_t(i);
        }
    }
}

// ------------------------------------------------------------------------------------------
// query grouping: the merge is a join between cell-sorted particles and range queries; sorting
// the QUERIES by the bucket of their own cell as well makes neighbouring lanes of the merge
// kernel walk identical ranges (their loads coalesce to broadcasts and their trip counts match).
// Pure scheduling: which particles a query visits and accepts is unchanged.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SVCM_BLOCK)
k_queryKeys(const __grid_constant__ QueryStore qs, const unsigned long long *vertexCount, const int32_t *bboxBits,
            float invCellSize, uint32_t numCells, uint32_t *key, uint32_t *cellCount, unsigned long long *counters,
            int spatial)
{
    const uint32_t n = uint32_t(min(*vertexCount, (unsigned long long)qs.capacity));
    const f3 bbMin = mk3(orderedToFloat(bboxBits[0]), orderedToFloat(bboxBits[1]), orderedToFloat(bboxBits[2]));
    const f3 bbMax = mk3(orderedToFloat(bboxBits[3]), orderedToFloat(bboxBits[4]), orderedToFloat(bboxBits[5]));
    // spatial order: 21-bit Morton code of the own cell, coarsened until 7 bits per axis suffice
    const int shift = spatial ? mortonShift(bbMin, bbMax, invCellSize) : 0;
    uint32_t nInside = 0u;
    for(uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x)
    {
        const uint32_t i = i0 + threadIdx.x;
        uint32_t k = 0xFFFFFFFFu;
        if(i < n && (__float_as_uint(qs.q1[i].w) & CAMV_MERGEABLE))
        {
            const float4 p = qs.q0[i];
            // HashGrid::Process early-out (src/hashgrid.hxx:116-122): such a vertex issues no query
            const bool inside = !(p.x - bbMin.x < 0.f || bbMax.x - p.x < 0.f || p.y - bbMin.y < 0.f ||
                                  bbMax.y - p.y < 0.f || p.z - bbMin.z < 0.f || bbMax.z - p.z < 0.f);
            if(inside)
            {
                const int32_t cx = int32_t(floorf(invCellSize * (p.x - bbMin.x))), cy = int32_t(floorf(invCellSize * (p.y - bbMin.y))),
                              cz = int32_t(floorf(invCellSize * (p.z - bbMin.z)));
                if(spatial) k = morton7(uint32_t(cx) >> shift) | (morton7(uint32_t(cy) >> shift) << 1) | (morton7(uint32_t(cz) >> shift) << 2);
                else        k = cellHash(cx, cy, cz, numCells);
                atomicAdd(cellCount + k, 1u);
                nInside++;
            }
        }
        if(i < n) key[i] = k;
    }
    warpCount(counters + CNT_QUERIES, nInside);
}

__global__ void __launch_bounds__(SVCM_BLOCK)
k_queryScatter(const unsigned long long *vertexCount, uint32_t capacity, const uint32_t *key, uint32_t *cellCount,
               uint32_t *order)
{
    const uint32_t n = uint32_t(min(*vertexCount, (unsigned long long)capacity));
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t k = key[i];
        if(k != 0xFFFFFFFFu) order[atomicAdd(cellCount + k, 1u)] = i;
    }
}

// ------------------------------------------------------------------------------------------
// merge: HashGrid::Process (reference src/hashgrid.hxx:110-169) with RangeQuery::Process
// (src/vertexcm.hxx:130-169) as the per-particle body. G lanes cooperate on one query: the
// 8 candidate cells' ranges are fetched by 8 lanes, prefix-summed with shuffles into ONE
// virtual candidate list, and the lanes stride over it — consecutive lanes read consecutive
// 16-byte positions of the cell-sorted array.
// ------------------------------------------------------------------------------------------
template<int G>
__global__ void __launch_bounds__(SVCM_BLOCK)
k_merge(const __grid_constant__ DevScene sc, const __grid_constant__ IterConst ic,
        const __grid_constant__ QueryStore qs, const unsigned long long *queryCountPtr,
        const uint32_t hostQueryCount, const __grid_constant__ GridBuffers g, float *accum,
        float *contribOut, int32_t *countsOut, unsigned long long *counters)
{
    static_assert(G == 8 || G == 16 || G == 32, "group size");
    uint32_t nq = hostQueryCount;
    if(queryCountPtr) nq = uint32_t(min(*queryCountPtr, (unsigned long long)qs.capacity));

    const uint32_t lane = laneId();
    const uint32_t gl = lane & (G - 1);                       // lane within the group
    const uint32_t gbase = lane & ~uint32_t(G - 1);           // first lane of the group
    const uint32_t gmask = (G == 32) ? FULL_MASK : (((1u << G) - 1u) << gbase);
    const uint32_t groupsPerBlock = SVCM_BLOCK / G;
    const uint32_t groupInBlock = threadIdx.x / G;

    const f3 bbMin = mk3(orderedToFloat(g.bboxBits[0]), orderedToFloat(g.bboxBits[1]), orderedToFloat(g.bboxBits[2]));
    const f3 bbMax = mk3(orderedToFloat(g.bboxBits[3]), orderedToFloat(g.bboxBits[4]), orderedToFloat(g.bboxBits[5]));
    uint32_t nCand = 0u, nAcc = 0u;

    for(uint32_t qi = blockIdx.x * groupsPerBlock + groupInBlock; qi < nq; qi += gridDim.x * groupsPerBlock)
    {
        const uint32_t q = qs.order ? qs.order[qi] : qi;
        const float4 q0 = qs.q0[q], q1 = qs.q1[q], q2 = qs.q2[q], q3 = qs.q3[q];
        const f3 qpos = xyz(q0);
        const uint32_t lenMat = __float_as_uint(q1.w);
        if(!(lenMat & CAMV_MERGEABLE)) continue;          // recorded for vertex connection only
        const uint32_t camLen = lenMat & 0xFFFFu;
        const int32_t camMat = int32_t((lenMat >> 16) & 0xFFu);
        const float camdVCM = q2.w, camdVM = q3.w;

        // camera BSDF: BSDF<false>::Setup from the stored normal / local direction / material
        Bsdf bsdf;
        bsdf.frame.setFromZ(xyz(q1));
        bsdf.localDirFix = xyz(q2);
        bsdf.diffProb = bsdf.phongProb = bsdf.reflProb = bsdf.refrProb = 0.f;
        bsdfFromLocal(sc, camMat, bsdf);
        const SvcmMaterial &mat = sc.mats[camMat];

        // the 8 cells: the query's cell and its neighbours on the near side of each axis
        const f3 cellPt = mk3(ic.invCellSize * (qpos.x - bbMin.x), ic.invCellSize * (qpos.y - bbMin.y),
                              ic.invCellSize * (qpos.z - bbMin.z));
        const f3 coordF = mk3(floorf(cellPt.x), floorf(cellPt.y), floorf(cellPt.z));
        const int px = int(coordF.x), py = int(coordF.y), pz = int(coordF.z);
        const int pxo = px + ((cellPt.x - coordF.x) < 0.5f ? -1 : +1);
        const int pyo = py + ((cellPt.y - coordF.y) < 0.5f ? -1 : +1);
        const int pzo = pz + ((cellPt.z - coordF.z) < 0.5f ? -1 : +1);

        // early-out of HashGrid::Process for a query outside the particle bbox
        // (src/hashgrid.hxx:116-122); the camera pass already filters, the kernel-level
        // entry point relies on this test
        const bool inside = !(qpos.x - bbMin.x < 0.f || bbMax.x - qpos.x < 0.f ||
                              qpos.y - bbMin.y < 0.f || bbMax.y - qpos.y < 0.f ||
                              qpos.z - bbMin.z < 0.f || bbMax.z - qpos.z < 0.f);

        uint32_t myStart = 0u, myCount = 0u;
        if(gl < 8u && inside)
        {
            const uint32_t h = cellHash((gl & 4u) ? pxo : px, (gl & 2u) ? pyo : py, (gl & 1u) ? pzo : pz,
                                        ic.numCells);
            myStart = (h == 0u) ? 0u : g.cellCount[h - 1u];   // GetCellRange, src/hashgrid.hxx:173-177
            myCount = g.cellCount[h] - myStart;
        }
        // inclusive prefix over the 8 cell counts, then broadcast all 8 (start, prefix) pairs
        uint32_t inc = myCount;
#pragma unroll
        for(int o = 1; o < 8; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(gmask, inc, o, G);
            if(gl >= uint32_t(o)) inc += t;
        }
        uint32_t st[8], pe[8];                                 // start, inclusive prefix (end)
#pragma unroll
        for(int j = 0; j < 8; j++)
        {
            st[j] = __shfl_sync(gmask, myStart, j, G);
            pe[j] = __shfl_sync(gmask, inc, j, G);
        }
        const uint32_t total = pe[7];

        f3 contrib = mk3(0.f);
        uint32_t accepted = 0u;
        for(uint32_t v = gl; v < total; v += G)
        {
            // which of the 8 ranges holds virtual candidate v
            uint32_t j = 0u;
#pragma unroll
            for(int k = 0; k < 7; k++) j += (v >= pe[k]) ? 1u : 0u;
            uint32_t sj = st[0], bj = 0u;
#pragma unroll
            for(int k = 1; k < 8; k++) if(j == uint32_t(k)) { sj = st[k]; bj = pe[k - 1]; }
            const uint32_t idx = sj + (v - bj);

            const float4 pl = g.sPosLen[idx];
            // |q - p|^2 <= r^2 with the reference's exact (unfused) evaluation order, so the
            // accept decision is bit-identical to src/hashgrid.hxx:162-165 for equal inputs
            const float dx = qpos.x - pl.x, dy = qpos.y - pl.y, dz = qpos.z - pl.z;
            const float distSqr = __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(dx, dx)), __fmul_rn(dy, dy)),
                                            __fmul_rn(dz, dz));
            if(distSqr <= ic.radiusSqr)
            {
                accepted++;
                // RangeQuery::Process, reference src/vertexcm.hxx:130-169
                const uint32_t lvLen = __float_as_uint(pl.w);
                if(!((lvLen + camLen > ic.maxLen) || (lvLen + camLen < ic.minLen)))
                {
                    const float4 dv = g.sDirVm[idx];
                    const f3 localDirGen = bsdf.frame.toLocal(xyz(dv));
                    if(!(localDirGen.z * bsdf.localDirFix.z < 0.f))
                    {
                        float camDirPdfW, camRevPdfW;
                        const f3 f = bsdfEvaluateLocal(mat, bsdf, localDirGen, camDirPdfW, camRevPdfW);
                        if(!iszero(f))
                        {
                            const float4 tv = g.sThrVcm[idx];
                            camDirPdfW *= bsdf.contProb;
                            camRevPdfW *= g.sContProb[idx];      // the LIGHT vertex's RR probability
                            const float wLight  = tv.w * ic.misVcWeightFactor + dv.w * camDirPdfW;       // (38)
                            const float wCamera = camdVCM * ic.misVcWeightFactor + camdVM * camRevPdfW;  // (39)
                            const float misWeight = ic.ppm ? 1.f : 1.f / (wLight + 1.f + wCamera);
                            contrib = contrib + (misWeight * f) * xyz(tv);
                        }
                    }
                }
            }
        }

        // reduce over the group
#pragma unroll
        for(int o = G / 2; o > 0; o >>= 1)
        {
            contrib.x += __shfl_xor_sync(gmask, contrib.x, o, G);
            contrib.y += __shfl_xor_sync(gmask, contrib.y, o, G);
            contrib.z += __shfl_xor_sync(gmask, contrib.z, o, G);
            accepted  += __shfl_xor_sync(gmask, accepted, o, G);
        }
        if(gl == 0u)
        {
            nCand += total; nAcc += accepted;
            if(contribOut)
            {
                contribOut[3ull * q + 0] = contrib.x; contribOut[3ull * q + 1] = contrib.y;
                contribOut[3ull * q + 2] = contrib.z;
            }
            if(countsOut) { countsOut[2ull * q] = int32_t(total); countsOut[2ull * q + 1] = int32_t(accepted); }
            // color += cameraState.mThroughput * mVmNormalization * query.GetContrib()  (:534)
            if(accum) addPixel(accum, __float_as_uint(q0.w), (xyz(q3) * ic.vmNormalization) * contrib);
        }
    }

    if(counters)
    {
        warpCount(counters + CNT_CANDIDATES, nCand);
        warpCount(counters + CNT_ACCEPTED, nAcc);
    }
}

// ------------------------------------------------------------------------------------------
// merge, lane-per-query variant (the default): every lane owns one range query and walks its 8
// cell ranges one candidate at a time; the accepted (query, particle) pairs — about one in
// eight — are compacted through a per-warp shared-memory queue with ballots, and each time 32
// pairs are waiting the whole warp evaluates RangeQuery::Process for them, one pair per lane.
// The distance test thus runs with whatever lanes still have candidates, and the expensive
// BSDF + MIS body always runs with a full warp, whatever the acceptance rate.
// ------------------------------------------------------------------------------------------
#define MERGE_WARPS (SVCM_BLOCK / 32)
#define MERGE_QWORDS 20     // per-query words parked in shared memory for the pair evaluation

struct MergeWarpSmem
{
    uint32_t queue[64];                 // ring of accepted pairs: sorted particle index ...
    uint8_t  queueOwner[64];            // ... and the lane whose query accepted it
    float    q[MERGE_QWORDS][32];       // [word][owner lane]: conflict-free for distinct owners
    float    acc[3][32];                // per-owner radiance sums
    uint32_t rangeStart[8][32];
    uint32_t rangeEnd[8][32];
};

SVCM_D void mergeEvaluatePairs(const DevScene &sc, const IterConst &ic, const GridBuffers &g,
                               MergeWarpSmem &w, uint32_t head, uint32_t tail)
{
    const uint32_t lane = laneId();
    if(head + lane < tail)
    {
        const uint32_t idx = w.queue[(head + lane) & 63u], owner = w.queueOwner[(head + lane) & 63u];
        // the particle's whole payload at once: one memory round trip per step instead of three
        // dependent ones (length -> direction -> throughput)
        const uint32_t lvLen = __float_as_uint(reinterpret_cast<const float*>(g.sPosLen + idx)[3]);
        const float4 dv = g.sDirVm[idx], tv = g.sThrVcm[idx];
        const float lightContProb = g.sContProb[idx];
        const uint32_t lenMat = __float_as_uint(w.q[12][owner]);
        const uint32_t camLen = lenMat & 0xFFFFu;
        const uint32_t camMat = (lenMat >> 16) & 0xFFu;
        // RangeQuery::Process, reference src/vertexcm.hxx:130-169
        if(!((lvLen + camLen > ic.maxLen) || (lvLen + camLen < ic.minLen)))
        {
            Bsdf b;
            b.frame.x = mk3(w.q[0][owner], w.q[1][owner], w.q[2][owner]);
            b.frame.y = mk3(w.q[3][owner], w.q[4][owner], w.q[5][owner]);
            b.frame.z = mk3(w.q[6][owner], w.q[7][owner], w.q[8][owner]);
            b.localDirFix = mk3(w.q[9][owner], w.q[10][owner], w.q[11][owner]);
            const f3 localDirGen = b.frame.toLocal(xyz(dv));
            if(!(localDirGen.z * b.localDirFix.z < 0.f))
            {
                b.diffProb = w.q[13][owner]; b.phongProb = w.q[14][owner]; b.contProb = w.q[15][owner];
                float camDirPdfW, camRevPdfW;
                const f3 f = bsdfEvaluateLocal(sc.mats[camMat], b, localDirGen, camDirPdfW, camRevPdfW);
                if(!iszero(f))
                {
                    camDirPdfW *= b.contProb;
                    camRevPdfW *= lightContProb;                       // the LIGHT vertex's RR probability
                    const float wLight  = tv.w * ic.misVcWeightFactor + dv.w * camDirPdfW;                     // (38)
                    const float wCamera = w.q[16][owner] * ic.misVcWeightFactor + w.q[17][owner] * camRevPdfW; // (39)
                    const float misWeight = ic.ppm ? 1.f : 1.f / (wLight + 1.f + wCamera);
                    const f3 c = (misWeight * f) * xyz(tv);
                    atomicAdd(&w.acc[0][owner], c.x);
                    atomicAdd(&w.acc[1][owner], c.y);
                    atomicAdd(&w.acc[2][owner], c.z);
                }
            }
        }
    }
    __syncwarp();
}

__global__ void SVCM_LB_MERGE
k_mergeLane(const __grid_constant__ DevScene sc, const __grid_constant__ IterConst ic,
            const __grid_constant__ QueryStore qs, const unsigned long long *queryCountPtr,
            const uint32_t hostQueryCount, const __grid_constant__ GridBuffers g, float *accum,
            float *contribOut, int32_t *countsOut, unsigned long long *counters)
{
    __shared__ MergeWarpSmem smem[MERGE_WARPS];
    MergeWarpSmem &w = smem[threadIdx.x >> 5];
    uint32_t nq = hostQueryCount;
    if(queryCountPtr) nq = uint32_t(min(*queryCountPtr, (unsigned long long)qs.capacity));
    const uint32_t lane = laneId();
    const uint32_t ltMask = (1u << lane) - 1u;

    const f3 bbMin = mk3(orderedToFloat(g.bboxBits[0]), orderedToFloat(g.bboxBits[1]), orderedToFloat(g.bboxBits[2]));
    const f3 bbMax = mk3(orderedToFloat(g.bboxBits[3]), orderedToFloat(g.bboxBits[4]), orderedToFloat(g.bboxBits[5]));
    uint32_t nCand = 0u, nAcc = 0u;

    for(uint32_t q0i = (blockIdx.x * MERGE_WARPS + (threadIdx.x >> 5)) * 32u; q0i < nq;
        q0i += gridDim.x * MERGE_WARPS * 32u)
    {
        bool valid = q0i + lane < nq;
        const uint32_t q = valid ? (qs.order ? qs.order[q0i + lane] : q0i + lane) : 0u;
        f3 qpos = mk3(0.f), scale = mk3(0.f);
        uint32_t pixel = 0u;
        uint32_t cand = 0u, accepted = 0u;
        w.acc[0][lane] = 0.f; w.acc[1][lane] = 0.f; w.acc[2][lane] = 0.f;
#pragma unroll
        for(int j = 0; j < 8; j++) { w.rangeStart[j][lane] = 0u; w.rangeEnd[j][lane] = 0u; }

        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;
        if(valid)
        {
            // the four words of the record at once (one memory round trip)
            q0 = qs.q0[q]; q1 = qs.q1[q]; q2 = qs.q2[q]; q3 = qs.q3[q];
            valid = (__float_as_uint(q1.w) & CAMV_MERGEABLE) != 0u;   // else: vertex connection only
        }
        if(valid)
        {
            qpos = xyz(q0); scale = xyz(q3) * ic.vmNormalization; pixel = __float_as_uint(q0.w);
            const uint32_t lenMat = __float_as_uint(q1.w);
            // camera BSDF: BSDF<false>::Setup from the stored normal / local direction / material
            Bsdf bsdf;
            bsdf.frame.setFromZ(xyz(q1));
            bsdf.localDirFix = xyz(q2);
            bsdf.diffProb = bsdf.phongProb = bsdf.reflProb = bsdf.refrProb = 0.f;
            bsdfFromLocal(sc, int32_t((lenMat >> 16) & 0xFFu), bsdf);
            w.q[0][lane] = bsdf.frame.x.x; w.q[1][lane] = bsdf.frame.x.y; w.q[2][lane] = bsdf.frame.x.z;
            w.q[3][lane] = bsdf.frame.y.x; w.q[4][lane] = bsdf.frame.y.y; w.q[5][lane] = bsdf.frame.y.z;
            w.q[6][lane] = bsdf.frame.z.x; w.q[7][lane] = bsdf.frame.z.y; w.q[8][lane] = bsdf.frame.z.z;
            w.q[9][lane] = bsdf.localDirFix.x; w.q[10][lane] = bsdf.localDirFix.y; w.q[11][lane] = bsdf.localDirFix.z;
            w.q[12][lane] = q1.w;
            w.q[13][lane] = bsdf.diffProb; w.q[14][lane] = bsdf.phongProb; w.q[15][lane] = bsdf.contProb;
            w.q[16][lane] = q2.w; w.q[17][lane] = q3.w;

            // HashGrid::Process: bbox early-out, then the 8 cells (src/hashgrid.hxx:116-155)
            const bool inside = !(qpos.x - bbMin.x < 0.f || bbMax.x - qpos.x < 0.f ||
                                  qpos.y - bbMin.y < 0.f || bbMax.y - qpos.y < 0.f ||
                                  qpos.z - bbMin.z < 0.f || bbMax.z - qpos.z < 0.f);
            if(inside)
            {
                const f3 cellPt = mk3(ic.invCellSize * (qpos.x - bbMin.x), ic.invCellSize * (qpos.y - bbMin.y),
                                      ic.invCellSize * (qpos.z - bbMin.z));
                const f3 coordF = mk3(floorf(cellPt.x), floorf(cellPt.y), floorf(cellPt.z));
                const int px = int(coordF.x), py = int(coordF.y), pz = int(coordF.z);
                const int pxo = px + ((cellPt.x - coordF.x) < 0.5f ? -1 : +1);
                const int pyo = py + ((cellPt.y - coordF.y) < 0.5f ? -1 : +1);
                const int pzo = pz + ((cellPt.z - coordF.z) < 0.5f ? -1 : +1);
#pragma unroll
                for(int j = 0; j < 8; j++)
                {
                    const uint32_t h = refHash((j & 4) ? pxo : px, (j & 2) ? pyo : py, (j & 1) ? pzo : pz, ic);
                    const uint32_t start = (h == 0u) ? 0u : g.cellCount[h - 1u];   // GetCellRange
                    const uint32_t end = g.cellCount[h];
                    w.rangeStart[j][lane] = start; w.rangeEnd[j][lane] = end;
                    cand += end - start;
                }
            }
        }
        __syncwarp();

        // walk the ranges; compact accepted pairs through the warp queue
        uint32_t j = 0u, cur = w.rangeStart[0][lane], end = w.rangeEnd[0][lane];
        uint32_t head = 0u, tail = 0u;
        // the position of the NEXT candidate is requested one step ahead: its latency overlaps the
        // test, the compaction and the pair evaluations of the current step
        while(cur == end && j < 7u) { j++; cur = w.rangeStart[j][lane]; end = w.rangeEnd[j][lane]; }
        bool has = cur < end;
        float4 plNext = make_float4(0.f, 0.f, 0.f, 0.f);
        if(has) plNext = g.sPosLen[cur];
        for(;;)
        {
            if(!__any_sync(FULL_MASK, has)) break;
            bool hitIt = false;
            uint32_t idx = 0u;
            const float4 pl = plNext;
            const bool hasNow = has;
            if(has)
            {
                idx = cur++;
                while(cur == end && j < 7u) { j++; cur = w.rangeStart[j][lane]; end = w.rangeEnd[j][lane]; }
                has = cur < end;
                if(has) plNext = g.sPosLen[cur];
            }
            if(hasNow)
            {
                // |q - p|^2 <= r^2 in the reference's exact (unfused) evaluation order
                // (src/hashgrid.hxx:162-165): identical accept decisions for identical inputs
                const float dx = qpos.x - pl.x, dy = qpos.y - pl.y, dz = qpos.z - pl.z;
                const float distSqr = __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(dx, dx)), __fmul_rn(dy, dy)),
                                                __fmul_rn(dz, dz));
                hitIt = distSqr <= ic.radiusSqr;
            }
            const uint32_t m = __ballot_sync(FULL_MASK, hitIt);
            if(m)
            {
                if(hitIt)
                {
                    const uint32_t slot = (tail + uint32_t(__popc(m & ltMask))) & 63u;
                    w.queue[slot] = idx; w.queueOwner[slot] = uint8_t(lane);
                    accepted++;
                }
                tail += uint32_t(__popc(m));
                __syncwarp();
                if(tail - head >= 32u)
                {
                    mergeEvaluatePairs(sc, ic, g, w, head, head + 32u);
                    head += 32u;
                }
            }
        }
        if(tail != head) mergeEvaluatePairs(sc, ic, g, w, head, tail);
        __syncwarp();

        if(valid)
        {
            const f3 contrib = mk3(w.acc[0][lane], w.acc[1][lane], w.acc[2][lane]);
            nCand += cand; nAcc += accepted;
            if(contribOut)
            {
                contribOut[3ull * q + 0] = contrib.x; contribOut[3ull * q + 1] = contrib.y;
                contribOut[3ull * q + 2] = contrib.z;
            }
            if(countsOut) { countsOut[2ull * q] = int32_t(cand); countsOut[2ull * q + 1] = int32_t(accepted); }
            // color += cameraState.mThroughput * mVmNormalization * query.GetContrib()  (:534)
            if(accum) addPixel(accum, pixel, scale * contrib);
        }
        __syncwarp();
    }

    if(counters)
    {
        warpCount(counters + CNT_CANDIDATES, nCand);
        warpCount(counters + CNT_ACCEPTED, nAcc);
    }
}

// ------------------------------------------------------------------------------------------
// kernel-level batch kernels (parity tests call the same device functions as the path)
// ------------------------------------------------------------------------------------------
template<bool kBvh>
__global__ void k_intersectBatch(const __grid_constant__ DevScene sc, const float *rays, int64_t count,
                                 int32_t *hitOut, float *dist, float *normal, int32_t *matID, int32_t *lightID)
{
    for(int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    {
        Hit hit;
        const bool found = sceneIntersect<kBvh>(sc, ld3(rays + 6 * i), ld3(rays + 6 * i + 3), hit);
        hitOut[i] = found ? hit.prim : -1;
        dist[i] = found ? hit.dist : 1e36f;
        normal[3 * i + 0] = found ? hit.normal.x : 0.f;
        normal[3 * i + 1] = found ? hit.normal.y : 0.f;
        normal[3 * i + 2] = found ? hit.normal.z : 0.f;
        matID[i] = found ? hit.matID : -1;
        lightID[i] = found ? sc.matToLight[hit.matID] : -1;
    }
}

template<bool kBvh>
__global__ void k_occludedBatch(const __grid_constant__ DevScene sc, const float *seg, int64_t count, int32_t *occ)
{
    for(int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
        occ[i] = sceneOccluded<kBvh>(sc, ld3(seg + 7 * i), ld3(seg + 7 * i + 3), seg[7 * i + 6]) ? 1 : 0;
}

__global__ void k_bsdfEvaluateBatch(const __grid_constant__ DevScene sc, const float *in, int64_t count, float *out)
{
    for(int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    {
        const float *p = in + 10 * i; float *r = out + 8 * i;
        Bsdf b;
        bsdfSetup(sc, -ld3(p + 3), ld3(p), __float_as_int(p[9]), b);
        for(int k = 0; k < 8; k++) r[k] = 0.f;
        if(!b.isValid()) continue;
        float cosGen, dirPdf, revPdf;
        const f3 f = bsdfEvaluate(sc, b, ld3(p + 6), cosGen, dirPdf, revPdf);
        r[0] = f.x; r[1] = f.y; r[2] = f.z;
        r[3] = iszero(f) ? 0.f : cosGen;
        r[4] = dirPdf; r[5] = revPdf; r[6] = b.contProb;
        r[7] = float(1 | (b.isDelta() ? 2 : 0));
    }
}

template<bool kLight>
__global__ void k_bsdfSampleBatch(const __grid_constant__ DevScene sc, const float *in, int64_t count,
                                  float *out, uint32_t *events)
{
    for(int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    {
        const float *p = in + 10 * i; float *r = out + 8 * i;
        Bsdf b;
        bsdfSetup(sc, -ld3(p + 3), ld3(p), __float_as_int(p[9]), b);
        for(int k = 0; k < 8; k++) r[k] = 0.f;
        events[i] = 0u;
        if(!b.isValid()) continue;
        f3 dirGen = mk3(0.f); float pdfW, cosGen; uint32_t ev;
        const f3 f = bsdfSample<kLight>(sc, b, p[6], p[7], p[8], dirGen, pdfW, cosGen, ev);
        if(iszero(f)) continue;
        r[0] = f.x; r[1] = f.y; r[2] = f.z; r[3] = dirGen.x; r[4] = dirGen.y; r[5] = dirGen.z;
        r[6] = pdfW; r[7] = cosGen;
        events[i] = ev;
    }
}

// AbstractLight::Emit / Illuminate / GetRadiance of light `lightIndex` on `count` inputs of 8
// floats, 12 floats out (layouts: include/svcm.h, svcm_light_*_batch)
__global__ void k_lightBatch(const __grid_constant__ DevScene sc, int op, int lightIndex, const float *in, int64_t count,
                             float *out)
{
    const DevLight &l = sc.lights[lightIndex];
    for(int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    {
        const float *p = in + 8 * i; float *r = out + 12 * i;
        for(int k = 0; k < 12; k++) r[k] = 0.f;
        if(op == 0)
        {
            f3 pos = mk3(0.f), dir = mk3(0.f); float emissionPdfW = 0.f, directPdfA = 0.f, cosLight = 0.f;
            const f3 e = lightEmit(sc, l, p[0], p[1], p[2], p[3], pos, dir, emissionPdfW, directPdfA, cosLight);
            r[0] = e.x; r[1] = e.y; r[2] = e.z; r[3] = pos.x; r[4] = pos.y; r[5] = pos.z;
            r[6] = dir.x; r[7] = dir.y; r[8] = dir.z; r[9] = emissionPdfW; r[10] = directPdfA; r[11] = cosLight;
        }
        else if(op == 1)
        {
            f3 dir = mk3(0.f); float dist = 0.f, directPdfW = 0.f, emissionPdfW = 0.f, cosAtLight = 0.f;
            const f3 e = lightIlluminate(sc, l, ld3(p), p[3], p[4], dir, dist, directPdfW, emissionPdfW, cosAtLight);
            r[0] = e.x; r[1] = e.y; r[2] = e.z; r[3] = dir.x; r[4] = dir.y; r[5] = dir.z; r[6] = dist;
            // the pdfs are only defined when radiance comes back (src/lights.hxx:150-152 returns early)
            if(!iszero(e)) { r[7] = directPdfW; r[8] = emissionPdfW; r[9] = cosAtLight; }
        }
        else
        {
            float directPdfA = 0.f, emissionPdfW = 0.f;
            const f3 e = lightRadiance(sc, l, ld3(p), directPdfA, emissionPdfW);
            r[0] = e.x; r[1] = e.y; r[2] = e.z; r[3] = directPdfA; r[4] = emissionPdfW;
        }
    }
}

// Camera::GenerateRay (src/camera.hxx:108-117) of a raster position and Camera::WorldToRaster +
// CheckRaster (:95-105) of a world point: 5 floats in, 6 out
__global__ void k_cameraBatch(const __grid_constant__ DevScene sc, const float *in, int64_t count, float *out)
{
    for(int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    {
        const float *p = in + 5 * i; float *r = out + 6 * i;
        const f3 camPos = ld3(sc.camPos);
        const f3 dir = normalize(transformPoint(sc.rasterToWorld, mk3(p[0], p[1], 0.f)) - camPos);
        const f3 raster = transformPoint(sc.worldToRaster, ld3(p + 2));
        r[0] = dir.x; r[1] = dir.y; r[2] = dir.z; r[3] = raster.x; r[4] = raster.y;
        r[5] = (raster.x >= 0.f && raster.y >= 0.f && raster.x < sc.resX && raster.y < sc.resY) ? 1.f : 0.f;
    }
}

__global__ void k_rngStream(uint32_t key, uint32_t firstPath, int paths, int draws, float *out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= paths) return;
    PathRng rng; rng.key = key; rng.path = firstPath + uint32_t(p); rng.draw = 0u;
    for(int k = 0; k < draws; k++) out[size_t(p) * draws + k] = rng.next();
}

__global__ void k_packMergeVertices(const SvcmMergeVertex *v, int64_t count, float4 *posLen, float4 *thrVcm,
                                    float4 *dirVm, float *contProb)
{
    for(int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    {
        posLen[i] = make_float4(v[i].position[0], v[i].position[1], v[i].position[2], __uint_as_float(v[i].pathLength));
        thrVcm[i] = make_float4(v[i].throughput[0], v[i].throughput[1], v[i].throughput[2], v[i].dVCM);
        dirVm[i]  = make_float4(v[i].dirFix[0], v[i].dirFix[1], v[i].dirFix[2], v[i].dVM);
        contProb[i] = v[i].contProb;
    }
}

__global__ void k_packMergeQueries(const __grid_constant__ DevScene sc, const SvcmMergeQuery *q, int64_t count,
                                   const __grid_constant__ QueryStore store)
{
    for(int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    {
        // BSDF<false>::Setup exactly as the camera pass does it at a hit (ray.dir = -dirFix)
        Bsdf b;
        bsdfSetup(sc, -ld3(q[i].dirFix), ld3(q[i].normal), q[i].matID, b);
        const bool usable = b.isValid() && !b.isDelta();
        // an unusable vertex never issues a query in the path; here it must yield zero:
        // park it outside any bbox
        const f3 pos = usable ? ld3(q[i].position) : mk3(3e37f);
        store.q0[i] = mk4(pos, __uint_as_float(uint32_t(i)));
        store.q1[i] = mk4(ld3(q[i].normal), __uint_as_float(q[i].pathLength | (uint32_t(q[i].matID) << 16) |
                                                             CAMV_MERGEABLE));
        store.q2[i] = mk4(b.localDirFix, q[i].dVCM);
        store.q3[i] = mk4(mk3(1.f), q[i].dVM);
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static int g_numSMs = 0;
int numSMs()
{
    if(g_numSMs == 0)
    {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_numSMs, cudaDevAttrMultiProcessorCount, dev);
        if(g_numSMs <= 0) g_numSMs = 148;
    }
    return g_numSMs;
}
// persistent-style grid: a multiple of the SM count, capped by the work available
// (blocksPerSM <= 0: one block per perBlock items, no grid-stride loop)
uint32_t gridFor(uint64_t items, int perBlock, int blocksPerSM)
{
    const uint64_t need = (items + uint64_t(perBlock) - 1) / uint64_t(perBlock);
    const uint64_t cap = blocksPerSM > 0 ? uint64_t(numSMs()) * uint64_t(blocksPerSM) : 0x7FFFFFFFull;
    return uint32_t(need < 1 ? 1 : (need < cap ? need : cap));
}
// blocks per SM of the hot kernels' grids; SVCM_GRID_<KERNEL>=n overrides the default (A/B)
static int gridBlocksPerSM(const char *env, int fallback)
{
    const char *v = getenv(env);
    return v ? atoi(v) : fallback;
}
#define SVCM_GRID_KNOB(name, env, fallback) static int name() { static const int v = gridBlocksPerSM(env, fallback); return v; }
SVCM_GRID_KNOB(gridLight,   "SVCM_GRID_LIGHT",   16)     // 16 (or no cap) vs 8: light pass alone 0.761 -> 0.739 ms at 1024^2
SVCM_GRID_KNOB(gridCamera,  "SVCM_GRID_CAMERA",  16)     // camera pass alone 1.324 -> 1.286 ms (profiles/r2_variants_grid_sizes.jsonl)
SVCM_GRID_KNOB(gridConnect, "SVCM_GRID_CONNECT", 8)
SVCM_GRID_KNOB(gridShadow,  "SVCM_GRID_SHADOW",  10)
SVCM_GRID_KNOB(gridMerge,   "SVCM_GRID_MERGE",   4)

void launchLightBounce(const DevScene &sc, const IterConst &ic, const TraceArgs &args, uint32_t bounce,
                       uint32_t maxItems, cudaStream_t stream)
{
    if(sc.bvh) k_lightBounce<true><<<gridFor(maxItems, SVCM_BLOCK, gridLight()), SVCM_BLOCK, 0, stream>>>(sc, ic, args, bounce);
    else       k_lightBounce<false><<<gridFor(maxItems, SVCM_BLOCK, gridLight()), SVCM_BLOCK, 0, stream>>>(sc, ic, args, bounce);
}

void launchPathTraceBounce(const DevScene &sc, const IterConst &ic, const TraceArgs &args, uint32_t bounce,
                           uint32_t maxItems, cudaStream_t stream)
{
    if(sc.bvh) k_pathTraceBounce<true><<<gridFor(maxItems, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, ic, args, bounce);
    else       k_pathTraceBounce<false><<<gridFor(maxItems, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, ic, args, bounce);
}

void launchEyeLight(const DevScene &sc, const IterConst &ic, const TraceArgs &args, uint32_t iteration,
                    cudaStream_t stream)
{
    if(sc.bvh) k_eyeLight<true><<<gridFor(ic.localPaths, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, ic, args, iteration);
    else       k_eyeLight<false><<<gridFor(ic.localPaths, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, ic, args, iteration);
}

void launchCameraBounce(const DevScene &sc, const IterConst &ic, const TraceArgs &args, uint32_t bounce,
                        uint32_t maxItems, cudaStream_t stream)
{
    if(sc.bvh) k_cameraBounce<true><<<gridFor(maxItems, SVCM_BLOCK, gridCamera()), SVCM_BLOCK, 0, stream>>>(sc, ic, args, bounce);
    else       k_cameraBounce<false><<<gridFor(maxItems, SVCM_BLOCK, gridCamera()), SVCM_BLOCK, 0, stream>>>(sc, ic, args, bounce);
}

int launchQuerySort(const IterConst &ic, const QueryStore &queries, const unsigned long long *vertexCount,
                    const int32_t *bboxBits, const QuerySortBuffers &qs, unsigned long long *counters, ScanGen *gen,
                    int spatial, cudaStream_t stream)
{
    // order of the queries: by the hash bucket of their own cell (round 1), or — spatial — by a
    // Morton code of the cell, so that consecutive warps of the merge work on neighbouring cells and
    // find the buckets they share still in the L2
    spatial = spatial && qs.cellEnd != NULL;
    const uint32_t n = spatial ? (1u << 21) + 1u : ic.numCells + 1u;
    uint32_t *table = spatial ? qs.cellEnd : qs.cellCount;
    cudaMemsetAsync(table, 0, sizeof(uint32_t) * n, stream);
    const uint32_t gridN = gridFor(queries.capacity, SVCM_BLOCK, 8);
    k_queryKeys<<<gridN, SVCM_BLOCK, 0, stream>>>(queries, vertexCount, bboxBits, ic.invCellSize, ic.numCells,
                                                  qs.key, table, counters, spatial);
    int launches = 3;
    if(gen && qs.scanState && (n + 4095u) / 4096u <= kScanStateWords)
        launchLookbackScan(table, NULL, n, qs.scanState, qs.scanCounter, *gen, stream);   // single pass
    else
    {
        const uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
        k_scanTiles<<<tiles, SVCM_BLOCK, 0, stream>>>(table, n, qs.blockSums);
        k_scanSums<<<1, SVCM_BLOCK, 0, stream>>>(qs.blockSums, tiles);
        k_scanAdd<<<tiles, SVCM_BLOCK, 0, stream>>>(table, n, qs.blockSums);
        launches += 2;
    }
    k_queryScatter<<<gridN, SVCM_BLOCK, 0, stream>>>(vertexCount, queries.capacity, qs.key, table, qs.order);
    return launches;
}

void launchConnect(const DevScene &sc, const IterConst &ic, const TraceArgs &args, cudaStream_t stream)
{
    k_makePairs<<<gridFor(args.queries.capacity, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(ic, args);
    if(sc.bvh) k_connect<true><<<gridFor(args.pairs.capacity, SVCM_BLOCK, gridConnect()), SVCM_BLOCK, 0, stream>>>(sc, ic, args);
    else       k_connect<false><<<gridFor(args.pairs.capacity, SVCM_BLOCK, gridConnect()), SVCM_BLOCK, 0, stream>>>(sc, ic, args);
}

void launchShadow(const DevScene &sc, const IterConst &ic, const TraceArgs &args, uint32_t which,
                  cudaStream_t stream)
{
    (void)ic;
    if(sc.bvh) k_shadow<true><<<gridFor(args.shadow.capacity, SVCM_BLOCK, gridShadow()), SVCM_BLOCK, 0, stream>>>(sc, args, which);
    else       k_shadow<false><<<gridFor(args.shadow.capacity, SVCM_BLOCK, gridShadow()), SVCM_BLOCK, 0, stream>>>(sc, args, which);
}

// HashGrid::Build in the reference's bucket order: bbox -> keys + histogram -> exclusive scan -> reorder.
// With `gen` (scan state allocated): single-pass look-back scan and the two-step reorder (4-byte
// index scatter + dense record gather); without: the three-kernel scan and the direct record scatter.
int launchGridBuild(const IterConst &ic, const MergeSet &set, const GridBuffers &grid, ScanGen *gen, cudaStream_t stream)
{
    const uint32_t n = ic.numCells + 1u;
    // bbox init: (+1e36, -1e36) like the reference (src/hashgrid.hxx:50-51); for all-gathered
    // blocks the bbox is already there (union of the block headers, launchGatherUnpackHeaders)
    const bool bboxReady = set.singleCount == NULL;
    if(!bboxReady) cudaMemcpyAsync(grid.bboxBits, grid.bboxInit, sizeof(int32_t) * 6, cudaMemcpyDeviceToDevice, stream);
    cudaMemsetAsync(grid.cellCount, 0, sizeof(uint32_t) * n, stream);
    int launches = 0;
    const uint32_t maxTotal = set.singleCount ? set.capacity : set.numSegs * set.segStride;
    if(maxTotal > 0)
    {
        const uint32_t gridN = gridFor(maxTotal, SVCM_BLOCK, 8);
        if(!bboxReady) launchParticleBBox(set, grid.bboxBits, stream);
        k_gridKeys<<<gridN, SVCM_BLOCK, 0, stream>>>(set, grid.bboxBits, ic.invCellSize, ic.numCells,
                                                     grid.cellKey, grid.cellCount);
        launches += 2;
    }
    if(gen && grid.scanState && (n + 4095u) / 4096u <= kScanStateWords)
    {
        launchLookbackScan(grid.cellCount, NULL, n, grid.scanState, grid.scanCounter, *gen, stream);
        launches += 1 + launchBucketReorder(set, grid, ic.numCells, stream);
        return launches;
    }
    const uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scanTiles<<<tiles, SVCM_BLOCK, 0, stream>>>(grid.cellCount, n, grid.blockSums);
    k_scanSums<<<1, SVCM_BLOCK, 0, stream>>>(grid.blockSums, tiles);
    k_scanAdd<<<tiles, SVCM_BLOCK, 0, stream>>>(grid.cellCount, n, grid.blockSums);
    launches += 3;
    if(maxTotal > 0)
    {
        k_gridScatter<<<gridFor((maxTotal + 3u) / 4u, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(set, grid);
        launches++;
    }
    return launches;
}

void launchMerge(const DevScene &sc, const IterConst &ic, const QueryStore &queries,
                 const unsigned long long *queryCount, uint32_t hostQueryCount, const GridBuffers &grid,
                 float *accum, float *contribOut, int32_t *countsOut, unsigned long long *counters,
                 int groupSize, cudaStream_t stream)
{
    const uint64_t maxQueries = queryCount ? queries.capacity : hostQueryCount;
    if(maxQueries == 0) return;
    if(groupSize == 1)
    {
        k_mergeLane<<<gridFor(maxQueries, SVCM_BLOCK, gridMerge()), SVCM_BLOCK, 0, stream>>>(
            sc, ic, queries, queryCount, hostQueryCount, grid, accum, contribOut, countsOut, counters);
        return;
    }
    const uint32_t gridN = gridFor(maxQueries * uint64_t(groupSize), SVCM_BLOCK, 8);
    if(groupSize == 32)
        k_merge<32><<<gridN, SVCM_BLOCK, 0, stream>>>(sc, ic, queries, queryCount, hostQueryCount, grid, accum,
                                                       contribOut, countsOut, counters);
    else if(groupSize == 16)
        k_merge<16><<<gridN, SVCM_BLOCK, 0, stream>>>(sc, ic, queries, queryCount, hostQueryCount, grid, accum,
                                                       contribOut, countsOut, counters);
    else
        k_merge<8><<<gridN, SVCM_BLOCK, 0, stream>>>(sc, ic, queries, queryCount, hostQueryCount, grid, accum,
                                                      contribOut, countsOut, counters);
}

void launchIntersectBatch(const DevScene &sc, const float *rays, int64_t count, int32_t *hit, float *dist,
                          float *normal, int32_t *matID, int32_t *lightID, cudaStream_t stream)
{
    if(count <= 0) return;
    if(sc.bvh) k_intersectBatch<true><<<gridFor(count, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, rays, count, hit, dist, normal, matID, lightID);
    else       k_intersectBatch<false><<<gridFor(count, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, rays, count, hit, dist, normal, matID, lightID);
}
void launchOccludedBatch(const DevScene &sc, const float *segments, int64_t count, int32_t *occluded, cudaStream_t stream)
{
    if(count <= 0) return;
    if(sc.bvh) k_occludedBatch<true><<<gridFor(count, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, segments, count, occluded);
    else       k_occludedBatch<false><<<gridFor(count, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, segments, count, occluded);
}
void launchBsdfEvaluateBatch(const DevScene &sc, const float *in, int64_t count, int fixIsLight, float *out,
                             cudaStream_t stream)
{
    (void)fixIsLight;   // Evaluate does not depend on the tracing direction (only SampleRefract does)
    if(count <= 0) return;
    k_bsdfEvaluateBatch<<<gridFor(count, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, in, count, out);
}
void launchBsdfSampleBatch(const DevScene &sc, const float *in, int64_t count, int fixIsLight, float *out,
                           uint32_t *events, cudaStream_t stream)
{
    if(count <= 0) return;
    const uint32_t gridN = gridFor(count, SVCM_BLOCK, 8);
    if(fixIsLight) k_bsdfSampleBatch<true><<<gridN, SVCM_BLOCK, 0, stream>>>(sc, in, count, out, events);
    else           k_bsdfSampleBatch<false><<<gridN, SVCM_BLOCK, 0, stream>>>(sc, in, count, out, events);
}
void launchLightBatch(const DevScene &sc, int op, int lightIndex, const float *in, int64_t count, float *out,
                      cudaStream_t stream)
{
    if(count <= 0) return;
    k_lightBatch<<<gridFor(count, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, op, lightIndex, in, count, out);
}
void launchCameraBatch(const DevScene &sc, const float *in, int64_t count, float *out, cudaStream_t stream)
{
    if(count <= 0) return;
    k_cameraBatch<<<gridFor(count, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, in, count, out);
}
// ------------------------------------------------------------------------------------------
// image conversions after the path: the 8-bit images main() writes
// (src/smallvcm.cxx:313-320) computed from the accumulated image where it lies, so that a
// progressive viewer downloads 3 or 4 bytes per pixel instead of 12
// ------------------------------------------------------------------------------------------

// x86 `byte(float)` as the reference's casts behave (cvttss2si, low byte; NaN / out of range -> 0)
__device__ __forceinline__ uint32_t byteOfFloat(float f)
{
    if(!(fabsf(f) < 2147483648.f)) return 0u;
    return uint32_t(__float2int_rz(f)) & 0xFFu;
}

// std::pow(float, float) of glibc evaluates in double and rounds once; the device's double pow
// rounded to float agrees with it except on sub-ulp ties (tests/test_gpu_image.py bounds those)
__device__ __forceinline__ uint32_t gammaByte(float c, float invGamma)
{
    const float g = __fmul_rn(float(pow(double(c), double(invGamma))), 255.f);
    const float lo = (0.f < g) ? g : 0.f;       // std::max(0.f, g): NaN -> 0
    const float hi = (lo < 255.f) ? lo : 255.f; // std::min(255.f, lo)
    return byteOfFloat(hi);
}

// Framebuffer::SaveBMP payload, src/framebuffer.hxx:196-215: thread per OUTPUT pixel, rows
// bottom-up, B G R. `scale` is GetFramebuffer's 1.f / mIterations (src/renderer.hxx:53-54).
__global__ void __launch_bounds__(SVCM_BLOCK)
k_imageBmp(const float *__restrict__ accum, int width, int height, float scale, int applyScale,
           float invGamma, uint8_t *__restrict__ out)
{
    const int64_t pixels = int64_t(width) * height;
    for(int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < pixels; i += int64_t(gridDim.x) * blockDim.x)
    {
        const int y = int(i / width), x = int(i - int64_t(y) * width);
        const float *src = accum + 3 * (int64_t(x) + int64_t(height - y - 1) * width);
        float r = src[0], g = src[1], b = src[2];
        if(applyScale) { r = __fmul_rn(r, scale); g = __fmul_rn(g, scale); b = __fmul_rn(b, scale); }
        uint8_t *dst = out + 3 * i;
        dst[0] = uint8_t(gammaByte(b, invGamma));
        dst[1] = uint8_t(gammaByte(g, invGamma));
        dst[2] = uint8_t(gammaByte(r, invGamma));
    }
}

// Framebuffer::SaveHDR payload, src/framebuffer.hxx:229-249: rows top-down, R G B E
__global__ void __launch_bounds__(SVCM_BLOCK)
k_imageRgbe(const float *__restrict__ accum, int64_t pixels, float scale, int applyScale, uchar4 *__restrict__ out)
{
    for(int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < pixels; i += int64_t(gridDim.x) * blockDim.x)
    {
        float r = accum[3 * i], g = accum[3 * i + 1], b = accum[3 * i + 2];
        if(applyScale) { r = __fmul_rn(r, scale); g = __fmul_rn(g, scale); b = __fmul_rn(b, scale); }
        const float gb = (g < b) ? b : g;           // std::max(g, b)
        float v = (r < gb) ? gb : r;                // std::max(r, max(g, b))
        uchar4 px = make_uchar4(0, 0, 0, 0);
        if(v >= 1e-32f)
        {
            int e;
            const float m = frexpf(v, &e);
            v = __fdiv_rn(__fmul_rn(m, 256.f), v);
            px.x = (unsigned char)byteOfFloat(__fmul_rn(r, v));
            px.y = (unsigned char)byteOfFloat(__fmul_rn(g, v));
            px.z = (unsigned char)byteOfFloat(__fmul_rn(b, v));
            px.w = (unsigned char)(uint32_t(e + 128) & 0xFFu);
        }
        out[i] = px;
    }
}

void launchImageConvert(const float *accum, int width, int height, int iterations, int format, float gamma,
                        uint8_t *out, cudaStream_t stream)
{
    const int64_t pixels = int64_t(width) * height;
    if(pixels <= 0) return;
    const float scale = iterations > 0 ? 1.f / float(iterations) : 1.f;
    const int grid = gridFor(pixels, SVCM_BLOCK, 8);
    if(format == SVCM_IMAGE_BMP24)
        k_imageBmp<<<grid, SVCM_BLOCK, 0, stream>>>(accum, width, height, scale, iterations > 0, 1.f / gamma, out);
    else
        k_imageRgbe<<<grid, SVCM_BLOCK, 0, stream>>>(accum, pixels, scale, iterations > 0, (uchar4 *)out);
}

void launchRngStream(uint32_t key, uint32_t firstPath, int paths, int draws, float *out, cudaStream_t stream)
{
    if(paths <= 0) return;
    k_rngStream<<<(paths + 127) / 128, 128, 0, stream>>>(key, firstPath, paths, draws, out);
}
void launchPackMergeVertices(const SvcmMergeVertex *verts, int64_t count, float4 *posLen, float4 *thrVcm,
                             float4 *dirVm, float *contProb, cudaStream_t stream)
{
    if(count <= 0) return;
    k_packMergeVertices<<<gridFor(count, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(verts, count, posLen, thrVcm, dirVm, contProb);
}
void launchPackMergeQueries(const DevScene &sc, const SvcmMergeQuery *queries, int64_t count, QueryStore store,
                            cudaStream_t stream)
{
    if(count <= 0) return;
    k_packMergeQueries<<<gridFor(count, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(sc, queries, count, store);
}

} // namespace svcm
