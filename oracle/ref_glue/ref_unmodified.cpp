// TEST INFRASTRUCTURE — not part of the product.
//
// Builds the UNMODIFIED reference (SmallVCM, /root/reference/src, compiled in place
// with -I, nothing copied) into oracle/_ref/libsvcm_ref.so behind a small C interface,
// because the BASELINE configurations (resolution, path length, radius, thread count)
// have no command-line switch in the reference (src/config.hxx:233-240).
//
// Uses: (1) pins oracle/vcm_oracle.cpp against the reference's own output
// (tests/test_oracle.py), (2) regenerates tests/golden/*, (3) is the
// `--impl reference` / cpu_baseline arm of bench.py (the reference's own OpenMP path,
// src/smallvcm.cxx:52-151, timed with a wall clock because render() returns CPU time).
//
// The reference headers define non-inline functions, so they can only be included into
// ONE translation unit: this file is that TU for this library.
#include <vector>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <chrono>

#define main smallvcm_reference_main
#include "smallvcm.cxx"
#undef main

#include "svcm.h"
#include "../../smallvcm_b200/host/flatten_scene.hxx"

#define REF_API extern "C" __attribute__((visibility("default")))

// Number of built-in scene configurations (src/config.hxx:146-151).
REF_API int ref_scene_count() { return SizeOfArray(g_SceneConfigs); }

REF_API unsigned ref_scene_mask(int sceneId)
{
    if(sceneId < 0 || sceneId >= SizeOfArray(g_SceneConfigs)) return 0;
    return g_SceneConfigs[sceneId];
}

REF_API int ref_num_procs() { return omp_get_num_procs(); }

// Runs the reference's own render() (src/smallvcm.cxx:52) on a programmatically filled
// Config. `algorithm` is Config::Algorithm (0 el, 1 pt, 2 lt, 3 ppm, 4 bpm, 5 bpt, 6 vcm).
// Writes W*H*3 floats (the averaged framebuffer render() produces) and returns the
// WALL seconds spent inside render(), or a negative value on error.
REF_API double ref_render(int sceneId, int algorithm, int width, int height,
                          int iterations, int threads, int baseSeed,
                          unsigned maxPathLength, unsigned minPathLength,
                          float radiusFactor, float radiusAlpha, float *outRgb)
{
    if(sceneId < 0 || sceneId >= SizeOfArray(g_SceneConfigs)) return -1.0;
    if(algorithm < 0 || algorithm >= (int)Config::kAlgorithmMax) return -1.0;

    Scene *scene = new Scene;
    scene->LoadCornellBox(Vec2i(width, height), g_SceneConfigs[sceneId]);
    scene->BuildSceneSphere();

    Framebuffer fb;
    Config config;
    config.mScene         = scene;
    config.mAlgorithm     = Config::Algorithm(algorithm);
    config.mIterations    = iterations;
    config.mMaxTime       = -1.f;
    config.mRadiusFactor  = radiusFactor;
    config.mRadiusAlpha   = radiusAlpha;
    config.mFramebuffer   = &fb;
    config.mNumThreads    = threads > 0 ? threads : std::max(1, omp_get_num_procs());
    config.mBaseSeed      = baseSeed;
    config.mMaxPathLength = maxPathLength;
    config.mMinPathLength = minPathLength;
    config.mOutputName    = "";
    config.mResolution    = Vec2i(width, height);
    config.mFullReport    = false;

    const auto t0 = std::chrono::steady_clock::now();
    render(config);
    const auto t1 = std::chrono::steady_clock::now();

    if(outRgb)
    {
        // Framebuffer::mColor is private; this TU is compiled with -fno-access-control
        memcpy(outRgb, &fb.mColor[0], sizeof(float) * 3 * size_t(width) * size_t(height));
    }

    delete scene;
    return std::chrono::duration<double>(t1 - t0).count();
}

// ref_render for a scene that is NOT one of the hard-coded boxes: g_SceneConfigs[sceneId] with
// `numExtra` more triangles appended to the GeometryList (vertices [n][3][3], material ids [n]),
// built by the reference's own Triangle constructor (src/geometry.hxx:112-123). The scene sphere
// is rebuilt by the reference (src/scene.hxx BuildSceneSphere) over the enlarged list. Optionally
// returns the flattened scene's primitives (product glue, flatten_scene.hxx) so that the caller
// can check its own way of building the same list, normals included.
REF_API double ref_render_extra(int sceneId, int algorithm, int width, int height, int iterations, int threads,
                                int baseSeed, unsigned maxPathLength, unsigned minPathLength, float radiusFactor,
                                float radiusAlpha, const float *extraVertices, const int *extraMaterials,
                                int numExtra, float *outRgb, SvcmPrimitive *outPrims, SvcmSceneSphere *outSphere)
{
    if(sceneId < 0 || sceneId >= SizeOfArray(g_SceneConfigs)) return -1.0;
    if(algorithm < 0 || algorithm >= (int)Config::kAlgorithmMax) return -1.0;

    Scene *scene = new Scene;
    scene->LoadCornellBox(Vec2i(width, height), g_SceneConfigs[sceneId]);
    GeometryList *list = dynamic_cast<GeometryList*>(scene->mGeometry);
    if(!list) { delete scene; return -1.0; }
    for(int i = 0; i < numExtra; i++)
    {
        const float *v = extraVertices + 9 * i;
        if(extraMaterials[i] < 0 || extraMaterials[i] >= scene->GetMaterialCount()) { delete scene; return -1.0; }
        list->mGeometry.push_back(new Triangle(Vec3f(v[0], v[1], v[2]), Vec3f(v[3], v[4], v[5]),
                                               Vec3f(v[6], v[7], v[8]), extraMaterials[i]));
    }
    scene->BuildSceneSphere();

    if(outPrims || outSphere)
    {
        svcm_host::FlatScene flat;
        if(!svcm_host::FlattenScene(*scene, flat)) { delete scene; return -1.0; }
        if(outPrims) memcpy(outPrims, flat.primitives.data(), sizeof(SvcmPrimitive) * flat.primitives.size());
        if(outSphere) *outSphere = flat.desc.sceneSphere;
    }

    Framebuffer fb;
    Config config;
    config.mScene         = scene;
    config.mAlgorithm     = Config::Algorithm(algorithm);
    config.mIterations    = iterations;
    config.mMaxTime       = -1.f;
    config.mRadiusFactor  = radiusFactor;
    config.mRadiusAlpha   = radiusAlpha;
    config.mFramebuffer   = &fb;
    config.mNumThreads    = threads > 0 ? threads : std::max(1, omp_get_num_procs());
    config.mBaseSeed      = baseSeed;
    config.mMaxPathLength = maxPathLength;
    config.mMinPathLength = minPathLength;
    config.mOutputName    = "";
    config.mResolution    = Vec2i(width, height);
    config.mFullReport    = false;

    const auto t0 = std::chrono::steady_clock::now();
    if(iterations > 0) render(config);
    const auto t1 = std::chrono::steady_clock::now();
    if(outRgb && iterations > 0)
        memcpy(outRgb, &fb.mColor[0], sizeof(float) * 3 * size_t(width) * size_t(height));
    delete scene;
    return std::chrono::duration<double>(t1 - t0).count();
}

// Framebuffer::TotalLuminance (src/framebuffer.hxx:89) of an image, float accumulation
// in the reference's pixel order — the KAT scalar of BASELINE.md §3.
REF_API float ref_total_luminance(const float *rgb, int width, int height)
{
    Framebuffer fb;
    fb.Setup(Vec2f(float(width), float(height)));
    memcpy(&fb.mColor[0], rgb, sizeof(float) * 3 * size_t(width) * size_t(height));
    return fb.TotalLuminance();
}

// Framebuffer::SaveBMP / SaveHDR (src/framebuffer.hxx:170-251) of a caller-supplied accumulated
// image after AbstractRenderer::GetFramebuffer's scaling (src/renderer.hxx:49-55): the reference
// writes the file itself; the tests read it back to pin the oracle's restatement of the pixel
// conversions and the product's file headers. format 0 = bmp, 1 = hdr.
REF_API int ref_save_image(const float *accum, int width, int height, int iterations, int format,
                           float gamma, const char *path)
{
    Framebuffer fb;
    fb.Setup(Vec2f(float(width), float(height)));
    memcpy(&fb.mColor[0], accum, sizeof(float) * 3 * size_t(width) * size_t(height));
    if(iterations > 0) fb.Scale(1.f / iterations);
    if(format == 0) fb.SaveBMP(path, gamma);
    else if(format == 1) fb.SaveHDR(path);
    else return -1;
    return 0;
}

// Flattens the reference's own Scene for `sceneId` with the product's host glue
// (smallvcm_b200/host/flatten_scene.hxx) into caller-provided arrays; used to check the
// library's built-in scene builder (svcm_scene_cornell_box) field by field.
REF_API int ref_flatten_scene(int sceneId, int width, int height,
                              SvcmPrimitive *prims, int *numPrims,
                              SvcmMaterial *mats, int *numMats,
                              SvcmLight *lights, int *numLights,
                              int *materialToLight, int *backgroundLight,
                              SvcmCamera *camera, SvcmSceneSphere *sphere)
{
    if(sceneId < 0 || sceneId >= SizeOfArray(g_SceneConfigs)) return -1;
    Scene scene;
    scene.LoadCornellBox(Vec2i(width, height), g_SceneConfigs[sceneId]);
    scene.BuildSceneSphere();

    svcm_host::FlatScene flat;
    if(!svcm_host::FlattenScene(scene, flat)) return -2;

    *numPrims  = (int)flat.primitives.size();
    *numMats   = (int)flat.materials.size();
    *numLights = (int)flat.lights.size();
    memcpy(prims,  flat.primitives.data(), sizeof(SvcmPrimitive) * flat.primitives.size());
    memcpy(mats,   flat.materials.data(),  sizeof(SvcmMaterial)  * flat.materials.size());
    memcpy(lights, flat.lights.data(),     sizeof(SvcmLight)     * flat.lights.size());
    memcpy(materialToLight, flat.materialToLight.data(), sizeof(int) * flat.materialToLight.size());
    *backgroundLight = flat.desc.backgroundLight;
    *camera = flat.desc.camera;
    *sphere = flat.desc.sceneSphere;
    return 0;
}
