// vertexcm_gpu.hxx — the reference-side plug: `class VertexCMGpu : public AbstractRenderer`.
//
// A drop-in for the reference's VertexCM (reference src/vertexcm.hxx:61-1031) behind the
// reference's own renderer interface (src/renderer.hxx:33-70): same nested AlgorithmType
// enum and numbering (src/vertexcm.hxx:182-204), same constructor signature
// (src/vertexcm.hxx:208-214), RunIteration(int) adds one full iteration into the inherited
// mFramebuffer and increments mIterations (src/vertexcm.hxx:544,547) so that the
// NON-virtual GetFramebuffer / WasUsed (src/renderer.hxx:49-58) keep working unchanged.
// All arithmetic happens in libsvcm.so (include/svcm.h) on the GPU; this header only
// flattens the Scene, forwards the calls and copies the accumulated image back.
//
// Header-only, to be included AFTER the reference headers in the single translation unit
// they allow — see smallvcm_gpu.cxx for the injection that leaves every
// reference file untouched. Errors follow the reference's convention: message + exit(2)
// (src/config.hxx:140-141). There is no CPU fallback for this class.
#ifndef SVCM_HOST_VERTEXCM_GPU_HXX_
#define SVCM_HOST_VERTEXCM_GPU_HXX_

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "svcm.h"
#include "flatten_scene.hxx"

class VertexCMGpu : public AbstractRenderer
{
public:

    // same names and values as VertexCM::AlgorithmType (src/vertexcm.hxx:182-204); the two
    // extra ids select the unidirectional integrators of the library (PathTracerGpu below)
    enum AlgorithmType
    {
        kLightTrace = 0,
        kPpm,
        kBpm,
        kBpt,
        kVcm,
        kPathTrace,
        kEyeLight
    };

    VertexCMGpu(
        const Scene&  aScene,
        AlgorithmType aAlgorithm,
        const float   aRadiusFactor,
        const float   aRadiusAlpha,
        int           aSeed = 1234
    ) :
        AbstractRenderer(aScene),
        mHandle(NULL),
        mRegistered(false)
    {
        if(!svcm_host::FlattenScene(aScene, mFlat))
            exit(2);

        memset(&mParams, 0, sizeof(mParams));
        mParams.algorithm    = int(aAlgorithm);
        mParams.radiusFactor = aRadiusFactor;
        mParams.radiusAlpha  = aRadiusAlpha;
        mParams.seed         = aSeed;
        mParams.device       = -1;      // chosen when the renderer is first used
        mParams.rank         = 0;
        mParams.worldSize    = 1;
    }

    virtual ~VertexCMGpu()
    {
        if(mRegistered)
            svcm_host_unregister(&mFramebuffer.mColor[0]);
        svcm_destroy(mHandle);
    }

    // One iteration = resX*resY light sub-paths + as many camera sub-paths, with the GLOBAL
    // iteration index the driver passes (the radius schedule needs it, src/vertexcm.hxx:295-298)
    // and the path-length limits the driver assigned after construction (src/smallvcm.cxx:70-71).
    virtual void RunIteration(int aIteration)
    {
        if(mHandle == NULL)
            Create();

        if(svcm_run_iterations(mHandle, aIteration, 1, mMinPathLength, mMaxPathLength) != SVCM_OK)
            Die();

        // mFramebuffer must hold the accumulated sum when we return: GetFramebuffer is not
        // virtual and reads it directly (src/renderer.hxx:49-55). Packed Vec3f == 3 floats.
        int iterationsDone = 0;
        if(svcm_download_accum(mHandle, reinterpret_cast<float*>(&mFramebuffer.mColor[0]),
               &iterationsDone) != SVCM_OK)
            Die();

        mIterations++;
    }

    svcm_renderer *Handle() const { return mHandle; }

    // What main() does with the result of ONE renderer (src/smallvcm.cxx:311-320:
    // GetFramebuffer, then SaveBMP with gamma 2.2 or SaveHDR by extension), with the pixel
    // conversion done on the device. Returns false for an unknown extension or an unused
    // renderer, after printing the reference's message for the former.
    bool SaveImage(const char *aFilename, float aGamma = 2.2f)
    {
        if(mHandle == NULL)
            return false;
        if(svcm_save_image(mHandle, aFilename, aGamma) == SVCM_OK)
            return true;
        printf("%s\n", svcm_last_error(mHandle));
        return false;
    }

private:

    // Device state is created on first use: the driver builds one renderer per hardware
    // thread up front (src/smallvcm.cxx:61-72) but with few iterations most are never run
    // (src/smallvcm.cxx:119-125), and an unused renderer should not hold GPU memory.
    void Create()
    {
        mParams.device = PickDevice();
        if(svcm_create(&mFlat.desc, &mParams, &mHandle) != SVCM_OK)
        {
            printf("VertexCMGpu: %s\n", svcm_last_error(NULL));
            exit(2);
        }

        // Pin the framebuffer's storage so the per-iteration copy-back runs at DMA speed.
        // Framebuffer::mColor is private (src/framebuffer.hxx:255): the wrapper TU is built
        // with -fno-access-control; the size is fixed by Setup() in the base constructor.
        if(svcm_host_register(&mFramebuffer.mColor[0],
               sizeof(Vec3f) * mFramebuffer.mColor.size()) == SVCM_OK)
            mRegistered = true;
    }

    void Die() const
    {
        printf("VertexCMGpu: %s\n", svcm_last_error(mHandle));
        exit(2);
    }

    // The driver creates one renderer per OpenMP thread (src/smallvcm.cxx:61-72); spread them
    // over the visible GPUs round-robin. SVCM_DEVICE pins everything to one ordinal.
    static int PickDevice()
    {
        const char *fixed = getenv("SVCM_DEVICE");
        if(fixed && fixed[0])
            return atoi(fixed);
        static int sNext = 0;
        int count = svcm_device_count();
        if(count <= 0)
            count = 1;
        int mine;
#pragma omp critical(svcm_pick_device)
        mine = sNext++;
        return mine % count;
    }

    svcm_host::FlatScene mFlat;
    SvcmParams           mParams;
    svcm_renderer       *mHandle;
    bool                 mRegistered;
};

// ---------------------------------------------------------------------------------------
// the reference's PathTracer / EyeLight on the GPU, opt-in.
//
// `-a pt` must keep working on a machine without a GPU (BASELINE config #1), so these two
// plugs choose at construction: with SVCM_GPU_PT=1 in the environment they run on the device
// (algorithm ids 5 / 6 of the same library), otherwise they delegate to the reference's own CPU
// class and mirror its framebuffer (GetFramebuffer is non-virtual and reads OUR members).
// Constructor signatures are the reference's: (const Scene&, int aSeed), src/pathtracer.hxx:37-42,
// src/eyelight.hxx:39-44.
// ---------------------------------------------------------------------------------------
template<class tCpuRenderer, int tAlgorithm>
class SvcmUnidirectionalGpu : public AbstractRenderer
{
public:

    SvcmUnidirectionalGpu(const Scene& aScene, int aSeed = 1234) :
        AbstractRenderer(aScene), mCpu(NULL), mGpu(NULL)
    {
        const char *useGpu = getenv("SVCM_GPU_PT");
        if(useGpu && useGpu[0] && useGpu[0] != '0')
            mGpu = new VertexCMGpu(aScene, VertexCMGpu::AlgorithmType(tAlgorithm), 0.003f, 0.75f, aSeed);
        else
            mCpu = new tCpuRenderer(aScene, aSeed);
    }

    virtual ~SvcmUnidirectionalGpu()
    {
        delete mCpu;
        delete mGpu;
    }

    virtual void RunIteration(int aIteration)
    {
        AbstractRenderer *inner = mGpu ? static_cast<AbstractRenderer*>(mGpu) : static_cast<AbstractRenderer*>(mCpu);
        inner->mMaxPathLength = mMaxPathLength;     // assigned by the driver after construction
        inner->mMinPathLength = mMinPathLength;
        inner->RunIteration(aIteration);
        mFramebuffer = inner->mFramebuffer;         // accumulated sum, as GetFramebuffer expects
        mIterations  = inner->mIterations;
    }

private:

    tCpuRenderer *mCpu;
    VertexCMGpu  *mGpu;
};

typedef SvcmUnidirectionalGpu<PathTracer, SVCM_ALG_PATH_TRACE> PathTracerGpu;
typedef SvcmUnidirectionalGpu<EyeLight,   SVCM_ALG_EYE_LIGHT>  EyeLightGpu;

#endif // SVCM_HOST_VERTEXCM_GPU_HXX_
