/*
 * svcm.h — C-ABI of the B200-native VCM integrator (libsvcm.so).
 *
 * This is the drop-in boundary for ONE hot path of SmallVCM: the body of
 * VertexCM::RunIteration (reference src/vertexcm.hxx:284-548) together with
 * HashGrid::Build / HashGrid::Process (reference src/hashgrid.hxx:41-169).
 * A host renderer object (smallvcm_b200/host/vertexcm_gpu.hxx, a subclass of the
 * reference's AbstractRenderer, src/renderer.hxx:33-70) flattens the reference
 * `Scene` into the POD arrays below and drives these entry points; nothing in this
 * header depends on C++, on the reference headers or on torch.
 *
 * Conventions
 *   - every function returns 0 on success and a negative SVCM_E_* code on failure;
 *     nothing throws across the boundary (the reference has no error channel either:
 *     printf + exit(2), src/config.hxx:140-141 — the host glue maps a failure to that);
 *   - svcm_last_error(r) gives a human-readable message for the last failure on `r`
 *     (or, with r == NULL, for the last failed svcm_create on this thread);
 *   - all pointers in SvcmSceneDesc are HOST pointers owned by the caller; the library
 *     copies what it needs inside svcm_create (the reference Scene outlives its
 *     renderers, src/smallvcm.cxx:323, but we do not rely on it);
 *   - a renderer handle is not entered concurrently (the reference never does,
 *     src/smallvcm.cxx:82-108); different handles may be driven from different
 *     threads (each call does its own cudaSetDevice).
 */
#ifndef SVCM_H_
#define SVCM_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SVCM_API __attribute__((visibility("default")))
#else
#define SVCM_API
#endif

/* ----------------------------------------------------------------------------------
 * error codes
 * -------------------------------------------------------------------------------- */
enum {
    SVCM_OK            = 0,
    SVCM_E_INVALID     = -1, /* bad argument / scene exceeds a compiled-in limit      */
    SVCM_E_NO_DEVICE   = -2, /* no usable CUDA device (the CPU integrators of the
                                reference stay usable: BASELINE config #1)            */
    SVCM_E_CUDA        = -3, /* a CUDA runtime call or kernel failed                   */
    SVCM_E_NCCL        = -4, /* NCCL not loadable or a collective failed               */
    SVCM_E_NOMEM       = -5, /* device/host allocation failed                          */
    SVCM_E_STATE       = -6  /* call not valid in the handle's current state           */
};

/* ----------------------------------------------------------------------------------
 * compiled-in scene limits: camera, materials, lights and up to SVCM_MAX_PRIMITIVES
 * primitives travel to every kernel as a __grid_constant__ parameter (constant bank,
 * ~5 KB), so concurrent handles with different scenes never share state. The reference's
 * scenes hold <= 22 primitives, 9 materials, 2 lights and are intersected by a loop over
 * the whole list (src/geometry.hxx:65-91), which that size keeps. Longer lists (up to
 * SVCM_MAX_PRIMITIVES_BVH) are copied to device memory under a bounding volume hierarchy
 * built at svcm_create; traversal returns exactly what the loop
 * would: same per-primitive test, first-in-list wins equal distances.
 * -------------------------------------------------------------------------------- */
#define SVCM_MAX_PRIMITIVES     64
#define SVCM_MAX_PRIMITIVES_BVH (1 << 24)
#define SVCM_MAX_MATERIALS      16
#define SVCM_MAX_LIGHTS         8

/* ----------------------------------------------------------------------------------
 * flattened scene (replaces the pointer graph of reference src/scene.hxx:474-485)
 * -------------------------------------------------------------------------------- */

/* Primitive kinds; list order is the reference's GeometryList order and must be kept
 * (first-in-list wins exact distance ties: strict `<` at src/geometry.hxx:146,226). */
enum { SVCM_PRIM_TRIANGLE = 0, SVCM_PRIM_SPHERE = 1 };

/* One entry of GeometryList::mGeometry (src/geometry.hxx:101).
 *   triangle: p0,p1,p2 and the precomputed unit normal (src/geometry.hxx:172-176)
 *   sphere  : v[0..2] = center, v[3] = radius            (src/geometry.hxx:261-265) */
typedef struct SvcmPrimitive {
    int32_t type;   /* SVCM_PRIM_*                                  */
    int32_t matID;  /* index into materials                         */
    float   v[12];  /* triangle: p0 p1 p2 normal; sphere: c.xyz r   */
} SvcmPrimitive;

/* reference `Material`, src/materials.hxx:36-66 (same field order, 44 bytes) */
typedef struct SvcmMaterial {
    float diffuse[3];
    float phong[3];
    float phongExponent;
    float mirror[3];
    float ior;       /* < 0: no Fresnel / no transmission */
} SvcmMaterial;

/* Light kinds — the four AbstractLight subclasses of src/lights.hxx */
enum {
    SVCM_LIGHT_AREA        = 0, /* AreaLight        src/lights.hxx:112-233 */
    SVCM_LIGHT_DIRECTIONAL = 1, /* DirectionalLight src/lights.hxx:236-316 */
    SVCM_LIGHT_POINT       = 2, /* PointLight       src/lights.hxx:320-397 */
    SVCM_LIGHT_BACKGROUND  = 3  /* BackgroundLight  src/lights.hxx:401-514 */
};

/* Tagged union of the public fields of the four light classes.
 *   area       : p0,e1,e2, frame(x,y,z), intensity, invArea
 *   directional: frame(x,y,z), intensity
 *   point      : p0 = position, intensity
 *   background : intensity = mBackgroundColor (already divided by 255), scale = mScale */
typedef struct SvcmLight {
    int32_t type;
    float   p0[3], e1[3], e2[3];
    float   frameX[3], frameY[3], frameZ[3];
    float   intensity[3];
    float   invArea;
    float   scale;
} SvcmLight;

/* reference `Camera` public fields, src/camera.hxx:121-126; matrices column-major
 * exactly as Mat4f stores them (src/math.hxx:270-274: element (r,c) at [r + 4*c]) */
typedef struct SvcmCamera {
    float position[3];
    float forward[3];
    float resolution[2];
    float rasterToWorld[16];
    float worldToRaster[16];
    float imagePlaneDist;
} SvcmCamera;

/* reference `SceneSphere`, src/lights.hxx:32-40 */
typedef struct SvcmSceneSphere {
    float center[3];
    float radius;
    float invRadiusSqr;
} SvcmSceneSphere;

typedef struct SvcmSceneDesc {
    const SvcmPrimitive *primitives;     int32_t numPrimitives;
    const SvcmMaterial  *materials;      int32_t numMaterials;
    const SvcmLight     *lights;         int32_t numLights;
    /* material -> light id (or -1): the std::map of src/scene.hxx:61-66, densified;
     * numMaterials entries */
    const int32_t       *materialToLight;
    /* index into lights of Scene::mBackground, or -1 (src/scene.hxx:112-115) */
    int32_t              backgroundLight;
    SvcmCamera           camera;
    SvcmSceneSphere      sceneSphere;
} SvcmSceneDesc;

/* ----------------------------------------------------------------------------------
 * renderer parameters — the arguments of VertexCM::VertexCM (src/vertexcm.hxx:208-214)
 * -------------------------------------------------------------------------------- */

/* same numbering as VertexCM::AlgorithmType, src/vertexcm.hxx:182-204 */
enum {
    SVCM_ALG_LIGHT_TRACE = 0,
    SVCM_ALG_PPM         = 1,
    SVCM_ALG_BPM         = 2,
    SVCM_ALG_BPT         = 3,
    SVCM_ALG_VCM         = 4,
    /* SURVEY §8(f) rank 1 — the reference's two other integrators on the same device functions:
     * PathTracer (src/pathtracer.hxx:45-215) and EyeLight (src/eyelight.hxx:47-78). They are
     * separate classes in the reference; here they are two more algorithm ids of one handle. */
    SVCM_ALG_PATH_TRACE  = 5,
    SVCM_ALG_EYE_LIGHT   = 6
};

typedef struct SvcmParams {
    int32_t algorithm;      /* SVCM_ALG_*                                              */
    float   radiusFactor;   /* Config::mRadiusFactor (src/config.hxx:239)              */
    float   radiusAlpha;    /* Config::mRadiusAlpha  (src/config.hxx:240)              */
    int32_t seed;           /* the aSeed of CreateRenderer (src/config.hxx:112)        */
    int32_t device;         /* CUDA device ordinal this handle runs on                 */
    /* Path sharding for one-process-per-GPU runs.  Path index p (= pixel
     * index = light sub-path index, src/vertexcm.hxx:321,415,504-506) is owned by the
     * rank (p / width) % worldSize (svcm_shard_owner below; DESIGN.md "sharding"); with
     * worldSize == 1 the handle owns every path. */
    int32_t rank;
    int32_t worldSize;
    int32_t reserved[5];    /* must be zero */
} SvcmParams;

/* Per-iteration workload counters and device timings of the LAST completed iteration;
 * feeds the roofline arithmetic of SURVEY §8(d). Times are CUDA-event milliseconds on
 * the handle's stream and are only collected when profiling is enabled
 * (svcm_set_option(r, SVCM_OPT_PROFILE, 1)) because recording them serialises nothing
 * but costs event calls. */
typedef struct SvcmStats {
    int64_t paths;            /* P = W*H                                               */
    int64_t localPaths;       /* paths owned by this rank                              */
    int64_t lightVertices;    /* N_v over all ranks (size of the gathered set)         */
    int64_t localLightVertices;
    int64_t mergeQueries;     /* N_q: non-delta camera vertices inside the vertex bbox */
    int64_t mergeCandidates;  /* N_cand: particles visited in the 8 buckets            */
    int64_t mergeAccepted;    /* N_acc: particles within the radius                    */
    int64_t cameraRays;       /* closest-hit rays traced by the camera pass            */
    int64_t lightRays;        /* closest-hit rays traced by the light pass             */
    int64_t shadowRays;       /* any-hit rays (both passes)                            */
    float   radius;           /* merge radius of that iteration                        */
    float   msLightTrace;
    float   msGather;
    float   msGridBuild;
    float   msCameraTrace;
    float   msMerge;
    float   msTotal;
    int32_t kernelLaunches;   /* this library's kernels launched for that iteration    */
    int32_t paramBytesPerLaunch; /* bytes of kernel parameters (scene + constants) a trace
                                    kernel launch carries host -> device               */
} SvcmStats;

/* Sums over every iteration run since the previous svcm_get_profile call (at most the last
 * SVCM_PROFILE_RING iterations are retained): workload counters always, stage times only
 * while SVCM_OPT_PROFILE is on. Lets a caller time K iterations without synchronising
 * between them and still get per-kernel averages for the roofline arithmetic. */
#define SVCM_PROFILE_RING 256
typedef struct SvcmProfile {
    int32_t iterations;       /* iterations summed here                                */
    int32_t timedIterations;  /* of which carried stage timings                        */
    int64_t lightVertices, mergeQueries, mergeCandidates, mergeAccepted;
    int64_t lightRays, cameraRays, shadowRays;
    int64_t kernelLaunches;
    double  msLightTrace, msGather, msGridBuild, msCameraTrace, msMerge, msTotal;
} SvcmProfile;

enum {
    SVCM_OPT_PROFILE       = 1, /* 0/1: record per-stage CUDA events                   */
    SVCM_OPT_MERGE_GROUP   = 2, /* merge kernel (bucket layout): 1 = one lane per range
                                   query straight from global memory; 2 = one lane per
                                   query, the buckets of a warp's 32 queries de-duplicated
                                   and staged into shared memory by bulk async copies;
                                   3 = like 2 with plain loads; 4, 5 = like 2 with smaller
                                   stages; 8, 16, 32 = that many lanes cooperating on one
                                   query                                               */
    SVCM_OPT_COUNTERS      = 3, /* 0/1: maintain N_cand/N_acc counters (small cost)    */
    SVCM_OPT_GRID_LAYOUT   = 4, /* particle order of the hash grid: 0 = the reference's
                                   hash buckets (mCellEnds / mIndices as they are);
                                   1 = true cell order (spatially coherent rows; the
                                   bucket collisions' double visits are reproduced
                                   arithmetically, results identical)                   */
    SVCM_OPT_MERGE_VARIANT = 5, /* cell layout only: 0 = lane per query from global
                                   memory, 1 = neighbourhood staged through shared
                                   memory by bulk async copies (default), 3 = like 1
                                   with smaller stages                                  */
    SVCM_OPT_OVERLAP       = 7, /* which stages run concurrently (bit mask, default 15):
                                   1 = camera walk with the light pass, 2 = grid build (and
                                   all-gather) with the camera walk, 4 = merge with vertex
                                   connection, 8 = the merge of one iteration with the light
                                   pass of the next (iterations enqueued back to back; every
                                   other entry point waits for the pending merge first).
                                   0 = one stage after the other: per-stage times are then
                                   those of the stage alone                               */
    SVCM_OPT_REF_CANDIDATES = 6 /* 0/1, cell layout: also histogram the reference's
                                   buckets so that mergeCandidates counts what the
                                   reference's walk would visit (statistics; off: the
                                   distance tests this library performs)                */
};

typedef struct svcm_renderer svcm_renderer;

/* ----------------------------------------------------------------------------------
 * life cycle
 * -------------------------------------------------------------------------------- */

/* Replaces `new VertexCM(scene, algorithm, radiusFactor, radiusAlpha, seed)`
 * (src/config.hxx:124-138). Copies the scene, allocates device state for W*H paths. */
SVCM_API int svcm_create(const SvcmSceneDesc *scene, const SvcmParams *params,
                         svcm_renderer **out);

/* Replaces the virtual destructor (src/renderer.hxx:45, src/smallvcm.cxx:146). */
SVCM_API void svcm_destroy(svcm_renderer *r);

SVCM_API const char *svcm_last_error(const svcm_renderer *r);

SVCM_API int svcm_set_option(svcm_renderer *r, int32_t option, int64_t value);

/* ----------------------------------------------------------------------------------
 * the hot path
 * -------------------------------------------------------------------------------- */

/* Replaces VertexCM::RunIteration(aIteration) (src/vertexcm.hxx:284-548) for `count`
 * consecutive global iteration indices first..first+count-1 (the radius schedule
 * depends on the global index, src/vertexcm.hxx:295-298).  minPathLength and
 * maxPathLength are AbstractRenderer::mMinPathLength / mMaxPathLength, which the
 * driver assigns AFTER construction (src/smallvcm.cxx:70-71) — hence arguments here.
 * Work is enqueued on the handle's stream; the call returns without waiting for the
 * GPU unless the handle is part of a multi-rank group (the vertex all-gather needs the
 * per-rank counts on the host). */
SVCM_API int svcm_run_iterations(svcm_renderer *r, int32_t firstIteration, int32_t count,
                                 uint32_t minPathLength, uint32_t maxPathLength);

/* Blocks until everything enqueued on the handle has finished. */
SVCM_API int svcm_synchronize(svcm_renderer *r);

/* Replaces the read of AbstractRenderer::mFramebuffer (src/renderer.hxx:49-55): copies
 * the ACCUMULATED (not yet divided by the iteration count) image, W*H*3 floats,
 * row-major x + y*W, into host memory and reports how many iterations it holds —
 * exactly the state `mFramebuffer` / `mIterations` would have (src/vertexcm.hxx:544,547).
 * With worldSize > 1 every rank holds the light-path splats of its own light sub-paths
 * over the whole frame plus the camera-path colours of its own pixels; the sum over
 * ranks is the image (svcm_reduce_accum does that sum on the GPUs). */
SVCM_API int svcm_download_accum(svcm_renderer *r, float *rgb, int32_t *iterationsDone);

/* Zeroes the accumulated image and the iteration count (Framebuffer::Clear,
 * src/framebuffer.hxx:70-73). */
SVCM_API int svcm_clear_accum(svcm_renderer *r);

/* The 8-bit images the reference's main() writes after the path (src/smallvcm.cxx:311-320),
 * converted on the device from the accumulated image so that a viewer moves 3 or 4 bytes per
 * pixel over PCIe instead of 12:
 *   SVCM_IMAGE_BMP24  Framebuffer::SaveBMP's payload (src/framebuffer.hxx:196-215): rows
 *                     bottom-up, B G R, no row padding, byte(min(255, max(0, pow(c, 1/gamma)*255)))
 *   SVCM_IMAGE_RGBE   Framebuffer::SaveHDR's payload (src/framebuffer.hxx:229-249): rows
 *                     top-down, shared-exponent R G B E (gamma ignored)
 * of GetFramebuffer's image (accumulated colour * 1.f/iterations, src/renderer.hxx:49-55).
 * svcm_image_bytes: payload size (-1 for an unknown format). svcm_image_header: writes the file
 * header the reference puts in front of the payload into `out` if `capacity` suffices and
 * returns its length. svcm_save_image = header + payload into `path`, format chosen from the
 * last three characters ("bmp" / "hdr") like main(); anything else is SVCM_E_INVALID with the
 * reference's "Used unknown extension" text. */
#define SVCM_IMAGE_BMP24 0
#define SVCM_IMAGE_RGBE  1
SVCM_API int64_t svcm_image_bytes(int32_t format, int32_t width, int32_t height);
SVCM_API int32_t svcm_image_header(int32_t format, int32_t width, int32_t height, uint8_t *out, int32_t capacity);
SVCM_API int svcm_download_image(svcm_renderer *r, int32_t format, float gamma, uint8_t *out, int32_t *iterationsDone);
SVCM_API int svcm_save_image(svcm_renderer *r, const char *path, float gamma);

SVCM_API int svcm_get_stats(svcm_renderer *r, SvcmStats *out);
SVCM_API int svcm_get_profile(svcm_renderer *r, SvcmProfile *out);

/* Page-locks / releases caller memory (cudaHostRegister) so that svcm_download_accum into it
 * runs at DMA speed; optional. svcm_device_count returns the number of CUDA devices (0 if
 * none or no driver). */
SVCM_API int svcm_host_register(void *ptr, size_t bytes);
SVCM_API int svcm_host_unregister(void *ptr);
SVCM_API int svcm_device_count(void);

/* CUDA-event timing on the handle's own stream (the library launches on a private
 * non-blocking stream, so events recorded elsewhere do not see its kernels):
 * svcm_event_record enqueues event `slot` (0..7) behind everything submitted so far;
 * svcm_event_elapsed_ms waits for both events and returns the device time between them. */
#define SVCM_NUM_USER_EVENTS 8
SVCM_API int svcm_event_record(svcm_renderer *r, int32_t slot);
SVCM_API int svcm_event_elapsed_ms(svcm_renderer *r, int32_t startSlot, int32_t stopSlot, float *ms);

/* ----------------------------------------------------------------------------------
 * multi-GPU (one process per GPU; no counterpart in the reference)
 * -------------------------------------------------------------------------------- */

#define SVCM_COMM_ID_BYTES 128

/* Rank 0 creates an id, the launcher broadcasts the bytes to the other ranks by any
 * means (torch.distributed, a file, MPI ...), then every rank calls svcm_comm_init. */
SVCM_API int svcm_comm_unique_id(uint8_t id[SVCM_COMM_ID_BYTES]);
SVCM_API int svcm_comm_init(svcm_renderer *r, const uint8_t id[SVCM_COMM_ID_BYTES]);

/* Sums the accumulated images of all ranks into rank `root`'s image (ncclReduce). After
 * it, svcm_download_accum on `root` returns the whole image. */
SVCM_API int svcm_reduce_accum(svcm_renderer *r, int32_t root);

/* The path -> rank map of a sharded run (host code, no GPU needed): image rows are dealt
 * round-robin, rank = (path / width) % worldSize; both sub-paths of an index live on the
 * same rank (vertex connection pairs them, src/vertexcm.hxx:504-506).
 *   svcm_shard_local_paths : number of paths rank owns
 *   svcm_shard_owner       : owning rank of global path index `path`
 *   svcm_shard_local_index : index of `path` inside its owner's local numbering
 *   svcm_shard_global_index: inverse of the above for (rank, local) */
SVCM_API int64_t svcm_shard_local_paths(int32_t width, int32_t height, int32_t rank, int32_t worldSize);
SVCM_API int32_t svcm_shard_owner(int32_t width, int32_t worldSize, int64_t path);
SVCM_API int64_t svcm_shard_local_index(int32_t width, int32_t worldSize, int64_t path);
SVCM_API int64_t svcm_shard_global_index(int32_t width, int32_t rank, int32_t worldSize, int64_t local);

/* ----------------------------------------------------------------------------------
 * kernel-level entry points (the stages of the path, callable on their own; used by
 * the parity tests and the micro-benchmarks — SURVEY §8c tier T1)
 * -------------------------------------------------------------------------------- */

/* A light vertex as the merge needs it: the 52 bytes of PathVertex<true> that
 * RangeQuery::Process reads (src/vertexcm.hxx:130-169). */
typedef struct SvcmMergeVertex {
    float    position[3];    /* mHitpoint                                   */
    float    throughput[3];  /* mThroughput                                 */
    float    dirFix[3];      /* mBsdf.WorldDirFix()                         */
    uint32_t pathLength;     /* mPathLength                                 */
    float    dVCM;
    float    dVM;
    float    contProb;       /* mBsdf.ContinuationProb()                    */
} SvcmMergeVertex;

/* A merge query: the camera vertex RangeQuery is built from (src/vertexcm.hxx:532). */
typedef struct SvcmMergeQuery {
    float    position[3];    /* hitPoint                                    */
    float    normal[3];      /* isect.normal (the BSDF frame is SetFromZ of it) */
    float    dirFix[3];      /* -ray.dir (world); BSDF::Setup makes it local */
    int32_t  matID;
    uint32_t pathLength;     /* cameraState.mPathLength                     */
    float    dVCM;
    float    dVM;
} SvcmMergeQuery;

typedef struct svcm_hashgrid svcm_hashgrid;

/* HashGrid::Reserve + HashGrid::Build (src/hashgrid.hxx:35-107) over `count` vertices
 * given in HOST memory: bbox, cell keys, per-cell histogram, exclusive scan, scatter
 * into the cell-sorted arrays — all on `device`. numCells is the argument of Reserve
 * (the reference passes the path count, src/vertexcm.hxx:406). */
SVCM_API int svcm_hashgrid_build(const SvcmMergeVertex *vertices, int64_t count,
                                 float radius, int32_t numCells, int32_t device,
                                 svcm_hashgrid **out);
SVCM_API void svcm_hashgrid_destroy(svcm_hashgrid *g);

/* Results of the build for exact comparison with the reference's members:
 *   bboxMin/bboxMax (6 floats), cellEnds (numCells ints: index one past the last
 *   particle of each cell, the final state of mCellEnds, src/hashgrid.hxx:90-91) and
 *   indices (count ints: original vertex index at each sorted position, mIndices —
 *   the order INSIDE a cell is unspecified here, compare cells as sets). Any output
 *   pointer may be NULL. */
SVCM_API int svcm_hashgrid_download(svcm_hashgrid *g, float bbox[6], int32_t *cellEnds,
                                    int32_t *indices);

/* HashGrid::Process + RangeQuery::Process (src/hashgrid.hxx:110-169,
 * src/vertexcm.hxx:130-169) for `count` queries given in HOST memory against scene
 * materials: writes the un-normalised `mContrib` (3 floats per query) and, if not
 * NULL, per-query candidate / accepted counts (2 ints per query). misVcWeightFactor is
 * mMisVcWeightFactor; ppm != 0 selects the "no MIS" branch (src/vertexcm.hxx:164-166). */
SVCM_API int svcm_hashgrid_query(svcm_hashgrid *g, const SvcmSceneDesc *scene,
                                 const SvcmMergeQuery *queries, int64_t count,
                                 uint32_t minPathLength, uint32_t maxPathLength,
                                 float misVcWeightFactor, int32_t ppm,
                                 float *contrib, int32_t *counts);

/* Scene::Intersect (src/scene.hxx:53-70) for `count` rays (6 floats each: org, dir;
 * tmin = 0, initial dist = 1e36 like src/vertexcm.hxx:333-335).  Outputs per ray:
 * hit[i] = primitive list index or -1, dist[i], normal (3 floats), matID, lightID. */
SVCM_API int svcm_intersect_batch(const SvcmSceneDesc *scene, const float *rays,
                                  int64_t count, int32_t device, int32_t *hit,
                                  float *dist, float *normal, int32_t *matID,
                                  int32_t *lightID);

/* Scene::Occluded (src/scene.hxx:72-85) for `count` segments (7 floats each: point,
 * dir, tMax). occluded[i] = 0/1. */
SVCM_API int svcm_occluded_batch(const SvcmSceneDesc *scene, const float *segments,
                                 int64_t count, int32_t device, int32_t *occluded);

/* BSDF<FixIsLight>::Setup + Evaluate (+ Pdf) (src/bsdf.hxx:95-180): per item 10 floats
 * in (normal, dirFix world = -ray.dir, dirGen world, as_float(matID)); 8 floats out:
 * f.rgb, cosThetaGen, directPdfW, reversePdfW, continuationProb, flags
 * (bit0 valid, bit1 delta). fixIsLight selects BSDF<true>/<false>. */
SVCM_API int svcm_bsdf_evaluate_batch(const SvcmSceneDesc *scene, const float *in,
                                      int64_t count, int32_t fixIsLight, int32_t device,
                                      float *out);

/* BSDF<FixIsLight>::Setup + Sample (src/bsdf.hxx:191-257): per item 10 floats in
 * (normal, dirFix world, rnd triplet, as_float(matID)); 8 floats out: f.rgb,
 * dirGen world xyz, pdfW, cosThetaGen; events[i] = sampled event bits or 0. */
SVCM_API int svcm_bsdf_sample_batch(const SvcmSceneDesc *scene, const float *in,
                                    int64_t count, int32_t fixIsLight, int32_t device,
                                    float *out, uint32_t *events);

/* The light classes, light `lightIndex` of the scene, 8 floats in and 12 floats out per item
 * (unused slots are ignored on input and zero on output):
 *   emit       AbstractLight::Emit (src/lights.hxx:168-196 area, :267-292 directional, :354-375
 *              point, :438-478 background). in: dirRnd.xy, posRnd.xy. out: energy.rgb, position.xyz,
 *              direction.xyz, emissionPdfW, directPdfA, cosThetaLight.
 *   illuminate AbstractLight::Illuminate (:131-166, :244-265, :329-352, :410-436). in:
 *              receivingPosition.xyz, rnd.xy. out: radiance.rgb, directionToLight.xyz, distance,
 *              directPdfW, emissionPdfW, cosAtLight (the three pdfs are zero when no radiance comes
 *              back: the reference leaves them unset there).
 *   radiance   AbstractLight::GetRadiance (:198-220, :296-304, :377-385, :480-502). in:
 *              rayDirection.xyz, hitPoint.xyz (unused by every light). out: radiance.rgb, directPdfA,
 *              emissionPdfW. */
SVCM_API int svcm_light_emit_batch(const SvcmSceneDesc *scene, int32_t lightIndex, const float *in,
                                   int64_t count, int32_t device, float *out);
SVCM_API int svcm_light_illuminate_batch(const SvcmSceneDesc *scene, int32_t lightIndex, const float *in,
                                         int64_t count, int32_t device, float *out);
SVCM_API int svcm_light_radiance_batch(const SvcmSceneDesc *scene, int32_t lightIndex, const float *in,
                                       int64_t count, int32_t device, float *out);

/* Camera::GenerateRay (src/camera.hxx:108-117) and Camera::WorldToRaster + CheckRaster (:95-105).
 * in, 5 floats: raster.xy of the ray, a world point.xyz. out, 6 floats: ray direction.xyz (the
 * origin is the camera position), raster.xy of the point, 1 or 0 = inside the image. */
SVCM_API int svcm_camera_ray_batch(const SvcmSceneDesc *scene, const float *in, int64_t count,
                                   int32_t device, float *out);

/* The hierarchy svcm_create builds for scenes of more than SVCM_MAX_PRIMITIVES primitives
 * (host code, no GPU needed; exposed for inspection and for the CPU check of the traversal
 * in tests/). Node 0 is the root. Inner node: count == 0, children are nodes `first` and
 * `first + 1`. Leaf: `count` slots starting at `first`; order[slot] is the primitive's index
 * in the caller's list. Boxes are padded (see smallvcm_b200/csrc/svcm_bvh.h). At most
 * 2 * numPrimitives - 1 nodes; SVCM_E_NOMEM if nodeCapacity is smaller (numNodes is set). */
typedef struct SvcmBvhNode {
    float   lo[3]; int32_t first;
    float   hi[3]; int32_t count;
} SvcmBvhNode;
SVCM_API int svcm_bvh_build(const SvcmPrimitive *primitives, int32_t numPrimitives, SvcmBvhNode *nodes,
                            int32_t nodeCapacity, int32_t *order, int32_t *numNodes);

/* How svcm_create cuts a list of at most SVCM_MAX_PRIMITIVES primitives (the reference's scenes:
 * GeometryList, src/geometry.hxx:65-91, tests every primitive for every ray) into contiguous runs
 * (host code, no GPU needed; exposed for inspection and tests). The list order is kept. Run i
 * covers primitives [runs[i].first, runs[i].first + runs[i].count). Bit i of *skippable set: the
 * run is consecutive small triangles inside the padded box runs[i].lo/hi and a ray that misses
 * the box skips it — like a leaf of the hierarchy above and under the same rules (no primitive the
 * arithmetic of the reference's test would hit is ever skipped); bit clear: always tested (large
 * triangles, spheres). At most SVCM_MAX_RUNS runs. */
#define SVCM_MAX_RUNS 8
SVCM_API int svcm_primitive_runs(const SvcmPrimitive *primitives, int32_t numPrimitives, SvcmBvhNode *runs,
                                 uint32_t *skippable, int32_t *numRuns);

/* The image conversions of svcm_download_image on a caller-supplied accumulated image
 * [height][width][3] (host memory); out holds svcm_image_bytes(format, width, height). */
SVCM_API int svcm_image_convert_batch(const float *accum, int32_t width, int32_t height, int32_t iterations,
                                      int32_t format, float gamma, int32_t device, uint8_t *out);

/* The per-path random stream shared by this library and the parity oracle (DESIGN.md
 * "random numbers"): fills out[i*draws + k] with draw k of path `firstPath + i` for
 * (seed, iteration, phase); phase 0 = light sub-path, 1 = camera sub-path. Host code,
 * no GPU needed — it pins the generator for the oracle. */
SVCM_API int svcm_rng_stream(int32_t seed, int32_t iteration, int32_t phase,
                             uint32_t firstPath, int32_t paths, int32_t draws,
                             float *out);
/* The same stream generated by the device code (one thread per path). */
SVCM_API int svcm_rng_stream_device(int32_t seed, int32_t iteration, int32_t phase,
                                    uint32_t firstPath, int32_t paths, int32_t draws,
                                    int32_t device, float *out);

/* ----------------------------------------------------------------------------------
 * built-in scenes (host code; replaces Scene::LoadCornellBox + BuildSceneSphere,
 * src/scene.hxx:132-398, for callers that do not link the reference headers — the
 * reference-side glue flattens the reference's own Scene instead)
 * -------------------------------------------------------------------------------- */

/* Scene::BoxMask bits, src/scene.hxx:116-130 */
enum {
    SVCM_BOX_LIGHT_CEILING       = 1,
    SVCM_BOX_LIGHT_SUN           = 2,
    SVCM_BOX_LIGHT_POINT         = 4,
    SVCM_BOX_LIGHT_BACKGROUND    = 8,
    SVCM_BOX_LARGE_MIRROR_SPHERE = 16,
    SVCM_BOX_LARGE_GLASS_SPHERE  = 32,
    SVCM_BOX_SMALL_MIRROR_SPHERE = 64,
    SVCM_BOX_SMALL_GLASS_SPHERE  = 128,
    SVCM_BOX_GLOSSY_FLOOR        = 256
};

typedef struct svcm_scene svcm_scene;

/* g_SceneConfigs[sceneID] (src/config.hxx:146-151); returns 0 for an invalid id. */
SVCM_API uint32_t svcm_scene_config_mask(int32_t sceneID);

/* Builds the Cornell box `boxMask` at resolution W x H; *desc points into memory owned
 * by *out until svcm_scene_free. */
SVCM_API int svcm_scene_cornell_box(uint32_t boxMask, int32_t width, int32_t height,
                                    svcm_scene **out, const SvcmSceneDesc **desc);
SVCM_API void svcm_scene_free(svcm_scene *s);

/* library / build identification: "svcm <version> sm_100a ..." */
SVCM_API const char *svcm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SVCM_H_ */
