// svcm_api.cu — host side of the C-ABI (include/svcm.h): handle life cycle, the per-iteration
// kernel schedule that replaces VertexCM::RunIteration (reference src/vertexcm.hxx:284-548),
// the NCCL all-gather of light vertices for pixel-sharded multi-GPU runs, and the kernel-level
// entry points. No reference types, no torch types, no CPU fallback: every entry point that
// computes needs a CUDA device and says so loudly if there is none.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <algorithm>
#include <string>
#include <vector>
#include <mutex>
#include <chrono>
#include "svcm.h"
#include "svcm_internal.h"

using namespace svcm;

// ------------------------------------------------------------------------------------------
// NCCL, resolved at run time (the process may already carry torch's libnccl.so.2; linking a
// second copy at build time would be wrong). Minimal declarations of the stable C API.
// ------------------------------------------------------------------------------------------
namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_UINT32 = 3, NCCL_FLOAT32 = 7 };
enum { NCCL_SUM = 0 };

struct NcclApi
{
    void *lib;
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
    int (*Reduce)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char *(*GetErrorString)(int);
    std::string error;
};

NcclApi *ncclApi()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char *names[] = { getenv("SVCM_NCCL_LIB"), "libnccl.so.2", "libnccl.so", NULL };
        api.lib = NULL;
        for(int i = 0; i < 3 && !api.lib; i++)
            if(names[i] && names[i][0]) api.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if(!api.lib) { api.error = std::string("cannot load NCCL: ") + (dlerror() ? dlerror() : "?"); return; }
#define SVCM_NCCL_SYM(field, name) \
        *(void**)(&api.field) = dlsym(api.lib, name); \
        if(!api.field) { api.error = std::string("NCCL symbol missing: ") + name; api.lib = NULL; return; }
        SVCM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        SVCM_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        SVCM_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        SVCM_NCCL_SYM(AllGather, "ncclAllGather")
        SVCM_NCCL_SYM(Reduce, "ncclReduce")
        SVCM_NCCL_SYM(GroupStart, "ncclGroupStart")
        SVCM_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        SVCM_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SVCM_NCCL_SYM
    });
    return &api;
}

thread_local std::string t_createError;

} // namespace

// Device copy of a scene too large for the constant bank: its primitives in hierarchy order and
// the hierarchy itself (svcm_bvh.h). Empty for scenes of up to SVCM_MAX_PRIMITIVES primitives.
struct SceneGeometry
{
    DevPrimG *prims;
    BvhNode  *nodes;
    SceneGeometry() : prims(NULL), nodes(NULL) {}
    void release() { if(prims) cudaFree(prims); if(nodes) cudaFree(nodes); prims = NULL; nodes = NULL; }
    ~SceneGeometry() { release(); }
};

// ------------------------------------------------------------------------------------------
// the renderer handle
// ------------------------------------------------------------------------------------------
struct svcm_renderer
{
    SvcmParams params;
    DevScene   scene;
    SceneGeometry geometry;         // primitives + hierarchy of a large scene (scene.bvh points here)
    int        device;
    cudaStream_t stream;
    bool useVC, useVM, ppm, lightTraceOnly, pathTrace, eyeLight;
    int  resX, resY;
    uint32_t paths, localPaths, localRows;
    float baseRadius;

    // device buffers
    PathQueue queue[4];             // [0], [1]: light pass (and the unidirectional walks); [2], [3]: camera pass
    uint32_t *queueCountsLight, *queueCountsCamera;   // [countsAlloc]
    uint32_t  countsAlloc;
    VertexStore verts;
    uint32_t  vertSlots;            // connection slots per path the store was sized for (maxPathLength - 1)
    uint8_t  *pathVertCount;
    QueryStore queries;             // camera vertices (connection + merge consumers)
    PairStore pairs;
    QuerySortBuffers qsort;
    bool sortQueries;
    ShadowQueue shadow;             // counts: [0] light pass, [1] camera pass
    unsigned long long *counters;   // CNT_WORDS
    unsigned long long *hostCounters; // pinned ring [SVCM_PROFILE_RING][CNT_WORDS]
    int       ringLaunches[SVCM_PROFILE_RING];
    bool      ringTimed[SVCM_PROFILE_RING];
    bool      ringGridSerial[SVCM_PROFILE_RING];   // the grid build ran on the main stream, between light and camera pass
    long long ringHead;             // iterations recorded so far
    long long ringTail;             // first iteration not yet returned by svcm_get_profile
    float    *accum;
    uint8_t  *image8 = NULL;        // W*H*4 bytes, allocated by the first svcm_download_image
    GridBuffers grid;
    uint32_t  gridCells, gridCapacity;
    int32_t  *bboxInit;
    // multi-rank: ONE buffer of worldSize blocks [GatherHeader | posLen | thrVcm | dirVm | contProb];
    // block `rank` is where the local light pass writes, ncclAllGather fills the others in place
    ncclComm_t comm;
    unsigned char *gatherBuf;
    size_t   gatherBytes;
    uint32_t gatherSlots;           // record slots per block the buffer was sized for
    uint32_t gatherStride;          // record slots per block in use (<= gatherSlots)
    uint32_t strideMaxLen;          // maxPathLength the stride was chosen for
    uint32_t lastMaxCount;          // largest per-rank vertex count seen at the last synchronisation (0: unknown)
    uint32_t *segCountDev;          // [8] valid records per block (device, from the headers)
    unsigned long long *devCounts64;   // [8] per-rank vertex counts (device, unclamped)
    unsigned long long *hostCounts64;  // pinned ring [SVCM_PROFILE_RING][8]
    float    *accumReduced;         // root of svcm_reduce_accum: sum over ranks (accum itself stays per-rank)
    bool      reducedValid;
    // the grid is built on its own stream, concurrently with the camera pass
    cudaStream_t gridStream;
    cudaEvent_t  evLightDone, evGridDone;
    // ... and the camera sub-paths on a third one, concurrently with the light pass (they meet at
    // the vertex connection)
    cudaStream_t cameraStream;
    cudaEvent_t  evIterStart, evCameraDone;
    bool overlapPasses;             // SVCM_NO_PASS_OVERLAP=1 serialises them again (A/B)
    bool overlapMerge;              // SVCM_NO_MERGE_OVERLAP=1: merge on the main stream behind the connection kernels
    // merge stage: its own stream behind the camera walk and the grid build. evMergeDone[b]: the merge
    // that read camera-vertex buffer b (and counter / bbox block b) is through, its counters are copied.
    cudaStream_t mergeStream;
    cudaEvent_t evMergeDone[2];
    int  lastMergeBlock;            // which of the two the most recent merge records
    // Camera vertices and the scratch of their sort exist twice when memory allows (queryBuffers == 2):
    // the camera walk of iteration i+1 then writes one set while the merge of iteration i still reads
    // the other. `queries` / `qsort` / `scanGenQuery` are the current iteration's; the *Alt members
    // hold the other set and the two are swapped at the start of every iteration.
    QueryStore queriesAlt;
    QuerySortBuffers qsortAlt;
    int  queryBuffers;
    // ... and the merge of one iteration is left running while the next iteration's light pass starts
    // (nothing it reads is written before the next grid build / camera walk, which wait for it).
    // tailPending: an iteration's merge + counter copy are still only ordered on the camera stream;
    // joinTail() orders the main stream behind them (every entry point but svcm_run_iterations calls it).
    bool overlapIterations;         // SVCM_NO_ITERATION_OVERLAP=1 joins at the end of every iteration again (A/B)
    bool tailPending;
    cudaEvent_t evMainDone;
    unsigned long long *countersBase;   // two blocks of CNT_WORDS; `counters` points at the current iteration's
    int32_t *bboxBlocks;                // two blocks of 6 words; grid.bboxBits points at the current iteration's (the
                                        // pending merge still reads the previous one while the all-gather headers of
                                        // the next are unpacked)
    bool overlapGrid;               // SVCM_NO_GRID_OVERLAP=1: grid build on the main stream (isolated stage times)
    bool spatialQueries;            // SVCM_QUERY_ORDER=bucket restores round 1's order (A/B)
    bool legacyBuild;               // SVCM_LEGACY_BUILD=1: round 1's three-kernel scan + direct record scatter (A/B)

    int  iterationsDone;
    int  mergeGroup;
    uint32_t queriesSizedFor, pairsSizedFor;   // maxPathLength the camera-vertex / pair stores were last sized for
    bool cellLayout;                // SVCM_OPT_GRID_LAYOUT
    int  mergeVariant;              // SVCM_OPT_MERGE_VARIANT
    bool refCandidates;             // SVCM_OPT_REF_CANDIDATES
    ScanGen scanGenGrid, scanGenQuery, scanGenQueryAlt;
    bool profile;
    cudaEvent_t (*ev)[11];         // [SVCM_PROFILE_RING][11], created when profiling is enabled:
                                   // 0 start, 1 light pass done, 4 camera pass done, 5 merge done (main
                                   // stream); 6 grid stream starts, 7 gather done, 8 build done; 9 / 10 merge
                                   // stage starts / ends (on whichever stream it runs)
    cudaEvent_t userEv[SVCM_NUM_USER_EVENTS];
    SvcmStats stats;
    bool statsPending;
    int  lastLaunches;
    std::string error;

    svcm_renderer() { memset(&params, 0, sizeof(params)); memset(&scene, 0, sizeof(scene)); }
};

namespace {

int fail(svcm_renderer *r, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if(r) r->error = buf; else t_createError = buf;
    return code;
}

#define SVCM_CUDA(r, call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) \
    return fail(r, e_ == cudaErrorMemoryAllocation ? SVCM_E_NOMEM : SVCM_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                cudaGetErrorString(e_), __FILE__, __LINE__); } while(0)

int checkDevice(svcm_renderer *r, int device)
{
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if(e != cudaSuccess || n <= 0)
        return fail(r, SVCM_E_NO_DEVICE, "no CUDA device available (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if(device < 0 || device >= n) return fail(r, SVCM_E_INVALID, "device %d out of range (have %d)", device, n);
    return SVCM_OK;
}

// flat scene -> the kernels' constant block; validates every limit and index
int buildDevScene(svcm_renderer *r, const SvcmSceneDesc *d, DevScene &s)
{
    if(!d) return fail(r, SVCM_E_INVALID, "scene is NULL");
    if(d->numPrimitives < 0 || d->numPrimitives > SVCM_MAX_PRIMITIVES_BVH ||
       d->numMaterials < 1 || d->numMaterials > SVCM_MAX_MATERIALS ||
       d->numLights < 1 || d->numLights > SVCM_MAX_LIGHTS)
        return fail(r, SVCM_E_INVALID, "scene exceeds limits: %d primitives (max %d), %d materials (1..%d), %d lights (1..%d)",
                    d->numPrimitives, SVCM_MAX_PRIMITIVES_BVH, d->numMaterials, SVCM_MAX_MATERIALS,
                    d->numLights, SVCM_MAX_LIGHTS);
    if((d->numPrimitives && !d->primitives) || !d->materials || !d->lights || !d->materialToLight)
        return fail(r, SVCM_E_INVALID, "scene arrays must not be NULL");
    memset(&s, 0, sizeof(s));
    s.numPrims = d->numPrimitives; s.numMats = d->numMaterials; s.numLights = d->numLights;
    s.background = d->backgroundLight;
    if(s.background >= s.numLights) return fail(r, SVCM_E_INVALID, "backgroundLight out of range");
    if(s.background >= 0 && d->lights[s.background].type != SVCM_LIGHT_BACKGROUND)
        return fail(r, SVCM_E_INVALID, "backgroundLight does not index a background light");
    for(int i = 0; i < 3; i++) { s.camPos[i] = d->camera.position[i]; s.camFwd[i] = d->camera.forward[i]; }
    s.resX = d->camera.resolution[0]; s.resY = d->camera.resolution[1];
    memcpy(s.rasterToWorld, d->camera.rasterToWorld, sizeof(float) * 16);
    memcpy(s.worldToRaster, d->camera.worldToRaster, sizeof(float) * 16);
    s.imagePlaneDist = d->camera.imagePlaneDist;
    for(int i = 0; i < 3; i++) s.sphereCenter[i] = d->sceneSphere.center[i];
    s.sphereRadius = d->sceneSphere.radius;
    s.sphereInvRadiusSqr = d->sceneSphere.invRadiusSqr;
    for(int i = 0; i < d->numMaterials; i++)
    {
        s.mats[i] = d->materials[i];
        s.matToLight[i] = d->materialToLight[i];
        if(s.matToLight[i] >= d->numLights) return fail(r, SVCM_E_INVALID, "materialToLight[%d] out of range", i);
    }
    for(int i = 0; i < d->numPrimitives; i++)
    {
        const SvcmPrimitive &p = d->primitives[i];
        if(p.type != SVCM_PRIM_TRIANGLE && p.type != SVCM_PRIM_SPHERE)
            return fail(r, SVCM_E_INVALID, "primitive %d has unknown type %d", i, p.type);
        if(p.matID < 0 || p.matID >= d->numMaterials)
            return fail(r, SVCM_E_INVALID, "primitive %d has material %d out of range", i, p.matID);
        if(d->numPrimitives > SVCM_MAX_PRIMITIVES) continue;      // goes to device memory: uploadGeometry
        memcpy(s.prims[i].v, p.v, sizeof(float) * 12);
        s.prims[i].type = p.type; s.prims[i].matID = p.matID;
    }
    s.numRuns = 0; s.runCull = 0u;
    if(d->numPrimitives > 0 && d->numPrimitives <= SVCM_MAX_PRIMITIVES)
    {
        const bool plain = getenv("SVCM_DEBUG_NO_RUNS") && atoi(getenv("SVCM_DEBUG_NO_RUNS")) != 0;   // A/B switch of the tests
        s.numRuns = primitiveRuns(d->primitives, d->numPrimitives, plain, s.runs, &s.runCull);
    }
    for(int i = 0; i < d->numLights; i++)
    {
        const SvcmLight &l = d->lights[i];
        if(l.type < SVCM_LIGHT_AREA || l.type > SVCM_LIGHT_BACKGROUND)
            return fail(r, SVCM_E_INVALID, "light %d has unknown type %d", i, l.type);
        DevLight &o = s.lights[i];
        o.type = l.type;
        memcpy(o.p0, l.p0, 12); memcpy(o.e1, l.e1, 12); memcpy(o.e2, l.e2, 12);
        memcpy(o.fx, l.frameX, 12); memcpy(o.fy, l.frameY, 12); memcpy(o.fz, l.frameZ, 12);
        memcpy(o.intensity, l.intensity, 12);
        if(l.type == SVCM_LIGHT_BACKGROUND)   // radiance = mBackgroundColor * mScale (src/lights.hxx:424)
            for(int k = 0; k < 3; k++) o.intensity[k] = l.intensity[k] * l.scale;
        o.invArea = l.invArea;
    }
    return SVCM_OK;
}

// For a scene buildDevScene accepted with more than SVCM_MAX_PRIMITIVES primitives: builds the
// hierarchy on the host, copies it and the reordered primitives to the current device and points
// `s` at them. No-op for smaller scenes.
int uploadGeometry(svcm_renderer *r, const SvcmSceneDesc *d, DevScene &s, SceneGeometry &g)
{
    s.bvh = NULL; s.bvhPrims = NULL;
    // test hook: SVCM_DEBUG_FORCE_BVH sends small scenes through the hierarchy too, so that the
    // two paths can be compared on the reference's own scenes
    if(d->numPrimitives <= SVCM_MAX_PRIMITIVES && !(getenv("SVCM_DEBUG_FORCE_BVH") && d->numPrimitives > 0)) return SVCM_OK;
    std::vector<BvhNode> nodes; std::vector<int32_t> order;
    if(bvhBuild(d->primitives, d->numPrimitives, nodes, order) != 0)
        return fail(r, SVCM_E_INVALID, "scene geometry is not finite");
    std::vector<DevPrimG> prims(order.size());
    for(size_t slot = 0; slot < order.size(); slot++)
    {
        const SvcmPrimitive &p = d->primitives[order[slot]];
        memcpy(prims[slot].v, p.v, sizeof(float) * 12);
        prims[slot].type = p.type; prims[slot].matID = p.matID; prims[slot].orig = order[slot]; prims[slot].pad = 0;
    }
    g.release();
    SVCM_CUDA(r, cudaMalloc((void**)&g.prims, sizeof(DevPrimG) * prims.size()));
    SVCM_CUDA(r, cudaMalloc((void**)&g.nodes, sizeof(BvhNode) * nodes.size()));
    SVCM_CUDA(r, cudaMemcpy(g.prims, prims.data(), sizeof(DevPrimG) * prims.size(), cudaMemcpyHostToDevice));
    SVCM_CUDA(r, cudaMemcpy(g.nodes, nodes.data(), sizeof(BvhNode) * nodes.size(), cudaMemcpyHostToDevice));
    s.bvh = g.nodes; s.bvhPrims = g.prims;
    return SVCM_OK;
}

template<typename T> int devAlloc(svcm_renderer *r, T **p, size_t count)
{
    *p = NULL;
    if(count == 0) count = 1;
    SVCM_CUDA(r, cudaMalloc((void**)p, sizeof(T) * count));
    return SVCM_OK;
}
template<typename T> void devFree(T *&p) { if(p) cudaFree(p); p = NULL; }

int allocQueue(svcm_renderer *r, PathQueue &q, size_t n)
{
    int rc;
    if((rc = devAlloc(r, &q.s0, n))) return rc;
    if((rc = devAlloc(r, &q.s1, n))) return rc;
    if((rc = devAlloc(r, &q.s2, n))) return rc;
    if((rc = devAlloc(r, &q.s3, n))) return rc;
    return SVCM_OK;
}
void freeQueue(PathQueue &q) { devFree(q.s0); devFree(q.s1); devFree(q.s2); devFree(q.s3); }

void freeGather(svcm_renderer *r)
{
    devFree(r->gatherBuf);
    r->gatherBytes = 0; r->gatherSlots = 0; r->gatherStride = 0;
}
void freeVerts(svcm_renderer *r)
{
    if(r->gatherBuf)
    {
        // the merge streams live inside the gather buffer
        r->verts.posLen = r->verts.thrVcm = r->verts.dirVm = NULL; r->verts.contProb = NULL;
        freeGather(r);
    }
    devFree(r->verts.posLen); devFree(r->verts.thrVcm); devFree(r->verts.dirVm); devFree(r->verts.contProb);
    devFree(r->verts.conn);
    r->verts.capacity = 0; r->vertSlots = 0; r->verts.connSlots = 0;
}

// bytes of one all-gather block holding `slots` records, and the stream offsets inside it
size_t gatherBlockBytes(uint32_t slots) { return (sizeof(GatherHeader) + size_t(slots) * 52u + 255u) & ~size_t(255u); }

// points the local merge streams at block `rank` of the gather buffer laid out for `stride` slots
void pointGatherStreams(svcm_renderer *r, uint32_t stride)
{
    unsigned char *block = r->gatherBuf + size_t(r->params.rank) * gatherBlockBytes(stride);
    r->verts.posLen   = reinterpret_cast<float4*>(block + sizeof(GatherHeader));
    r->verts.thrVcm   = r->verts.posLen + stride;
    r->verts.dirVm    = r->verts.thrVcm + stride;
    r->verts.contProb = reinterpret_cast<float*>(r->verts.dirVm + stride);
    r->verts.capacity = stride;
    r->gatherStride   = stride;
}
void freeQueries(svcm_renderer *r)
{
    devFree(r->queries.q0); devFree(r->queries.q1); devFree(r->queries.q2); devFree(r->queries.q3);
    devFree(r->queries.q4); devFree(r->pairs.pairs);
    devFree(r->qsort.key); devFree(r->qsort.order);
    devFree(r->queriesAlt.q0); devFree(r->queriesAlt.q1); devFree(r->queriesAlt.q2); devFree(r->queriesAlt.q3);
    devFree(r->queriesAlt.q4); devFree(r->qsortAlt.key); devFree(r->qsortAlt.order);
    r->queries.capacity = 0; r->queriesAlt.capacity = 0; r->pairs.capacity = 0; r->pairsSizedFor = 0;
    r->queryBuffers = 1;
}
void freeGridSorted(svcm_renderer *r)
{
    devFree(r->grid.cellKey); devFree(r->grid.sPosLen); devFree(r->grid.sThrVcm); devFree(r->grid.sDirVm);
    devFree(r->grid.sContProb); devFree(r->grid.sIndex);
    r->gridCapacity = 0;
}
int ensureGridSorted(svcm_renderer *r, uint32_t capacity)
{
    if(capacity <= r->gridCapacity) return SVCM_OK;
    freeGridSorted(r);
    int rc;
    if((rc = devAlloc(r, &r->grid.cellKey, capacity))) return rc;
    if((rc = devAlloc(r, &r->grid.sPosLen, capacity))) return rc;
    if((rc = devAlloc(r, &r->grid.sThrVcm, capacity))) return rc;
    if((rc = devAlloc(r, &r->grid.sDirVm, capacity))) return rc;
    if((rc = devAlloc(r, &r->grid.sContProb, capacity))) return rc;
    if((rc = devAlloc(r, &r->grid.sIndex, capacity))) return rc;
    r->gridCapacity = capacity;
    return SVCM_OK;
}

// key space of the cell layout: 2^21 storage cells (8 MB per table), independent of the resolution
// because the cell edge follows the radius, not the pixel count
static const uint32_t kCellTableCapacity = 1u << 21;

int allocCellTables(svcm_renderer *r, GridBuffers &grid, QuerySortBuffers *qs, uint32_t numCells)
{
    int rc;
    grid.tcap = kCellTableCapacity;
    if(!grid.cellEnd)
    {
        if((rc = devAlloc(r, &grid.cellEnd, size_t(grid.tcap) + 2 + 64))) return rc;
        if((rc = devAlloc(r, &grid.dims, 1))) return rc;
        if((rc = devAlloc(r, &grid.scanState, kScanStateWords))) return rc;
        if((rc = devAlloc(r, &grid.scanCounter, 1))) return rc;
        SVCM_CUDA(r, cudaMemset(grid.scanState, 0, sizeof(unsigned long long) * kScanStateWords));
        SVCM_CUDA(r, cudaMemset(grid.scanCounter, 0, sizeof(uint32_t)));
        SVCM_CUDA(r, cudaMemset(grid.dims, 0, sizeof(GridDims)));
    }
    if(!grid.refCount) if((rc = devAlloc(r, &grid.refCount, size_t(numCells)))) return rc;
    if(qs && !qs->cellEnd)
    {
        if((rc = devAlloc(r, &qs->cellEnd, size_t(grid.tcap) + 2 + 64))) return rc;
        if((rc = devAlloc(r, &qs->scanState, kScanStateWords))) return rc;
        if((rc = devAlloc(r, &qs->scanCounter, 1))) return rc;
        SVCM_CUDA(r, cudaMemset(qs->scanState, 0, sizeof(unsigned long long) * kScanStateWords));
        SVCM_CUDA(r, cudaMemset(qs->scanCounter, 0, sizeof(uint32_t)));
    }
    return SVCM_OK;
}

int ensureGridCells(svcm_renderer *r, uint32_t numCells)
{
    if(numCells == r->gridCells) return SVCM_OK;
    devFree(r->grid.cellCount); devFree(r->grid.blockSums); devFree(r->grid.refCount);
    int rc;
    if((rc = devAlloc(r, &r->grid.cellCount, size_t(numCells) + 1))) return rc;
    if((rc = devAlloc(r, &r->grid.blockSums, (size_t(numCells) + 1) / 2048 + 2))) return rc;
    if((rc = allocCellTables(r, r->grid, &r->qsort, numCells))) return rc;
    r->gridCells = numCells;
    return SVCM_OK;
}

// Sizes everything that depends on the path-length limit. Worst case is one stored vertex
// per light bounce and one query per camera bounce; if that does not fit comfortably the
// stores are bounded and an overflow is reported as an error at the next synchronisation.
int ensureCapacity(svcm_renderer *r, uint32_t maxLen)
{
    const uint32_t lightBounces = std::max(1u, maxLen > 0 ? maxLen - 1u : 0u);
    const uint32_t camBounces = std::max(1u, maxLen);
    if(lightBounces > 255u) return fail(r, SVCM_E_INVALID, "maxPathLength %u too large (max 256)", maxLen);
    int rc;

    const uint32_t needCounts = camBounces + 3u;
    if(needCounts > r->countsAlloc)
    {
        devFree(r->queueCountsLight); devFree(r->queueCountsCamera);
        if((rc = devAlloc(r, &r->queueCountsLight, needCounts))) return rc;
        if((rc = devAlloc(r, &r->queueCountsCamera, needCounts))) return rc;
        r->countsAlloc = needCounts;
    }
    if(!r->shadow.counts) if((rc = devAlloc(r, &r->shadow.counts, 2))) return rc;

    size_t freeB = 0, totalB = 0;      // queried only where something is (re)allocated

    const bool needVerts = r->useVC || r->useVM;
    if(needVerts && (r->vertSlots < lightBounces || r->verts.capacity == 0))
    {
        freeVerts(r);
        freeGridSorted(r);
        cudaMemGetInfo(&freeB, &totalB);
        const uint64_t worst = uint64_t(r->localPaths) * lightBounces;
        const uint32_t ws = uint32_t(r->params.worldSize);
        // the path-major connection records have a slot for every possible vertex, whatever is stored
        if(r->useVC)
        {
            if((rc = devAlloc(r, &r->verts.conn, size_t(r->localPaths) * lightBounces * 4u))) return rc;
            r->verts.connSlots = lightBounces;
            cudaMemGetInfo(&freeB, &totalB);
        }
        // bytes per stored vertex: merge record 52, plus per rank the gathered copy (sharded runs)
        // and the cell-sorted copy with its key
        const uint64_t perVertex = 52u + (r->useVM ? (ws > 1u ? 112u * ws : 60u) : 0u) + 6u;
        const uint64_t budget = uint64_t(double(freeB) * 0.45);
        uint64_t cap = std::min<uint64_t>(worst, budget / perVertex);
        cap = std::min<uint64_t>(cap, 0xFFFFFF00u / ws);
        cap = std::max<uint64_t>(cap, std::min<uint64_t>(worst, uint64_t(r->localPaths)));
        cap = (cap + 255u) & ~uint64_t(255u);
        if(ws > 1u && r->useVM)
        {
            // all-gather blocks, sized once (no reallocation inside the iteration loop)
            r->gatherBytes = gatherBlockBytes(uint32_t(cap)) * ws;
            if((rc = devAlloc(r, &r->gatherBuf, r->gatherBytes))) return rc;
            r->gatherSlots = uint32_t(cap);
            pointGatherStreams(r, uint32_t(cap));
            r->lastMaxCount = 0; r->strideMaxLen = 0;
        }
        else
        {
            if((rc = devAlloc(r, &r->verts.posLen, cap))) return rc;
            if((rc = devAlloc(r, &r->verts.thrVcm, cap))) return rc;
            if((rc = devAlloc(r, &r->verts.dirVm, cap))) return rc;
            if((rc = devAlloc(r, &r->verts.contProb, cap))) return rc;
        }
        r->verts.capacity = uint32_t(cap);
        r->vertSlots = lightBounces;
        if(r->useVM)
            if((rc = ensureGridSorted(r, uint32_t(std::min<uint64_t>(cap * ws, 0xFFFFFF00u))))) return rc;
    }
    if(r->gatherBuf)
    {
        // Slots per block for this call: the whole capacity until a synchronisation has shown the real
        // counts, then the largest count + 12.5 % (they wobble by a fraction of a percent between
        // iterations). Every rank sees the same headers, hence decides the same stride.
        uint32_t stride = r->gatherSlots;
        if(r->lastMaxCount && maxLen <= r->strideMaxLen)
            stride = uint32_t(std::min<uint64_t>(r->gatherSlots, (uint64_t(r->lastMaxCount) + r->lastMaxCount / 8u + 4096u + 255u) & ~uint64_t(255u)));
        else { r->strideMaxLen = maxLen; r->lastMaxCount = 0; }
        if(stride != r->gatherStride) pointGatherStreams(r, stride);
    }

    // camera vertices: at most one per camera bounce that can still connect or merge
    if(!r->lightTraceOnly && !r->pathTrace && !r->eyeLight)
    {
        const uint64_t worst = uint64_t(r->localPaths) * (r->ppm ? 1u : std::max(1u, camBounces - 1u));
        // (a store bounded by free memory stays smaller than the worst case for good: only a longer
        // path limit is a reason to size it again, not every call)
        if(r->queries.capacity < worst && r->queries.capacity < 0xFFFFFF00u && maxLen > r->queriesSizedFor)
        {
            r->queriesSizedFor = maxLen;
            const uint32_t old = r->queries.capacity;
            freeQueries(r);
            cudaMemGetInfo(&freeB, &totalB);
            const uint64_t perVertex = 80u + (r->useVC ? 8u * 12u : 0u);
            uint64_t cap = std::min<uint64_t>(worst, uint64_t(double(freeB) * 0.4) / perVertex);
            cap = std::min<uint64_t>(cap, 0xFFFFFF00u);
            cap = std::max<uint64_t>(cap, old);
            // a second set for the iteration overlap, only when the worst case fits twice with room to spare
            const bool twice = r->useVM && cap == worst && 2u * worst * (80u + 8u) < uint64_t(double(freeB) * 0.25) &&
                               !(getenv("SVCM_SINGLE_QUERY_BUFFER") && atoi(getenv("SVCM_SINGLE_QUERY_BUFFER")) != 0);
            for(int b = 0; b < (twice ? 2 : 1); b++)
            {
                QueryStore &q = b ? r->queriesAlt : r->queries;
                QuerySortBuffers &qs = b ? r->qsortAlt : r->qsort;
                if((rc = devAlloc(r, &q.q0, cap))) return rc;
                if((rc = devAlloc(r, &q.q1, cap))) return rc;
                if((rc = devAlloc(r, &q.q2, cap))) return rc;
                if((rc = devAlloc(r, &q.q3, cap))) return rc;
                if(r->useVC) if((rc = devAlloc(r, &q.q4, cap))) return rc;
                if(r->useVM)
                {
                    if((rc = devAlloc(r, &qs.key, cap))) return rc;
                    if((rc = devAlloc(r, &qs.order, cap))) return rc;
                }
                q.capacity = uint32_t(cap);
            }
            r->queryBuffers = twice ? 2 : 1;
        }
    }
    // (camera vertex, light vertex) pairs: a camera vertex of length c pairs with at most
    // maxLen-1-c light vertices (see k_cameraBounce), i.e. sum_{c=1}^{maxLen-2} (maxLen-1-c) per path
    if(r->useVC)
    {
        uint64_t perPath = 0;
        for(uint32_t c = 1; c + 1u < maxLen; c++) perPath += std::min<uint64_t>(lightBounces, maxLen - 1u - c);
        const uint64_t worst = std::max<uint64_t>(1, uint64_t(r->localPaths) * perPath);
        if(r->pairs.capacity < worst && r->pairs.capacity < 0xFFFFFF00u && maxLen > r->pairsSizedFor)
        {
            r->pairsSizedFor = maxLen;
            const uint32_t old = r->pairs.capacity;
            devFree(r->pairs.pairs);
            cudaMemGetInfo(&freeB, &totalB);
            uint64_t cap = std::min<uint64_t>(worst, uint64_t(double(freeB) * 0.25) / 8u);
            cap = std::min<uint64_t>(cap, 0xFFFFFF00u);
            cap = std::max<uint64_t>(cap, old);
            // test hook: a deliberately small pair list must surface as an error, not as a dim image
            if(getenv("SVCM_DEBUG_PAIR_CAP")) cap = std::max<uint64_t>(32u, strtoull(getenv("SVCM_DEBUG_PAIR_CAP"), NULL, 10));
            if((rc = devAlloc(r, &r->pairs.pairs, cap))) return rc;
            r->pairs.capacity = uint32_t(cap);
        }
    }
    // visibility segments: one per light vertex splat, one per direct-illumination sample and
    // one per pair at most; sized for the measured mix (scene 1 vcm: 7.9 per path), a full
    // queue makes producers trace inline
    if((r->useVC || r->lightTraceOnly || r->pathTrace) && r->shadow.capacity == 0)
    {
        cudaMemGetInfo(&freeB, &totalB);
        uint64_t cap = uint64_t(r->localPaths) * ((r->lightTraceOnly || r->pathTrace) ? std::min(camBounces, 4u) : 12u);
        cap = std::min<uint64_t>(cap, uint64_t(double(freeB) * 0.3) / 48u);
        cap = std::min<uint64_t>(std::max<uint64_t>(cap, 1024u), 0xFFFFFF00u);
        // test hook: a deliberately small queue exercises the producers' inline fallback
        if(getenv("SVCM_DEBUG_SHADOW_CAP")) cap = std::max<uint64_t>(32u, strtoull(getenv("SVCM_DEBUG_SHADOW_CAP"), NULL, 10));
        if((rc = devAlloc(r, &r->shadow.r0, cap))) return rc;
        if((rc = devAlloc(r, &r->shadow.r1, cap))) return rc;
        if((rc = devAlloc(r, &r->shadow.r2, cap))) return rc;
        r->shadow.capacity = uint32_t(cap);
    }
    if(r->useVM)
    {
        if((rc = ensureGridCells(r, r->paths))) return rc;
        for(int b = 0; b < r->queryBuffers; b++)
        {
            QuerySortBuffers &qs = b ? r->qsortAlt : r->qsort;
            if(b && (rc = allocCellTables(r, r->grid, &qs, r->paths))) return rc;
            if(!qs.cellCount)
            {
                if((rc = devAlloc(r, &qs.cellCount, size_t(r->paths) + 1))) return rc;
                if((rc = devAlloc(r, &qs.blockSums, (size_t(r->paths) + 1) / 2048 + 2))) return rc;
            }
        }
    }
    return SVCM_OK;
}

// the scalars of reference src/vertexcm.hxx:288-308, computed on the host in the same order
void iterationConstants(const svcm_renderer *r, int iteration, uint32_t minLen, uint32_t maxLen, IterConst &ic)
{
    memset(&ic, 0, sizeof(ic));
    const float lightSubPathCount = float(r->resX * r->resY);
    float radius = r->baseRadius;
    radius /= std::pow(float(iteration + 1), 0.5f * (1 - r->params.radiusAlpha));
    radius = std::max(radius, 1e-7f);
    const float radiusSqr = radius * radius;
    const float kPi = 3.14159265358979f;
    ic.radius = radius;
    ic.radiusSqr = radiusSqr;
    ic.invCellSize = 1.f / (radius * 2.f);                       // src/hashgrid.hxx:47-48
    ic.vmNormalization = 1.f / (radiusSqr * kPi * lightSubPathCount);
    const float etaVCM = (kPi * radiusSqr) * lightSubPathCount;
    ic.misVmWeightFactor = r->useVM ? etaVCM : 0.f;
    ic.misVcWeightFactor = r->useVC ? 1.f / etaVCM : 0.f;
    ic.lightSubPathCount = lightSubPathCount;
    ic.minLen = minLen; ic.maxLen = maxLen;
    ic.keyLight  = rngKey(r->params.seed, iteration, 0);
    ic.keyCamera = rngKey(r->params.seed, iteration, 1);
    ic.useVC = r->useVC; ic.useVM = r->useVM; ic.ppm = r->ppm; ic.lightTraceOnly = r->lightTraceOnly;
    ic.resX = r->resX; ic.resY = r->resY;
    ic.rank = r->params.rank; ic.worldSize = r->params.worldSize;
    ic.localPaths = r->localPaths;
    ic.numCells = r->paths;
    ic.hashMagic = hashMagicFor(ic.numCells);
}

int ncclCheck(svcm_renderer *r, int rc, const char *what)
{
    if(rc == NCCL_SUCCESS) return SVCM_OK;
    return fail(r, SVCM_E_NCCL, "%s failed: %s", what, ncclApi()->GetErrorString ? ncclApi()->GetErrorString(rc) : "?");
}

// All-gather of the merge records of every rank's light vertices: ONE ncclAllGather of
// fixed-size blocks, in place (block `rank` is where the local light pass wrote). Count and bbox of
// every rank ride in the block headers, so there is no host round trip and no second collective;
// everything is enqueued on `st` and the call returns at once.
int gatherLightVertices(svcm_renderer *r, MergeSet &set, cudaStream_t st, int slot)
{
    NcclApi *n = ncclApi();
    const int ws = r->params.worldSize;
    if(!r->comm) return fail(r, SVCM_E_STATE, "worldSize %d but svcm_comm_init was not called", ws);
    const uint32_t stride = r->gatherStride;
    const size_t blockBytes = gatherBlockBytes(stride);

    MergeSet local; memset(&local, 0, sizeof(local));
    local.posLen = r->verts.posLen; local.thrVcm = r->verts.thrVcm; local.dirVm = r->verts.dirVm;
    local.contProb = r->verts.contProb;
    local.singleCount = r->counters + CNT_LIGHT_VERTS;
    local.capacity = stride; local.segStride = stride; local.numSegs = 1;
    unsigned char *mine = r->gatherBuf + size_t(r->params.rank) * blockBytes;
    launchGatherPackHeader(local, r->bboxInit, reinterpret_cast<GatherHeader*>(mine), st);

    int rc;
    if((rc = ncclCheck(r, n->AllGather(mine, r->gatherBuf, blockBytes, NCCL_UINT8, r->comm, st),
                       "ncclAllGather(light vertices)"))) return rc;
    launchGatherUnpackHeaders(r->gatherBuf, blockBytes, uint32_t(ws), stride, r->segCountDev, r->grid.bboxBits,
                              r->devCounts64, st);
    SVCM_CUDA(r, cudaMemcpyAsync(r->hostCounts64 + size_t(slot) * 8, r->devCounts64, sizeof(unsigned long long) * 8,
                                 cudaMemcpyDeviceToHost, st));

    memset(&set, 0, sizeof(set));
    set.posLen   = reinterpret_cast<const float4*>(r->gatherBuf + sizeof(GatherHeader));
    set.thrVcm   = set.posLen + stride;
    set.dirVm    = set.thrVcm + stride;
    set.contProb = reinterpret_cast<const float*>(set.dirVm + stride);
    set.singleCount = NULL;
    set.capacity = stride * uint32_t(ws);
    set.segStride = stride;
    set.numSegs = uint32_t(ws);
    set.segBytes = blockBytes;
    set.segCountDev = r->segCountDev;
    return SVCM_OK;
}

// The visibility queue is one allocation; the light pass (one segment per splatted light vertex)
// and the camera pass (direct illumination + vertex connections) fill disjoint parts of it, so the
// two passes can run concurrently. counts[0] / counts[1] are their fill levels.
ShadowQueue lightShadow(const svcm_renderer *r)
{
    ShadowQueue q = r->shadow;
    if(!r->lightTraceOnly && !r->pathTrace) q.capacity = r->shadow.capacity / 4u;
    return q;
}
ShadowQueue cameraShadow(const svcm_renderer *r)
{
    ShadowQueue q = r->shadow;
    if(r->lightTraceOnly || r->pathTrace) return q;
    const uint32_t skip = r->shadow.capacity / 4u;
    q.r0 += skip; q.r1 += skip; q.r2 += skip;
    q.capacity = r->shadow.capacity - skip;
    return q;
}

// the camera walk (src/vertexcm.hxx:415-545 without connection and merging): maxLen bounce kernels
int enqueueCameraBounces(svcm_renderer *r, const IterConst &ic, const TraceArgs &base, uint32_t maxLen, cudaStream_t cs,
                         cudaEvent_t *ev)
{
    TraceArgs c = base;
    c.queueCounts = r->queueCountsCamera;
    c.shadow = cameraShadow(r);
    const uint32_t camBounces = std::max(1u, maxLen);
    if(ev) cudaEventRecord(ev[2], cs);
    for(uint32_t b = 1; b <= camBounces; b++)
    {
        c.in = r->queue[2 + ((b + 1) & 1)]; c.out = r->queue[2 + (b & 1)];
        launchCameraBounce(r->scene, ic, c, b, r->localPaths, cs);
    }
    if(ev) cudaEventRecord(ev[3], cs);
    return int(camBounces);
}

// The deferred merging of all recorded camera vertices: query sort (bbox early-out, spatial order,
// count) + merge kernel, on stream `ms`, which must already be ordered behind the camera walk and
// the grid build. Returns the number of kernels launched.
int enqueueMerge(svcm_renderer *r, const IterConst &ic, cudaStream_t ms)
{
    int launches = 0;
    QueryStore qs = r->queries;
    // the queries take a cell-coherent order; the keying kernel applies HashGrid::Process' bbox
    // early-out and counts the queries
    const unsigned long long *count = r->counters + CNT_QUERIES;
    if(r->cellLayout)
    {
        launches += launchCellQuerySort(ic, r->queries, r->counters + CNT_CAM_VERTS, 0, r->grid, r->qsort,
                                        r->scanGenQuery, r->counters, ms);
        qs.order = r->qsort.order;
        GridBuffers grid = r->grid;
        if(!r->refCandidates) grid.refCount = NULL;
        launchCellMerge(r->scene, ic, qs, count, 0, grid, r->accum, NULL, NULL, r->counters, r->mergeVariant, ms);
    }
    else
    {
        launches += launchQuerySort(ic, r->queries, r->counters + CNT_CAM_VERTS, r->grid.bboxBits, r->qsort,
                                    r->counters, r->legacyBuild ? NULL : &r->scanGenQuery, r->spatialQueries, ms);
        qs.order = r->qsort.order;
        if(r->mergeGroup >= 2 && r->mergeGroup <= 5)
            launchBucketStagedMerge(r->scene, ic, qs, count, 0, r->grid, r->accum, NULL, NULL, r->counters,
                                    r->mergeGroup, ms);
        else
            launchMerge(r->scene, ic, qs, count, 0, r->grid, r->accum, NULL, NULL,
                        r->counters, r->mergeGroup, ms);
    }
    return launches + 1;
}

// Orders the main stream behind an iteration tail left on the camera stream (overlapIterations).
int joinTail(svcm_renderer *r)
{
    if(r->tailPending)
    {
        // the merge stream runs in order: the most recent merge is the last to finish
        SVCM_CUDA(r, cudaStreamWaitEvent(r->stream, r->evMergeDone[r->lastMergeBlock], 0));
        r->tailPending = false;
    }
    return SVCM_OK;
}

// One iteration: the kernel schedule that replaces the two serial per-pixel loops
static double hostNowMs()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int runOneIteration(svcm_renderer *r, int iteration, uint32_t minLen, uint32_t maxLen)
{
    static const bool traceHost = getenv("SVCM_TRACE_HOST") != NULL;
    double th[8] = { 0 }; th[0] = hostNowMs();
    IterConst ic;
    iterationConstants(r, iteration, minLen, maxLen, ic);
    cudaStream_t st = r->stream;
    int launches = 0;
    const int slot = int(r->ringHead % SVCM_PROFILE_RING);
    const bool timed = r->profile && r->ev != NULL;
    cudaEvent_t *ev = timed ? r->ev[slot] : NULL;

    // the counters of two consecutive iterations are distinct blocks: the tail of the previous one may
    // still be counting (overlapIterations)
    const int blk = int(r->ringHead & 1);
    r->counters = r->countersBase + size_t(blk) * CNT_WORDS;
    r->grid.bboxBits = r->bboxBlocks + size_t(blk) * 8;
    if(r->queryBuffers == 2)
    {
        // this iteration's camera vertices go to the set the merge of TWO iterations ago read (it also
        // counted into this counter block): wait for it, whatever else is still pending
        if(r->tailPending) SVCM_CUDA(r, cudaStreamWaitEvent(st, r->evMergeDone[blk], 0));
        std::swap(r->queries, r->queriesAlt); std::swap(r->qsort, r->qsortAlt); std::swap(r->scanGenQuery, r->scanGenQueryAlt);
    }
    SVCM_CUDA(r, cudaMemsetAsync(r->counters, 0, sizeof(unsigned long long) * CNT_WORDS, st));
    SVCM_CUDA(r, cudaMemsetAsync(r->queueCountsLight, 0, sizeof(uint32_t) * r->countsAlloc, st));
    SVCM_CUDA(r, cudaMemsetAsync(r->queueCountsCamera, 0, sizeof(uint32_t) * r->countsAlloc, st));
    SVCM_CUDA(r, cudaMemsetAsync(r->shadow.counts, 0, sizeof(uint32_t) * 2, st));
    if(r->useVC) SVCM_CUDA(r, cudaMemsetAsync(r->pathVertCount, 0, r->localPaths, st));
    if(timed) cudaEventRecord(ev[0], st);
    th[1] = hostNowMs();

    TraceArgs a;
    memset(&a, 0, sizeof(a));
    a.verts = r->verts;
    a.pathVertCount = r->pathVertCount;
    a.queries = r->queries;
    a.pairs = r->pairs;
    a.shadow = r->shadow;
    a.counters = r->counters;
    a.accum = r->accum;
    a.bboxBits = r->grid.bboxBits;

    // ---- the unidirectional integrators ---------------------------------
    if(r->pathTrace || r->eyeLight)
    {
        if(timed) { cudaEventRecord(ev[1], st); cudaEventRecord(ev[6], st); cudaEventRecord(ev[7], st); cudaEventRecord(ev[8], st); }
        a.queueCounts = r->queueCountsCamera;
        if(r->eyeLight)
        {
            launchEyeLight(r->scene, ic, a, uint32_t(iteration), st);
            launches++;
        }
        else
        {
            const uint32_t bounces = std::max(1u, maxLen);
            for(uint32_t b = 1; b <= bounces; b++)
            {
                a.in = r->queue[(b + 1) & 1]; a.out = r->queue[b & 1];
                launchPathTraceBounce(r->scene, ic, a, b, r->localPaths, st);
                launches++;
            }
            launchShadow(r->scene, ic, a, 1, st);      // next-event visibility, then splat
            launches++;
        }
        if(timed) { cudaEventRecord(ev[4], st); cudaEventRecord(ev[9], st); cudaEventRecord(ev[10], st); cudaEventRecord(ev[5], st); }
        SVCM_CUDA(r, cudaMemcpyAsync(r->hostCounters + size_t(slot) * CNT_WORDS, r->counters,
                                     sizeof(unsigned long long) * CNT_WORDS, cudaMemcpyDeviceToHost, st));
        SVCM_CUDA(r, cudaGetLastError());
        r->ringLaunches[slot] = launches;
        r->ringTimed[slot] = timed;
        r->ringHead++;
        if(r->ringHead - r->ringTail > SVCM_PROFILE_RING) r->ringTail = r->ringHead - SVCM_PROFILE_RING;
        r->stats.radius = 0.f;
        r->statsPending = true;
        r->lastLaunches = launches;
        r->iterationsDone++;
        return SVCM_OK;
    }

    // ---- fork: the camera walk runs on its own stream, concurrently with the light pass ---------
    if(!r->lightTraceOnly && r->overlapPasses)
    {
        SVCM_CUDA(r, cudaEventRecord(r->evIterStart, st));
        SVCM_CUDA(r, cudaStreamWaitEvent(r->cameraStream, r->evIterStart, 0));
        // one set of camera vertices: the previous iteration's merge still reads what this walk overwrites
        if(r->tailPending && r->queryBuffers != 2) SVCM_CUDA(r, cudaStreamWaitEvent(r->cameraStream, r->evMergeDone[blk ^ 1], 0));
        launches += enqueueCameraBounces(r, ic, a, maxLen, r->cameraStream, timed ? ev : NULL);
        SVCM_CUDA(r, cudaEventRecord(r->evCameraDone, r->cameraStream));
    }

    // ---- light sub-paths (reference src/vertexcm.hxx:321-396) -----------------------------
    const uint32_t lightBounces = std::max(1u, maxLen > 0 ? maxLen - 1u : 0u);
    a.queueCounts = r->queueCountsLight;
    a.shadow = lightShadow(r);
    for(uint32_t b = 1; b <= lightBounces; b++)
    {
        a.in = r->queue[(b + 1) & 1]; a.out = r->queue[b & 1];
        launchLightBounce(r->scene, ic, a, b, r->localPaths, st);
        launches++;
    }
    if(r->useVC || r->lightTraceOnly)
    {
        launchShadow(r->scene, ic, a, 0, st);     // visibility of all camera connections, then splat
        launches++;
    }
    if(timed) cudaEventRecord(ev[1], st);
    th[2] = hostNowMs();

    // ---- hash grid over ALL light vertices (src/vertexcm.hxx:403-408) ----------------------
    // On its own stream: the camera pass needs nothing from it (its vertices are tested against the
    // particle bbox when they are keyed for the merge), so the all-gather and the build run
    // concurrently with the camera sub-paths and only the merge waits for them.
    if(r->useVM)
    {
        cudaStream_t sg = r->overlapGrid ? r->gridStream : st;      // same stream: strictly after the light pass, before the camera pass joins
        SVCM_CUDA(r, cudaEventRecord(r->evLightDone, st));
        SVCM_CUDA(r, cudaStreamWaitEvent(sg, r->evLightDone, 0));
        if(timed) cudaEventRecord(ev[6], sg);
        MergeSet set;
        if(r->params.worldSize > 1)
        {
            const int rc = gatherLightVertices(r, set, sg, slot);
            if(rc) return rc;
            launches += 4;
        }
        else
        {
            memset(&set, 0, sizeof(set));
            set.posLen = r->verts.posLen; set.thrVcm = r->verts.thrVcm; set.dirVm = r->verts.dirVm;
            set.contProb = r->verts.contProb;
            set.singleCount = r->counters + CNT_LIGHT_VERTS;
            set.capacity = r->verts.capacity;
            set.segStride = r->verts.capacity; set.numSegs = 1;
        }
        if(timed) cudaEventRecord(ev[7], sg);
        // the previous iteration's merge still reads the grid this build overwrites (the all-gather
        // above touches nothing of it: the bounding box lives in the other block)
        if(r->tailPending) SVCM_CUDA(r, cudaStreamWaitEvent(sg, r->evMergeDone[blk ^ 1], 0));
        if(r->cellLayout) launches += launchCellGridBuild(ic, set, r->grid, r->scanGenGrid, r->refCandidates, sg);
        else launches += launchGridBuild(ic, set, r->grid, r->legacyBuild ? NULL : &r->scanGenGrid, sg);
        if(timed) cudaEventRecord(ev[8], sg);
        SVCM_CUDA(r, cudaEventRecord(r->evGridDone, sg));
    }
    else if(timed) { cudaEventRecord(ev[6], st); cudaEventRecord(ev[7], st); cudaEventRecord(ev[8], st); }
    // the merge goes behind the camera walk on its stream as soon as the grid is enqueued
    bool mergeForked = false;
    if(r->useVM && !r->lightTraceOnly && r->overlapPasses && r->overlapMerge)
    {
        cudaStream_t ms = r->mergeStream;
        SVCM_CUDA(r, cudaStreamWaitEvent(ms, r->evCameraDone, 0));
        SVCM_CUDA(r, cudaStreamWaitEvent(ms, r->evGridDone, 0));
        if(timed) cudaEventRecord(ev[9], ms);
        launches += enqueueMerge(r, ic, ms);
        if(timed) cudaEventRecord(ev[10], ms);
        SVCM_CUDA(r, cudaEventRecord(r->evMergeDone[blk], ms));
        r->lastMergeBlock = blk;
        mergeForked = true;
    }
    th[3] = hostNowMs();

    // ---- camera sub-paths (src/vertexcm.hxx:415-545) ---------------------------------------
    // The walk itself needs nothing from the light pass: its bounces were enqueued on the camera
    // stream before the light pass (below `forkCamera`); here the two passes meet for the vertex
    // connection of every (camera vertex, stored light vertex of the same path) pair.
    if(!r->lightTraceOnly)
    {
        if(r->overlapPasses) SVCM_CUDA(r, cudaStreamWaitEvent(st, r->evCameraDone, 0));
        else launches += enqueueCameraBounces(r, ic, a, maxLen, st, timed ? ev : NULL);
        if(r->useVC)
        {
            TraceArgs c = a;
            c.queueCounts = r->queueCountsCamera;
            c.shadow = cameraShadow(r);
            launchConnect(r->scene, ic, c, st);       // pair list, then every (camera vertex, light vertex) pair
            launchShadow(r->scene, ic, c, 1, st);     // direct-illumination + connection visibility
            launches += 3;
        }
    }
    if(timed) cudaEventRecord(ev[4], st);
    th[4] = hostNowMs();

    // ---- deferred merging of all queued camera vertices ------------------------------------
    // Merging reads the camera vertices and the grid and adds to the image with atomics: it depends on
    // neither vertex connection nor its visibility tests. With the passes overlapped it was enqueued on
    // the camera stream right behind the walk (`mergeForked` below) and runs concurrently with the
    // connection kernels of this stream; here the two meet again.
    // With overlapIterations the merge is NOT joined here: the main stream goes on to the next
    // iteration's light pass, the camera stream finishes this one (merge, then — once the main stream's
    // part is through as well — the counter copy) and whoever needs the result joins (joinTail, the next
    // grid build, the next camera walk by stream order).
    bool leaveTail = false;
    if(r->useVM && !r->lightTraceOnly)
    {
        if(mergeForked && r->overlapIterations) leaveTail = true;
        else if(mergeForked) SVCM_CUDA(r, cudaStreamWaitEvent(st, r->evMergeDone[blk], 0));
        else
        {
            SVCM_CUDA(r, cudaStreamWaitEvent(st, r->evGridDone, 0));
            if(timed) cudaEventRecord(ev[9], st);
            launches += enqueueMerge(r, ic, st);
            if(timed) cudaEventRecord(ev[10], st);
        }
    }
    else if(r->useVM) SVCM_CUDA(r, cudaStreamWaitEvent(st, r->evGridDone, 0));
    if(timed && !(r->useVM && !r->lightTraceOnly)) { cudaEventRecord(ev[9], st); cudaEventRecord(ev[10], st); }
    cudaStream_t last = st;
    if(leaveTail)
    {
        last = r->mergeStream;
        // the grid build read the light vertices the next light pass overwrites: it is long done when the
        // camera walk is, but nothing else orders the main stream behind it
        SVCM_CUDA(r, cudaStreamWaitEvent(st, r->evGridDone, 0));
        SVCM_CUDA(r, cudaEventRecord(r->evMainDone, st));
        SVCM_CUDA(r, cudaStreamWaitEvent(last, r->evMainDone, 0));
    }
    if(timed) cudaEventRecord(ev[5], last);

    SVCM_CUDA(r, cudaMemcpyAsync(r->hostCounters + size_t(slot) * CNT_WORDS, r->counters,
                                 sizeof(unsigned long long) * CNT_WORDS, cudaMemcpyDeviceToHost, last));
    if(leaveTail) SVCM_CUDA(r, cudaEventRecord(r->evMergeDone[blk], last));
    r->tailPending = leaveTail;
    SVCM_CUDA(r, cudaGetLastError());
    th[5] = hostNowMs();
    if(traceHost)
        fprintf(stderr, "[svcm host] it %d: setup %.3f light %.3f grid %.3f camera %.3f merge+copy %.3f ms\n", iteration,
                th[1] - th[0], th[2] - th[1], th[3] - th[2], th[4] - th[3], th[5] - th[4]);
    r->ringLaunches[slot] = launches;
    r->ringTimed[slot] = timed;
    r->ringGridSerial[slot] = r->useVM && !r->overlapGrid;
    r->ringHead++;
    if(r->ringHead - r->ringTail > SVCM_PROFILE_RING) r->ringTail = r->ringHead - SVCM_PROFILE_RING;
    r->stats.radius = ic.radius;
    r->statsPending = true;
    r->lastLaunches = launches;
    r->iterationsDone++;
    return SVCM_OK;
}

int finishStats(svcm_renderer *r)
{
    if(!r->statsPending) return SVCM_OK;
    const int slot = int((r->ringHead - 1) % SVCM_PROFILE_RING);
    const unsigned long long *c = r->hostCounters + size_t(slot) * CNT_WORDS;
    SvcmStats &s = r->stats;
    s.paths = r->paths; s.localPaths = r->localPaths;
    s.localLightVertices = (int64_t)c[CNT_LIGHT_VERTS];
    s.lightVertices = s.localLightVertices;
    if(r->params.worldSize > 1 && r->useVM)
    {
        const unsigned long long *per = r->hostCounts64 + size_t(slot) * 8;
        s.lightVertices = 0;
        unsigned long long mx = 0;
        for(int i = 0; i < r->params.worldSize; i++) { s.lightVertices += (int64_t)per[i]; mx = std::max(mx, per[i]); }
        r->lastMaxCount = uint32_t(std::min<unsigned long long>(mx, 0xFFFFFFFFull));   // sizes the next call's blocks
    }
    s.mergeQueries = (int64_t)c[CNT_QUERIES];
    s.mergeCandidates = (int64_t)c[CNT_CANDIDATES];
    s.mergeAccepted = (int64_t)c[CNT_ACCEPTED];
    s.lightRays = (int64_t)c[CNT_LIGHT_RAYS];
    s.cameraRays = (int64_t)c[CNT_CAMERA_RAYS];
    s.shadowRays = (int64_t)c[CNT_SHADOW_RAYS];
    s.kernelLaunches = r->lastLaunches;
    s.paramBytesPerLaunch = int32_t(sizeof(DevScene) + sizeof(IterConst) + sizeof(TraceArgs) + 4);
    if(r->ringTimed[slot])
    {
        cudaEvent_t *ev = r->ev[slot];
        // gather and build run on the grid stream, concurrently with the camera pass: the stage
        // times no longer add up to the total, that is the point
        cudaEventElapsedTime(&s.msLightTrace, ev[0], ev[1]);
        cudaEventElapsedTime(&s.msGather, ev[6], ev[7]);
        cudaEventElapsedTime(&s.msGridBuild, ev[7], ev[8]);
        cudaEventElapsedTime(&s.msCameraTrace, ev[1], ev[4]);
        if(r->ringGridSerial[slot]) { float g = 0.f; cudaEventElapsedTime(&g, ev[6], ev[8]); s.msCameraTrace -= g; }
        cudaEventElapsedTime(&s.msMerge, ev[9], ev[10]);
        cudaEventElapsedTime(&s.msTotal, ev[0], ev[5]);
    }
    r->statsPending = false;
    // any iteration of the ring that dropped records poisons the image: report it
    unsigned long long overflow = 0;
    for(long long it = r->ringTail; it < r->ringHead; it++)
        overflow += r->hostCounters[size_t(it % SVCM_PROFILE_RING) * CNT_WORDS + CNT_OVERFLOW];
    if(overflow)
        return fail(r, SVCM_E_NOMEM, "light-vertex or query store overflowed (%llu records dropped): "
                    "not enough device memory for this resolution / path length", overflow);
    return SVCM_OK;
}

} // namespace

// ------------------------------------------------------------------------------------------
// C-ABI: life cycle
// ------------------------------------------------------------------------------------------
extern "C" {

const char *svcm_version(void) { return "svcm 0.1 (sm_100a; CUDA wavefront VCM; C-ABI 1)"; }

const char *svcm_last_error(const svcm_renderer *r) { return r ? r->error.c_str() : t_createError.c_str(); }

int svcm_create(const SvcmSceneDesc *scene, const SvcmParams *params, svcm_renderer **out)
{
    if(!out) return fail(NULL, SVCM_E_INVALID, "out is NULL");
    *out = NULL;
    if(!params) return fail(NULL, SVCM_E_INVALID, "params is NULL");
    if(params->algorithm < SVCM_ALG_LIGHT_TRACE || params->algorithm > SVCM_ALG_EYE_LIGHT)
        return fail(NULL, SVCM_E_INVALID, "unknown algorithm %d", params->algorithm);
    if(params->worldSize < 1 || params->worldSize > 8 || params->rank < 0 || params->rank >= params->worldSize)
        return fail(NULL, SVCM_E_INVALID, "bad rank/worldSize %d/%d (1..8 ranks)", params->rank, params->worldSize);
    int rc = checkDevice(NULL, params->device);
    if(rc) return rc;

    svcm_renderer *r = new svcm_renderer;
    // zero the POD part by hand (the struct holds std::string members)
    r->device = params->device; r->stream = NULL;
    memset(r->queue, 0, sizeof(r->queue));
    r->queueCountsLight = r->queueCountsCamera = NULL; r->countsAlloc = 0;
    memset(&r->verts, 0, sizeof(r->verts)); r->vertSlots = 0; r->pathVertCount = NULL;
    memset(&r->queries, 0, sizeof(r->queries)); memset(&r->queriesAlt, 0, sizeof(r->queriesAlt));
    memset(&r->qsortAlt, 0, sizeof(r->qsortAlt)); r->queryBuffers = 1; r->scanGenQueryAlt.value = 0u;
    r->mergeStream = NULL; r->evMergeDone[0] = r->evMergeDone[1] = NULL; r->lastMergeBlock = 0;
    memset(&r->shadow, 0, sizeof(r->shadow));
    memset(&r->pairs, 0, sizeof(r->pairs));
    memset(&r->qsort, 0, sizeof(r->qsort)); r->sortQueries = getenv("SVCM_NO_QUERY_SORT") == NULL;
    r->bboxBlocks = NULL;
    r->counters = NULL; r->countersBase = NULL; r->hostCounters = NULL; r->accum = NULL;
    r->tailPending = false; r->evMainDone = NULL;
    memset(&r->grid, 0, sizeof(r->grid)); r->gridCells = 0; r->gridCapacity = 0; r->bboxInit = NULL;
    r->comm = NULL; r->gatherBuf = NULL; r->gatherBytes = 0; r->gatherSlots = 0; r->gatherStride = 0;
    r->strideMaxLen = 0; r->lastMaxCount = 0; r->segCountDev = NULL;
    r->devCounts64 = NULL; r->hostCounts64 = NULL;
    r->accumReduced = NULL; r->reducedValid = false;
    r->gridStream = NULL; r->evLightDone = NULL; r->evGridDone = NULL;
    r->cameraStream = NULL; r->evIterStart = NULL; r->evCameraDone = NULL;
    r->overlapPasses = !(getenv("SVCM_NO_PASS_OVERLAP") && atoi(getenv("SVCM_NO_PASS_OVERLAP")) != 0);
    r->legacyBuild = getenv("SVCM_LEGACY_BUILD") && atoi(getenv("SVCM_LEGACY_BUILD")) != 0;
    r->overlapGrid = !(getenv("SVCM_NO_GRID_OVERLAP") && atoi(getenv("SVCM_NO_GRID_OVERLAP")) != 0);
    r->overlapMerge = !(getenv("SVCM_NO_MERGE_OVERLAP") && atoi(getenv("SVCM_NO_MERGE_OVERLAP")) != 0);
    r->overlapIterations = !(getenv("SVCM_NO_ITERATION_OVERLAP") && atoi(getenv("SVCM_NO_ITERATION_OVERLAP")) != 0);
    r->spatialQueries = !(getenv("SVCM_QUERY_ORDER") && strcmp(getenv("SVCM_QUERY_ORDER"), "bucket") == 0);
    r->iterationsDone = 0; r->profile = false;
    r->mergeGroup = getenv("SVCM_MERGE_KERNEL") ? atoi(getenv("SVCM_MERGE_KERNEL")) : 1;
    r->queriesSizedFor = 0; r->pairsSizedFor = 0;
    {
        // defaults; SVCM_GRID_LAYOUT / SVCM_MERGE_VARIANT override them for A/B measurements
        const char *layout = getenv("SVCM_GRID_LAYOUT");
        r->cellLayout = layout ? (strcmp(layout, "bucket") != 0 && strcmp(layout, "0") != 0) : false;
        const char *variant = getenv("SVCM_MERGE_VARIANT");
        r->mergeVariant = variant ? atoi(variant) : 1;
        r->refCandidates = getenv("SVCM_REF_CANDIDATES") != NULL && atoi(getenv("SVCM_REF_CANDIDATES")) != 0;
        r->scanGenGrid.value = 0; r->scanGenQuery.value = 0;
    }
    r->ev = NULL; r->ringHead = r->ringTail = 0;
    memset(r->ringLaunches, 0, sizeof(r->ringLaunches)); memset(r->ringTimed, 0, sizeof(r->ringTimed));
    memset(r->ringGridSerial, 0, sizeof(r->ringGridSerial));
    memset(r->userEv, 0, sizeof(r->userEv)); memset(&r->stats, 0, sizeof(r->stats));
    r->statsPending = false; r->lastLaunches = 0;
    r->params = *params;

#define SVCM_CREATE_FAIL(code) do { t_createError = r->error; svcm_destroy(r); return (code); } while(0)
    if((rc = buildDevScene(r, scene, r->scene))) SVCM_CREATE_FAIL(rc);

    // algorithm flags, reference src/vertexcm.hxx:222-244
    r->useVC = r->useVM = r->ppm = r->lightTraceOnly = r->pathTrace = r->eyeLight = false;
    switch(params->algorithm)
    {
    case SVCM_ALG_PATH_TRACE: r->pathTrace = true; break;
    case SVCM_ALG_EYE_LIGHT: r->eyeLight = true; break;
    case SVCM_ALG_LIGHT_TRACE: r->lightTraceOnly = true; break;
    case SVCM_ALG_PPM: r->ppm = true; r->useVM = true; break;
    case SVCM_ALG_BPM: r->useVM = true; break;
    case SVCM_ALG_BPT: r->useVC = true; break;
    case SVCM_ALG_VCM: r->useVC = true; r->useVM = true; break;
    }
    if(r->ppm)
    {
        // PPM cannot handle materials mixing specular and non-specular lobes: switch to BPM
        // like the reference does (src/vertexcm.hxx:246-278)
        for(int i = 0; i < r->scene.numMats; i++)
        {
            const SvcmMaterial &m = r->scene.mats[i];
            const bool nonSpec = std::max(std::max(m.diffuse[0], m.diffuse[1]), m.diffuse[2]) > 0 ||
                                 std::max(std::max(m.phong[0], m.phong[1]), m.phong[2]) > 0;
            const bool spec = std::max(std::max(m.mirror[0], m.mirror[1]), m.mirror[2]) > 0 || m.ior > 0;
            if(nonSpec && spec) { r->ppm = false; break; }
        }
    }

    r->resX = int(r->scene.resX); r->resY = int(r->scene.resY);
    if(r->resX <= 0 || r->resY <= 0 || uint64_t(r->resX) * uint64_t(r->resY) > 0x7FFFFFFFull)
    { fail(r, SVCM_E_INVALID, "bad resolution %d x %d", r->resX, r->resY); SVCM_CREATE_FAIL(SVCM_E_INVALID); }
    r->paths = uint32_t(r->resX) * uint32_t(r->resY);
    const uint32_t ws = uint32_t(params->worldSize), rk = uint32_t(params->rank);
    r->localRows = (uint32_t(r->resY) + ws - 1u - rk) / ws;     // rows y with y % ws == rank
    r->localPaths = r->localRows * uint32_t(r->resX);
    r->baseRadius = params->radiusFactor * r->scene.sphereRadius;  // src/vertexcm.hxx:280

#define SVCM_TRY(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { \
        fail(r, e_ == cudaErrorMemoryAllocation ? SVCM_E_NOMEM : SVCM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
        SVCM_CREATE_FAIL(e_ == cudaErrorMemoryAllocation ? SVCM_E_NOMEM : SVCM_E_CUDA); } } while(0)
    SVCM_TRY(cudaSetDevice(r->device));
    if((rc = uploadGeometry(r, scene, r->scene, r->geometry))) SVCM_CREATE_FAIL(rc);
    SVCM_TRY(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    for(int i = 0; i < SVCM_NUM_USER_EVENTS; i++) SVCM_TRY(cudaEventCreate(&r->userEv[i]));
    const bool bidirectional = params->algorithm != SVCM_ALG_PATH_TRACE && params->algorithm != SVCM_ALG_EYE_LIGHT &&
                               params->algorithm != SVCM_ALG_LIGHT_TRACE;
    for(int i = 0; i < (bidirectional ? 4 : 2); i++) if((rc = allocQueue(r, r->queue[i], r->localPaths))) SVCM_CREATE_FAIL(rc);
    if((rc = devAlloc(r, &r->pathVertCount, r->localPaths))) SVCM_CREATE_FAIL(rc);
    if((rc = devAlloc(r, &r->countersBase, 2 * CNT_WORDS))) SVCM_CREATE_FAIL(rc);
    r->counters = r->countersBase;
    if((rc = devAlloc(r, &r->accum, size_t(r->paths) * 3))) SVCM_CREATE_FAIL(rc);
    if((rc = devAlloc(r, &r->bboxBlocks, 16))) SVCM_CREATE_FAIL(rc);
    r->grid.bboxBits = r->bboxBlocks;
    if((rc = devAlloc(r, &r->bboxInit, 6))) SVCM_CREATE_FAIL(rc);
    if((rc = devAlloc(r, &r->devCounts64, 8))) SVCM_CREATE_FAIL(rc);
    if((rc = devAlloc(r, &r->segCountDev, 8))) SVCM_CREATE_FAIL(rc);
    SVCM_TRY(cudaStreamCreateWithFlags(&r->gridStream, cudaStreamNonBlocking));
    SVCM_TRY(cudaEventCreateWithFlags(&r->evLightDone, cudaEventDisableTiming));
    SVCM_TRY(cudaEventCreateWithFlags(&r->evGridDone, cudaEventDisableTiming));
    SVCM_TRY(cudaStreamCreateWithFlags(&r->cameraStream, cudaStreamNonBlocking));
    SVCM_TRY(cudaEventCreateWithFlags(&r->evIterStart, cudaEventDisableTiming));
    SVCM_TRY(cudaEventCreateWithFlags(&r->evCameraDone, cudaEventDisableTiming));
    SVCM_TRY(cudaStreamCreateWithFlags(&r->mergeStream, cudaStreamNonBlocking));
    SVCM_TRY(cudaEventCreateWithFlags(&r->evMergeDone[0], cudaEventDisableTiming));
    SVCM_TRY(cudaEventCreateWithFlags(&r->evMergeDone[1], cudaEventDisableTiming));
    SVCM_TRY(cudaEventCreateWithFlags(&r->evMainDone, cudaEventDisableTiming));
    SVCM_TRY(cudaMallocHost((void**)&r->hostCounters, sizeof(unsigned long long) * CNT_WORDS * SVCM_PROFILE_RING));
    SVCM_TRY(cudaMallocHost((void**)&r->hostCounts64, sizeof(unsigned long long) * 8 * SVCM_PROFILE_RING));
    memset(r->hostCounters, 0, sizeof(unsigned long long) * CNT_WORDS * SVCM_PROFILE_RING);
    memset(r->hostCounts64, 0, sizeof(unsigned long long) * 8 * SVCM_PROFILE_RING);
    const int32_t init[6] = { floatToOrdered(1e36f), floatToOrdered(1e36f), floatToOrdered(1e36f),
                              floatToOrdered(-1e36f), floatToOrdered(-1e36f), floatToOrdered(-1e36f) };
    SVCM_TRY(cudaMemcpy(r->bboxInit, init, sizeof(init), cudaMemcpyHostToDevice));
    SVCM_TRY(cudaMemcpy(r->bboxBlocks, init, sizeof(init), cudaMemcpyHostToDevice));
    SVCM_TRY(cudaMemcpy(r->bboxBlocks + 8, init, sizeof(init), cudaMemcpyHostToDevice));
    r->grid.bboxInit = r->bboxInit;
    SVCM_TRY(cudaMemset(r->accum, 0, sizeof(float) * 3 * size_t(r->paths)));
    SVCM_TRY(cudaMemset(r->devCounts64, 0, sizeof(unsigned long long) * 8));
    SVCM_TRY(cudaMemset(r->segCountDev, 0, sizeof(uint32_t) * 8));
#undef SVCM_TRY
#undef SVCM_CREATE_FAIL
    *out = r;
    return SVCM_OK;
}

void svcm_destroy(svcm_renderer *r)
{
    if(!r) return;
    cudaSetDevice(r->device);
    if(r->stream) { joinTail(r); cudaStreamSynchronize(r->stream); }
    if(r->comm && ncclApi()->lib) ncclApi()->CommDestroy(r->comm);
    for(int i = 0; i < 4; i++) freeQueue(r->queue[i]);
    if(r->cameraStream) { cudaStreamSynchronize(r->cameraStream); cudaStreamDestroy(r->cameraStream); }
    if(r->evIterStart) cudaEventDestroy(r->evIterStart);
    if(r->evCameraDone) cudaEventDestroy(r->evCameraDone);
    if(r->mergeStream) { cudaStreamSynchronize(r->mergeStream); cudaStreamDestroy(r->mergeStream); }
    for(int b = 0; b < 2; b++) if(r->evMergeDone[b]) cudaEventDestroy(r->evMergeDone[b]);
    if(r->evMainDone) cudaEventDestroy(r->evMainDone);
    devFree(r->queueCountsLight); devFree(r->queueCountsCamera);
    devFree(r->shadow.r0); devFree(r->shadow.r1); devFree(r->shadow.r2); devFree(r->shadow.counts);
    freeVerts(r); freeQueries(r); freeGridSorted(r);
    devFree(r->pathVertCount); devFree(r->countersBase); r->counters = NULL; devFree(r->accum); devFree(r->image8);
    r->geometry.release();
    devFree(r->grid.cellCount); devFree(r->grid.blockSums); devFree(r->bboxBlocks); r->grid.bboxBits = NULL;
    devFree(r->grid.cellEnd); devFree(r->grid.dims); devFree(r->grid.refCount);
    devFree(r->grid.scanState); devFree(r->grid.scanCounter);
    devFree(r->qsort.cellCount); devFree(r->qsort.blockSums);
    devFree(r->qsort.cellEnd); devFree(r->qsort.scanState); devFree(r->qsort.scanCounter);
    devFree(r->qsortAlt.cellCount); devFree(r->qsortAlt.blockSums);
    devFree(r->qsortAlt.cellEnd); devFree(r->qsortAlt.scanState); devFree(r->qsortAlt.scanCounter);
    devFree(r->bboxInit); devFree(r->devCounts64); devFree(r->segCountDev); devFree(r->accumReduced);
    if(r->gridStream) { cudaStreamSynchronize(r->gridStream); cudaStreamDestroy(r->gridStream); }
    if(r->evLightDone) cudaEventDestroy(r->evLightDone);
    if(r->evGridDone) cudaEventDestroy(r->evGridDone);
    if(r->hostCounters) cudaFreeHost(r->hostCounters);
    if(r->hostCounts64) cudaFreeHost(r->hostCounts64);
    if(r->ev)
    {
        for(int k = 0; k < SVCM_PROFILE_RING; k++) for(int i = 0; i < 11; i++) if(r->ev[k][i]) cudaEventDestroy(r->ev[k][i]);
        delete[] r->ev;
    }
    for(int i = 0; i < SVCM_NUM_USER_EVENTS; i++) if(r->userEv[i]) cudaEventDestroy(r->userEv[i]);
    if(r->stream) cudaStreamDestroy(r->stream);
    delete r;
}

int svcm_set_option(svcm_renderer *r, int32_t option, int64_t value)
{
    if(!r) return SVCM_E_INVALID;
    if(r->tailPending) { cudaSetDevice(r->device); const int rc = joinTail(r); if(rc) return rc; }
    switch(option)
    {
    case SVCM_OPT_PROFILE:
        if(value != 0 && !r->ev)
        {
            SVCM_CUDA(r, cudaSetDevice(r->device));
            r->ev = new cudaEvent_t[SVCM_PROFILE_RING][11];
            memset(r->ev, 0, sizeof(cudaEvent_t) * 11 * SVCM_PROFILE_RING);
            for(int k = 0; k < SVCM_PROFILE_RING; k++) for(int i = 0; i < 11; i++) SVCM_CUDA(r, cudaEventCreate(&r->ev[k][i]));
        }
        r->profile = value != 0;
        return SVCM_OK;
    case SVCM_OPT_MERGE_GROUP:
        if(!((value >= 1 && value <= 5) || value == 8 || value == 16 || value == 32))
            return fail(r, SVCM_E_INVALID, "merge kernel must be 1 (lane per query), 2..5 (staged buckets), 8, 16 or 32");
        r->mergeGroup = int(value); return SVCM_OK;
    case SVCM_OPT_COUNTERS: return SVCM_OK;   // counters are always maintained
    case SVCM_OPT_GRID_LAYOUT:
        if(value != 0 && value != 1) return fail(r, SVCM_E_INVALID, "grid layout must be 0 (hash buckets) or 1 (cell order)");
        r->cellLayout = value != 0; return SVCM_OK;
    case SVCM_OPT_MERGE_VARIANT:
        if(value != 0 && value != 1 && value != 3) return fail(r, SVCM_E_INVALID, "merge variant must be 0, 1 or 3");
        r->mergeVariant = int(value); return SVCM_OK;
    case SVCM_OPT_REF_CANDIDATES:
        r->refCandidates = value != 0; return SVCM_OK;
    case SVCM_OPT_OVERLAP:
        if(value < 0 || value > 15) return fail(r, SVCM_E_INVALID, "overlap mask must be 0..15");
        r->overlapPasses = (value & 1) != 0; r->overlapGrid = (value & 2) != 0; r->overlapMerge = (value & 4) != 0;
        r->overlapIterations = (value & 8) != 0;
        return SVCM_OK;
    default: return fail(r, SVCM_E_INVALID, "unknown option %d", option);
    }
}

// ------------------------------------------------------------------------------------------
// C-ABI: the hot path
// ------------------------------------------------------------------------------------------
int svcm_run_iterations(svcm_renderer *r, int32_t firstIteration, int32_t count, uint32_t minPathLength,
                        uint32_t maxPathLength)
{
    if(!r) return SVCM_E_INVALID;
    if(count < 0 || firstIteration < 0) return fail(r, SVCM_E_INVALID, "bad iteration range %d+%d", firstIteration, count);
    SVCM_CUDA(r, cudaSetDevice(r->device));
    int rc = ensureCapacity(r, maxPathLength);
    if(rc) return rc;
    if(count > 0) r->reducedValid = false;
    for(int i = 0; i < count; i++)
    {
        // per-iteration counters of the previous iteration are overwritten: check overflow first
        // the per-iteration counters live in a ring of SVCM_PROFILE_RING slots: before it wraps,
        // wait and look at them, so that no iteration's overflow report is ever overwritten unseen
        if(r->ringHead - r->ringTail >= SVCM_PROFILE_RING - 1)
        {
            if((rc = joinTail(r))) return rc;
            SVCM_CUDA(r, cudaStreamSynchronize(r->stream));
            if((rc = finishStats(r))) return rc;
            r->ringTail = r->ringHead - 1;       // keep the last iteration's record for svcm_get_profile
        }
        if((rc = runOneIteration(r, firstIteration + i, minPathLength, maxPathLength))) return rc;
    }
    return SVCM_OK;
}

int svcm_synchronize(svcm_renderer *r)
{
    if(!r) return SVCM_E_INVALID;
    SVCM_CUDA(r, cudaSetDevice(r->device));
    { const int rc = joinTail(r); if(rc) return rc; }
    SVCM_CUDA(r, cudaStreamSynchronize(r->stream));
    return finishStats(r);
}

int svcm_download_accum(svcm_renderer *r, float *rgb, int32_t *iterationsDone)
{
    if(!r || !rgb) return SVCM_E_INVALID;
    SVCM_CUDA(r, cudaSetDevice(r->device));
    // after svcm_reduce_accum the root serves the sum over ranks (until more iterations are rendered)
    const float *src = (r->reducedValid && r->accumReduced) ? r->accumReduced : r->accum;
    { const int rc = joinTail(r); if(rc) return rc; }
    SVCM_CUDA(r, cudaMemcpyAsync(rgb, src, sizeof(float) * 3 * size_t(r->paths), cudaMemcpyDeviceToHost, r->stream));
    SVCM_CUDA(r, cudaStreamSynchronize(r->stream));
    if(iterationsDone) *iterationsDone = r->iterationsDone;
    return finishStats(r);
}

int64_t svcm_image_bytes(int32_t format, int32_t width, int32_t height)
{
    if(width < 0 || height < 0) return -1;
    if(format == SVCM_IMAGE_BMP24) return int64_t(width) * height * 3;
    if(format == SVCM_IMAGE_RGBE)  return int64_t(width) * height * 4;
    return -1;
}

int32_t svcm_image_header(int32_t format, int32_t width, int32_t height, uint8_t *out, int32_t capacity)
{
    if(width < 0 || height < 0 || (!out && capacity > 0)) return -1;
    if(format == SVCM_IMAGE_BMP24)
    {
        // "BM" + the 52-byte BmpHeader of src/framebuffer.hxx:149-193 (little endian)
        const uint32_t payload = uint32_t(width) * uint32_t(height) * 3u;
        const uint32_t words[13] = { 54u + payload, 0u, 54u, 40u, uint32_t(width), uint32_t(height),
                                     1u | (24u << 16), 0u, payload, 2953u, 2953u, 0u, 0u };
        if(capacity >= 54)
        {
            out[0] = 'B'; out[1] = 'M';
            for(int w = 0; w < 13; w++)
                for(int b = 0; b < 4; b++) out[2 + 4 * w + b] = uint8_t(words[w] >> (8 * b));
        }
        return 54;
    }
    if(format == SVCM_IMAGE_RGBE)
    {
        // src/framebuffer.hxx:224-227
        char text[128];
        const int n = snprintf(text, sizeof(text), "#?RADIANCE\n# SmallVCM\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", height, width);
        if(capacity >= n) memcpy(out, text, size_t(n));
        return n;
    }
    return -1;
}

int svcm_download_image(svcm_renderer *r, int32_t format, float gamma, uint8_t *out, int32_t *iterationsDone)
{
    if(!r || !out) return SVCM_E_INVALID;
    if(format != SVCM_IMAGE_BMP24 && format != SVCM_IMAGE_RGBE) return fail(r, SVCM_E_INVALID, "unknown image format %d", format);
    if(format == SVCM_IMAGE_BMP24 && !(gamma > 0.f)) return fail(r, SVCM_E_INVALID, "gamma must be positive");
    SVCM_CUDA(r, cudaSetDevice(r->device));
    if(!r->image8) { const int rc = devAlloc(r, &r->image8, size_t(r->paths) * 4u); if(rc) return rc; }
    { const int rc = joinTail(r); if(rc) return rc; }
    launchImageConvert(r->accum, r->resX, r->resY, r->iterationsDone, format, gamma, r->image8, r->stream);
    SVCM_CUDA(r, cudaGetLastError());
    SVCM_CUDA(r, cudaMemcpyAsync(out, r->image8, size_t(svcm_image_bytes(format, r->resX, r->resY)), cudaMemcpyDeviceToHost, r->stream));
    SVCM_CUDA(r, cudaStreamSynchronize(r->stream));
    if(iterationsDone) *iterationsDone = r->iterationsDone;
    return finishStats(r);
}

int svcm_save_image(svcm_renderer *r, const char *path, float gamma)
{
    if(!r || !path) return SVCM_E_INVALID;
    // main() decides by the last three characters of the name (src/smallvcm.cxx:311-320)
    const size_t len = strlen(path);
    const char *ext = len >= 3 ? path + len - 3 : "";
    int format;
    if(!strcmp(ext, "bmp")) format = SVCM_IMAGE_BMP24;
    else if(!strcmp(ext, "hdr")) format = SVCM_IMAGE_RGBE;
    else return fail(r, SVCM_E_INVALID, "Used unknown extension %s", ext);
    uint8_t header[128];
    const int32_t headerBytes = svcm_image_header(format, r->resX, r->resY, header, int32_t(sizeof(header)));
    std::vector<uint8_t> payload(size_t(svcm_image_bytes(format, r->resX, r->resY)));
    const int rc = svcm_download_image(r, format, gamma, payload.data(), NULL);
    if(rc) return rc;
    FILE *f = fopen(path, "wb");
    if(!f) return fail(r, SVCM_E_INVALID, "cannot open %s for writing", path);
    const bool ok = fwrite(header, 1, size_t(headerBytes), f) == size_t(headerBytes) &&
                    fwrite(payload.data(), 1, payload.size(), f) == payload.size();
    if(fclose(f) != 0 || !ok) return fail(r, SVCM_E_INVALID, "short write to %s", path);
    return SVCM_OK;
}

int svcm_clear_accum(svcm_renderer *r)
{
    if(!r) return SVCM_E_INVALID;
    SVCM_CUDA(r, cudaSetDevice(r->device));
    { const int rc = joinTail(r); if(rc) return rc; }
    SVCM_CUDA(r, cudaMemsetAsync(r->accum, 0, sizeof(float) * 3 * size_t(r->paths), r->stream));
    r->iterationsDone = 0; r->reducedValid = false;
    return SVCM_OK;
}

int svcm_get_stats(svcm_renderer *r, SvcmStats *out)
{
    if(!r || !out) return SVCM_E_INVALID;
    const int rc = svcm_synchronize(r);
    *out = r->stats;
    return rc;
}

int svcm_get_profile(svcm_renderer *r, SvcmProfile *out)
{
    if(!r || !out) return SVCM_E_INVALID;
    const int rc = svcm_synchronize(r);
    memset(out, 0, sizeof(*out));
    for(long long it = r->ringTail; it < r->ringHead; it++)
    {
        const int slot = int(it % SVCM_PROFILE_RING);
        const unsigned long long *c = r->hostCounters + size_t(slot) * CNT_WORDS;
        out->iterations++;
        out->lightVertices += (int64_t)c[CNT_LIGHT_VERTS];
        out->mergeQueries += (int64_t)c[CNT_QUERIES];
        out->mergeCandidates += (int64_t)c[CNT_CANDIDATES];
        out->mergeAccepted += (int64_t)c[CNT_ACCEPTED];
        out->lightRays += (int64_t)c[CNT_LIGHT_RAYS];
        out->cameraRays += (int64_t)c[CNT_CAMERA_RAYS];
        out->shadowRays += (int64_t)c[CNT_SHADOW_RAYS];
        out->kernelLaunches += r->ringLaunches[slot];
        if(r->ringTimed[slot])
        {
            cudaEvent_t *ev = r->ev[slot];
            float ms[6] = { 0, 0, 0, 0, 0, 0 };
            cudaEventElapsedTime(&ms[0], ev[0], ev[1]); cudaEventElapsedTime(&ms[1], ev[6], ev[7]);
            cudaEventElapsedTime(&ms[2], ev[7], ev[8]); cudaEventElapsedTime(&ms[3], ev[1], ev[4]);
            if(r->ringGridSerial[slot]) { float g = 0.f; cudaEventElapsedTime(&g, ev[6], ev[8]); ms[3] -= g; }
            cudaEventElapsedTime(&ms[4], ev[9], ev[10]); cudaEventElapsedTime(&ms[5], ev[0], ev[5]);
            out->msLightTrace += ms[0]; out->msGather += ms[1]; out->msGridBuild += ms[2];
            out->msCameraTrace += ms[3]; out->msMerge += ms[4]; out->msTotal += ms[5];
            out->timedIterations++;
        }
    }
    r->ringTail = r->ringHead;
    return rc;
}

int svcm_host_register(void *ptr, size_t bytes)
{
    if(!ptr || !bytes) return SVCM_E_INVALID;
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if(e != cudaSuccess) { cudaGetLastError(); return fail(NULL, SVCM_E_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e)); }
    return SVCM_OK;
}

int svcm_host_unregister(void *ptr)
{
    if(!ptr) return SVCM_E_INVALID;
    const cudaError_t e = cudaHostUnregister(ptr);
    if(e != cudaSuccess) { cudaGetLastError(); return fail(NULL, SVCM_E_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e)); }
    return SVCM_OK;
}

int svcm_device_count(void)
{
    int n = 0;
    if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int svcm_event_record(svcm_renderer *r, int32_t slot)
{
    if(!r || slot < 0 || slot >= SVCM_NUM_USER_EVENTS) return r ? fail(r, SVCM_E_INVALID, "bad event slot") : SVCM_E_INVALID;
    SVCM_CUDA(r, cudaSetDevice(r->device));
    { const int rc = joinTail(r); if(rc) return rc; }     // an event marks the end of everything enqueued so far
    SVCM_CUDA(r, cudaEventRecord(r->userEv[slot], r->stream));
    return SVCM_OK;
}

int svcm_event_elapsed_ms(svcm_renderer *r, int32_t startSlot, int32_t stopSlot, float *ms)
{
    if(!r || !ms || startSlot < 0 || startSlot >= SVCM_NUM_USER_EVENTS || stopSlot < 0 || stopSlot >= SVCM_NUM_USER_EVENTS)
        return r ? fail(r, SVCM_E_INVALID, "bad event arguments") : SVCM_E_INVALID;
    SVCM_CUDA(r, cudaSetDevice(r->device));
    SVCM_CUDA(r, cudaEventSynchronize(r->userEv[stopSlot]));
    SVCM_CUDA(r, cudaEventElapsedTime(ms, r->userEv[startSlot], r->userEv[stopSlot]));
    return SVCM_OK;
}

// ------------------------------------------------------------------------------------------
// C-ABI: multi-GPU
// ------------------------------------------------------------------------------------------
int64_t svcm_shard_local_paths(int32_t width, int32_t height, int32_t rank, int32_t worldSize)
{
    if(width <= 0 || height <= 0 || worldSize <= 0 || rank < 0 || rank >= worldSize) return -1;
    const int64_t rows = (int64_t(height) + worldSize - 1 - rank) / worldSize;
    return rows * width;
}
int32_t svcm_shard_owner(int32_t width, int32_t worldSize, int64_t path)
{
    if(width <= 0 || worldSize <= 0 || path < 0) return -1;
    return int32_t((path / width) % worldSize);
}
int64_t svcm_shard_local_index(int32_t width, int32_t worldSize, int64_t path)
{
    if(width <= 0 || worldSize <= 0 || path < 0) return -1;
    const int64_t y = path / width, x = path - y * width;
    return (y / worldSize) * width + x;
}
int64_t svcm_shard_global_index(int32_t width, int32_t rank, int32_t worldSize, int64_t local)
{
    if(width <= 0 || worldSize <= 0 || rank < 0 || rank >= worldSize || local < 0) return -1;
    const int64_t ly = local / width, x = local - ly * width;
    return (ly * worldSize + rank) * width + x;
}

int svcm_comm_unique_id(uint8_t id[SVCM_COMM_ID_BYTES])
{
    NcclApi *n = ncclApi();
    if(!n->lib) return fail(NULL, SVCM_E_NCCL, "%s", n->error.c_str());
    ncclUniqueId uid;
    const int rc = n->GetUniqueId(&uid);
    if(rc != NCCL_SUCCESS) return fail(NULL, SVCM_E_NCCL, "ncclGetUniqueId failed: %s", n->GetErrorString(rc));
    memcpy(id, uid.internal, SVCM_COMM_ID_BYTES);
    return SVCM_OK;
}

int svcm_comm_init(svcm_renderer *r, const uint8_t id[SVCM_COMM_ID_BYTES])
{
    if(!r || !id) return SVCM_E_INVALID;
    NcclApi *n = ncclApi();
    if(!n->lib) return fail(r, SVCM_E_NCCL, "%s", n->error.c_str());
    if(r->comm) return fail(r, SVCM_E_STATE, "communicator already initialised");
    SVCM_CUDA(r, cudaSetDevice(r->device));
    ncclUniqueId uid; memcpy(uid.internal, id, SVCM_COMM_ID_BYTES);
    return ncclCheck(r, n->CommInitRank(&r->comm, r->params.worldSize, uid, r->params.rank), "ncclCommInitRank");
}

int svcm_reduce_accum(svcm_renderer *r, int32_t root)
{
    if(!r) return SVCM_E_INVALID;
    if(r->params.worldSize == 1) return SVCM_OK;
    if(!r->comm) return fail(r, SVCM_E_STATE, "svcm_comm_init was not called");
    if(root < 0 || root >= r->params.worldSize) return fail(r, SVCM_E_INVALID, "root %d out of range", root);
    SVCM_CUDA(r, cudaSetDevice(r->device));
    // The sum goes to a buffer of its own on the root: every rank's accum stays what that rank
    // rendered, so rendering more iterations and reducing again counts nothing twice.
    if(r->params.rank == root && !r->accumReduced)
    {
        const int rc = devAlloc(r, &r->accumReduced, size_t(r->paths) * 3);
        if(rc) return rc;
    }
    { const int rcj = joinTail(r); if(rcj) return rcj; }
    const int rc = ncclCheck(r, ncclApi()->Reduce(r->accum, r->params.rank == root ? r->accumReduced : r->accum,
                                                  size_t(r->paths) * 3, NCCL_FLOAT32, NCCL_SUM, root, r->comm, r->stream),
                             "ncclReduce(accum)");
    if(rc == SVCM_OK && r->params.rank == root) r->reducedValid = true;
    return rc;
}

// ------------------------------------------------------------------------------------------
// C-ABI: kernel-level entry points
// ------------------------------------------------------------------------------------------
} // extern "C"

struct svcm_hashgrid
{
    int device;
    cudaStream_t stream;
    int64_t count;
    uint32_t numCells;
    float radius;
    float4 *posLen, *thrVcm, *dirVm; float *contProb;
    unsigned long long *countDev;
    GridBuffers grid;
    int32_t *bboxInit;
    bool cellLayout;                // SVCM_GRID_LAYOUT at build time
    ScanGen scanGen;
    std::string error;
};

namespace {

// RAII device scratch for the batch entry points
struct Scratch
{
    std::vector<void*> ptrs;
    ~Scratch() { for(size_t i = 0; i < ptrs.size(); i++) cudaFree(ptrs[i]); }
    template<typename T> T *alloc(size_t n, cudaError_t &err)
    {
        void *p = NULL;
        if(err == cudaSuccess) { err = cudaMalloc(&p, std::max<size_t>(1, n) * sizeof(T)); if(err == cudaSuccess) ptrs.push_back(p); }
        return (T*)p;
    }
};

// `geo` owns the device copy of a large scene's geometry for the duration of the batch call;
// NULL for stages that never trace a ray (BSDF, merge)
int batchPrologue(const SvcmSceneDesc *scene, int device, DevScene &sc, SceneGeometry *geo)
{
    int rc = checkDevice(NULL, device);
    if(rc) return rc;
    if((rc = buildDevScene(NULL, scene, sc))) return rc;
    if(cudaSetDevice(device) != cudaSuccess) return fail(NULL, SVCM_E_CUDA, "cudaSetDevice(%d) failed", device);
    if(geo) return uploadGeometry(NULL, scene, sc, *geo);
    return SVCM_OK;
}

int cudaDone(cudaError_t e, const char *what)
{
    if(e == cudaSuccess) e = cudaDeviceSynchronize();
    if(e == cudaSuccess) e = cudaGetLastError();
    if(e != cudaSuccess) return fail(NULL, e == cudaErrorMemoryAllocation ? SVCM_E_NOMEM : SVCM_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return SVCM_OK;
}

} // namespace

extern "C" {

int svcm_hashgrid_build(const SvcmMergeVertex *vertices, int64_t count, float radius, int32_t numCells,
                        int32_t device, svcm_hashgrid **out)
{
    if(!out) return fail(NULL, SVCM_E_INVALID, "out is NULL");
    *out = NULL;
    if(count < 0 || count > 0x7FFFFF00ll || numCells <= 0 || !(radius > 0.f) || (count && !vertices))
        return fail(NULL, SVCM_E_INVALID, "bad hash grid arguments");
    int rc = checkDevice(NULL, device);
    if(rc) return rc;
    cudaSetDevice(device);

    svcm_hashgrid *g = new svcm_hashgrid;
    g->device = device; g->stream = NULL; g->count = count; g->numCells = uint32_t(numCells); g->radius = radius;
    g->posLen = g->thrVcm = g->dirVm = NULL; g->contProb = NULL; g->countDev = NULL; g->bboxInit = NULL;
    memset(&g->grid, 0, sizeof(g->grid));
    {
        const char *layout = getenv("SVCM_GRID_LAYOUT");
        g->cellLayout = layout ? (strcmp(layout, "bucket") != 0 && strcmp(layout, "0") != 0) : false;
        g->scanGen.value = 0;
    }

    cudaError_t e = cudaSuccess;
    SvcmMergeVertex *staging = NULL;
    const size_t n = size_t(std::max<int64_t>(count, 1));
#define G_ALLOC(p, cnt) if(e == cudaSuccess) e = cudaMalloc((void**)&(p), sizeof(*(p)) * (cnt))
    G_ALLOC(staging, n);
    G_ALLOC(g->posLen, n); G_ALLOC(g->thrVcm, n); G_ALLOC(g->dirVm, n); G_ALLOC(g->contProb, n);
    G_ALLOC(g->countDev, 1); G_ALLOC(g->bboxInit, 6);
    G_ALLOC(g->grid.cellKey, n); G_ALLOC(g->grid.cellCount, size_t(numCells) + 1);
    G_ALLOC(g->grid.blockSums, (size_t(numCells) + 1) / 2048 + 2); G_ALLOC(g->grid.bboxBits, 6);
    G_ALLOC(g->grid.sPosLen, n); G_ALLOC(g->grid.sThrVcm, n); G_ALLOC(g->grid.sDirVm, n);
    G_ALLOC(g->grid.sContProb, n); G_ALLOC(g->grid.sIndex, n);
#undef G_ALLOC
    if(e == cudaSuccess)
    {
        const int32_t init[6] = { floatToOrdered(1e36f), floatToOrdered(1e36f), floatToOrdered(1e36f),
                                  floatToOrdered(-1e36f), floatToOrdered(-1e36f), floatToOrdered(-1e36f) };
        e = cudaMemcpy(g->bboxInit, init, sizeof(init), cudaMemcpyHostToDevice);
        g->grid.bboxInit = g->bboxInit;
    }
    const unsigned long long c64 = (unsigned long long)count;
    if(e == cudaSuccess) e = cudaMemcpy(g->countDev, &c64, 8, cudaMemcpyHostToDevice);
    if(e == cudaSuccess && count) e = cudaMemcpy(staging, vertices, sizeof(SvcmMergeVertex) * size_t(count), cudaMemcpyHostToDevice);
    if(e == cudaSuccess)
    {
        launchPackMergeVertices(staging, count, g->posLen, g->thrVcm, g->dirVm, g->contProb, 0);
        IterConst ic; memset(&ic, 0, sizeof(ic));
        ic.radius = radius; ic.radiusSqr = radius * radius; ic.invCellSize = 1.f / (radius * 2.f);
        ic.numCells = uint32_t(numCells); ic.hashMagic = hashMagicFor(ic.numCells);
        MergeSet set; memset(&set, 0, sizeof(set));
        set.posLen = g->posLen; set.thrVcm = g->thrVcm; set.dirVm = g->dirVm; set.contProb = g->contProb;
        set.singleCount = g->countDev; set.capacity = uint32_t(count); set.segStride = uint32_t(count); set.numSegs = 1;
        if(g->cellLayout)
        {
            // the reference-bucket histogram rides along: svcm_hashgrid_download and the per-query
            // candidate counts are stated in the reference's convention
            if(allocCellTables(NULL, g->grid, NULL, uint32_t(numCells)) != SVCM_OK) e = cudaErrorMemoryAllocation;
            else launchCellGridBuild(ic, set, g->grid, g->scanGen, true, 0);
        }
        else
        {
            // scan state for the single-pass scan (the product's build); SVCM_LEGACY_BUILD=1: round 1's
            const bool legacy = getenv("SVCM_LEGACY_BUILD") && atoi(getenv("SVCM_LEGACY_BUILD")) != 0;
            if(!legacy)
            {
                if(e == cudaSuccess) e = cudaMalloc((void**)&g->grid.scanState, sizeof(unsigned long long) * kScanStateWords);
                if(e == cudaSuccess) e = cudaMalloc((void**)&g->grid.scanCounter, sizeof(uint32_t));
                if(e == cudaSuccess) e = cudaMemset(g->grid.scanState, 0, sizeof(unsigned long long) * kScanStateWords);
                if(e == cudaSuccess) e = cudaMemset(g->grid.scanCounter, 0, sizeof(uint32_t));
            }
            if(e == cudaSuccess) launchGridBuild(ic, set, g->grid, legacy ? NULL : &g->scanGen, 0);
        }
    }
    rc = cudaDone(e, "svcm_hashgrid_build");
    if(staging) cudaFree(staging);
    if(rc) { svcm_hashgrid_destroy(g); return rc; }
    *out = g;
    return SVCM_OK;
}

void svcm_hashgrid_destroy(svcm_hashgrid *g)
{
    if(!g) return;
    cudaSetDevice(g->device);
    cudaFree(g->posLen); cudaFree(g->thrVcm); cudaFree(g->dirVm); cudaFree(g->contProb); cudaFree(g->countDev);
    cudaFree(g->bboxInit); cudaFree(g->grid.cellKey); cudaFree(g->grid.cellCount); cudaFree(g->grid.blockSums);
    cudaFree(g->grid.bboxBits); cudaFree(g->grid.sPosLen); cudaFree(g->grid.sThrVcm); cudaFree(g->grid.sDirVm);
    cudaFree(g->grid.sContProb); cudaFree(g->grid.sIndex);
    cudaFree(g->grid.cellEnd); cudaFree(g->grid.dims); cudaFree(g->grid.refCount);
    cudaFree(g->grid.scanState); cudaFree(g->grid.scanCounter);
    delete g;
}

int svcm_hashgrid_download(svcm_hashgrid *g, float bbox[6], int32_t *cellEnds, int32_t *indices)
{
    if(!g) return SVCM_E_INVALID;
    cudaSetDevice(g->device);
    cudaError_t e = cudaSuccess;
    if(bbox)
    {
        int32_t bits[6];
        e = cudaMemcpy(bits, g->grid.bboxBits, sizeof(bits), cudaMemcpyDeviceToHost);
        for(int i = 0; i < 6; i++) bbox[i] = orderedToFloat(bits[i]);
    }
    if(g->cellLayout)
    {
        // The particles lie in cell order; mCellEnds / mIndices are restated from it: the device gives
        // the reference bucket of the particle at every sorted slot and the per-bucket histogram, the
        // grouping by bucket happens here (a stable counting sort of the slots).
        if(e == cudaSuccess && (cellEnds || indices))
        {
            const size_t n = size_t(g->count);
            std::vector<uint32_t> refKey(n), counts(g->numCells);
            std::vector<int32_t> perm(n);
            uint32_t *dKeys = NULL;
            if(n) e = cudaMalloc((void**)&dKeys, sizeof(uint32_t) * n);
            if(e == cudaSuccess && n)
            {
                IterConst ic; memset(&ic, 0, sizeof(ic));
                ic.invCellSize = 1.f / (g->radius * 2.f); ic.numCells = g->numCells; ic.hashMagic = hashMagicFor(g->numCells);
                launchCellRefKeys(ic, g->grid, uint32_t(n), dKeys, 0);
                e = cudaMemcpy(refKey.data(), dKeys, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost);
                if(e == cudaSuccess) e = cudaMemcpy(perm.data(), g->grid.sIndex, sizeof(int32_t) * n, cudaMemcpyDeviceToHost);
            }
            if(dKeys) cudaFree(dKeys);
            if(e == cudaSuccess) e = cudaMemcpy(counts.data(), g->grid.refCount, sizeof(uint32_t) * g->numCells, cudaMemcpyDeviceToHost);
            if(e == cudaSuccess)
            {
                std::vector<uint32_t> cursor(g->numCells);
                uint32_t run = 0;
                for(uint32_t c = 0; c < g->numCells; c++) { cursor[c] = run; run += counts[c]; if(cellEnds) cellEnds[c] = int32_t(run); }
                if(run != n) return fail(NULL, SVCM_E_STATE, "bucket histogram (%u) does not add up to the particle count (%zu)", run, n);
                if(indices) for(size_t slot = 0; slot < n; slot++)
                {
                    if(refKey[slot] >= g->numCells) return fail(NULL, SVCM_E_STATE, "bucket out of range");
                    indices[cursor[refKey[slot]]++] = perm[slot];
                }
            }
        }
        return cudaDone(e, "svcm_hashgrid_download");
    }
    if(e == cudaSuccess && cellEnds)
        e = cudaMemcpy(cellEnds, g->grid.cellCount, sizeof(int32_t) * g->numCells, cudaMemcpyDeviceToHost);
    if(e == cudaSuccess && indices && g->count)
        e = cudaMemcpy(indices, g->grid.sIndex, sizeof(int32_t) * size_t(g->count), cudaMemcpyDeviceToHost);
    return cudaDone(e, "svcm_hashgrid_download");
}

int svcm_hashgrid_query(svcm_hashgrid *g, const SvcmSceneDesc *scene, const SvcmMergeQuery *queries, int64_t count,
                        uint32_t minPathLength, uint32_t maxPathLength, float misVcWeightFactor, int32_t ppm,
                        float *contrib, int32_t *counts)
{
    if(!g || !scene || count < 0 || count > 0x7FFFFF00ll || (count && (!queries || !contrib)))
        return fail(NULL, SVCM_E_INVALID, "bad hash grid query arguments");
    for(int64_t i = 0; i < count; i++)
        if(queries[i].matID < 0 || queries[i].matID >= scene->numMaterials || queries[i].pathLength > 0xFFFFu)
            return fail(NULL, SVCM_E_INVALID, "query %lld: material or path length out of range", (long long)i);
    DevScene sc;
    int rc = batchPrologue(scene, g->device, sc, NULL);
    if(rc) return rc;
    if(count == 0) return SVCM_OK;

    Scratch s; cudaError_t e = cudaSuccess;
    SvcmMergeQuery *dq = s.alloc<SvcmMergeQuery>(size_t(count), e);
    QueryStore qs;
    qs.q0 = s.alloc<float4>(size_t(count), e); qs.q1 = s.alloc<float4>(size_t(count), e);
    qs.q2 = s.alloc<float4>(size_t(count), e); qs.q3 = s.alloc<float4>(size_t(count), e);
    qs.q4 = NULL; qs.order = NULL;
    qs.capacity = uint32_t(count);
    float *dContrib = s.alloc<float>(size_t(count) * 3, e);
    int32_t *dCounts = s.alloc<int32_t>(size_t(count) * 2, e);
    if(e == cudaSuccess) e = cudaMemcpy(dq, queries, sizeof(SvcmMergeQuery) * size_t(count), cudaMemcpyHostToDevice);
    if(e == cudaSuccess)
    {
        IterConst ic; memset(&ic, 0, sizeof(ic));
        ic.radius = g->radius; ic.radiusSqr = g->radius * g->radius; ic.invCellSize = 1.f / (g->radius * 2.f);
        ic.numCells = g->numCells; ic.minLen = minPathLength; ic.maxLen = maxPathLength;
        ic.hashMagic = hashMagicFor(ic.numCells);
        ic.misVcWeightFactor = misVcWeightFactor; ic.ppm = ppm != 0; ic.useVM = 1;
        ic.vmNormalization = 1.f;             // the entry point returns the un-normalised mContrib
        launchPackMergeQueries(sc, dq, count, qs, 0);
        if(g->cellLayout)
        {
            // like the path: queries sorted into the particles' order, then the merge variant the
            // environment picks (SVCM_MERGE_VARIANT, default 1 = staged with bulk copies)
            QuerySortBuffers sort; memset(&sort, 0, sizeof(sort));
            sort.key = s.alloc<uint32_t>(size_t(count), e); sort.order = s.alloc<uint32_t>(size_t(count), e);
            sort.cellEnd = s.alloc<uint32_t>(size_t(g->grid.tcap) + 2 + 64, e);
            sort.scanState = s.alloc<unsigned long long>(kScanStateWords, e);
            sort.scanCounter = s.alloc<uint32_t>(1, e);
            if(e == cudaSuccess) e = cudaMemset(sort.scanState, 0, sizeof(unsigned long long) * kScanStateWords);
            if(e == cudaSuccess) e = cudaMemset(sort.scanCounter, 0, sizeof(uint32_t));
            if(e == cudaSuccess)
            {
                ScanGen gen; gen.value = 0;
                launchCellQuerySort(ic, qs, NULL, uint32_t(count), g->grid, sort, gen, NULL, 0);
                qs.order = sort.order;
                const char *variant = getenv("SVCM_MERGE_VARIANT");
                launchCellMerge(sc, ic, qs, NULL, uint32_t(count), g->grid, NULL, dContrib, dCounts, NULL,
                                variant ? atoi(variant) : 1, 0);
            }
        }
        else
        {
        // the parity tests exercise every merge kernel: the environment picks the variant
        const char *variant = getenv("SVCM_MERGE_GROUP");
        const int group = variant ? atoi(variant) : 1;
        if(group >= 2 && group <= 5)
            launchBucketStagedMerge(sc, ic, qs, NULL, uint32_t(count), g->grid, NULL, dContrib, dCounts, NULL, group, 0);
        else
            launchMerge(sc, ic, qs, NULL, uint32_t(count), g->grid, NULL, dContrib, dCounts, NULL,
                        (group == 8 || group == 16 || group == 32) ? group : 1, 0);
        }
        e = cudaMemcpy(contrib, dContrib, sizeof(float) * 3 * size_t(count), cudaMemcpyDeviceToHost);
        if(e == cudaSuccess && counts) e = cudaMemcpy(counts, dCounts, sizeof(int32_t) * 2 * size_t(count), cudaMemcpyDeviceToHost);
    }
    return cudaDone(e, "svcm_hashgrid_query");
}

int svcm_intersect_batch(const SvcmSceneDesc *scene, const float *rays, int64_t count, int32_t device, int32_t *hit,
                         float *dist, float *normal, int32_t *matID, int32_t *lightID)
{
    if(count < 0 || (count && (!rays || !hit || !dist || !normal || !matID || !lightID)))
        return fail(NULL, SVCM_E_INVALID, "bad arguments");
    SceneGeometry geo; DevScene sc; int rc = batchPrologue(scene, device, sc, &geo); if(rc) return rc;
    if(count == 0) return SVCM_OK;
    Scratch s; cudaError_t e = cudaSuccess; const size_t n = size_t(count);
    float *dRays = s.alloc<float>(n * 6, e); int32_t *dHit = s.alloc<int32_t>(n, e); float *dDist = s.alloc<float>(n, e);
    float *dN = s.alloc<float>(n * 3, e); int32_t *dMat = s.alloc<int32_t>(n, e); int32_t *dLight = s.alloc<int32_t>(n, e);
    if(e == cudaSuccess) e = cudaMemcpy(dRays, rays, sizeof(float) * 6 * n, cudaMemcpyHostToDevice);
    if(e == cudaSuccess)
    {
        launchIntersectBatch(sc, dRays, count, dHit, dDist, dN, dMat, dLight, 0);
        e = cudaMemcpy(hit, dHit, 4 * n, cudaMemcpyDeviceToHost);
        if(e == cudaSuccess) e = cudaMemcpy(dist, dDist, 4 * n, cudaMemcpyDeviceToHost);
        if(e == cudaSuccess) e = cudaMemcpy(normal, dN, 12 * n, cudaMemcpyDeviceToHost);
        if(e == cudaSuccess) e = cudaMemcpy(matID, dMat, 4 * n, cudaMemcpyDeviceToHost);
        if(e == cudaSuccess) e = cudaMemcpy(lightID, dLight, 4 * n, cudaMemcpyDeviceToHost);
    }
    return cudaDone(e, "svcm_intersect_batch");
}

int svcm_occluded_batch(const SvcmSceneDesc *scene, const float *segments, int64_t count, int32_t device, int32_t *occluded)
{
    if(count < 0 || (count && (!segments || !occluded))) return fail(NULL, SVCM_E_INVALID, "bad arguments");
    SceneGeometry geo; DevScene sc; int rc = batchPrologue(scene, device, sc, &geo); if(rc) return rc;
    if(count == 0) return SVCM_OK;
    Scratch s; cudaError_t e = cudaSuccess; const size_t n = size_t(count);
    float *dSeg = s.alloc<float>(n * 7, e); int32_t *dOcc = s.alloc<int32_t>(n, e);
    if(e == cudaSuccess) e = cudaMemcpy(dSeg, segments, sizeof(float) * 7 * n, cudaMemcpyHostToDevice);
    if(e == cudaSuccess)
    {
        launchOccludedBatch(sc, dSeg, count, dOcc, 0);
        e = cudaMemcpy(occluded, dOcc, 4 * n, cudaMemcpyDeviceToHost);
    }
    return cudaDone(e, "svcm_occluded_batch");
}

static int bsdfBatchCheck(const SvcmSceneDesc *scene, const float *in, int64_t count)
{
    for(int64_t i = 0; i < count; i++)
    {
        int32_t m; memcpy(&m, in + 10 * i + 9, 4);
        if(m < 0 || m >= scene->numMaterials) return fail(NULL, SVCM_E_INVALID, "item %lld: material %d out of range", (long long)i, m);
    }
    return SVCM_OK;
}

int svcm_bsdf_evaluate_batch(const SvcmSceneDesc *scene, const float *in, int64_t count, int32_t fixIsLight,
                             int32_t device, float *out)
{
    if(count < 0 || !scene || (count && (!in || !out))) return fail(NULL, SVCM_E_INVALID, "bad arguments");
    DevScene sc; int rc = batchPrologue(scene, device, sc, NULL); if(rc) return rc;
    if((rc = bsdfBatchCheck(scene, in, count))) return rc;
    if(count == 0) return SVCM_OK;
    Scratch s; cudaError_t e = cudaSuccess; const size_t n = size_t(count);
    float *dIn = s.alloc<float>(n * 10, e); float *dOut = s.alloc<float>(n * 8, e);
    if(e == cudaSuccess) e = cudaMemcpy(dIn, in, sizeof(float) * 10 * n, cudaMemcpyHostToDevice);
    if(e == cudaSuccess)
    {
        launchBsdfEvaluateBatch(sc, dIn, count, fixIsLight, dOut, 0);
        e = cudaMemcpy(out, dOut, sizeof(float) * 8 * n, cudaMemcpyDeviceToHost);
    }
    return cudaDone(e, "svcm_bsdf_evaluate_batch");
}

int svcm_bsdf_sample_batch(const SvcmSceneDesc *scene, const float *in, int64_t count, int32_t fixIsLight,
                           int32_t device, float *out, uint32_t *events)
{
    if(count < 0 || !scene || (count && (!in || !out || !events))) return fail(NULL, SVCM_E_INVALID, "bad arguments");
    DevScene sc; int rc = batchPrologue(scene, device, sc, NULL); if(rc) return rc;
    if((rc = bsdfBatchCheck(scene, in, count))) return rc;
    if(count == 0) return SVCM_OK;
    Scratch s; cudaError_t e = cudaSuccess; const size_t n = size_t(count);
    float *dIn = s.alloc<float>(n * 10, e); float *dOut = s.alloc<float>(n * 8, e); uint32_t *dEv = s.alloc<uint32_t>(n, e);
    if(e == cudaSuccess) e = cudaMemcpy(dIn, in, sizeof(float) * 10 * n, cudaMemcpyHostToDevice);
    if(e == cudaSuccess)
    {
        launchBsdfSampleBatch(sc, dIn, count, fixIsLight, dOut, dEv, 0);
        e = cudaMemcpy(out, dOut, sizeof(float) * 8 * n, cudaMemcpyDeviceToHost);
        if(e == cudaSuccess) e = cudaMemcpy(events, dEv, 4 * n, cudaMemcpyDeviceToHost);
    }
    return cudaDone(e, "svcm_bsdf_sample_batch");
}

static int lightBatch(const SvcmSceneDesc *scene, int op, int32_t lightIndex, const float *in, int64_t count,
                      int32_t device, float *out, const char *what)
{
    if(count < 0 || !scene || (count && (!in || !out))) return fail(NULL, SVCM_E_INVALID, "bad arguments");
    if(lightIndex < 0 || lightIndex >= scene->numLights) return fail(NULL, SVCM_E_INVALID, "light %d out of range", lightIndex);
    DevScene sc; int rc = batchPrologue(scene, device, sc, NULL); if(rc) return rc;
    if(count == 0) return SVCM_OK;
    Scratch s; cudaError_t e = cudaSuccess; const size_t n = size_t(count);
    float *dIn = s.alloc<float>(n * 8, e); float *dOut = s.alloc<float>(n * 12, e);
    if(e == cudaSuccess) e = cudaMemcpy(dIn, in, sizeof(float) * 8 * n, cudaMemcpyHostToDevice);
    if(e == cudaSuccess)
    {
        launchLightBatch(sc, op, lightIndex, dIn, count, dOut, 0);
        e = cudaMemcpy(out, dOut, sizeof(float) * 12 * n, cudaMemcpyDeviceToHost);
    }
    return cudaDone(e, what);
}

int svcm_light_emit_batch(const SvcmSceneDesc *scene, int32_t lightIndex, const float *in, int64_t count,
                          int32_t device, float *out)
{ return lightBatch(scene, 0, lightIndex, in, count, device, out, "svcm_light_emit_batch"); }

int svcm_light_illuminate_batch(const SvcmSceneDesc *scene, int32_t lightIndex, const float *in, int64_t count,
                                int32_t device, float *out)
{ return lightBatch(scene, 1, lightIndex, in, count, device, out, "svcm_light_illuminate_batch"); }

int svcm_light_radiance_batch(const SvcmSceneDesc *scene, int32_t lightIndex, const float *in, int64_t count,
                              int32_t device, float *out)
{ return lightBatch(scene, 2, lightIndex, in, count, device, out, "svcm_light_radiance_batch"); }

int svcm_camera_ray_batch(const SvcmSceneDesc *scene, const float *in, int64_t count, int32_t device, float *out)
{
    if(count < 0 || !scene || (count && (!in || !out))) return fail(NULL, SVCM_E_INVALID, "bad arguments");
    DevScene sc; int rc = batchPrologue(scene, device, sc, NULL); if(rc) return rc;
    if(count == 0) return SVCM_OK;
    Scratch s; cudaError_t e = cudaSuccess; const size_t n = size_t(count);
    float *dIn = s.alloc<float>(n * 5, e); float *dOut = s.alloc<float>(n * 6, e);
    if(e == cudaSuccess) e = cudaMemcpy(dIn, in, sizeof(float) * 5 * n, cudaMemcpyHostToDevice);
    if(e == cudaSuccess)
    {
        launchCameraBatch(sc, dIn, count, dOut, 0);
        e = cudaMemcpy(out, dOut, sizeof(float) * 6 * n, cudaMemcpyDeviceToHost);
    }
    return cudaDone(e, "svcm_camera_ray_batch");
}

int svcm_rng_stream(int32_t seed, int32_t iteration, int32_t phase, uint32_t firstPath, int32_t paths, int32_t draws,
                    float *out)
{
    if(paths < 0 || draws < 0 || !out) return fail(NULL, SVCM_E_INVALID, "bad arguments");
    for(int p = 0; p < paths; p++)
    {
        PathRng rng; rng.key = rngKey(seed, iteration, phase); rng.path = firstPath + uint32_t(p); rng.draw = 0u;
        for(int k = 0; k < draws; k++) out[size_t(p) * draws + k] = rng.next();
    }
    return SVCM_OK;
}

int svcm_rng_stream_device(int32_t seed, int32_t iteration, int32_t phase, uint32_t firstPath, int32_t paths,
                           int32_t draws, int32_t device, float *out)
{
    if(paths < 0 || draws < 0 || !out) return fail(NULL, SVCM_E_INVALID, "bad arguments");
    int rc = checkDevice(NULL, device); if(rc) return rc;
    cudaSetDevice(device);
    if(paths == 0 || draws == 0) return SVCM_OK;
    Scratch s; cudaError_t e = cudaSuccess; const size_t n = size_t(paths) * size_t(draws);
    float *d = s.alloc<float>(n, e);
    if(e == cudaSuccess)
    {
        launchRngStream(rngKey(seed, iteration, phase), firstPath, paths, draws, d, 0);
        e = cudaMemcpy(out, d, sizeof(float) * n, cudaMemcpyDeviceToHost);
    }
    return cudaDone(e, "svcm_rng_stream_device");
}

int svcm_image_convert_batch(const float *accum, int32_t width, int32_t height, int32_t iterations, int32_t format,
                             float gamma, int32_t device, uint8_t *out)
{
    const int64_t bytes = svcm_image_bytes(format, width, height);
    if(!accum || !out || bytes < 0 || (format == SVCM_IMAGE_BMP24 && !(gamma > 0.f)))
        return fail(NULL, SVCM_E_INVALID, "bad arguments");
    int rc = checkDevice(NULL, device); if(rc) return rc;
    cudaSetDevice(device);
    if(bytes == 0) return SVCM_OK;
    Scratch s; cudaError_t e = cudaSuccess; const size_t n = size_t(width) * size_t(height);
    float *dIn = s.alloc<float>(3 * n, e);
    uint8_t *dOut = s.alloc<uint8_t>(4 * n, e);
    if(e == cudaSuccess) e = cudaMemcpy(dIn, accum, sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
    if(e == cudaSuccess)
    {
        launchImageConvert(dIn, width, height, iterations, format, gamma, dOut, 0);
        e = cudaMemcpy(out, dOut, size_t(bytes), cudaMemcpyDeviceToHost);
    }
    return cudaDone(e, "svcm_image_convert_batch");
}

} // extern "C"
