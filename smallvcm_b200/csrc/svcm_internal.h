// svcm_internal.h — structures shared by the kernels (svcm_kernels.cu) and the host side
// of the C-ABI (svcm_api.cu). Not installed; include/svcm.h is the public face.
#ifndef SVCM_INTERNAL_H_
#define SVCM_INTERNAL_H_

#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "svcm.h"
#include "svcm_device.cuh"

namespace svcm {

// Per-iteration constants: the scalars VertexCM::RunIteration derives at its top
// (reference src/vertexcm.hxx:288-308) plus launch geometry. Travels as a kernel parameter.
struct IterConst
{
    float    radius, radiusSqr, invCellSize;
    float    vmNormalization;      // 1 / (pi r^2 N)
    float    misVmWeightFactor;    // eta_VCM  (0 unless merging)
    float    misVcWeightFactor;    // 1 / eta_VCM (0 unless connecting)
    float    lightSubPathCount;    // N = W*H (global, also with sharding)
    uint32_t minLen, maxLen;
    uint32_t keyLight, keyCamera;  // Philox keys for phase 0 / phase 1 of this iteration
    int32_t  useVC, useVM, ppm, lightTraceOnly;
    int32_t  resX, resY;
    int32_t  rank, worldSize;      // path p is owned by rank (p / resX) % worldSize
    uint32_t localPaths;
    uint32_t numCells;             // hash-grid cell count (= W*H, src/vertexcm.hxx:406)
    unsigned long long hashMagic;  // ceil(2^64 / numCells): exact x % numCells without a division (0: numCells == 1)
};

// SoA path state, one slot per queued sub-path; four 16-byte streams so that every
// access is one coalesced 128-bit load/store per lane.
//   s0 = origin.xyz      | dVCM
//   s1 = direction.xyz   | dVC
//   s2 = throughput.xyz  | dVM
//   s3 = path index | pathLength + flags | rng draw counter | light: stored-vertex count,
//                                                             camera: target pixel
struct PathQueue
{
    float4 *s0, *s1, *s2;
    uint4  *s3;
};
enum { PATH_FLAG_FINITE = 1u << 30, PATH_FLAG_SPECULAR = 1u << 31, PATH_LEN_MASK = (1u << 30) - 1u };

// Light vertices of the LOCAL light sub-paths (PathVertex<true>, src/vertexcm.hxx:79-98), stored
// twice, once per consumer:
//   merge record (52 B, all-gathered and cell-sorted; only when merging): posLen, thrVcm, dirVm,
//     contProb, in the order the vertices were created
//   connection record (64 B, local; only when connecting): `conn`, PATH-MAJOR — the k-th stored
//     vertex of local path l lies at conn[(l * connSlots + k) * 4 .. +3], so the vertices one
//     camera vertex connects to (its own light sub-path, src/vertexcm.hxx:504-506) are contiguous:
//       [0] hitpoint.xyz | pathLength   [1] throughput.xyz | dVCM
//       [2] shading normal.xyz | dVC    [3] local fixed direction.xyz | matID
//     pathVertCount[l] says how many there are (replaces mPathEnds)
struct VertexStore
{
    float4  *posLen;    // hitpoint.xyz | pathLength (bits)
    float4  *thrVcm;    // throughput.xyz | dVCM
    float4  *dirVm;     // WorldDirFix().xyz | dVM
    float   *contProb;  // light BSDF continuation probability
    float4  *conn;      // connection records, 4 x float4 per slot
    uint32_t connSlots; // slots per path (maxPathLength - 1)
    uint32_t capacity;
};

// The particle set of one grid build: the merge records of ALL light vertices.
//   single rank : the local streams, `*singleCount` records (device counter, clamped to capacity).
//   all ranks   : numSegs blocks of one all-gather, `segBytes` apart; the four pointers address the
//                 streams of block 0, every block holds `segStride` record slots of which the
//                 first segCountDev[seg] are valid (device array filled from the block headers).
struct MergeSet
{
    const float4 *posLen; const float4 *thrVcm; const float4 *dirVm; const float *contProb;
    const unsigned long long *singleCount;
    uint32_t capacity;
    uint32_t segStride;
    uint32_t numSegs;
    uint64_t segBytes;
    const uint32_t *segCountDev;
};

// Header of one rank's block in the light-vertex all-gather ("padded to equal per-rank
// capacity with the true count in a header slot"): the count and the rank's bbox, so that neither
// needs a host round trip or a second collective.
struct GatherHeader
{
    unsigned long long count;      // light vertices this rank stored (may exceed the slots: overflow)
    int32_t bbox[6];               // ordered-int min xyz / max xyz of them
    uint32_t pad[56];              // 256 bytes: the record streams behind it stay 256-byte aligned
};

// Geometry of the cell-layout grid (svcm_grid.cu), derived on the device from the particle bbox:
// storage cells are the reference's cells (edge 2r, src/hashgrid.hxx:47-48) with `m` of them
// merged along x when the grid is extremely fine; a storage row (fixed y, z) is CONTIGUOUS in key
// space with one empty guard cell either side, and rows are folded modulo (fy, fz) into a table of
// at most `tcap` keys. Folding only aliases cells >= 4 cells apart, which the distance test rejects.
//   key(sx, cy, cz) = ((cz mod fz) * fy + (cy mod fy)) * nxp + (sx + 1) + 1        (key 0: empty dummy)
struct GridDims
{
    int32_t  m;            // reference cells per storage cell along x (1 unless the grid is extremely fine)
    int32_t  nxp;          // storage cells along x + 2 guard cells
    int32_t  fy, fz;       // fold moduli (>= 4, or the whole extent + 2 when that is smaller)
    uint32_t numKeys;      // fz * fy * nxp  (<= tcap); keys run 1..numKeys
    int32_t  nx, ny, nz;   // reference cells spanned by the bbox
};

struct GridBuffers
{
    uint32_t *cellKey;    // [virtual index] hash bucket of each vertex (or ~0 for padding)
    uint32_t *cellCount;  // [numCells + 1] histogram -> exclusive scan (cell starts) -> cell ENDS
                          // after the scatter (the reference's mCellEnds, src/hashgrid.hxx:90-91)
    uint32_t *blockSums;  // scan scratch
    int32_t  *bboxBits;   // 6 ordered-int encoded floats: min xyz, max xyz
    const int32_t *bboxInit; // device copy of the empty box (+1e36 x3, -1e36 x3)
    // cell-sorted merge records
    float4   *sPosLen; float4 *sThrVcm; float4 *sDirVm; float *sContProb;
    int32_t  *sIndex;     // original (virtual) index at each sorted position  (mIndices)
    // ---- cell layout (svcm_grid.cu); NULL / 0 in the bucket layout ----
    GridDims *dims;       // device
    uint32_t *cellEnd;    // [tcap + 2] END offset of every storage cell; cellEnd[0] = 0 (dummy cell 0)
    uint32_t  tcap;
    uint32_t *refCount;   // optional [numCells]: particles per REFERENCE hash bucket (statistics / T1 only)
    unsigned long long *scanState;  // decoupled look-back scan: one word per tile
    uint32_t *scanCounter;          // dynamic tile ids (self-resetting)
};

// Camera vertices: every non-delta vertex of a camera sub-path that can still be connected or
// merged is appended here once by its bounce kernel and consumed twice afterwards — by
// k_connect (vertex connection, src/vertexcm.hxx:498-526) through the pair list and by the
// merge kernels as a RangeQuery (src/vertexcm.hxx:109-178, 530-538).
//   q0 = position.xyz | pixel (bits)      q1 = normal.xyz | pathLength | matID << 16 | flags << 24
//   q2 = local fixed dir.xyz | dVCM       q3 = throughput.xyz | dVM
//   q4 = dVC | local path index (bits) | - | -          (vertex connection only)
struct QueryStore
{
    float4 *q0, *q1, *q2, *q3, *q4;
    uint32_t capacity;
    // optional processing order for the merge: indices of the CAMV_MERGEABLE vertices grouped by
    // the hash bucket of their own cell (NULL: process 0..count-1 and skip the others)
    const uint32_t *order;
};

// scratch of the query sort (a second counting sort over the same bucket space as the grid)
struct QuerySortBuffers
{
    uint32_t *cellCount;   // [numCells + 1]
    uint32_t *key;         // [query capacity]
    uint32_t *order;       // [query capacity]
    uint32_t *blockSums;
    // cell layout: the queries are counting-sorted over the same key space as the particles
    uint32_t *cellEnd;     // [tcap + 2]
    unsigned long long *scanState;
    uint32_t *scanCounter;
};
enum { CAMV_MERGEABLE = 1u << 24 };   // inside the light-vertex bbox and merging is on

// (camera vertex index, k): connect camera vertex to the k-th stored vertex of its light path
struct PairStore
{
    uint2   *pairs;
    uint32_t capacity;
};

// Deferred shadow rays (Scene::Occluded calls of ConnectToCamera, DirectIllumination and
// ConnectVertices: src/vertexcm.hxx:928,734,805). The bounce kernels only EVALUATE the
// connection and append (segment, contribution, pixel); k_shadow then runs the any-hit test
// for a whole bounce with every lane busy and splats the unoccluded contributions.
//   r0 = point.xyz | tMax      r1 = direction.xyz | pixel (bits)      r2 = contribution.xyz | -
struct ShadowQueue
{
    float4 *r0, *r1, *r2;
    uint32_t *counts;       // [0] rays queued by the light pass, [1] by the camera pass
    uint32_t capacity;      // a full queue makes the producer trace its ray inline instead
};

// device-side counters (one block of 64-bit words, zeroed per iteration)
enum {
    CNT_LIGHT_VERTS = 0, CNT_QUERIES, CNT_CANDIDATES, CNT_ACCEPTED, CNT_LIGHT_RAYS,
    CNT_CAMERA_RAYS, CNT_SHADOW_RAYS, CNT_OVERFLOW, CNT_CAM_VERTS, CNT_PAIRS, CNT_WORDS
};

struct TraceArgs
{
    PathQueue   in, out;
    uint32_t   *queueCounts;     // [bounce index] number of live sub-paths entering that bounce
    VertexStore verts;
    uint8_t    *pathVertCount;   // [local path] number of stored vertices
    QueryStore  queries;
    PairStore   pairs;
    ShadowQueue shadow;
    unsigned long long *counters; // CNT_*
    float      *accum;           // W*H*3 accumulated image
    const int32_t *bboxBits;     // light-vertex bbox (camera pass: early reject of queries)
};

// ordered-int encoding of floats for atomicMin/atomicMax
SVCM_HD int32_t floatToOrdered(float f)
{
    int32_t i;
#ifdef __CUDA_ARCH__
    i = __float_as_int(f);
#else
    memcpy(&i, &f, 4);
#endif
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
SVCM_HD float orderedToFloat(int32_t i)
{
    i = i >= 0 ? i : i ^ 0x7FFFFFFF;
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}

// Generation counter of the look-back scan state words; bump once per scan launch.
struct ScanGen { uint32_t value; };

// ---- launchers implemented in svcm_kernels.cu (all asynchronous on `stream`) -------------
void launchLightBounce(const DevScene &sc, const IterConst &ic, const TraceArgs &args,
                       uint32_t bounce, uint32_t maxItems, cudaStream_t stream);
void launchCameraBounce(const DevScene &sc, const IterConst &ic, const TraceArgs &args,
                        uint32_t bounce, uint32_t maxItems, cudaStream_t stream);
// groups the mergeable camera vertices by cell (counting sort of indices): neighbouring lanes of
// the merge kernel then walk the same cell ranges. Returns the number of kernels launched.
int launchQuerySort(const IterConst &ic, const QueryStore &queries, const unsigned long long *vertexCount,
                    const int32_t *bboxBits, const QuerySortBuffers &qs, unsigned long long *counters, ScanGen *gen,
                    int spatial, cudaStream_t stream);
// vertex connection of every queued (camera vertex, light vertex) pair
void launchConnect(const DevScene &sc, const IterConst &ic, const TraceArgs &args, cudaStream_t stream);
// any-hit tests + splats for the rays queued by pass `which` (0 light, 1 camera; count read on
// the device)
void launchShadow(const DevScene &sc, const IterConst &ic, const TraceArgs &args, uint32_t which,
                  cudaStream_t stream);
// PathTracer (src/pathtracer.hxx:45-215) bounce and EyeLight (src/eyelight.hxx:47-78) pass
void launchPathTraceBounce(const DevScene &sc, const IterConst &ic, const TraceArgs &args,
                           uint32_t bounce, uint32_t maxItems, cudaStream_t stream);
void launchEyeLight(const DevScene &sc, const IterConst &ic, const TraceArgs &args, uint32_t iteration,
                    cudaStream_t stream);
// HashGrid::Build: bbox -> keys + histogram -> exclusive scan -> scatter. Returns the
// number of kernels launched.
int  launchGridBuild(const IterConst &ic, const MergeSet &set, const GridBuffers &grid, ScanGen *gen,
                     cudaStream_t stream);
// HashGrid::Process + RangeQuery::Process over the queued camera vertices flagged
// CAMV_MERGEABLE (their number is read from *queryCount on the device when queryCount != NULL,
// else the host value).
void launchMerge(const DevScene &sc, const IterConst &ic, const QueryStore &queries,
                 const unsigned long long *queryCount, uint32_t hostQueryCount,
                 const GridBuffers &grid, float *accum, float *contribOut, int32_t *countsOut,
                 unsigned long long *counters, int groupSize, cudaStream_t stream);

// ---- cell layout (svcm_grid.cu) ------------------------------------------------------------
// HashGrid::Build in the cell layout: bbox -> grid dims -> keys + histogram -> single-pass
// look-back scan -> index scatter -> record gather. `withRefCount`: also histogram the
// reference's hash buckets into grid.refCount (numCells entries). Returns kernels launched.
int  launchCellGridBuild(const IterConst &ic, const MergeSet &set, const GridBuffers &grid, ScanGen &gen,
                         bool withRefCount, cudaStream_t stream);
// counting sort of the mergeable camera vertices over the grid's key space -> qs.order
// `counters` != NULL (the path): vertices outside the particle bbox issue no query and are left out;
// the number of queries is added to counters[CNT_QUERIES]. NULL (kernel-level entry point): they
// are kept, parked in the dummy cell, so that every input receives its (zero) result.
int  launchCellQuerySort(const IterConst &ic, const QueryStore &queries, const unsigned long long *vertexCount,
                         uint32_t hostVertexCount, const GridBuffers &grid, const QuerySortBuffers &qs, ScanGen &gen,
                         unsigned long long *counters, cudaStream_t stream);
// the all-gather block headers (GatherHeader): bbox + count of the local vertices into `header`,
// and after the collective the per-block counts, the union bbox and the total out of all headers
// bbox of the particles into bboxBits (ordered-int atomics; the caller initialised it)
void launchParticleBBox(const MergeSet &set, int32_t *bboxBits, cudaStream_t stream);
void launchGatherPackHeader(const MergeSet &localSet, const int32_t *bboxInit, GatherHeader *header, cudaStream_t stream);
void launchGatherUnpackHeaders(const unsigned char *blocks, uint64_t blockBytes, uint32_t numSegs, uint32_t segStride,
                               uint32_t *segCountDev, int32_t *bboxBits, unsigned long long *totals, cudaStream_t stream);
// merge over the cell layout. variant: 0 = lane per query straight from global memory, 1 = staged
// through shared memory with bulk async copies, 3 = the same with smaller stages
void launchCellMerge(const DevScene &sc, const IterConst &ic, const QueryStore &queries,
                     const unsigned long long *queryCount, uint32_t hostQueryCount,
                     const GridBuffers &grid, float *accum, float *contribOut, int32_t *countsOut,
                     unsigned long long *counters, int variant, cudaStream_t stream);
// single-pass decoupled look-back exclusive scan of data[0..n) (dims: optional cap on n, cell layout);
// needs (n + 4095) / 4096 state words and a zeroed counter
void launchLookbackScan(uint32_t *data, const GridDims *dims, uint32_t n, unsigned long long *state,
                        uint32_t *counter, ScanGen &gen, cudaStream_t stream);
static const uint32_t kScanStateWords = 8192;      // scans of up to 2^25 entries
int launchBucketReorder(const MergeSet &set, const GridBuffers &grid, uint32_t numCells, cudaStream_t stream);
// merge over the BUCKET layout with a warp's buckets staged into shared memory. variant: 2 = bulk
// async copies (default), 3 = plain loads, 4 = bulk with smaller stages
void launchBucketStagedMerge(const DevScene &sc, const IterConst &ic, const QueryStore &queries,
                             const unsigned long long *queryCount, uint32_t hostQueryCount,
                             const GridBuffers &grid, float *accum, float *contribOut, int32_t *countsOut,
                             unsigned long long *counters, int variant, cudaStream_t stream);
// per sorted slot: the reference hash bucket of that particle (T1 export of the cell layout)
void launchCellRefKeys(const IterConst &ic, const GridBuffers &grid, uint32_t count, uint32_t *out, cudaStream_t stream);
uint64_t hashMagicFor(uint32_t numCells);

// kernel-level batch entry points (device pointers)
void launchIntersectBatch(const DevScene &sc, const float *rays, int64_t count, int32_t *hit,
                          float *dist, float *normal, int32_t *matID, int32_t *lightID,
                          cudaStream_t stream);
void launchOccludedBatch(const DevScene &sc, const float *segments, int64_t count,
                         int32_t *occluded, cudaStream_t stream);
void launchBsdfEvaluateBatch(const DevScene &sc, const float *in, int64_t count, int fixIsLight,
                             float *out, cudaStream_t stream);
void launchBsdfSampleBatch(const DevScene &sc, const float *in, int64_t count, int fixIsLight,
                           float *out, uint32_t *events, cudaStream_t stream);
// op 0 Emit, 1 Illuminate, 2 GetRadiance of light `lightIndex`: 8 floats in, 12 out per item
void launchLightBatch(const DevScene &sc, int op, int lightIndex, const float *in, int64_t count, float *out,
                      cudaStream_t stream);
// Camera::GenerateRay + WorldToRaster: 5 floats in, 6 out per item
void launchCameraBatch(const DevScene &sc, const float *in, int64_t count, float *out, cudaStream_t stream);
// svcm_bvh.cpp: hierarchy over a primitive list; order[slot] = original index. 0 on success.
int bvhBuild(const SvcmPrimitive *prims, int32_t count, std::vector<BvhNode> &nodes, std::vector<int32_t> &order);
int primitiveRuns(const SvcmPrimitive *prims, int32_t count, bool plain, BvhNode *runs, uint32_t *cullMask);

// SaveBMP / SaveHDR pixel payloads of the accumulated image (src/framebuffer.hxx:170-251)
void launchImageConvert(const float *accum, int width, int height, int iterations, int format, float gamma,
                        uint8_t *out, cudaStream_t stream);
void launchRngStream(uint32_t key, uint32_t firstPath, int paths, int draws, float *out,
                     cudaStream_t stream);
// packs host-layout records into the device streams
void launchPackMergeVertices(const SvcmMergeVertex *verts, int64_t count, float4 *posLen,
                             float4 *thrVcm, float4 *dirVm, float *contProb, cudaStream_t stream);
void launchPackMergeQueries(const DevScene &sc, const SvcmMergeQuery *queries, int64_t count,
                            QueryStore store, cudaStream_t stream);

} // namespace svcm
#endif
