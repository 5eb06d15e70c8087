// svcm_grid.cu — the CELL-LAYOUT hash grid (sm_100a): HashGrid::Build and HashGrid::Process
// (reference src/hashgrid.hxx:41-169) with RangeQuery::Process (src/vertexcm.hxx:130-169) as the
// per-particle body, re-designed around spatial coherence.
//
// The reference keeps its particles in the order of `hash(cell) % (W*H)` buckets, so the 8 cells a
// query visits are 8 unrelated places of memory. Which particles a query ACCEPTS does not depend
// on that order: a particle within the radius lies in one of the query's 8 cells, and everything
// else a bucket holds (aliases of far-away cells) fails the distance test. The only visible trace
// of the hashing is that a bucket shared by two of the 8 cells is visited twice
// (src/hashgrid.hxx:142-155, not de-duplicated), which counts its in-range particles twice.
//
// So the particles are stored here in TRUE cell order — key = (folded row (y, z), x), a storage
// row contiguous in memory — and so are the queries. A warp takes 32 consecutive sorted queries,
// i.e. queries of one row and of a few neighbouring cells; the 3x3 rows around them are NINE
// contiguous runs of the sorted position stream, which are staged into shared memory with bulk
// asynchronous copies (cp.async.bulk + mbarrier, double-buffered when a neighbourhood exceeds one
// stage) and tested from there. The reference's double visits are reproduced arithmetically: the
// 8 reference hashes of a query are compared with each other and an accepted particle is counted
// as often as the bucket of ITS reference cell occurs among them. Accepted pairs go to per-lane
// lists and are evaluated flattened, 32 pairs per warp step, whatever the acceptance rate.
//
// Build = device counting sort over the key space: keys + histogram, ONE single-pass
// decoupled-look-back scan, a 4-byte index scatter and a gather that writes the four record
// streams densely (random 16-byte reads instead of random 16-byte writes of half-filled sectors).
#include <cuda_runtime.h>
#include <stdint.h>
#include "svcm_internal.h"
#include "svcm_kutil.cuh"

namespace svcm {

// ------------------------------------------------------------------------------------------
// keys
// ------------------------------------------------------------------------------------------

// floor(c / m) for c >= -1
SVCM_D int32_t storageX(int32_t cx, int32_t m) { return (m == 1) ? cx : (cx + m) / m - 1; }
// c mod f for c >= -1
SVCM_D int32_t foldCoord(int32_t c, int32_t f) { return (c < 0) ? f - 1 : (c < f ? c : c % f); }
// key of the storage cell (sx, folded row ry, rz); sx in [-1, nxp - 2]
SVCM_D uint32_t rowBaseKey(const GridDims &d, int32_t ry, int32_t rz)
{
    return uint32_t(rz * d.fy + ry) * uint32_t(d.nxp) + 2u;   // + (sx + 1) + 1 with sx = 0 folded in below
}
SVCM_D uint32_t cellKey(const GridDims &d, int32_t sx, int32_t ry, int32_t rz)
{
    return uint32_t(int32_t(rowBaseKey(d, ry, rz)) + sx);     // = row * nxp + (sx + 1) + 1
}

// reference cell coordinate of a point (src/hashgrid.hxx:189-202), clamped so that the int
// conversion is defined for any input
SVCM_D int32_t cellCoord(float p, float bbMin, float invCellSize)
{
    const float c = floorf(invCellSize * (p - bbMin));
    return int32_t(fminf(fmaxf(c, -2.f), 1073741824.f));
}

// grid geometry from the bbox; one thread
__global__ void k_cellDims(const int32_t *bboxBits, float invCellSize, uint32_t tcap, GridDims *out)
{
    if(blockIdx.x != 0 || threadIdx.x != 0) return;
    const float mn[3] = { orderedToFloat(bboxBits[0]), orderedToFloat(bboxBits[1]), orderedToFloat(bboxBits[2]) };
    const float mx[3] = { orderedToFloat(bboxBits[3]), orderedToFloat(bboxBits[4]), orderedToFloat(bboxBits[5]) };
    int32_t n[3];
    for(int a = 0; a < 3; a++)
    {
        n[a] = 1;
        if(mx[a] >= mn[a]) n[a] = cellCoord(mx[a], mn[a], invCellSize) + 1;
        if(n[a] < 1) n[a] = 1;
    }
    GridDims d;
    d.nx = n[0]; d.ny = n[1]; d.nz = n[2];
    const int32_t maxRow = int32_t(tcap / 16u) - 2;            // a row may use 1/16 of the table at most
    d.m = (n[0] + maxRow - 1) / maxRow;
    if(d.m < 1) d.m = 1;
    d.nxp = (n[0] + d.m - 1) / d.m + 2;
    int32_t f = int32_t(sqrtf(float(tcap / uint32_t(d.nxp))));
    while(uint64_t(f + 1) * uint64_t(f + 1) * uint64_t(d.nxp) <= uint64_t(tcap)) f++;
    while(f > 1 && uint64_t(f) * uint64_t(f) * uint64_t(d.nxp) > uint64_t(tcap)) f--;
    d.fy = min(f, n[1] + 2);
    d.fz = min(f, n[2] + 2);
    d.numKeys = uint32_t(d.fz) * uint32_t(d.fy) * uint32_t(d.nxp);
    *out = d;
}

// key of every particle + histogram (the counting half of the sort)
__global__ void __launch_bounds__(SVCM_BLOCK)
k_cellKeys(const __grid_constant__ MergeSet set, const __grid_constant__ IterConst ic, const int32_t *bboxBits,
           const GridDims *dims, uint32_t *key, uint32_t *cellEnd, uint32_t *refCount)
{
    const GridDims d = *dims;
    const f3 bbMin = mk3(orderedToFloat(bboxBits[0]), orderedToFloat(bboxBits[1]), orderedToFloat(bboxBits[2]));
    const uint32_t total = mergeSetTotal(set);
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    {
        uint32_t k = 0xFFFFFFFFu;
        const MergeSlot ms = mergeSetSlot(set, i);
        if(mergeSetValid(set, ms))
        {
            const float4 p = *mergeSetStream(set, set.posLen, ms);
            const int32_t cx = cellCoord(p.x, bbMin.x, ic.invCellSize), cy = cellCoord(p.y, bbMin.y, ic.invCellSize),
                          cz = cellCoord(p.z, bbMin.z, ic.invCellSize);
            k = cellKey(d, storageX(cx, d.m), foldCoord(cy, d.fy), foldCoord(cz, d.fz));
            atomicAdd(cellEnd + k, 1u);
            if(refCount) atomicAdd(refCount + refHash(cx, cy, cz, ic), 1u);
        }
        key[i] = k;
    }
}

// key of every mergeable camera vertex (its own cell) + histogram
__global__ void __launch_bounds__(SVCM_BLOCK)
k_cellQueryKeys(const __grid_constant__ QueryStore qs, const unsigned long long *vertexCount, uint32_t hostCount,
                const __grid_constant__ IterConst ic, const int32_t *bboxBits, const GridDims *dims, uint32_t *key,
                uint32_t *cellEnd, unsigned long long *counters)
{
    const GridDims d = *dims;
    const uint32_t n = vertexCount ? uint32_t(min(*vertexCount, (unsigned long long)qs.capacity)) : hostCount;
    const f3 bbMin = mk3(orderedToFloat(bboxBits[0]), orderedToFloat(bboxBits[1]), orderedToFloat(bboxBits[2]));
    const f3 bbMax = mk3(orderedToFloat(bboxBits[3]), orderedToFloat(bboxBits[4]), orderedToFloat(bboxBits[5]));
    uint32_t nInside = 0u;
    for(uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x)
    {
        const uint32_t i = i0 + threadIdx.x;
        uint32_t k = 0xFFFFFFFFu;
        if(i < n && (__float_as_uint(qs.q1[i].w) & CAMV_MERGEABLE))
        {
            const float4 p = qs.q0[i];
            // HashGrid::Process early-out (src/hashgrid.hxx:116-122): a vertex outside the particle
            // bbox issues no query. In the path it is left out of the order; the kernel-level entry
            // point parks it in the dummy cell so that it still receives its (zero) result.
            const bool inside = !(p.x - bbMin.x < 0.f || bbMax.x - p.x < 0.f || p.y - bbMin.y < 0.f ||
                                  bbMax.y - p.y < 0.f || p.z - bbMin.z < 0.f || bbMax.z - p.z < 0.f);
            if(inside)
            {
                k = cellKey(d, storageX(cellCoord(p.x, bbMin.x, ic.invCellSize), d.m),
                            foldCoord(cellCoord(p.y, bbMin.y, ic.invCellSize), d.fy),
                            foldCoord(cellCoord(p.z, bbMin.z, ic.invCellSize), d.fz));
                nInside++;
            }
            else if(!counters) k = 0u;
            if(k != 0xFFFFFFFFu) atomicAdd(cellEnd + k, 1u);
        }
        if(i < n) key[i] = k;
    }
    if(counters) warpCount(counters + CNT_QUERIES, nInside);
}

// bbox + count of the LOCAL light vertices into this rank's all-gather header
__global__ void k_gatherHeaderCount(const unsigned long long *count, GatherHeader *header)
{
    if(blockIdx.x == 0 && threadIdx.x == 0) header->count = *count;
}

// after the all-gather: per-block counts (clamped to the slots a block has), the union of the
// blocks' bboxes = the bbox of all particles (src/hashgrid.hxx:50-61), and the totals
__global__ void k_gatherUnpackHeaders(const unsigned char *blocks, uint64_t blockBytes, uint32_t numSegs, uint32_t segStride,
                                      uint32_t *segCountDev, int32_t *bboxBits, unsigned long long *totals)
{
    const uint32_t lane = threadIdx.x;
    int32_t bb[6] = { floatToOrdered(1e36f), floatToOrdered(1e36f), floatToOrdered(1e36f),
                      floatToOrdered(-1e36f), floatToOrdered(-1e36f), floatToOrdered(-1e36f) };
    unsigned long long cnt = 0ull;
    if(lane < numSegs)
    {
        const GatherHeader *h = reinterpret_cast<const GatherHeader*>(blocks + uint64_t(lane) * blockBytes);
        cnt = h->count;
        if(cnt > 0ull) for(int k = 0; k < 6; k++) bb[k] = h->bbox[k];
        segCountDev[lane] = uint32_t(min(cnt, (unsigned long long)segStride));
        totals[lane] = cnt;
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
        for(int k = 0; k < 3; k++) bb[k] = min(bb[k], __shfl_xor_sync(FULL_MASK, bb[k], o));
        for(int k = 3; k < 6; k++) bb[k] = max(bb[k], __shfl_xor_sync(FULL_MASK, bb[k], o));
    }
    if(lane < 6u) bboxBits[lane] = bb[lane];
}

// ------------------------------------------------------------------------------------------
// single-pass exclusive scan with decoupled look-back (reference src/hashgrid.hxx:75-81)
// state word: generation << 34 | flag << 32 | value;  flag 1 = tile aggregate, 2 = inclusive prefix
// ------------------------------------------------------------------------------------------
SVCM_D unsigned long long globalTimerNs()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// a wait that cannot hang the device: two seconds without progress is a bug, trap and report it
#define SVCM_SPIN_LIMIT_NS 2000000000ull
#define LB_ITEMS 16
#define LB_TILE (SVCM_BLOCK * LB_ITEMS)

SVCM_D unsigned long long lbPack(uint32_t gen, uint32_t flag, uint32_t value)
{
    return ((unsigned long long)gen << 34) | ((unsigned long long)flag << 32) | (unsigned long long)value;
}

__global__ void __launch_bounds__(SVCM_BLOCK)
k_scanLookback(uint32_t *data, const GridDims *dims, uint32_t nHost, volatile unsigned long long *state,
               uint32_t *tileCounter, uint32_t numTiles, uint32_t gen)
{
    __shared__ uint32_t sTile;
    __shared__ uint32_t sWarp[SVCM_BLOCK / 32 + 1];
    __shared__ uint32_t sPrefix;
    // tiles are handed out in launch order so that a tile's predecessors are always running or done;
    // the counter wraps back to zero with the last tile (no memset between launches)
    if(threadIdx.x == 0) sTile = atomicInc(tileCounter, numTiles - 1u);
    __syncthreads();
    const uint32_t tile = sTile;
    // everything beyond the keys in use is zero and needs no result; the tile still publishes
    const uint32_t n = dims ? min(nHost, dims->numKeys + 2u) : nHost;
    const uint32_t base = tile * LB_TILE + threadIdx.x * LB_ITEMS;

    uint32_t v[LB_ITEMS], sum = 0u;
    if(base + LB_ITEMS <= n)
    {
        const uint4 *p = reinterpret_cast<const uint4*>(data + base);
#pragma unroll
        for(int k = 0; k < LB_ITEMS / 4; k++)
        {
            const uint4 t = p[k];
            v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        }
    }
    else
    {
#pragma unroll
        for(int k = 0; k < LB_ITEMS; k++) v[k] = (base + k < n) ? data[base + k] : 0u;
    }
#pragma unroll
    for(int k = 0; k < LB_ITEMS; k++) sum += v[k];

    // block scan of the per-thread sums
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(FULL_MASK, inc, o);
        if(lane >= uint32_t(o)) inc += t;
    }
    if(lane == 31u) sWarp[warp] = inc;
    __syncthreads();
    if(warp == 0u)
    {
        const uint32_t wv = (lane < SVCM_BLOCK / 32) ? sWarp[lane] : 0u;
        uint32_t winc = wv;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(FULL_MASK, winc, o);
            if(lane >= uint32_t(o)) winc += t;
        }
        if(lane < SVCM_BLOCK / 32) sWarp[lane] = winc - wv;          // exclusive warp offsets
        const uint32_t aggregate = __shfl_sync(FULL_MASK, winc, 31);

        // publish the aggregate, look back for the exclusive prefix of this tile, publish the inclusive
        if(lane == 0u) state[tile] = lbPack(gen, tile == 0u ? 2u : 1u, aggregate);
        uint32_t prefix = 0u;
        if(tile > 0u)
        {
            int32_t lookBase = int32_t(tile) - 1;
            for(;;)
            {
                const int32_t t = lookBase - int32_t(lane);
                unsigned long long s;
                uint32_t flag;
                uint32_t spins = 0u;
                unsigned long long t0 = 0ull;
                do
                {
                    s = (t >= 0) ? state[t] : lbPack(gen, 2u, 0u);
                    flag = (uint32_t(s >> 34) == gen) ? (uint32_t(s >> 32) & 3u) : 0u;
                    if((++spins & 1023u) == 0u)                      // never hang the device
                    {
                        if(t0 == 0ull) t0 = globalTimerNs();
                        else if(globalTimerNs() - t0 > SVCM_SPIN_LIMIT_NS) __trap();
                    }
                } while(__any_sync(FULL_MASK, flag == 0u));
                const uint32_t inclMask = __ballot_sync(FULL_MASK, flag == 2u);
                const int first = __ffs(inclMask) - 1;               // nearest predecessor with an inclusive prefix
                uint32_t c = (first < 0 || int(lane) <= first) ? uint32_t(s) : 0u;
#pragma unroll
                for(int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
                prefix += c;
                if(first >= 0) break;
                lookBase -= 32;
            }
            if(lane == 0u) state[tile] = lbPack(gen, 2u, prefix + aggregate);
        }
        if(lane == 0u) sPrefix = prefix;
    }
    __syncthreads();

    uint32_t run = sPrefix + sWarp[warp] + inc - sum;
    if(base + LB_ITEMS <= n)
    {
        uint4 *p = reinterpret_cast<uint4*>(data + base);
#pragma unroll
        for(int k = 0; k < LB_ITEMS / 4; k++)
        {
            uint4 t;
            t.x = run; run += v[4 * k];
            t.y = run; run += v[4 * k + 1];
            t.z = run; run += v[4 * k + 2];
            t.w = run; run += v[4 * k + 3];
            p[k] = t;
        }
    }
    else
    {
#pragma unroll
        for(int k = 0; k < LB_ITEMS; k++) { if(base + k < n) data[base + k] = run; run += v[k]; }
    }
}

// scatter of the 4-byte source index into cell order; like the reference the scanned array is
// the cursor, so afterwards cellEnd[k] is the END of cell k (src/hashgrid.hxx:83-91)
__global__ void __launch_bounds__(SVCM_BLOCK)
k_cellScatterIndex(const uint32_t *key, const unsigned long long *countPtr, uint32_t hostCount, uint32_t capacity,
                   uint32_t *cellEnd, uint32_t *perm)
{
    const uint32_t total = countPtr ? uint32_t(min(*countPtr, (unsigned long long)capacity)) : hostCount;
    for(uint32_t i0 = blockIdx.x * blockDim.x * 4u + threadIdx.x; i0 < total; i0 += gridDim.x * blockDim.x * 4u)
    {
        uint32_t k[4], dst[4];
#pragma unroll
        for(int j = 0; j < 4; j++)
        {
            const uint32_t i = i0 + uint32_t(j) * blockDim.x;
            k[j] = (i < total) ? key[i] : 0xFFFFFFFFu;
        }
#pragma unroll
        for(int j = 0; j < 4; j++) dst[j] = (k[j] != 0xFFFFFFFFu) ? atomicAdd(cellEnd + k[j], 1u) : 0u;
#pragma unroll
        for(int j = 0; j < 4; j++)
            if(k[j] != 0xFFFFFFFFu) perm[dst[j]] = i0 + uint32_t(j) * blockDim.x;
    }
}

// slot-ordered gather of the records: dense 16-byte stores, random 16-byte loads
__global__ void __launch_bounds__(SVCM_BLOCK)
k_cellGather(const __grid_constant__ MergeSet set, const __grid_constant__ GridBuffers g, const uint32_t *totalPtr)
{
    // cell layout: the end of the last key; bucket layout: the scan's last entry (never a cursor)
    const uint32_t total = totalPtr ? *totalPtr : g.cellEnd[g.dims->numKeys];
    for(uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += gridDim.x * blockDim.x)
    {
        const MergeSlot ms = mergeSetSlot(set, uint32_t(g.sIndex[slot]));
        g.sPosLen[slot]   = *mergeSetStream(set, set.posLen, ms);
        g.sThrVcm[slot]   = *mergeSetStream(set, set.thrVcm, ms);
        g.sDirVm[slot]    = *mergeSetStream(set, set.dirVm, ms);
        g.sContProb[slot] = *mergeSetStream(set, set.contProb, ms);
    }
}

// T1 export: the reference hash bucket of the particle at each sorted slot
__global__ void __launch_bounds__(SVCM_BLOCK)
k_cellRefKeys(const __grid_constant__ IterConst ic, const __grid_constant__ GridBuffers g, uint32_t count, uint32_t *out)
{
    const f3 bbMin = mk3(orderedToFloat(g.bboxBits[0]), orderedToFloat(g.bboxBits[1]), orderedToFloat(g.bboxBits[2]));
    for(uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < count; s += gridDim.x * blockDim.x)
    {
        const float4 p = g.sPosLen[s];
        out[s] = refHash(cellCoord(p.x, bbMin.x, ic.invCellSize), cellCoord(p.y, bbMin.y, ic.invCellSize),
                         cellCoord(p.z, bbMin.z, ic.invCellSize), ic);
    }
}

// bbox of the particles (reference src/hashgrid.hxx:50-61): per-thread min/max, shuffle reduction
// per warp, shared-memory reduction per block, then six ordered-int atomics per BLOCK (atomics on
// six addresses from every warp were most of this kernel's time)
__global__ void __launch_bounds__(SVCM_BLOCK)
k_cellBBox(const __grid_constant__ MergeSet set, int32_t *bboxBits)
{
    __shared__ float sMin[3][SVCM_BLOCK / 32], sMax[3][SVCM_BLOCK / 32];
    f3 mn = mk3(1e36f), mx = mk3(-1e36f);
    const uint32_t total = mergeSetTotal(set);
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    {
        const MergeSlot ms = mergeSetSlot(set, i);
        if(!mergeSetValid(set, ms)) continue;
        const float4 p = *mergeSetStream(set, set.posLen, ms);
        mn = mk3(fminf(mn.x, p.x), fminf(mn.y, p.y), fminf(mn.z, p.z));
        mx = mk3(fmaxf(mx.x, p.x), fmaxf(mx.y, p.y), fmaxf(mx.z, p.z));
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
        mn.x = fminf(mn.x, __shfl_xor_sync(FULL_MASK, mn.x, o));
        mn.y = fminf(mn.y, __shfl_xor_sync(FULL_MASK, mn.y, o));
        mn.z = fminf(mn.z, __shfl_xor_sync(FULL_MASK, mn.z, o));
        mx.x = fmaxf(mx.x, __shfl_xor_sync(FULL_MASK, mx.x, o));
        mx.y = fmaxf(mx.y, __shfl_xor_sync(FULL_MASK, mx.y, o));
        mx.z = fmaxf(mx.z, __shfl_xor_sync(FULL_MASK, mx.z, o));
    }
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
    if(lane == 0u)
    {
        sMin[0][warp] = mn.x; sMin[1][warp] = mn.y; sMin[2][warp] = mn.z;
        sMax[0][warp] = mx.x; sMax[1][warp] = mx.y; sMax[2][warp] = mx.z;
    }
    __syncthreads();
    if(warp == 0u && lane < 6u)
    {
        const bool isMax = lane >= 3u;
        const float (*src)[SVCM_BLOCK / 32] = isMax ? sMax : sMin;
        const uint32_t c = isMax ? lane - 3u : lane;
        float v = src[c][0];
        for(uint32_t k = 1; k < blockDim.x / 32u; k++) v = isMax ? fmaxf(v, src[c][k]) : fminf(v, src[c][k]);
        // an empty block holds (+1e36, -1e36): the atomics leave the box as it is
        if(isMax) atomicMax(bboxBits + lane, floatToOrdered(v));
        else      atomicMin(bboxBits + lane, floatToOrdered(v));
    }
}

// ------------------------------------------------------------------------------------------
// the query side shared by the merge kernels
// ------------------------------------------------------------------------------------------

// The 8 cells of HashGrid::Process (src/hashgrid.hxx:124-155) and how often the reference visits
// the bucket of each: nibble j of `mults` = number of i with hash(cell_i) == hash(cell_j), cell j =
// (j & 4 ? pxo : px, j & 2 ? pyo : py, j & 1 ? pzo : pz).
struct QueryCells
{
    int32_t px, py, pz, pxo, pyo, pzo;
    uint32_t mults;
};

SVCM_D QueryCells queryCells(const IterConst &ic, f3 qpos, f3 bbMin, uint32_t *refCandidates, const uint32_t *refCount)
{
    QueryCells c;
    const f3 cellPt = mk3(ic.invCellSize * (qpos.x - bbMin.x), ic.invCellSize * (qpos.y - bbMin.y),
                          ic.invCellSize * (qpos.z - bbMin.z));
    const f3 coordF = mk3(floorf(cellPt.x), floorf(cellPt.y), floorf(cellPt.z));
    c.px = int(coordF.x); c.py = int(coordF.y); c.pz = int(coordF.z);
    c.pxo = c.px + ((cellPt.x - coordF.x) < 0.5f ? -1 : +1);
    c.pyo = c.py + ((cellPt.y - coordF.y) < 0.5f ? -1 : +1);
    c.pzo = c.pz + ((cellPt.z - coordF.z) < 0.5f ? -1 : +1);
    uint32_t h[8];
#pragma unroll
    for(int j = 0; j < 8; j++)
        h[j] = refHash((j & 4) ? c.pxo : c.px, (j & 2) ? c.pyo : c.py, (j & 1) ? c.pzo : c.pz, ic);
    bool any = false;
#pragma unroll
    for(int j = 1; j < 8; j++)
#pragma unroll
        for(int i = 0; i < j; i++) any = any || (h[i] == h[j]);
    c.mults = 0x11111111u;
    if(any)
    {
        c.mults = 0u;
#pragma unroll
        for(int j = 0; j < 8; j++)
        {
            uint32_t n = 0u;
#pragma unroll
            for(int i = 0; i < 8; i++) n += (h[i] == h[j]) ? 1u : 0u;
            c.mults |= n << (4 * j);
        }
    }
    if(refCount)
    {
        uint32_t n = 0u;
#pragma unroll
        for(int j = 0; j < 8; j++) n += refCount[h[j]];
        *refCandidates = n;
    }
    return c;
}

// how often the reference counts an in-range particle at `p`: the multiplicity of the bucket of
// the particle's own cell if that cell is one of the query's 8, else 0
SVCM_D uint32_t acceptWeight(const QueryCells &c, const IterConst &ic, f3 bbMin, float4 p)
{
    // (a particle within the radius always lies in one of the 8 cells up to rounding at a cell face;
    // checked anyway so that the result is the reference's in every case)
    const int32_t cx = cellCoord(p.x, bbMin.x, ic.invCellSize), cy = cellCoord(p.y, bbMin.y, ic.invCellSize),
                  cz = cellCoord(p.z, bbMin.z, ic.invCellSize);
    const int jx = (cx == c.px) ? 0 : ((cx == c.pxo) ? 4 : -1);
    const int jy = (cy == c.py) ? 0 : ((cy == c.pyo) ? 2 : -1);
    const int jz = (cz == c.pz) ? 0 : ((cz == c.pzo) ? 1 : -1);
    if((jx | jy | jz) < 0) return 0u;
    return (c.mults >> (4 * (jx | jy | jz))) & 15u;
}

// per-query words parked in shared memory for the pair evaluation: [word][owner lane]
#define CQ_WORDS 18

SVCM_D void parkQuery(float (*q)[32], uint32_t lane, const Bsdf &bsdf, float lenMatBits, float dVCM, float dVM)
{
    q[0][lane] = bsdf.frame.x.x; q[1][lane] = bsdf.frame.x.y; q[2][lane] = bsdf.frame.x.z;
    q[3][lane] = bsdf.frame.y.x; q[4][lane] = bsdf.frame.y.y; q[5][lane] = bsdf.frame.y.z;
    q[6][lane] = bsdf.frame.z.x; q[7][lane] = bsdf.frame.z.y; q[8][lane] = bsdf.frame.z.z;
    q[9][lane] = bsdf.localDirFix.x; q[10][lane] = bsdf.localDirFix.y; q[11][lane] = bsdf.localDirFix.z;
    q[12][lane] = lenMatBits;
    q[13][lane] = bsdf.diffProb; q[14][lane] = bsdf.phongProb; q[15][lane] = bsdf.contProb;
    q[16][lane] = dVCM; q[17][lane] = dVM;
}

// RangeQuery::Process (reference src/vertexcm.hxx:130-169) for the particle at sorted slot `idx`
// and the query parked by lane `owner`; the result is added to the owner's sums
// (the particle's payload is loaded by the caller, all four words at once: one memory round trip
// per evaluation step instead of three dependent ones)
SVCM_D f3 pairContribution(const DevScene &sc, const IterConst &ic, float (*q)[32], uint32_t owner,
                           float lenBits, float4 dv, float4 tv, float lightContProb)
{
    const uint32_t lvLen = __float_as_uint(lenBits);
    const uint32_t lenMat = __float_as_uint(q[12][owner]);
    const uint32_t camLen = lenMat & 0xFFFFu;
    const uint32_t camMat = (lenMat >> 16) & 0xFFu;
    if((lvLen + camLen > ic.maxLen) || (lvLen + camLen < ic.minLen)) return mk3(0.f);
    Bsdf b;
    b.frame.x = mk3(q[0][owner], q[1][owner], q[2][owner]);
    b.frame.y = mk3(q[3][owner], q[4][owner], q[5][owner]);
    b.frame.z = mk3(q[6][owner], q[7][owner], q[8][owner]);
    b.localDirFix = mk3(q[9][owner], q[10][owner], q[11][owner]);
    const f3 localDirGen = b.frame.toLocal(xyz(dv));
    if(localDirGen.z * b.localDirFix.z < 0.f) return mk3(0.f);
    b.diffProb = q[13][owner]; b.phongProb = q[14][owner]; b.contProb = q[15][owner];
    float camDirPdfW, camRevPdfW;
    const f3 f = bsdfEvaluateLocal(sc.mats[camMat], b, localDirGen, camDirPdfW, camRevPdfW);
    if(iszero(f)) return mk3(0.f);
    camDirPdfW *= b.contProb;
    camRevPdfW *= lightContProb;                                 // the LIGHT vertex's RR probability
    const float wLight  = tv.w * ic.misVcWeightFactor + dv.w * camDirPdfW;                 // (38)
    const float wCamera = q[16][owner] * ic.misVcWeightFactor + q[17][owner] * camRevPdfW; // (39)
    const float misWeight = ic.ppm ? 1.f : 1.f / (wLight + 1.f + wCamera);
    return (misWeight * f) * xyz(tv);
}

SVCM_D void evaluatePairLoaded(const DevScene &sc, const IterConst &ic, float (*q)[32], float (*acc)[32], uint32_t owner,
                               float lenBits, float4 dv, float4 tv, float lightContProb, float weight)
{
    const uint32_t lvLen = __float_as_uint(lenBits);
    const uint32_t lenMat = __float_as_uint(q[12][owner]);
    const uint32_t camLen = lenMat & 0xFFFFu;
    const uint32_t camMat = (lenMat >> 16) & 0xFFu;
    if((lvLen + camLen > ic.maxLen) || (lvLen + camLen < ic.minLen)) return;
    Bsdf b;
    b.frame.x = mk3(q[0][owner], q[1][owner], q[2][owner]);
    b.frame.y = mk3(q[3][owner], q[4][owner], q[5][owner]);
    b.frame.z = mk3(q[6][owner], q[7][owner], q[8][owner]);
    b.localDirFix = mk3(q[9][owner], q[10][owner], q[11][owner]);
    const f3 localDirGen = b.frame.toLocal(xyz(dv));
    if(localDirGen.z * b.localDirFix.z < 0.f) return;
    b.diffProb = q[13][owner]; b.phongProb = q[14][owner]; b.contProb = q[15][owner];
    float camDirPdfW, camRevPdfW;
    const f3 f = bsdfEvaluateLocal(sc.mats[camMat], b, localDirGen, camDirPdfW, camRevPdfW);
    if(iszero(f)) return;
    camDirPdfW *= b.contProb;
    camRevPdfW *= lightContProb;                                 // the LIGHT vertex's RR probability
    const float wLight  = tv.w * ic.misVcWeightFactor + dv.w * camDirPdfW;                 // (38)
    const float wCamera = q[16][owner] * ic.misVcWeightFactor + q[17][owner] * camRevPdfW; // (39)
    const float misWeight = ic.ppm ? 1.f : 1.f / (wLight + 1.f + wCamera);
    const f3 c = ((misWeight * weight) * f) * xyz(tv);
    atomicAdd(&acc[0][owner], c.x);
    atomicAdd(&acc[1][owner], c.y);
    atomicAdd(&acc[2][owner], c.z);
}

SVCM_D void evaluatePair(const DevScene &sc, const IterConst &ic, const GridBuffers &g, float (*q)[32],
                         float (*acc)[32], uint32_t idx, uint32_t owner, float lenBits, float weight)
{
    const uint32_t lvLen = __float_as_uint(lenBits);
    const uint32_t lenMat = __float_as_uint(q[12][owner]);
    const uint32_t camLen = lenMat & 0xFFFFu;
    const uint32_t camMat = (lenMat >> 16) & 0xFFu;
    if((lvLen + camLen > ic.maxLen) || (lvLen + camLen < ic.minLen)) return;
    const float4 dv = g.sDirVm[idx];
    Bsdf b;
    b.frame.x = mk3(q[0][owner], q[1][owner], q[2][owner]);
    b.frame.y = mk3(q[3][owner], q[4][owner], q[5][owner]);
    b.frame.z = mk3(q[6][owner], q[7][owner], q[8][owner]);
    b.localDirFix = mk3(q[9][owner], q[10][owner], q[11][owner]);
    const f3 localDirGen = b.frame.toLocal(xyz(dv));
    if(localDirGen.z * b.localDirFix.z < 0.f) return;
    b.diffProb = q[13][owner]; b.phongProb = q[14][owner]; b.contProb = q[15][owner];
    float camDirPdfW, camRevPdfW;
    const f3 f = bsdfEvaluateLocal(sc.mats[camMat], b, localDirGen, camDirPdfW, camRevPdfW);
    if(iszero(f)) return;
    const float4 tv = g.sThrVcm[idx];
    camDirPdfW *= b.contProb;
    camRevPdfW *= g.sContProb[idx];                              // the LIGHT vertex's RR probability
    const float wLight  = tv.w * ic.misVcWeightFactor + dv.w * camDirPdfW;                 // (38)
    const float wCamera = q[16][owner] * ic.misVcWeightFactor + q[17][owner] * camRevPdfW; // (39)
    const float misWeight = ic.ppm ? 1.f : 1.f / (wLight + 1.f + wCamera);
    const f3 c = ((misWeight * weight) * f) * xyz(tv);   // `weight` visits of the reference add `weight` times this
    atomicAdd(&acc[0][owner], c.x);
    atomicAdd(&acc[1][owner], c.y);
    atomicAdd(&acc[2][owner], c.z);
}

// |q - p|^2 <= r^2 in the reference's exact (unfused) evaluation order (src/hashgrid.hxx:162-165):
// identical accept decisions for identical inputs
SVCM_D bool inRange(f3 qpos, float4 pl, float radiusSqr)
{
    const float dx = qpos.x - pl.x, dy = qpos.y - pl.y, dz = qpos.z - pl.z;
    const float distSqr = __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(dx, dx)), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return distSqr <= radiusSqr;
}

// ------------------------------------------------------------------------------------------
// merge, variant 0: lane per query straight from global memory (the bucket-layout kernel's walk
// on the cell layout). Kept as the attribution point "what spatial order alone buys".
// ------------------------------------------------------------------------------------------
#define CM_WARPS (SVCM_BLOCK / 32)

struct CellLaneSmem
{
    uint32_t queue[64];
    uint8_t  queueOwner[64];
    float    q[CQ_WORDS][32];
    float    acc[3][32];
};

__global__ void __launch_bounds__(SVCM_BLOCK)
k_mergeCellLane(const __grid_constant__ DevScene sc, const __grid_constant__ IterConst ic,
                const __grid_constant__ QueryStore qs, const unsigned long long *queryCountPtr,
                const uint32_t hostQueryCount, const __grid_constant__ GridBuffers g, float *accum,
                float *contribOut, int32_t *countsOut, unsigned long long *counters)
{
    __shared__ CellLaneSmem smem[CM_WARPS];
    CellLaneSmem &w = smem[threadIdx.x >> 5];
    uint32_t nq = hostQueryCount;
    if(queryCountPtr) nq = uint32_t(min(*queryCountPtr, (unsigned long long)qs.capacity));
    const uint32_t lane = laneId();
    const uint32_t ltMask = (1u << lane) - 1u;
    const GridDims d = *g.dims;
    const f3 bbMin = mk3(orderedToFloat(g.bboxBits[0]), orderedToFloat(g.bboxBits[1]), orderedToFloat(g.bboxBits[2]));
    const f3 bbMax = mk3(orderedToFloat(g.bboxBits[3]), orderedToFloat(g.bboxBits[4]), orderedToFloat(g.bboxBits[5]));
    uint32_t nCand = 0u, nAcc = 0u;

    for(uint32_t q0i = (blockIdx.x * CM_WARPS + (threadIdx.x >> 5)) * 32u; q0i < nq; q0i += gridDim.x * CM_WARPS * 32u)
    {
        bool valid = q0i + lane < nq;
        const uint32_t q = valid ? (qs.order ? qs.order[q0i + lane] : q0i + lane) : 0u;
        f3 qpos = mk3(0.f), scale = mk3(0.f);
        uint32_t pixel = 0u, cand = 0u, accepted = 0u, refCand = 0u;
        w.acc[0][lane] = 0.f; w.acc[1][lane] = 0.f; w.acc[2][lane] = 0.f;
        QueryCells qc; qc.px = qc.py = qc.pz = qc.pxo = qc.pyo = qc.pzo = 0; qc.mults = 0x11111111u;
        // the (up to) 8 storage cells: start slot and end slot of each
        uint32_t rs[8], re[8];
#pragma unroll
        for(int j = 0; j < 8; j++) { rs[j] = 0u; re[j] = 0u; }

        float4 q1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if(valid)
        {
            q1 = qs.q1[q];
            valid = (__float_as_uint(q1.w) & CAMV_MERGEABLE) != 0u;
        }
        if(valid)
        {
            const float4 q0 = qs.q0[q], q2 = qs.q2[q], q3 = qs.q3[q];
            qpos = xyz(q0); scale = xyz(q3) * ic.vmNormalization; pixel = __float_as_uint(q0.w);
            const uint32_t lenMat = __float_as_uint(q1.w);
            Bsdf bsdf;
            bsdf.frame.setFromZ(xyz(q1));
            bsdf.localDirFix = xyz(q2);
            bsdf.diffProb = bsdf.phongProb = bsdf.reflProb = bsdf.refrProb = 0.f;
            bsdfFromLocal(sc, int32_t((lenMat >> 16) & 0xFFu), bsdf);
            parkQuery(w.q, lane, bsdf, q1.w, q2.w, q3.w);
            const bool inside = !(qpos.x - bbMin.x < 0.f || bbMax.x - qpos.x < 0.f || qpos.y - bbMin.y < 0.f ||
                                  bbMax.y - qpos.y < 0.f || qpos.z - bbMin.z < 0.f || bbMax.z - qpos.z < 0.f);
            if(inside)
            {
                qc = queryCells(ic, qpos, bbMin, &refCand, g.refCount);
                const int32_t sx0 = storageX(qc.px, d.m), sx1 = storageX(qc.pxo, d.m);
                const int32_t ry0 = foldCoord(qc.py, d.fy), ry1 = foldCoord(qc.pyo, d.fy);
                const int32_t rz0 = foldCoord(qc.pz, d.fz), rz1 = foldCoord(qc.pzo, d.fz);
#pragma unroll
                for(int j = 0; j < 8; j++)
                {
                    if((j & 4) && sx1 == sx0) continue;            // both reference cells in one storage cell
                    const uint32_t k = cellKey(d, (j & 4) ? sx1 : sx0, (j & 2) ? ry1 : ry0, (j & 1) ? rz1 : rz0);
                    rs[j] = g.cellEnd[k - 1u]; re[j] = g.cellEnd[k];
                }
            }
        }
        __syncwarp();

        uint32_t j = 0u, cur = rs[0], end = re[0];
        uint32_t head = 0u, tail = 0u;
        for(;;)
        {
            while(cur == end && j < 7u)
            {
                j++;
                // (select instead of dynamic indexing keeps the ranges in registers)
                cur = j == 1u ? rs[1] : j == 2u ? rs[2] : j == 3u ? rs[3] : j == 4u ? rs[4] : j == 5u ? rs[5] : j == 6u ? rs[6] : rs[7];
                end = j == 1u ? re[1] : j == 2u ? re[2] : j == 3u ? re[3] : j == 4u ? re[4] : j == 5u ? re[5] : j == 6u ? re[6] : re[7];
            }
            const bool has = cur < end;
            if(!__any_sync(FULL_MASK, has)) break;
            uint32_t weight = 0u, idx = 0u;
            if(has)
            {
                idx = cur++;
                cand++;
                const float4 pl = g.sPosLen[idx];
                if(inRange(qpos, pl, ic.radiusSqr)) weight = acceptWeight(qc, ic, bbMin, pl);
            }
            accepted += weight;
            // a bucket the reference visits n times contributes its particles n times
            while(__any_sync(FULL_MASK, weight != 0u))
            {
                const bool hitIt = weight != 0u;
                const uint32_t m = __ballot_sync(FULL_MASK, hitIt);
                if(hitIt)
                {
                    const uint32_t slot = (tail + uint32_t(__popc(m & ltMask))) & 63u;
                    w.queue[slot] = idx; w.queueOwner[slot] = uint8_t(lane);
                    weight--;
                }
                tail += uint32_t(__popc(m));
                __syncwarp();
                if(tail - head >= 32u)
                {
                    {
                        const uint32_t pi = w.queue[(head + lane) & 63u];
                        evaluatePair(sc, ic, g, w.q, w.acc, pi, w.queueOwner[(head + lane) & 63u],
                                     reinterpret_cast<const float*>(g.sPosLen + pi)[3], 1.f);
                    }
                    __syncwarp();
                    head += 32u;
                }
            }
        }
        if(tail != head)
        {
            if(head + lane < tail)
            {
                const uint32_t pi = w.queue[(head + lane) & 63u];
                evaluatePair(sc, ic, g, w.q, w.acc, pi, w.queueOwner[(head + lane) & 63u],
                             reinterpret_cast<const float*>(g.sPosLen + pi)[3], 1.f);
            }
        }
        __syncwarp();

        if(valid)
        {
            const f3 contrib = mk3(w.acc[0][lane], w.acc[1][lane], w.acc[2][lane]);
            if(g.refCount) cand = refCand;
            nCand += cand; nAcc += accepted;
            if(contribOut)
            {
                contribOut[3ull * q + 0] = contrib.x; contribOut[3ull * q + 1] = contrib.y;
                contribOut[3ull * q + 2] = contrib.z;
            }
            if(countsOut) { countsOut[2ull * q] = int32_t(cand); countsOut[2ull * q + 1] = int32_t(accepted); }
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
// merge, variants 1 / 2: neighbourhood staged through shared memory
// ------------------------------------------------------------------------------------------
SVCM_D uint32_t smemAddr(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
SVCM_D void mbarInit(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(bar)), "r"(count) : "memory");
}
SVCM_D void mbarInitFence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SVCM_D void mbarArriveExpectTx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemAddr(bar)), "r"(bytes) : "memory");
}
SVCM_D bool mbarTryWait(unsigned long long *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smemAddr(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
SVCM_D void mbarWait(unsigned long long *bar, uint32_t parity)
{
    if(mbarTryWait(bar, parity)) return;
    const unsigned long long t0 = globalTimerNs();
    while(!mbarTryWait(bar, parity)) if(globalTimerNs() - t0 > SVCM_SPIN_LIMIT_NS) __trap();
}
// global -> shared bulk asynchronous copy (TMA engine, 1-D), completion counted on `bar`
SVCM_D void bulkLoad(void *dstSmem, const void *srcGlobal, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smemAddr(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

#define ST_WARPS 8          // warps per block; every warp is an independent pipeline
#define ST_XSPAN 6          // storage cells along x one sub-chunk may span
#define ST_RUNS 9           // 3 x 3 rows around the sub-chunk's row

// the staged kernel parks five more words per query: its cell (px, py, pz), the side of the second
// cell on every axis, and the bucket multiplicities — the accept weight of a pair is worked out when
// the pair is evaluated (all lanes busy), not inside the candidate walk
#define SQ_WORDS (CQ_WORDS + 5)

template<int PCAP, int LCAP>
struct alignas(128) StagedWarpSmem
{
    float4   stage[2][PCAP];            // staged positions | path length
    uint32_t list[LCAP][32];            // in-range sorted slots, [entry][lane]
    float    q[SQ_WORDS][32];
    float    acc[3][32];
    uint32_t nacc[32];                  // accepted pairs per owner, counted like the reference does
    uint32_t pref[32];
    unsigned long long bar[2];
};

// evaluates every listed pair of the warp, 32 per step; empties the lists. A listed pair is in
// range; how often the reference counts it (the bucket multiplicity of the particle's own cell, 0 if
// that cell is not one of the query's 8) is worked out here from the parked cell words.
template<int PCAP, int LCAP>
SVCM_D void flushLists(const DevScene &sc, const IterConst &ic, const GridBuffers &g, StagedWarpSmem<PCAP, LCAP> &w,
                       f3 bbMin, uint32_t &cnt)
{
    const uint32_t lane = laneId();
    uint32_t inc = cnt;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(FULL_MASK, inc, o);
        if(lane >= uint32_t(o)) inc += t;
    }
    const uint32_t total = __shfl_sync(FULL_MASK, inc, 31);
    if(total == 0u) return;
    w.pref[lane] = inc - cnt;
    __syncwarp();
    for(uint32_t base = 0u; base < total; base += 32u)
    {
        const uint32_t p = base + lane;
        if(p < total)
        {
            uint32_t o = 0u;                      // the lane whose list holds pair p: pref[o] <= p < pref[o + 1]
#pragma unroll
            for(int s = 16; s > 0; s >>= 1) if(w.pref[o + uint32_t(s)] <= p) o += uint32_t(s);
            const uint32_t idx = w.list[p - w.pref[o]][o];
            const float4 pl = g.sPosLen[idx];
            QueryCells c;
            c.px = __float_as_int(w.q[CQ_WORDS + 0][o]); c.py = __float_as_int(w.q[CQ_WORDS + 1][o]);
            c.pz = __float_as_int(w.q[CQ_WORDS + 2][o]);
            const uint32_t side = __float_as_uint(w.q[CQ_WORDS + 3][o]);
            c.pxo = c.px + ((side & 4u) ? 1 : -1); c.pyo = c.py + ((side & 2u) ? 1 : -1); c.pzo = c.pz + ((side & 1u) ? 1 : -1);
            c.mults = __float_as_uint(w.q[CQ_WORDS + 4][o]);
            const uint32_t weight = acceptWeight(c, ic, bbMin, pl);
            if(weight)
            {
                atomicAdd(&w.nacc[o], weight);
                evaluatePair(sc, ic, g, w.q, w.acc, idx, o, pl.w, float(weight));
            }
        }
    }
    cnt = 0u;
    __syncwarp();
}

template<int PCAP, int LCAP, bool kBulk>
__global__ void __launch_bounds__(ST_WARPS * 32)
k_mergeStaged(const __grid_constant__ DevScene sc, const __grid_constant__ IterConst ic,
              const __grid_constant__ QueryStore qs, const unsigned long long *queryCountPtr,
              const uint32_t hostQueryCount, const __grid_constant__ GridBuffers g, float *accum,
              float *contribOut, int32_t *countsOut, unsigned long long *counters)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    typedef StagedWarpSmem<PCAP, LCAP> WarpSmem;
    WarpSmem &w = reinterpret_cast<WarpSmem*>(smemRaw)[threadIdx.x >> 5];

    uint32_t nq = hostQueryCount;
    if(queryCountPtr) nq = uint32_t(min(*queryCountPtr, (unsigned long long)qs.capacity));
    const uint32_t lane = laneId();
    const GridDims d = *g.dims;
    const f3 bbMin = mk3(orderedToFloat(g.bboxBits[0]), orderedToFloat(g.bboxBits[1]), orderedToFloat(g.bboxBits[2]));
    const f3 bbMax = mk3(orderedToFloat(g.bboxBits[3]), orderedToFloat(g.bboxBits[4]), orderedToFloat(g.bboxBits[5]));
    uint32_t nCand = 0u, nAcc = 0u;

    uint32_t parity0 = 0u, parity1 = 0u;          // phase each stage buffer's barrier completes next
    if(kBulk)
    {
        if(lane == 0u) { mbarInit(&w.bar[0], 1u); mbarInit(&w.bar[1], 1u); mbarInitFence(); }
        __syncwarp();
    }

    const uint32_t warpGlobal = blockIdx.x * ST_WARPS + (threadIdx.x >> 5);
    for(uint32_t q0i = warpGlobal * 32u; q0i < nq; q0i += gridDim.x * ST_WARPS * 32u)
    {
        bool valid = q0i + lane < nq;
        const uint32_t q = valid ? (qs.order ? qs.order[q0i + lane] : q0i + lane) : 0u;
        f3 qpos = mk3(0.f), scale = mk3(0.f);
        uint32_t pixel = 0u, cand = 0u, refCand = 0u;
        w.acc[0][lane] = 0.f; w.acc[1][lane] = 0.f; w.acc[2][lane] = 0.f; w.nacc[lane] = 0u;
        QueryCells qc; qc.px = qc.py = qc.pz = qc.pxo = qc.pyo = qc.pzo = 0; qc.mults = 0x11111111u;
        // own storage cell: x, folded row; the x interval and the second row of each axis
        int32_t sxq = 0, sxlo = 0, sxhi = 0, ry0 = 0, ry1 = 0, rz0 = 0, rz1 = 0;
        int32_t dyo = 0, dzo = 0;                 // second row relative to the own row: -1 / +1
        bool pending = false;

        float4 q1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if(valid)
        {
            q1 = qs.q1[q];
            valid = (__float_as_uint(q1.w) & CAMV_MERGEABLE) != 0u;
        }
        if(valid)
        {
            const float4 q0 = qs.q0[q], q2 = qs.q2[q], q3 = qs.q3[q];
            qpos = xyz(q0); scale = xyz(q3) * ic.vmNormalization; pixel = __float_as_uint(q0.w);
            const uint32_t lenMat = __float_as_uint(q1.w);
            Bsdf bsdf;                             // BSDF<false>::Setup from normal / local direction / material
            bsdf.frame.setFromZ(xyz(q1));
            bsdf.localDirFix = xyz(q2);
            bsdf.diffProb = bsdf.phongProb = bsdf.reflProb = bsdf.refrProb = 0.f;
            bsdfFromLocal(sc, int32_t((lenMat >> 16) & 0xFFu), bsdf);
            parkQuery(w.q, lane, bsdf, q1.w, q2.w, q3.w);
            // HashGrid::Process: bbox early-out (src/hashgrid.hxx:116-122)
            const bool inside = !(qpos.x - bbMin.x < 0.f || bbMax.x - qpos.x < 0.f || qpos.y - bbMin.y < 0.f ||
                                  bbMax.y - qpos.y < 0.f || qpos.z - bbMin.z < 0.f || bbMax.z - qpos.z < 0.f);
            if(inside)
            {
                qc = queryCells(ic, qpos, bbMin, &refCand, g.refCount);
                sxq = storageX(qc.px, d.m);
                const int32_t sxo = storageX(qc.pxo, d.m);
                sxlo = min(sxq, sxo); sxhi = max(sxq, sxo);
                ry0 = foldCoord(qc.py, d.fy); ry1 = foldCoord(qc.pyo, d.fy);
                rz0 = foldCoord(qc.pz, d.fz); rz1 = foldCoord(qc.pzo, d.fz);
                dyo = qc.pyo - qc.py; dzo = qc.pzo - qc.pz;
                w.q[CQ_WORDS + 0][lane] = __int_as_float(qc.px); w.q[CQ_WORDS + 1][lane] = __int_as_float(qc.py);
                w.q[CQ_WORDS + 2][lane] = __int_as_float(qc.pz);
                w.q[CQ_WORDS + 3][lane] = __uint_as_float((qc.pxo > qc.px ? 4u : 0u) | (qc.pyo > qc.py ? 2u : 0u) |
                                                          (qc.pzo > qc.pz ? 1u : 0u));
                w.q[CQ_WORDS + 4][lane] = __uint_as_float(qc.mults);
                pending = true;
            }
        }
        const int32_t rowId = rz0 * d.fy + ry0;
        uint32_t cnt = 0u;                        // entries in this lane's list
        __syncwarp();

        // ---- sub-chunks: queries of one row whose cells lie within ST_XSPAN of the first ----
        for(;;)
        {
            const uint32_t pm = __ballot_sync(FULL_MASK, pending);
            if(pm == 0u) break;
            const int leader = __ffs(pm) - 1;
            const int32_t lrow = __shfl_sync(FULL_MASK, rowId, leader);
            const int32_t lsx = __shfl_sync(FULL_MASK, sxq, leader);
            const int32_t lry = __shfl_sync(FULL_MASK, ry0, leader), lrz = __shfl_sync(FULL_MASK, rz0, leader);
            const bool inSub = pending && rowId == lrow && uint32_t(sxq - lsx) <= uint32_t(ST_XSPAN);
            const int32_t xlo = __reduce_min_sync(FULL_MASK, inSub ? sxlo : 0x7FFFFFFF);
            const int32_t xhi = __reduce_max_sync(FULL_MASK, inSub ? sxhi : -0x7FFFFFFF);

            // the 9 runs: row (lry + dy, lrz + dz), storage cells xlo .. xhi; lanes 0..8 hold one each
            uint32_t runStart = 0u, runLen = 0u;
            if(lane < uint32_t(ST_RUNS))
            {
                const int32_t dy = int32_t(lane) / 3 - 1, dz = int32_t(lane) % 3 - 1;
                int32_t ry = lry + dy, rz = lrz + dz;
                ry = ry < 0 ? ry + d.fy : (ry >= d.fy ? ry - d.fy : ry);
                rz = rz < 0 ? rz + d.fz : (rz >= d.fz ? rz - d.fz : rz);
                const uint32_t kLo = cellKey(d, xlo, ry, rz), kHi = cellKey(d, xhi, ry, rz);
                runStart = g.cellEnd[kLo - 1u];
                runLen = g.cellEnd[kHi] - runStart;
            }
            // virtual concatenation of the runs: P = start of a run in it, L = its length
            uint32_t incl = runLen;
#pragma unroll
            for(int o = 1; o < 16; o <<= 1)
            {
                const uint32_t t = __shfl_up_sync(FULL_MASK, incl, o);
                if(lane >= uint32_t(o)) incl += t;
            }
            const uint32_t runP = incl - runLen;
            const uint32_t L = __shfl_sync(FULL_MASK, incl, ST_RUNS - 1);

            // this lane's (up to) 4 ranges in virtual coordinates; slot = v + delta
            uint32_t vs[4], ve[4], dl[4];
#pragma unroll
            for(int r = 0; r < 4; r++)
            {
                const int32_t dy = (r & 2) ? dyo : 0, dz = (r & 1) ? dzo : 0;
                const int run = (dy + 1) * 3 + (dz + 1);
                const uint32_t rP = __shfl_sync(FULL_MASK, runP, run), rS = __shfl_sync(FULL_MASK, runStart, run);
                vs[r] = 0u; ve[r] = 0u; dl[r] = 0u;
                if(inSub)
                {
                    const int32_t ry = (r & 2) ? ry1 : ry0, rz = (r & 1) ? rz1 : rz0;
                    const uint32_t kLo = cellKey(d, sxlo, ry, rz), kHi = cellKey(d, sxhi, ry, rz);
                    const uint32_t gs = g.cellEnd[kLo - 1u], ge = g.cellEnd[kHi];
                    vs[r] = rP + (gs - rS); ve[r] = rP + (ge - rS); dl[r] = gs - vs[r];
                    cand += ge - gs;
                }
            }

            const uint32_t numStages = (L + uint32_t(PCAP) - 1u) / uint32_t(PCAP);
            // copies of stage k into buffer k & 1
            auto issue = [&](uint32_t k)
            {
                const uint32_t b = k & 1u;
                const uint32_t w0 = k * uint32_t(PCAP), w1 = min(L, w0 + uint32_t(PCAP));
                if(kBulk)
                {
                    if(lane == 0u) mbarArriveExpectTx(&w.bar[b], (w1 - w0) * 16u);
                    __syncwarp();
                    if(lane < uint32_t(ST_RUNS))
                    {
                        const uint32_t lo = max(runP, w0), hi = min(runP + runLen, w1);
                        if(lo < hi) bulkLoad(&w.stage[b][lo - w0], g.sPosLen + (runStart + (lo - runP)), (hi - lo) * 16u, &w.bar[b]);
                    }
                }
                else
                {
#pragma unroll 1
                    for(int run = 0; run < ST_RUNS; run++)
                    {
                        const uint32_t rP = __shfl_sync(FULL_MASK, runP, run), rL = __shfl_sync(FULL_MASK, runLen, run),
                                       rS = __shfl_sync(FULL_MASK, runStart, run);
                        const uint32_t lo = max(rP, w0), hi = min(rP + rL, w1);
                        for(uint32_t v = lo + lane; v < hi; v += 32u) w.stage[b][v - w0] = g.sPosLen[rS + (v - rP)];
                    }
                }
            };

            if(numStages > 0u) issue(0u);
            for(uint32_t k = 0u; k < numStages; k++)
            {
                const uint32_t b = k & 1u;
                // the other buffer was consumed in the previous round (all lanes passed the __syncwarp below)
                if(k + 1u < numStages) issue(k + 1u);
                if(kBulk)
                {
                    mbarWait(&w.bar[b], b ? parity1 : parity0);
                    if(b) parity1 ^= 1u; else parity0 ^= 1u;
                }
                else __syncwarp();
                const uint32_t w0 = k * uint32_t(PCAP), w1 = min(L, w0 + uint32_t(PCAP));
                const float4 *stage = w.stage[b];

                // walk this lane's ranges inside the window: the loop body is the distance test and
                // nothing else — an in-range slot goes to the private list, a full list anywhere in
                // the warp flushes all of them (evaluation with every lane busy)
                uint32_t r = 0u, v = max(vs[0], w0), vend = min(ve[0], w1), delta = dl[0];
                for(;;)
                {
                    bool exhausted = false;                            // else: stopped at a full list
                    for(;;)
                    {
                        while(v >= vend && r < 3u)
                        {
                            r++;
                            const uint32_t s = r == 1u ? vs[1] : r == 2u ? vs[2] : vs[3];
                            const uint32_t e = r == 1u ? ve[1] : r == 2u ? ve[2] : ve[3];
                            delta = r == 1u ? dl[1] : r == 2u ? dl[2] : dl[3];
                            v = max(s, w0); vend = min(e, w1);
                        }
                        if(v >= vend) { exhausted = true; break; }      // nothing left in this window
                        if(cnt == uint32_t(LCAP)) break;
                        if(inRange(qpos, stage[v - w0], ic.radiusSqr)) w.list[cnt++][lane] = v + delta;
                        v++;
                    }
                    if(!__any_sync(FULL_MASK, !exhausted)) break;
                    flushLists<PCAP, LCAP>(sc, ic, g, w, bbMin, cnt);
                }
                __syncwarp();
            }
            pending = pending && !inSub;
        }

        flushLists<PCAP, LCAP>(sc, ic, g, w, bbMin, cnt);

        if(valid)
        {
            const f3 contrib = mk3(w.acc[0][lane], w.acc[1][lane], w.acc[2][lane]);
            const uint32_t accepted = w.nacc[lane];
            if(g.refCount) cand = refCand;
            nCand += cand; nAcc += accepted;
            if(contribOut)
            {
                contribOut[3ull * q + 0] = contrib.x; contribOut[3ull * q + 1] = contrib.y;
                contribOut[3ull * q + 2] = contrib.z;
            }
            if(countsOut) { countsOut[2ull * q] = int32_t(cand); countsOut[2ull * q + 1] = int32_t(accepted); }
            // color += cameraState.mThroughput * mVmNormalization * query.GetContrib()  (src/vertexcm.hxx:534)
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
// merge over the reference's BUCKET layout with the buckets of a warp's 32 queries staged into shared
// memory (HashGrid::Process src/hashgrid.hxx:110-169 + RangeQuery::Process src/vertexcm.hxx:130-169).
//
// A warp takes 32 consecutive queries of the bucket-sorted query order — queries of one or two
// cells. Their 8 x 32 bucket ids are de-duplicated through a small per-warp hash set in shared
// memory (neighbouring queries share most of their neighbour cells); the lane that inserts an id
// fetches that bucket's range once and copies its positions into the warp's stage with ONE bulk
// asynchronous copy (cp.async.bulk, completion on an mbarrier). Every lane then walks its 8 ranges
// — in the reference's order, a bucket shared by two of its cells twice, exactly like the
// reference — from shared memory; a bucket that does not fit the stage is walked from global
// memory instead (dense grids: the stage is an accelerator, never a limit). In-range slots go to
// per-lane lists and are evaluated flattened, 32 pairs per warp step.
// ------------------------------------------------------------------------------------------
#define BS_WARPS 4          // warps per block: small blocks pack the shared memory of an SM better
#define BS_GLOBAL 0xFFFFFFFFu

template<int PCAP, int LCAP, int TSLOTS>
struct alignas(128) BucketWarpSmem
{
    float4   stage[PCAP];               // staged positions | path length
    uint32_t list[LCAP][32];            // in-range sorted slots, [entry][lane]
    float    q[CQ_WORDS][32];
    float    acc[3][32];
    uint32_t pref[32];
    uint32_t slotKey[TSLOTS];           // bucket id + 1 (0: empty)
    uint32_t slotStart[TSLOTS];         // first sorted slot of the bucket
    uint32_t slotCount[TSLOTS];         // particles in it
    uint32_t slotOff[TSLOTS];           // offset in `stage`, or BS_GLOBAL: walk it from global memory
    uint32_t stagedTotal;               // particles reserved in `stage` for this chunk
    unsigned long long bar;
};

// Evaluates every listed pair of the warp, 32 per step, and empties the lists. Pairs are taken in
// owner-major order, so the lanes of one step hold runs of equal owners: their contributions are
// summed with a segmented shuffle scan and the last lane of a run adds the run's sum to its owner
// (shared-memory float atomics compile to compare-and-swap loops and were 11 % of this kernel's
// instructions).
template<int PCAP, int LCAP, int TSLOTS>
SVCM_D void flushBucketLists(const DevScene &sc, const IterConst &ic, const GridBuffers &g,
                             BucketWarpSmem<PCAP, LCAP, TSLOTS> &w, uint32_t &cnt)
{
    const uint32_t lane = laneId();
    uint32_t inc = cnt;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(FULL_MASK, inc, o);
        if(lane >= uint32_t(o)) inc += t;
    }
    const uint32_t total = __shfl_sync(FULL_MASK, inc, 31);
    if(total == 0u) return;
    w.pref[lane] = inc - cnt;
    __syncwarp();
    for(uint32_t base = 0u; base < total; base += 32u)
    {
        const uint32_t p = base + lane;
        uint32_t o = 0xFFFFFFFFu;                 // owner; lanes past the end form a run of their own
        f3 c = mk3(0.f);
        if(p < total)
        {
            o = 0u;                               // the lane whose list holds pair p: pref[o] <= p < pref[o + 1]
#pragma unroll
            for(int s = 16; s > 0; s >>= 1) if(w.pref[o + uint32_t(s)] <= p) o += uint32_t(s);
            const uint32_t idx = w.list[p - w.pref[o]][o];
            const float lenBits = reinterpret_cast<const float*>(g.sPosLen + idx)[3];
            const float4 dv = g.sDirVm[idx], tv = g.sThrVcm[idx];
            const float cp = g.sContProb[idx];
            c = pairContribution(sc, ic, w.q, o, lenBits, dv, tv, cp);
        }
#pragma unroll
        for(int d = 1; d < 32; d <<= 1)
        {
            const uint32_t ou = __shfl_up_sync(FULL_MASK, o, d);
            const float cx = __shfl_up_sync(FULL_MASK, c.x, d), cy = __shfl_up_sync(FULL_MASK, c.y, d),
                        cz = __shfl_up_sync(FULL_MASK, c.z, d);
            if(lane >= uint32_t(d) && ou == o) { c.x += cx; c.y += cy; c.z += cz; }
        }
        const uint32_t on = __shfl_down_sync(FULL_MASK, o, 1);
        if(p < total && (lane == 31u || on != o))
        {
            w.acc[0][o] += c.x; w.acc[1][o] += c.y; w.acc[2][o] += c.z;
        }
        __syncwarp();
    }
    cnt = 0u;
    __syncwarp();
}

template<int PCAP, int LCAP, int TSLOTS, bool kBulk>
__global__ void __launch_bounds__(BS_WARPS * 32)
k_mergeBucketStaged(const __grid_constant__ DevScene sc, const __grid_constant__ IterConst ic,
                    const __grid_constant__ QueryStore qs, const unsigned long long *queryCountPtr,
                    const uint32_t hostQueryCount, const __grid_constant__ GridBuffers g, float *accum,
                    float *contribOut, int32_t *countsOut, unsigned long long *counters)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    typedef BucketWarpSmem<PCAP, LCAP, TSLOTS> WarpSmem;
    WarpSmem &w = reinterpret_cast<WarpSmem*>(smemRaw)[threadIdx.x >> 5];

    uint32_t nq = hostQueryCount;
    if(queryCountPtr) nq = uint32_t(min(*queryCountPtr, (unsigned long long)qs.capacity));
    const uint32_t lane = laneId();
    const f3 bbMin = mk3(orderedToFloat(g.bboxBits[0]), orderedToFloat(g.bboxBits[1]), orderedToFloat(g.bboxBits[2]));
    const f3 bbMax = mk3(orderedToFloat(g.bboxBits[3]), orderedToFloat(g.bboxBits[4]), orderedToFloat(g.bboxBits[5]));
    uint32_t nCand = 0u, nAcc = 0u;
    uint32_t parity = 0u;
    if(kBulk)
    {
        if(lane == 0u) { mbarInit(&w.bar, 1u); mbarInitFence(); }
        __syncwarp();
    }

    const uint32_t warpGlobal = blockIdx.x * BS_WARPS + (threadIdx.x >> 5);
    for(uint32_t q0i = warpGlobal * 32u; q0i < nq; q0i += gridDim.x * BS_WARPS * 32u)
    {
        bool valid = q0i + lane < nq;
        const uint32_t q = valid ? (qs.order ? qs.order[q0i + lane] : q0i + lane) : 0u;
        f3 qpos = mk3(0.f), scale = mk3(0.f);
        uint32_t pixel = 0u, cand = 0u, accepted = 0u;
        w.acc[0][lane] = 0.f; w.acc[1][lane] = 0.f; w.acc[2][lane] = 0.f;
        // the table is emptied by everybody, the stage starts empty
#pragma unroll
        for(int k = 0; k < TSLOTS / 32; k++) w.slotKey[k * 32 + int(lane)] = 0u;
        if(lane == 0u) w.stagedTotal = 0u;
        bool inside = false;
        uint32_t h[8];
#pragma unroll
        for(int j = 0; j < 8; j++) h[j] = 0u;

        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;
        if(valid)
        {
            // all four words at once (one memory round trip); a vertex recorded for connection only is skipped
            q0 = qs.q0[q]; q1 = qs.q1[q]; q2 = qs.q2[q]; q3 = qs.q3[q];
            valid = (__float_as_uint(q1.w) & CAMV_MERGEABLE) != 0u;
        }
        if(valid)
        {
            qpos = xyz(q0); scale = xyz(q3) * ic.vmNormalization; pixel = __float_as_uint(q0.w);
            const uint32_t lenMat = __float_as_uint(q1.w);
            Bsdf bsdf;                             // BSDF<false>::Setup from normal / local direction / material
            bsdf.frame.setFromZ(xyz(q1));
            bsdf.localDirFix = xyz(q2);
            bsdf.diffProb = bsdf.phongProb = bsdf.reflProb = bsdf.refrProb = 0.f;
            bsdfFromLocal(sc, int32_t((lenMat >> 16) & 0xFFu), bsdf);
            parkQuery(w.q, lane, bsdf, q1.w, q2.w, q3.w);
            // HashGrid::Process: bbox early-out, then the 8 cells (src/hashgrid.hxx:116-155)
            inside = !(qpos.x - bbMin.x < 0.f || bbMax.x - qpos.x < 0.f || qpos.y - bbMin.y < 0.f ||
                       bbMax.y - qpos.y < 0.f || qpos.z - bbMin.z < 0.f || bbMax.z - qpos.z < 0.f);
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
                    h[j] = refHash((j & 4) ? pxo : px, (j & 2) ? pyo : py, (j & 1) ? pzo : pz, ic);
            }
        }
        __syncwarp();

        // ---- de-duplicate the warp's bucket ids; the inserting lane owns the bucket ----------------
        uint32_t slots = 0u, slotsHi = 0u;        // 8 x 8-bit table positions of this lane's buckets
        uint32_t owned = 0u;                      // bit j: this lane inserted bucket j (fetches and stages it)
        if(inside)
        {
#pragma unroll
            for(int j = 0; j < 8; j++)
            {
                const uint32_t key = h[j] + 1u;
                uint32_t s = (h[j] * 2654435761u) >> (32 - 8);        // TSLOTS == 256
                s &= uint32_t(TSLOTS - 1);
                for(;;)
                {
                    const uint32_t old = atomicCAS(&w.slotKey[s], 0u, key);
                    if(old == 0u) { owned |= 1u << j; break; }
                    if(old == key) break;
                    s = (s + 1u) & uint32_t(TSLOTS - 1);
                }
                if(j < 4) slots |= s << (8 * j); else slotsHi |= s << (8 * (j - 4));
            }
            // the owner fetches the range (GetCellRange, src/hashgrid.hxx:173-177) and reserves stage room
#pragma unroll
            for(int j = 0; j < 8; j++)
            {
                if(!(owned & (1u << j))) continue;
                const uint32_t s = ((j < 4 ? slots >> (8 * j) : slotsHi >> (8 * (j - 4))) & 0xFFu);
                const uint32_t start = (h[j] == 0u) ? 0u : g.cellCount[h[j] - 1u];
                const uint32_t count = g.cellCount[h[j]] - start;
                uint32_t off = BS_GLOBAL;
                if(count > 0u && count <= uint32_t(PCAP))
                {
                    off = atomicAdd(&w.stagedTotal, count);
                    if(off + count > uint32_t(PCAP)) { atomicSub(&w.stagedTotal, count); off = BS_GLOBAL; }
                }
                w.slotStart[s] = start; w.slotCount[s] = count; w.slotOff[s] = off;
            }
        }
        __syncwarp();
        // (a reservation that was taken back may have pushed a later one over the limit spuriously;
        // that bucket is then walked from global memory — slower, never wrong)

        // ---- stage: one bulk copy per owned bucket that fits -------------------------------------
        const uint32_t stagedTotal = w.stagedTotal;
        if(kBulk)
        {
            if(stagedTotal > 0u)
            {
                if(lane == 0u) mbarArriveExpectTx(&w.bar, stagedTotal * 16u);
                __syncwarp();
#pragma unroll
                for(int j = 0; j < 8; j++)
                {
                    if(!(owned & (1u << j))) continue;
                    const uint32_t s = ((j < 4 ? slots >> (8 * j) : slotsHi >> (8 * (j - 4))) & 0xFFu);
                    const uint32_t off = w.slotOff[s];
                    if(off != BS_GLOBAL) bulkLoad(&w.stage[off], g.sPosLen + w.slotStart[s], w.slotCount[s] * 16u, &w.bar);
                }
                mbarWait(&w.bar, parity);
                parity ^= 1u;
            }
        }
        else
        {
#pragma unroll
            for(int j = 0; j < 8; j++)
            {
                // plain-load variant: every lane copies the buckets it owns itself
                if(!(owned & (1u << j))) continue;
                const uint32_t s = ((j < 4 ? slots >> (8 * j) : slotsHi >> (8 * (j - 4))) & 0xFFu);
                const uint32_t off = w.slotOff[s];
                if(off != BS_GLOBAL)
                    for(uint32_t i = 0u; i < w.slotCount[s]; i++) w.stage[off + i] = g.sPosLen[w.slotStart[s] + i];
            }
            __syncwarp();
        }

        // ---- walk the 8 ranges in the reference's order, the whole warp in step ------------------
        // Range j of every lane together, then candidate i of every lane together: the loop
        // bookkeeping runs once per step for 32 lanes (a lane-private walk spent 40 % of the kernel's
        // instructions hopping between its short ranges with 4 of 32 lanes active).
        uint32_t cnt = 0u;                        // entries in this lane's list
#pragma unroll 1
        for(uint32_t j = 0u; j < 8u; j++)
        {
            uint32_t n = 0u, start = 0u;
            const float4 *base = g.sPosLen;
            if(inside)
            {
                const uint32_t s = ((j < 4u ? slots >> (8u * j) : slotsHi >> (8u * (j - 4u))) & 0xFFu);
                n = w.slotCount[s]; start = w.slotStart[s];
                const uint32_t off = w.slotOff[s];
                base = (off != BS_GLOBAL) ? &w.stage[off] : g.sPosLen + start;    // generic pointer: either space
                cand += n;
            }
            const uint32_t nmax = __reduce_max_sync(FULL_MASK, n);
            for(uint32_t i = 0u; i < nmax; i++)
            {
                if(i < n && inRange(qpos, base[i], ic.radiusSqr)) { w.list[cnt++][lane] = start + i; accepted++; }
                if(__any_sync(FULL_MASK, cnt == uint32_t(LCAP))) flushBucketLists<PCAP, LCAP, TSLOTS>(sc, ic, g, w, cnt);
            }
        }
        flushBucketLists<PCAP, LCAP, TSLOTS>(sc, ic, g, w, cnt);

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
            // color += cameraState.mThroughput * mVmNormalization * query.GetContrib()  (src/vertexcm.hxx:534)
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

template<int PCAP, int LCAP, bool kBulk>
static void launchBucketStagedT(const DevScene &sc, const IterConst &ic, const QueryStore &queries,
                                const unsigned long long *queryCount, uint32_t hostQueryCount, const GridBuffers &grid,
                                float *accum, float *contribOut, int32_t *countsOut, unsigned long long *counters,
                                uint64_t maxQueries, cudaStream_t stream)
{
    typedef BucketWarpSmem<PCAP, LCAP, 256> WarpSmem;
    const size_t smem = sizeof(WarpSmem) * BS_WARPS;
    static bool configured = false;
    if(!configured)
    {
        cudaFuncSetAttribute(k_mergeBucketStaged<PCAP, LCAP, 256, kBulk>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        configured = true;
    }
    int perSM = int((227u * 1024u) / (smem + 1024u));
    if(perSM < 1) perSM = 1;
    if(perSM > 8) perSM = 8;
    k_mergeBucketStaged<PCAP, LCAP, 256, kBulk><<<gridFor(maxQueries, BS_WARPS * 32, perSM), BS_WARPS * 32, smem, stream>>>(
        sc, ic, queries, queryCount, hostQueryCount, grid, accum, contribOut, countsOut, counters);
}

void launchBucketStagedMerge(const DevScene &sc, const IterConst &ic, const QueryStore &queries,
                             const unsigned long long *queryCount, uint32_t hostQueryCount, const GridBuffers &grid,
                             float *accum, float *contribOut, int32_t *countsOut, unsigned long long *counters,
                             int variant, cudaStream_t stream)
{
    const uint64_t maxQueries = queryCount ? queries.capacity : hostQueryCount;
    if(maxQueries == 0) return;
    if(variant == 3)
        launchBucketStagedT<512, 16, false>(sc, ic, queries, queryCount, hostQueryCount, grid, accum, contribOut, countsOut,
                                            counters, maxQueries, stream);
    else if(variant == 4)
        launchBucketStagedT<128, 8, true>(sc, ic, queries, queryCount, hostQueryCount, grid, accum, contribOut, countsOut,
                                          counters, maxQueries, stream);
    else if(variant == 5)
        launchBucketStagedT<256, 16, true>(sc, ic, queries, queryCount, hostQueryCount, grid, accum, contribOut, countsOut,
                                           counters, maxQueries, stream);
    else
        launchBucketStagedT<512, 16, true>(sc, ic, queries, queryCount, hostQueryCount, grid, accum, contribOut, countsOut,
                                           counters, maxQueries, stream);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
uint64_t hashMagicFor(uint32_t numCells)
{
    if(numCells <= 1u) return 0ull;
    return 0xFFFFFFFFFFFFFFFFull / uint64_t(numCells) + 1ull;
}

void launchLookbackScan(uint32_t *data, const GridDims *dims, uint32_t n, unsigned long long *state,
                        uint32_t *counter, ScanGen &gen, cudaStream_t stream)
{
    const uint32_t tiles = (n + LB_TILE - 1) / LB_TILE;
    gen.value = (gen.value + 1u) & 0x3FFFFFFFu;
    if(gen.value == 0u) gen.value = 1u;          // generation 0 is what freshly zeroed state words carry
    k_scanLookback<<<tiles, SVCM_BLOCK, 0, stream>>>(data, dims, n, state, counter, tiles, gen.value);
}

int launchCellGridBuild(const IterConst &ic, const MergeSet &set, const GridBuffers &grid, ScanGen &gen,
                        bool withRefCount, cudaStream_t stream)
{
    int launches = 0;
    const uint32_t n = grid.tcap + 2u;
    // all-gathered blocks: the bbox is the union of the block headers' (launchGatherUnpackHeaders)
    const bool bboxReady = set.singleCount == NULL;
    if(!bboxReady) cudaMemcpyAsync(grid.bboxBits, grid.bboxInit, sizeof(int32_t) * 6, cudaMemcpyDeviceToDevice, stream);
    cudaMemsetAsync(grid.cellEnd, 0, sizeof(uint32_t) * n, stream);
    if(withRefCount && grid.refCount) cudaMemsetAsync(grid.refCount, 0, sizeof(uint32_t) * ic.numCells, stream);
    const uint32_t maxTotal = set.singleCount ? set.capacity : set.numSegs * set.segStride;
    const uint32_t gridN = gridFor(maxTotal ? maxTotal : 1u, SVCM_BLOCK, 8);
    if(maxTotal > 0u && !bboxReady) { k_cellBBox<<<gridN, SVCM_BLOCK, 0, stream>>>(set, grid.bboxBits); launches++; }
    k_cellDims<<<1, 32, 0, stream>>>(grid.bboxBits, ic.invCellSize, grid.tcap, grid.dims);
    launches++;
    if(maxTotal > 0u)
    {
        k_cellKeys<<<gridN, SVCM_BLOCK, 0, stream>>>(set, ic, grid.bboxBits, grid.dims, grid.cellKey, grid.cellEnd,
                                                     withRefCount ? grid.refCount : NULL);
        launches++;
    }
    launchLookbackScan(grid.cellEnd, grid.dims, n, grid.scanState, grid.scanCounter, gen, stream);
    launches++;
    if(maxTotal > 0u)
    {
        k_cellScatterIndex<<<gridFor((maxTotal + 3u) / 4u, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(
            grid.cellKey, set.singleCount, maxTotal, maxTotal, grid.cellEnd, reinterpret_cast<uint32_t*>(grid.sIndex));
        k_cellGather<<<gridN, SVCM_BLOCK, 0, stream>>>(set, grid, NULL);
        launches += 2;
    }
    return launches;
}

int launchCellQuerySort(const IterConst &ic, const QueryStore &queries, const unsigned long long *vertexCount,
                        uint32_t hostVertexCount, const GridBuffers &grid, const QuerySortBuffers &qs, ScanGen &gen,
                        unsigned long long *counters, cudaStream_t stream)
{
    const uint32_t n = grid.tcap + 2u;
    cudaMemsetAsync(qs.cellEnd, 0, sizeof(uint32_t) * n, stream);
    const uint32_t maxCount = vertexCount ? queries.capacity : hostVertexCount;
    if(maxCount == 0u) return 0;
    const uint32_t gridN = gridFor(maxCount, SVCM_BLOCK, 8);
    k_cellQueryKeys<<<gridN, SVCM_BLOCK, 0, stream>>>(queries, vertexCount, hostVertexCount, ic, grid.bboxBits, grid.dims,
                                                      qs.key, qs.cellEnd, counters);
    launchLookbackScan(qs.cellEnd, grid.dims, n, qs.scanState, qs.scanCounter, gen, stream);
    k_cellScatterIndex<<<gridFor((maxCount + 3u) / 4u, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(
        qs.key, vertexCount, hostVertexCount, queries.capacity, qs.cellEnd, qs.order);
    return 3;
}

template<int PCAP, int LCAP, bool kBulk>
static void launchStaged(const DevScene &sc, const IterConst &ic, const QueryStore &queries,
                         const unsigned long long *queryCount, uint32_t hostQueryCount, const GridBuffers &grid,
                         float *accum, float *contribOut, int32_t *countsOut, unsigned long long *counters,
                         uint64_t maxQueries, cudaStream_t stream)
{
    const size_t smem = sizeof(StagedWarpSmem<PCAP, LCAP>) * ST_WARPS;
    static bool configured = false;            // per instantiation; the attribute is per device function
    if(!configured)
    {
        cudaFuncSetAttribute(k_mergeStaged<PCAP, LCAP, kBulk>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        configured = true;
    }
    int perSM = int((227u * 1024u) / (smem + 1024u));
    if(perSM < 1) perSM = 1;
    if(perSM > 8) perSM = 8;
    k_mergeStaged<PCAP, LCAP, kBulk><<<gridFor(maxQueries, ST_WARPS * 32, perSM), ST_WARPS * 32, smem, stream>>>(
        sc, ic, queries, queryCount, hostQueryCount, grid, accum, contribOut, countsOut, counters);
}

void launchCellMerge(const DevScene &sc, const IterConst &ic, const QueryStore &queries,
                     const unsigned long long *queryCount, uint32_t hostQueryCount, const GridBuffers &grid,
                     float *accum, float *contribOut, int32_t *countsOut, unsigned long long *counters,
                     int variant, cudaStream_t stream)
{
    const uint64_t maxQueries = queryCount ? queries.capacity : hostQueryCount;
    if(maxQueries == 0) return;
    switch(variant)
    {
    case 0:
        k_mergeCellLane<<<gridFor(maxQueries, SVCM_BLOCK, 4), SVCM_BLOCK, 0, stream>>>(
            sc, ic, queries, queryCount, hostQueryCount, grid, accum, contribOut, countsOut, counters);
        break;
    case 3:
        launchStaged<128, 8, true>(sc, ic, queries, queryCount, hostQueryCount, grid, accum, contribOut, countsOut,
                                   counters, maxQueries, stream);
        break;
    default:
        launchStaged<256, 16, true>(sc, ic, queries, queryCount, hostQueryCount, grid, accum, contribOut, countsOut,
                                    counters, maxQueries, stream);
        break;
    }
}

void launchParticleBBox(const MergeSet &set, int32_t *bboxBits, cudaStream_t stream)
{
    const uint32_t maxTotal = set.singleCount ? set.capacity : set.numSegs * set.segStride;
    if(maxTotal > 0u) k_cellBBox<<<gridFor(maxTotal, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(set, bboxBits);
}

void launchGatherPackHeader(const MergeSet &localSet, const int32_t *bboxInit, GatherHeader *header, cudaStream_t stream)
{
    cudaMemcpyAsync(header->bbox, bboxInit, sizeof(int32_t) * 6, cudaMemcpyDeviceToDevice, stream);
    if(localSet.capacity > 0u)
        k_cellBBox<<<gridFor(localSet.capacity, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(localSet, header->bbox);
    k_gatherHeaderCount<<<1, 32, 0, stream>>>(localSet.singleCount, header);
}

void launchGatherUnpackHeaders(const unsigned char *blocks, uint64_t blockBytes, uint32_t numSegs, uint32_t segStride,
                               uint32_t *segCountDev, int32_t *bboxBits, unsigned long long *totals, cudaStream_t stream)
{
    k_gatherUnpackHeaders<<<1, 32, 0, stream>>>(blocks, blockBytes, numSegs, segStride, segCountDev, bboxBits, totals);
}

// second half of HashGrid::Build in the BUCKET layout (after keys + histogram + scan): 4-byte index
// scatter through the cursor, then the slot-ordered record gather
int launchBucketReorder(const MergeSet &set, const GridBuffers &grid, uint32_t numCells, cudaStream_t stream)
{
    const uint32_t maxTotal = set.singleCount ? set.capacity : set.numSegs * set.segStride;
    if(maxTotal == 0u) return 0;
    k_cellScatterIndex<<<gridFor((maxTotal + 3u) / 4u, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(
        grid.cellKey, set.singleCount, maxTotal, maxTotal, grid.cellCount, reinterpret_cast<uint32_t*>(grid.sIndex));
    k_cellGather<<<gridFor(maxTotal, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(set, grid, grid.cellCount + numCells);
    return 2;
}

void launchCellRefKeys(const IterConst &ic, const GridBuffers &grid, uint32_t count, uint32_t *out, cudaStream_t stream)
{
    if(count == 0u) return;
    k_cellRefKeys<<<gridFor(count, SVCM_BLOCK, 8), SVCM_BLOCK, 0, stream>>>(ic, grid, count, out);
}

} // namespace svcm
