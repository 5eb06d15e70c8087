// svcm_kutil.cuh — warp-level utilities shared by the kernel translation units
// (svcm_kernels.cu: trace kernels + the bucket-layout grid; svcm_grid.cu: the cell-layout grid,
// its sort and the staged merge).
#ifndef SVCM_KUTIL_CUH_
#define SVCM_KUTIL_CUH_
#include <cuda_runtime.h>
#include <stdint.h>
#include "svcm_internal.h"

namespace svcm {

#define SVCM_BLOCK 256
#define FULL_MASK 0xffffffffu

// ------------------------------------------------------------------------------------------
// small device utilities
// ------------------------------------------------------------------------------------------
SVCM_D uint32_t laneId() { return threadIdx.x & 31u; }

// Warp-aggregated append: every lane of the warp must call it; lanes with pred get a
// distinct slot from one atomicAdd issued by the first such lane.
SVCM_D uint32_t warpAppend(uint32_t *counter, bool pred)
{
    const uint32_t m = __ballot_sync(FULL_MASK, pred);
    if(m == 0u) return 0u;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0u;
    if(int(laneId()) == leader) base = atomicAdd(counter, uint32_t(__popc(m)));
    base = __shfl_sync(FULL_MASK, base, leader);
    return base + uint32_t(__popc(m & ((1u << laneId()) - 1u)));
}
SVCM_D uint32_t warpAppend64(unsigned long long *counter, bool pred)
{
    const uint32_t m = __ballot_sync(FULL_MASK, pred);
    if(m == 0u) return 0u;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0ull;
    if(int(laneId()) == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
    return uint32_t(base) + uint32_t(__popc(m & ((1u << laneId()) - 1u)));
}

// Opportunistic aggregation for appends issued from divergent code: the lanes that happen to
// be converged here share one atomicAdd.
SVCM_D uint32_t activeAppend(uint32_t *counter)
{
    const uint32_t m = __activemask();
    const int leader = __ffs(m) - 1;
    uint32_t base = 0u;
    if(int(laneId()) == leader) base = atomicAdd(counter, uint32_t(__popc(m)));
    base = __shfl_sync(m, base, leader);
    return base + uint32_t(__popc(m & ((1u << laneId()) - 1u)));
}

// adds a per-thread count into a 64-bit device counter with one atomic per warp
SVCM_D void warpCount(unsigned long long *counter, uint32_t mine)
{
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(FULL_MASK, mine, o);
    if(laneId() == 0u && mine) atomicAdd(counter, (unsigned long long)mine);
}

SVCM_D void addPixel(float *accum, uint32_t pixel, f3 c)
{
    float *p = accum + 3ull * pixel;
    if(c.x != 0.f) atomicAdd(p + 0, c.x);
    if(c.y != 0.f) atomicAdd(p + 1, c.y);
    if(c.z != 0.f) atomicAdd(p + 2, c.z);
}

SVCM_D float4 mk4(f3 a, float w) { return make_float4(a.x, a.y, a.z, w); }
SVCM_D f3 xyz(float4 a) { return mk3(a.x, a.y, a.z); }


// the particle set of one grid build: a single device-counted array, or the padded segments of
// an all-gather
SVCM_D uint32_t mergeSetTotal(const MergeSet &set)
{
    if(set.singleCount) return uint32_t(min(*set.singleCount, (unsigned long long)set.capacity));
    return set.numSegs * set.segStride;
}
// virtual index -> (block, slot); one block in the single-rank case
struct MergeSlot { uint32_t seg, j; };
SVCM_D MergeSlot mergeSetSlot(const MergeSet &set, uint32_t i)
{
    MergeSlot s; s.seg = 0u; s.j = i;
    if(!set.singleCount) { s.seg = i / set.segStride; s.j = i - s.seg * set.segStride; }
    return s;
}
SVCM_D bool mergeSetValid(const MergeSet &set, MergeSlot s)
{
    return set.singleCount ? true : s.j < set.segCountDev[s.seg];
}
template<typename T> SVCM_D const T *mergeSetStream(const MergeSet &set, const T *base, MergeSlot s)
{
    return reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(base) + uint64_t(s.seg) * set.segBytes) + s.j;
}

// HashGrid::GetCellIndex(Vec3i), reference src/hashgrid.hxx:179-187
SVCM_D uint32_t cellHash(int32_t cx, int32_t cy, int32_t cz, uint32_t numCells)
{
    return ((uint32_t(cx) * 73856093u) ^ (uint32_t(cy) * 19349663u) ^ (uint32_t(cz) * 83492791u)) % numCells;
}


// x % numCells for the reference hash (src/hashgrid.hxx:179-187) without an integer division:
// floor(a / n) = mulhi64(a, ceil(2^64 / n)) exactly for a < 2^32
SVCM_D uint32_t refHash(int32_t cx, int32_t cy, int32_t cz, const IterConst &ic)
{
    const uint32_t a = (uint32_t(cx) * 73856093u) ^ (uint32_t(cy) * 19349663u) ^ (uint32_t(cz) * 83492791u);
    if(ic.hashMagic == 0ull) return 0u;                       // numCells == 1
    const uint32_t q = uint32_t(__umul64hi((unsigned long long)a, ic.hashMagic));
    return a - q * ic.numCells;
}

// ---- spatial order of the merge queries: 21-bit Morton code of the query's own cell, coarsened by
// `shift` bits per axis until 7 bits per axis suffice. All queries of one code form a GROUP: the
// cells of a cube of 2^shift cells per axis.
// 7 bits -> every third bit of 21
SVCM_D uint32_t morton7(uint32_t v)
{
    v &= 0x7Fu;
    v = (v | (v << 8)) & 0x0000700Fu;     // ---- ---- ---- ---- -654 ---- ---- 3210
    v = (v | (v << 4)) & 0x000430C3u;     // ---- ---- ---- -6-- --54 ---- 32-- --10
    v = (v | (v << 2)) & 0x00049249u;     // ---- ---- ---- -6-- 5--4 --3- -2-- 1--0
    return v;
}
// every third bit of 21 -> 7 bits
SVCM_D uint32_t unmorton7(uint32_t v)
{
    v &= 0x00049249u;
    v = (v | (v >> 2)) & 0x000430C3u;
    v = (v | (v >> 4)) & 0x0000700Fu;
    v = (v | (v >> 8)) & 0x0000007Fu;
    return v;
}
SVCM_D int mortonShift(f3 bbMin, f3 bbMax, float invCellSize)
{
    const float ext = fmaxf(fmaxf(bbMax.x - bbMin.x, bbMax.y - bbMin.y), bbMax.z - bbMin.z);
    const float cells = fminf(fmaxf(floorf(invCellSize * ext), 0.f), 1073741824.f);
    uint32_t top = uint32_t(cells) >> 7;
    int shift = 0;
    while(top) { shift++; top >>= 1; }
    return shift;
}

// persistent-style grid: a multiple of the SM count, capped by the work available
int numSMs();
uint32_t gridFor(uint64_t items, int perBlock, int blocksPerSM);

} // namespace svcm
#endif
