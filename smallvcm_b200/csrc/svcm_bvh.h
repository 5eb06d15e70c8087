// svcm_bvh.h — bounding volume hierarchy over the scene's primitives.
//
// The reference intersects a ray with EVERY primitive (GeometryList::Intersect / IntersectP,
// reference src/geometry.hxx:65-91) — its README names the missing acceleration structure as
// the limitation. Scenes of up to SVCM_MAX_PRIMITIVES primitives keep that loop (the list sits
// in the constant bank); larger scenes live in global memory under this hierarchy. The
// traversal must return what the loop returns, so:
//   * a primitive's hit distance is computed by the same test, independent of traversal order;
//   * the loop keeps the FIRST primitive among equal distances (strict <, list order); here the
//     winner of a tie is the lower ORIGINAL index, which is the same primitive;
//   * boxes are padded and the slab test is slack on both ends, so that no primitive the
//     arithmetic of the test would hit is ever culled.
//
// The traversal is a template over the primitive test so that the same code runs in the CUDA
// kernels and, with plain C++ tests, in the CPU check of tests/native/bvh_check.cpp.
#ifndef SVCM_BVH_H_
#define SVCM_BVH_H_

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SVCM_BVH_HD __host__ __device__ __forceinline__
#else
#define SVCM_BVH_HD inline
#endif

namespace svcm {

// 32 bytes = two 16-byte loads. Inner node: count == 0, children at first and first + 1.
// Leaf: `count` primitive slots starting at `first` (slots index the reordered primitive array).
struct alignas(16) BvhNode
{
    float   lo[3]; int32_t first;
    float   hi[3]; int32_t count;
};

#ifndef SVCM_MAX_RUNS
#define SVCM_MAX_RUNS  8           // runs of the constant-bank primitive list (primitiveRuns, svcm_bvh.cpp; also in svcm.h)
#endif
#define SVCM_BVH_STACK 64          // the builder never produces a deeper tree (bvhBuild checks)
#define SVCM_BVH_LEAF  4           // primitives per leaf the builder aims for

// Conservative ray / box test for t in [0, tLimit]: true if the ray MAY cross the box.
// inv* = 1 / direction (infinite for zero components: the slab then only tests the origin).
// fminf / fmaxf drop NaNs (0 * inf when the origin lies on a slab plane of a zero component).
SVCM_BVH_HD bool bvhSlab(const BvhNode &n, float ox, float oy, float oz, float ix, float iy, float iz,
                         float tLimit, float &tEnter)
{
    const float ax = (n.lo[0] - ox) * ix, bx = (n.hi[0] - ox) * ix;
    const float ay = (n.lo[1] - oy) * iy, by = (n.hi[1] - oy) * iy;
    const float az = (n.lo[2] - oz) * iz, bz = (n.hi[2] - oz) * iz;
    const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
    const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tLimit));
    tEnter = t0;
    // a few ulps of slack on the exit side; the boxes themselves are padded by the builder
    return t0 <= t1 * 1.0000005f;
}

// Closest hit. hit(slot, dist, orig): runs the primitive test of `slot` with no upper distance
// bound, returns whether it hit, its distance and the primitive's ORIGINAL list index.
// On return best / bestOrig / bestSlot describe the winner (bestSlot < 0: nothing hit).
template<typename HitFn>
SVCM_BVH_HD void bvhClosest(const BvhNode *nodes, float ox, float oy, float oz, float dx, float dy, float dz,
                            float &best, int32_t &bestOrig, int32_t &bestSlot, HitFn hit)
{
    const float ix = 1.f / dx, iy = 1.f / dy, iz = 1.f / dz;
    int32_t stack[SVCM_BVH_STACK];
    int sp = 0;
    int32_t node = 0;
    float tEnter;
    if(!bvhSlab(nodes[0], ox, oy, oz, ix, iy, iz, best, tEnter)) return;
    for(;;)
    {
        const BvhNode n = nodes[node];
        if(n.count > 0)
        {
            for(int32_t k = 0; k < n.count; k++)
            {
                float d; int32_t orig;
                if(hit(n.first + k, d, orig) && (d < best || (d == best && orig < bestOrig)))
                { best = d; bestOrig = orig; bestSlot = n.first + k; }
            }
        }
        else
        {
            float tl, tr;
            const bool hl = bvhSlab(nodes[n.first], ox, oy, oz, ix, iy, iz, best, tl);
            const bool hr = bvhSlab(nodes[n.first + 1], ox, oy, oz, ix, iy, iz, best, tr);
            if(hl && hr)
            {
                // nearer child first; the other one waits on the stack
                const bool leftFirst = tl <= tr;
                stack[sp++] = leftFirst ? n.first + 1 : n.first;
                node = leftFirst ? n.first : n.first + 1;
                continue;
            }
            if(hl) { node = n.first; continue; }
            if(hr) { node = n.first + 1; continue; }
        }
        // pop; boxes that fell behind the current best since they were pushed are skipped
        for(;;)
        {
            if(sp == 0) return;
            node = stack[--sp];
            if(bvhSlab(nodes[node], ox, oy, oz, ix, iy, iz, best, tEnter)) break;
        }
    }
}

// Any hit with distance in (0, tLimit). hit(slot): the primitive test of `slot` against tLimit.
template<typename HitFn>
SVCM_BVH_HD bool bvhAny(const BvhNode *nodes, float ox, float oy, float oz, float dx, float dy, float dz,
                        float tLimit, HitFn hit)
{
    const float ix = 1.f / dx, iy = 1.f / dy, iz = 1.f / dz;
    int32_t stack[SVCM_BVH_STACK];
    int sp = 0;
    int32_t node = 0;
    float tEnter;
    if(!bvhSlab(nodes[0], ox, oy, oz, ix, iy, iz, tLimit, tEnter)) return false;
    for(;;)
    {
        const BvhNode n = nodes[node];
        if(n.count > 0)
        {
            for(int32_t k = 0; k < n.count; k++)
                if(hit(n.first + k)) return true;
        }
        else
        {
            float tl, tr;
            const bool hl = bvhSlab(nodes[n.first], ox, oy, oz, ix, iy, iz, tLimit, tl);
            const bool hr = bvhSlab(nodes[n.first + 1], ox, oy, oz, ix, iy, iz, tLimit, tr);
            if(hl && hr) { stack[sp++] = n.first + 1; node = n.first; continue; }
            if(hl) { node = n.first; continue; }
            if(hr) { node = n.first + 1; continue; }
        }
        if(sp == 0) return false;
        node = stack[--sp];
    }
}

} // namespace svcm

#endif // SVCM_BVH_H_
