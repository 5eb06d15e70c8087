// bvh_check.cpp — CPU check of the hierarchy traversal (test infrastructure, built by
// tests/test_bvh.py with g++ -ffp-contract=off). It instantiates the traversal templates of
// smallvcm_b200/csrc/svcm_bvh.h — the code the CUDA kernels run — with plain C++ primitive
// tests and compares every ray with a loop over the whole list in list order, i.e. with the
// semantics of the reference's GeometryList::Intersect / IntersectP (src/geometry.hxx:65-91):
// closest hit, first in the list wins equal distances; any hit below a limit.
// The primitive tests here only need to be deterministic functions of (primitive, ray): the
// claim checked is "traversal == loop for the same test", for arbitrary geometry and rays.
#include <cmath>
#include <cstdint>
#include <vector>

#include "svcm.h"
#include "svcm_bvh.h"

namespace {

struct V { float x, y, z; };
inline V sub(V a, V b) { return V{ a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V cross(V a, V b) { return V{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V ld(const float *p) { return V{ p[0], p[1], p[2] }; }

// the edge-function triangle test and the two-root sphere test, shaped like the reference's
bool hitPrim(const SvcmPrimitive &p, V org, V dir, float tmax, float &dist)
{
    if(p.type == SVCM_PRIM_TRIANGLE)
    {
        const V ao = sub(ld(p.v), org), bo = sub(ld(p.v + 3), org), co = sub(ld(p.v + 6), org);
        const float v0 = dot(cross(co, bo), dir), v1 = dot(cross(bo, ao), dir), v2 = dot(cross(ao, co), dir);
        if((v0 < 0.f && v1 < 0.f && v2 < 0.f) || (v0 >= 0.f && v1 >= 0.f && v2 >= 0.f))
        {
            const V n = ld(p.v + 9);
            const float d = dot(n, ao) / dot(n, dir);
            if(d > 0.f && d < tmax) { dist = d; return true; }
        }
        return false;
    }
    const V to = sub(org, ld(p.v));
    const float A = dot(dir, dir), B = 2.f * dot(dir, to), Cc = dot(to, to) - p.v[3] * p.v[3];
    const float disc = B * B - 4.f * A * Cc;
    if(disc < 0.f) return false;
    const double s = std::sqrt(double(disc));
    const double q = (B < 0.f) ? (-double(B) - s) / 2.0 : (-double(B) + s) / 2.0;
    double t0 = q / double(A), t1 = double(Cc) / q;
    if(t0 > t1) { const double t = t0; t0 = t1; t1 = t; }
    if(t0 > 0.0 && t0 < double(tmax)) { dist = float(t0); return true; }
    if(t1 > 0.0 && t1 < double(tmax)) { dist = float(t1); return true; }
    return false;
}

} // namespace

// rays: [n][7] = origin, direction, tLimit (for the any-hit query). Results per ray:
// closest-hit primitive (original index, -1 = none) and distance by loop and by traversal,
// any-hit by loop and by traversal. Returns the number of rays on which the two disagree.
extern "C" __attribute__((visibility("default")))
int bvh_check(const SvcmPrimitive *prims, int32_t numPrims, const SvcmBvhNode *nodes, const int32_t *order,
              const float *rays, int32_t numRays, int32_t *loopPrim, float *loopDist, int32_t *bvhPrim,
              float *bvhDist, int32_t *loopAny, int32_t *bvhAny)
{
    const svcm::BvhNode *tree = reinterpret_cast<const svcm::BvhNode*>(nodes);
    int bad = 0;
    for(int32_t r = 0; r < numRays; r++)
    {
        const V org = ld(rays + 7 * r), dir = ld(rays + 7 * r + 3);
        const float limit = rays[7 * r + 6];
        // GeometryList::Intersect: running best as the upper bound, strict <
        float best = 1e36f; int32_t prim = -1;
        for(int32_t i = 0; i < numPrims; i++) { float d; if(hitPrim(prims[i], org, dir, best, d)) { best = d; prim = i; } }
        int32_t any = 0;
        for(int32_t i = 0; i < numPrims && !any; i++) { float d; any = hitPrim(prims[i], org, dir, limit, d) ? 1 : 0; }

        float tBest = 1e36f; int32_t tOrig = 0x7FFFFFFF, tSlot = -1;
        svcm::bvhClosest(tree, org.x, org.y, org.z, dir.x, dir.y, dir.z, tBest, tOrig, tSlot,
                         [&](int32_t slot, float &d, int32_t &orig) -> bool
                         { orig = order[slot]; return hitPrim(prims[orig], org, dir, 1e36f, d); });
        const int32_t tAny = svcm::bvhAny(tree, org.x, org.y, org.z, dir.x, dir.y, dir.z, limit,
                                          [&](int32_t slot) -> bool
                                          { float d; return hitPrim(prims[order[slot]], org, dir, limit, d); }) ? 1 : 0;
        const int32_t tPrim = tSlot < 0 ? -1 : tOrig;
        loopPrim[r] = prim; loopDist[r] = prim < 0 ? 0.f : best;
        bvhPrim[r] = tPrim; bvhDist[r] = tPrim < 0 ? 0.f : tBest;
        loopAny[r] = any; bvhAny[r] = tAny;
        if(prim != tPrim || (prim >= 0 && best != tBest) || any != tAny) bad++;
    }
    return bad;
}
