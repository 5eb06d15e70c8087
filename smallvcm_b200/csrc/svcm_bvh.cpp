// svcm_bvh.cpp — host-side builder of the hierarchy in svcm_bvh.h (binned surface-area
// heuristic, median split when the bins cannot separate the centroids). Plain C++: it runs once
// per scene at svcm_create, on the flattened primitive list.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "svcm.h"
#include "svcm_bvh.h"

namespace svcm {

namespace {

struct Box
{
    float lo[3], hi[3];
    void clear() { for(int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; } }
    void grow(const Box &o) { for(int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], o.lo[a]); hi[a] = std::max(hi[a], o.hi[a]); } }
    void grow(const float *p) { for(int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    float halfArea() const
    {
        const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return (x < 0.f) ? 0.f : x * y + y * z + z * x;
    }
};

struct Item { Box box; float c[3]; int32_t orig; };

struct Job { int32_t node, begin, end, depth; };

const int kBins = 16;
const int kMaxDepth = SVCM_BVH_STACK - 8;   // deepest leaf; the traversal stack holds one entry per level

} // namespace

// Builds the hierarchy. nodes[0] is the root; order[slot] = original index of the primitive in
// `slot`. Returns 0, or -1 for non-finite geometry / an unknown primitive type.
int bvhBuild(const SvcmPrimitive *prims, int32_t count, std::vector<BvhNode> &nodes, std::vector<int32_t> &order)
{
    nodes.clear(); order.clear();
    if(count <= 0) return -1;
    std::vector<Item> items;
    items.resize(size_t(count));
    Box scene; scene.clear();
    for(int32_t i = 0; i < count; i++)
    {
        Item &it = items[size_t(i)];
        it.orig = i; it.box.clear();
        const float *v = prims[i].v;
        const int used = (prims[i].type == SVCM_PRIM_TRIANGLE) ? 9 : 4;
        for(int k = 0; k < used; k++)
            if(!std::isfinite(v[k])) return -1;                 // std::min / std::max would silently drop a NaN
        if(prims[i].type == SVCM_PRIM_TRIANGLE)
        {
            it.box.grow(v); it.box.grow(v + 3); it.box.grow(v + 6);
        }
        else if(prims[i].type == SVCM_PRIM_SPHERE)
        {
            const float r = std::fabs(v[3]);
            const float lo[3] = { v[0] - r, v[1] - r, v[2] - r }, hi[3] = { v[0] + r, v[1] + r, v[2] + r };
            it.box.grow(lo); it.box.grow(hi);
        }
        else return -1;
        for(int a = 0; a < 3; a++)
        {
            if(!std::isfinite(it.box.lo[a]) || !std::isfinite(it.box.hi[a])) return -1;
            it.c[a] = 0.5f * it.box.lo[a] + 0.5f * it.box.hi[a];
        }
        scene.grow(it.box);
    }

    nodes.reserve(size_t(count) * 2u);
    nodes.push_back(BvhNode());
    std::vector<Job> jobs;
    jobs.push_back(Job{ 0, 0, count, 0 });
    while(!jobs.empty())
    {
        const Job job = jobs.back(); jobs.pop_back();
        Box bounds, cbounds; bounds.clear(); cbounds.clear();
        for(int32_t i = job.begin; i < job.end; i++) { bounds.grow(items[size_t(i)].box); cbounds.grow(items[size_t(i)].c); }
        BvhNode &self = nodes[size_t(job.node)];
        for(int a = 0; a < 3; a++) { self.lo[a] = bounds.lo[a]; self.hi[a] = bounds.hi[a]; }
        const int32_t n = job.end - job.begin;
        if(n <= SVCM_BVH_LEAF || job.depth >= kMaxDepth)
        {
            self.first = job.begin; self.count = n;
            continue;
        }

        // binned SAH over the three axes
        int bestAxis = -1, bestBin = -1;
        float bestCost = float(n) * bounds.halfArea();       // cost of staying a leaf (up to a constant)
        for(int axis = 0; axis < 3; axis++)
        {
            const float lo = cbounds.lo[axis], ext = cbounds.hi[axis] - lo;
            if(!(ext > 0.f)) continue;
            Box binBox[kBins]; int32_t binCount[kBins];
            for(int b = 0; b < kBins; b++) { binBox[b].clear(); binCount[b] = 0; }
            const float scale = float(kBins) / ext;
            for(int32_t i = job.begin; i < job.end; i++)
            {
                const Item &it = items[size_t(i)];
                const int b = std::min(kBins - 1, std::max(0, int((it.c[axis] - lo) * scale)));
                binBox[b].grow(it.box); binCount[b]++;
            }
            float rightArea[kBins]; int32_t rightCount[kBins];
            Box acc; acc.clear(); int32_t cnt = 0;
            for(int b = kBins - 1; b > 0; b--)
            {
                acc.grow(binBox[b]); cnt += binCount[b];
                rightArea[b] = acc.halfArea(); rightCount[b] = cnt;
            }
            acc.clear(); cnt = 0;
            for(int b = 0; b < kBins - 1; b++)
            {
                acc.grow(binBox[b]); cnt += binCount[b];
                if(cnt == 0 || rightCount[b + 1] == 0) continue;
                const float cost = 0.5f * bounds.halfArea() + acc.halfArea() * float(cnt) + rightArea[b + 1] * float(rightCount[b + 1]);
                if(cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = b; }
            }
        }

        int32_t mid;
        if(bestAxis >= 0)
        {
            const float lo = cbounds.lo[bestAxis], scale = float(kBins) / (cbounds.hi[bestAxis] - lo);
            Item *first = &items[size_t(job.begin)], *last = first + n;
            Item *m = std::partition(first, last, [&](const Item &it) {
                return std::min(kBins - 1, std::max(0, int((it.c[bestAxis] - lo) * scale))) <= bestBin; });
            mid = job.begin + int32_t(m - first);
        }
        else
            mid = job.begin;                                   // force the fallback below
        if(mid == job.begin || mid == job.end)
        {
            // no useful SAH split (identical centroids, or a leaf looked cheaper than any split of
            // more than SVCM_BVH_LEAF primitives): halve by the median of the widest centroid axis
            int axis = 0;
            for(int a = 1; a < 3; a++)
                if(cbounds.hi[a] - cbounds.lo[a] > cbounds.hi[axis] - cbounds.lo[axis]) axis = a;
            mid = job.begin + n / 2;
            std::nth_element(items.begin() + job.begin, items.begin() + mid, items.begin() + job.end,
                             [axis](const Item &x, const Item &y) { return x.c[axis] < y.c[axis] || (x.c[axis] == y.c[axis] && x.orig < y.orig); });
        }

        const int32_t left = int32_t(nodes.size());
        nodes.push_back(BvhNode()); nodes.push_back(BvhNode());
        nodes[size_t(job.node)].first = left; nodes[size_t(job.node)].count = 0;    // `self` may dangle after push_back
        jobs.push_back(Job{ left + 1, mid, job.end, job.depth + 1 });
        jobs.push_back(Job{ left, job.begin, mid, job.depth + 1 });
    }

    // pad every box: the slab test and the primitive tests round independently, a primitive
    // the test would hit must never be culled (svcm_bvh.h). 1e-5 of the largest coordinate
    // is ~40 float ulps at that magnitude.
    float maxAbs = 1e-30f;
    for(int a = 0; a < 3; a++) maxAbs = std::max(maxAbs, std::max(std::fabs(scene.lo[a]), std::fabs(scene.hi[a])));
    const float pad = 1e-5f * maxAbs;
    for(size_t i = 0; i < nodes.size(); i++)
        for(int a = 0; a < 3; a++) { nodes[i].lo[a] -= pad; nodes[i].hi[a] += pad; }

    order.resize(size_t(count));
    for(int32_t i = 0; i < count; i++) order[size_t(i)] = items[size_t(i)].orig;
    return 0;
}

// Cuts a primitive list small enough for the constant bank into at most SVCM_MAX_RUNS contiguous
// runs (DevScene::runs, svcm_device.cuh); the list order is kept. Consecutive triangles whose common
// box stays small against the scene's (surface area at most kRunAreaFraction of it) form a run that
// a ray skips when it misses the box; everything else — large triangles, spheres (their own
// discriminant test already rejects early) — goes into runs that are always tested (bit clear in
// cullMask, box = everything). Boxes are padded like the hierarchy's. A skippable run of fewer than
// three triangles is not worth its box test and is marked always-tested. `plain`: one always-tested
// run with the whole list. Returns the number of runs.
int primitiveRuns(const SvcmPrimitive *prims, int32_t n, bool plain, BvhNode *runs, uint32_t *cullMask)
{
    const float kRunAreaFraction = 0.05f;
    *cullMask = 0u;
    if(n <= 0) return 0;
    std::vector<Box> boxes((size_t)n);
    Box scene; scene.clear();
    float maxAbs = 1e-30f; bool finite = true;
    for(int32_t i = 0; i < n; i++)
    {
        const SvcmPrimitive &p = prims[i];
        Box &b = boxes[(size_t)i];
        b.clear();
        if(p.type == SVCM_PRIM_TRIANGLE) { b.grow(p.v); b.grow(p.v + 3); b.grow(p.v + 6); }
        else for(int a = 0; a < 3; a++) { b.lo[a] = p.v[a] - p.v[3]; b.hi[a] = p.v[a] + p.v[3]; }
        for(int a = 0; a < 3; a++)
        {
            finite = finite && std::isfinite(b.lo[a]) && std::isfinite(b.hi[a]);
            maxAbs = std::max(maxAbs, std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])));
        }
        scene.grow(b);
    }
    const float pad = 1e-5f * maxAbs;
    const float limit = kRunAreaFraction * scene.halfArea();
    if(!finite) plain = true;

    int numRuns = 0;
    auto close = [&](int32_t first, int32_t end, bool cull, const Box &b)
    {
        cull = cull && (end - first) >= 3;
        BvhNode &r = runs[numRuns];
        r.first = first; r.count = end - first;
        for(int a = 0; a < 3; a++) { r.lo[a] = cull ? b.lo[a] - pad : -3.0e38f; r.hi[a] = cull ? b.hi[a] + pad : 3.0e38f; }
        if(cull) *cullMask |= 1u << numRuns;
        numRuns++;
    };
    auto isSmall = [&](int32_t i) { return !plain && prims[i].type == SVCM_PRIM_TRIANGLE && boxes[(size_t)i].halfArea() <= limit; };
    int32_t first = 0; bool cull = isSmall(0); Box cur = boxes[0];
    for(int32_t i = 1; i < n; i++)
    {
        const bool small = isSmall(i);
        Box joined = cur; joined.grow(boxes[(size_t)i]);
        const bool extend = (small == cull) && (!cull || joined.halfArea() <= limit);
        if(extend) { cur = joined; continue; }
        if(numRuns == SVCM_MAX_RUNS - 1) { cull = false; cur = joined; continue; }   // the last slot takes the rest, always tested
        close(first, i, cull, cur);
        first = i; cull = small; cur = boxes[(size_t)i];
    }
    close(first, n, cull, cur);
    return numRuns;
}

} // namespace svcm

extern "C" {

SVCM_API int svcm_bvh_build(const SvcmPrimitive *primitives, int32_t numPrimitives, SvcmBvhNode *nodes,
                            int32_t nodeCapacity, int32_t *order, int32_t *numNodes)
{
    if(!primitives || numPrimitives <= 0 || !nodes || !order || !numNodes) return SVCM_E_INVALID;
    std::vector<svcm::BvhNode> built; std::vector<int32_t> perm;
    if(svcm::bvhBuild(primitives, numPrimitives, built, perm) != 0) return SVCM_E_INVALID;
    *numNodes = int32_t(built.size());
    if(int32_t(built.size()) > nodeCapacity) return SVCM_E_NOMEM;
    static_assert(sizeof(SvcmBvhNode) == sizeof(svcm::BvhNode), "node layouts must match");
    memcpy(nodes, built.data(), sizeof(svcm::BvhNode) * built.size());
    memcpy(order, perm.data(), sizeof(int32_t) * perm.size());
    return SVCM_OK;
}

SVCM_API int svcm_primitive_runs(const SvcmPrimitive *primitives, int32_t numPrimitives, SvcmBvhNode *runs,
                                 uint32_t *skippable, int32_t *numRuns)
{
    if(!primitives || numPrimitives <= 0 || numPrimitives > SVCM_MAX_PRIMITIVES || !runs || !skippable || !numRuns)
        return SVCM_E_INVALID;
    svcm::BvhNode built[SVCM_MAX_RUNS];
    *numRuns = svcm::primitiveRuns(primitives, numPrimitives, false, built, skippable);
    memcpy(runs, built, sizeof(svcm::BvhNode) * size_t(*numRuns));
    return SVCM_OK;
}

} // extern "C"
