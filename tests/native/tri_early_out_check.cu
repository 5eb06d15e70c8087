// tri_early_out_check.cu — host-side check of hitTriangle's plane-distance early-out
// (smallvcm_b200/csrc/svcm_device.cuh). Compiled by tests/test_bvh.py with nvcc as plain host code.
//
// hitTriangle decides "the plane distance cannot lie in (0, tmax)" before the edge tests and without
// the division. That is only legal if it never rejects a ray the reference's order of evaluation
// (edge tests, then distance = num / den, then 0 < distance < tmax; reference src/geometry.hxx:125-156)
// would accept. fullTest below is that order, written with the same expressions; the program runs
// both over random and adversarial rays (zero components, rays in the triangle's plane, rays whose
// distance is within ulps of tmax and of 0, subnormal and huge scales, negative and infinite tmax)
// and counts disagreements. Exit code 0 and "disagreements 0" on success.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "svcm_device.cuh"

using namespace svcm;

static bool fullTest(const float *v, f3 org, f3 dir, float tmax, float &oDist)
{
    const f3 ao = ld3(v) - org;
    const f3 bo = ld3(v + 3) - org;
    const f3 co = ld3(v + 6) - org;
    const float v0d = dot(cross(co, bo), dir);
    const float v1d = dot(cross(bo, ao), dir);
    const float v2d = dot(cross(ao, co), dir);
    if(((v0d < 0.f) && (v1d < 0.f) && (v2d < 0.f)) || ((v0d >= 0.f) && (v1d >= 0.f) && (v2d >= 0.f)))
    {
        const f3 n = ld3(v + 9);
        const float distance = dot(n, ao) / dot(n, dir);
        if((distance > 0.f) && (distance < tmax)) { oDist = distance; return true; }
    }
    return false;
}

static uint64_t state = 0x9E3779B97F4A7C15ull;
static uint32_t next32() { state ^= state << 13; state ^= state >> 7; state ^= state << 17; return uint32_t(state >> 32); }
static float uni() { return float(next32() >> 8) * (1.f / 16777216.f); }
static float sym(float s) { return (uni() * 2.f - 1.f) * s; }
static float nudge(float x, int ulps) { uint32_t b; memcpy(&b, &x, 4); b += uint32_t(ulps); float r; memcpy(&r, &b, 4); return r; }

int main(int argc, char **argv)
{
    const long rounds = argc > 1 ? atol(argv[1]) : 2000000;
    long bad = 0, hits = 0, early = 0, total = 0;
    const float scales[6] = { 1.f, 1e-3f, 1e3f, 1e-18f, 1e15f, 3e-37f };
    for(long r = 0; r < rounds; r++)
    {
        const float s = scales[r % 6];
        float v[12];
        for(int k = 0; k < 9; k++) v[k] = sym(s);
        if(r % 5 == 0) { v[2] = v[5] = v[8] = sym(s); }                    // axis-aligned, like the reference's walls
        {
            // the stored normal: normalized cross product of the edges, like the scene builder
            const f3 e1 = ld3(v + 3) - ld3(v), e2 = ld3(v + 6) - ld3(v);
            f3 n = cross(e1, e2);
            const float len = sqrtf(dot(n, n));
            if(!(len > 0.f)) n = mk3(0.f, 0.f, 1.f); else n = n / len;
            v[9] = n.x; v[10] = n.y; v[11] = n.z;
        }
        const f3 centre = (ld3(v) + ld3(v + 3) + ld3(v + 6)) / 3.f;
        for(int c = 0; c < 24; c++)
        {
            f3 org = mk3(sym(2.f * s), sym(2.f * s), sym(2.f * s));
            f3 target = centre + mk3(sym(s), sym(s), sym(s)) * ((c & 1) ? 0.2f : 1.5f);
            if(c % 6 == 2) org = ld3(v + 3 * (c % 3)) + mk3(sym(1e-6f * s), sym(1e-6f * s), sym(1e-6f * s));   // starting on a vertex
            if(c % 6 == 3) { const f3 n = ld3(v + 9); org = org - n * dot(n, org - ld3(v)); }                       // starting in the plane
            f3 dir = target - org;
            const float len = sqrtf(dot(dir, dir));
            if(len > 0.f && c % 4 != 3) dir = dir / len;
            if(c % 6 == 4) { const f3 n = ld3(v + 9); dir = dir - n * dot(n, dir); }                                // direction in the plane
            if(c % 8 == 5) dir.x = 0.f;
            if(c % 8 == 6) { dir.y = 0.f; dir.z = -0.f; }
            const f3 n = ld3(v + 9);
            const float d = dot(n, ld3(v) - org) / dot(n, dir);
            float tmaxs[10] = { 1e36f, d, nudge(d, 1), nudge(d, -1), nudge(d, 40), nudge(d, -40), len - 2e-3f, 0.f, -1.f, INFINITY };
            for(int t = 0; t < 10; t++)
            {
                float a = -1.f, b = -1.f;
                const bool ha = hitTriangle(v, org, dir, tmaxs[t], a);
                const bool hb = fullTest(v, org, dir, tmaxs[t], b);
                total++;
                if(ha != hb || (ha && memcmp(&a, &b, 4) != 0))
                {
                    if(bad < 10)
                        fprintf(stderr, "mismatch: early-out %d (%.9g) vs full %d (%.9g), tmax %.9g, num/den %.9g\n", ha, a, hb, b, tmaxs[t], d);
                    bad++;
                }
                hits += hb;
                {
                    // how often the early-out fires (informational)
                    const float num = dot(n, ld3(v) - org), den = dot(n, dir);
                    const float sn = (den < 0.f) ? -num : num, q = tmaxs[t] * fabsf(den);
                    early += (!(sn > 0.f) || (q >= 1e-30f && sn > q * 1.00001f));
                }
            }
        }
    }
    printf("tests %ld hits %ld early-outs %ld disagreements %ld\n", total, hits, early, bad);
    return bad ? 1 : 0;
}
