// svcm_scene_builtin.cpp — host-side builder of the reference's four Cornell-box scenes as
// flat SvcmSceneDesc data (include/svcm.h), for callers that do not link the reference
// headers (bench.py, the tests on the GPU box, the Python host layer).
//
// It follows Scene::LoadCornellBox + Scene::BuildSceneSphere (reference src/scene.hxx:132-398),
// Camera::Setup (src/camera.hxx:36-80), the AreaLight / DirectionalLight constructors
// (src/lights.hxx:116-129,239-242) and Triangle's constructor (src/geometry.hxx:109-123),
// with the same float arithmetic in the same order, so that the result is BIT-IDENTICAL to
// flattening the reference's own Scene (checked by tests/test_scene_builder.py against
// fixtures made from the compiled reference). Plain host C++: no CUDA here.
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "svcm.h"

namespace {

const float kPi    = 3.14159265358979f;   // src/math.hxx:29
const float kInvPi = 1.f / kPi;

struct V3 { float x, y, z; };
inline V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
inline V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
// Dot accumulates from zero like src/math.hxx:129-130
inline float dot(V3 a, V3 b) { float r = 0.f; r += a.x * b.x; r += a.y * b.y; r += a.z * b.z; return r; }
inline V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline V3 normalize(V3 a) { return a / std::sqrt(dot(a, a)); }
inline void put(float *d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

// Frame::SetFromZ, src/frame.hxx:53-59
void frameFromZ(V3 z, V3 &ox, V3 &oy, V3 &oz)
{
    const V3 tz = oz = normalize(z);
    const V3 tx = (std::abs(tz.x) > 0.99f) ? mk(0, 1, 0) : mk(1, 0, 0);
    oy = normalize(cross(tz, tx));
    ox = cross(oy, tz);
}

// Mat4f, column major: element (r,c) at m[r + 4c] (src/math.hxx:163-274)
struct M4 { float m[16]; float &at(int r, int c) { return m[r + 4 * c]; } float at(int r, int c) const { return m[r + 4 * c]; } };
M4 zero4() { M4 r; for(int i = 0; i < 16; i++) r.m[i] = 0.f; return r; }
M4 identity4() { M4 r = zero4(); for(int i = 0; i < 4; i++) r.at(i, i) = 1.f; return r; }
M4 mul4(const M4 &l, const M4 &r)                       // operator*, src/math.hxx:276-286
{
    M4 res = zero4();
    for(int row = 0; row < 4; row++)
        for(int col = 0; col < 4; col++)
            for(int i = 0; i < 4; i++)
                res.at(row, col) += l.at(row, i) * r.at(i, col);
    return res;
}
M4 scale4(V3 s) { M4 r = identity4(); r.at(0, 0) = s.x; r.at(1, 1) = s.y; r.at(2, 2) = s.z; r.at(3, 3) = 1; return r; }
M4 translate4(V3 t) { M4 r = identity4(); r.at(0, 3) = t.x; r.at(1, 3) = t.y; r.at(2, 3) = t.z; r.at(3, 3) = 1; return r; }
M4 perspective4(float fov, float nearZ, float farZ)     // src/math.hxx:248-265
{
    const float f = 1.f / (std::tan(fov * kPi / 360.0f));
    const float d = 1.f / (nearZ - farZ);
    M4 r = zero4();
    r.at(0, 0) = f;
    r.at(1, 1) = -f;
    r.at(2, 2) = (nearZ + farZ) * d;  r.at(2, 3) = 2.0f * nearZ * farZ * d;
    r.at(3, 2) = -1.0f;
    return r;
}
// Invert, src/math.hxx:290-419 — the classic cofactor expansion. Each inv[] entry is a sum
// of six triple products; the helper keeps the reference's left-to-right evaluation
// ((((t0 + t1) + t2) + t3) + t4) + t5 with the signs folded into the terms.
inline float six(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2,
                 float d0, float d1, float d2, float e0, float e1, float e2, float f0, float f1, float f2,
                 int s0, int s1, int s2, int s3, int s4, int s5)
{
    float r = (s0 > 0) ? (a0 * a1 * a2) : (-a0 * a1 * a2);
    r = (s1 > 0) ? r + b0 * b1 * b2 : r - b0 * b1 * b2;
    r = (s2 > 0) ? r + c0 * c1 * c2 : r - c0 * c1 * c2;
    r = (s3 > 0) ? r + d0 * d1 * d2 : r - d0 * d1 * d2;
    r = (s4 > 0) ? r + e0 * e1 * e2 : r - e0 * e1 * e2;
    r = (s5 > 0) ? r + f0 * f1 * f2 : r - f0 * f1 * f2;
    return r;
}
M4 invert4(const M4 &a)
{
    const float *m = a.m;
    float inv[16];
    inv[0]  = six(m[5],m[10],m[15],  m[5],m[11],m[14],  m[9],m[6],m[15],  m[9],m[7],m[14],  m[13],m[6],m[11],  m[13],m[7],m[10],  +1,-1,-1,+1,+1,-1);
    inv[4]  = six(m[4],m[10],m[15],  m[4],m[11],m[14],  m[8],m[6],m[15],  m[8],m[7],m[14],  m[12],m[6],m[11],  m[12],m[7],m[10],  -1,+1,+1,-1,-1,+1);
    inv[8]  = six(m[4],m[9],m[15],   m[4],m[11],m[13],  m[8],m[5],m[15],  m[8],m[7],m[13],  m[12],m[5],m[11],  m[12],m[7],m[9],   +1,-1,-1,+1,+1,-1);
    inv[12] = six(m[4],m[9],m[14],   m[4],m[10],m[13],  m[8],m[5],m[14],  m[8],m[6],m[13],  m[12],m[5],m[10],  m[12],m[6],m[9],   -1,+1,+1,-1,-1,+1);
    inv[1]  = six(m[1],m[10],m[15],  m[1],m[11],m[14],  m[9],m[2],m[15],  m[9],m[3],m[14],  m[13],m[2],m[11],  m[13],m[3],m[10],  -1,+1,+1,-1,-1,+1);
    inv[5]  = six(m[0],m[10],m[15],  m[0],m[11],m[14],  m[8],m[2],m[15],  m[8],m[3],m[14],  m[12],m[2],m[11],  m[12],m[3],m[10],  +1,-1,-1,+1,+1,-1);
    inv[9]  = six(m[0],m[9],m[15],   m[0],m[11],m[13],  m[8],m[1],m[15],  m[8],m[3],m[13],  m[12],m[1],m[11],  m[12],m[3],m[9],   -1,+1,+1,-1,-1,+1);
    inv[13] = six(m[0],m[9],m[14],   m[0],m[10],m[13],  m[8],m[1],m[14],  m[8],m[2],m[13],  m[12],m[1],m[10],  m[12],m[2],m[9],   +1,-1,-1,+1,+1,-1);
    inv[2]  = six(m[1],m[6],m[15],   m[1],m[7],m[14],   m[5],m[2],m[15],  m[5],m[3],m[14],  m[13],m[2],m[7],   m[13],m[3],m[6],   +1,-1,-1,+1,+1,-1);
    inv[6]  = six(m[0],m[6],m[15],   m[0],m[7],m[14],   m[4],m[2],m[15],  m[4],m[3],m[14],  m[12],m[2],m[7],   m[12],m[3],m[6],   -1,+1,+1,-1,-1,+1);
    inv[10] = six(m[0],m[5],m[15],   m[0],m[7],m[13],   m[4],m[1],m[15],  m[4],m[3],m[13],  m[12],m[1],m[7],   m[12],m[3],m[5],   +1,-1,-1,+1,+1,-1);
    inv[14] = six(m[0],m[5],m[14],   m[0],m[6],m[13],   m[4],m[1],m[14],  m[4],m[2],m[13],  m[12],m[1],m[6],   m[12],m[2],m[5],   -1,+1,+1,-1,-1,+1);
    inv[3]  = six(m[1],m[6],m[11],   m[1],m[7],m[10],   m[5],m[2],m[11],  m[5],m[3],m[10],  m[9],m[2],m[7],    m[9],m[3],m[6],    -1,+1,+1,-1,-1,+1);
    inv[7]  = six(m[0],m[6],m[11],   m[0],m[7],m[10],   m[4],m[2],m[11],  m[4],m[3],m[10],  m[8],m[2],m[7],    m[8],m[3],m[6],    +1,-1,-1,+1,+1,-1);
    inv[11] = six(m[0],m[5],m[11],   m[0],m[7],m[9],    m[4],m[1],m[11],  m[4],m[3],m[9],   m[8],m[1],m[7],    m[8],m[3],m[5],    -1,+1,+1,-1,-1,+1);
    inv[15] = six(m[0],m[5],m[10],   m[0],m[6],m[9],    m[4],m[1],m[10],  m[4],m[2],m[9],   m[8],m[1],m[6],    m[8],m[2],m[5],    +1,-1,-1,+1,+1,-1);

    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if(det == 0) return identity4();
    det = 1.f / det;
    M4 res;
    for(int i = 0; i < 16; i++) res.m[i] = inv[i] * det;
    return res;
}

struct Builder
{
    std::vector<SvcmPrimitive> prims;
    std::vector<SvcmMaterial>  mats;
    std::vector<SvcmLight>     lights;
    std::vector<int32_t>       mat2light;
    SvcmSceneDesc              desc;

    void triangle(V3 p0, V3 p1, V3 p2, int mat)          // Triangle::Triangle, src/geometry.hxx:109-123
    {
        SvcmPrimitive p; memset(&p, 0, sizeof(p));
        p.type = SVCM_PRIM_TRIANGLE; p.matID = mat;
        put(p.v, p0); put(p.v + 3, p1); put(p.v + 6, p2);
        put(p.v + 9, normalize(cross(p1 - p0, p2 - p0)));
        prims.push_back(p);
    }
    void sphere(V3 c, float r, int mat)
    {
        SvcmPrimitive p; memset(&p, 0, sizeof(p));
        p.type = SVCM_PRIM_SPHERE; p.matID = mat;
        put(p.v, c); p.v[3] = r;
        prims.push_back(p);
    }
    void material(V3 diffuse, V3 phong, float exponent, V3 mirror, float ior)
    {
        SvcmMaterial m;
        put(m.diffuse, diffuse); put(m.phong, phong); m.phongExponent = exponent;
        put(m.mirror, mirror); m.ior = ior;
        mats.push_back(m);
    }
    void areaLight(V3 p0, V3 p1, V3 p2, float intensity)  // AreaLight::AreaLight, src/lights.hxx:116-129
    {
        SvcmLight l; memset(&l, 0, sizeof(l));
        l.type = SVCM_LIGHT_AREA;
        const V3 e1 = p1 - p0, e2 = p2 - p0;
        const V3 normal = cross(e1, e2);
        const float len = std::sqrt(dot(normal, normal));
        l.invArea = 2.f / len;
        V3 fx, fy, fz; frameFromZ(normal, fx, fy, fz);
        put(l.p0, p0); put(l.e1, e1); put(l.e2, e2);
        put(l.frameX, fx); put(l.frameY, fy); put(l.frameZ, fz);
        put(l.intensity, mk(intensity, intensity, intensity));
        lights.push_back(l);
    }
};

} // namespace

struct svcm_scene { Builder b; };

extern "C" {

// g_SceneConfigs, reference src/config.hxx:146-151
uint32_t svcm_scene_config_mask(int32_t sceneID)
{
    static const uint32_t configs[4] = {
        SVCM_BOX_GLOSSY_FLOOR | SVCM_BOX_SMALL_MIRROR_SPHERE | SVCM_BOX_SMALL_GLASS_SPHERE | SVCM_BOX_LIGHT_SUN,
        SVCM_BOX_GLOSSY_FLOOR | SVCM_BOX_LARGE_MIRROR_SPHERE | SVCM_BOX_LIGHT_CEILING,
        SVCM_BOX_GLOSSY_FLOOR | SVCM_BOX_SMALL_MIRROR_SPHERE | SVCM_BOX_SMALL_GLASS_SPHERE | SVCM_BOX_LIGHT_POINT,
        SVCM_BOX_GLOSSY_FLOOR | SVCM_BOX_SMALL_MIRROR_SPHERE | SVCM_BOX_SMALL_GLASS_SPHERE | SVCM_BOX_LIGHT_BACKGROUND,
    };
    if(sceneID < 0 || sceneID >= 4) return 0u;
    return configs[sceneID];
}

int svcm_scene_cornell_box(uint32_t mask, int32_t width, int32_t height, svcm_scene **out,
                           const SvcmSceneDesc **descOut)
{
    if(!out || !descOut || width <= 0 || height <= 0) return SVCM_E_INVALID;
    svcm_scene *s = new svcm_scene;
    Builder &b = s->b;

    // src/scene.hxx:137-141: both large spheres requested -> keep the mirror one
    if((mask & (SVCM_BOX_LARGE_MIRROR_SPHERE | SVCM_BOX_LARGE_GLASS_SPHERE)) ==
       (SVCM_BOX_LARGE_MIRROR_SPHERE | SVCM_BOX_LARGE_GLASS_SPHERE))
        mask &= ~uint32_t(SVCM_BOX_LARGE_GLASS_SPHERE);

    const bool lightCeiling    = (mask & SVCM_BOX_LIGHT_CEILING) != 0;
    const bool lightSun        = (mask & SVCM_BOX_LIGHT_SUN) != 0;
    const bool lightPoint      = (mask & SVCM_BOX_LIGHT_POINT) != 0;
    const bool lightBackground = (mask & SVCM_BOX_LIGHT_BACKGROUND) != 0;
    const bool lightBox        = !lightPoint;                                  // :148-152

    // ---- materials, src/scene.hxx:162-206 ------------------------------------------------
    const V3 z3 = mk(0, 0, 0);
    b.material(z3, z3, 1.f, z3, -1.f);                                         // 0 light1
    b.material(z3, z3, 1.f, z3, -1.f);                                         // 1 light2
    b.material(mk(0.1f, 0.1f, 0.1f), mk(0.7f, 0.7f, 0.7f), 90.f, z3, -1.f);    // 2 glossy floor
    b.material(mk(0.156863f, 0.803922f, 0.172549f), z3, 1.f, z3, -1.f);        // 3 green wall
    b.material(mk(0.803922f, 0.152941f, 0.152941f), z3, 1.f, z3, -1.f);        // 4 red wall
    b.material(mk(0.803922f, 0.803922f, 0.803922f), z3, 1.f, z3, -1.f);        // 5 white
    b.material(z3, z3, 1.f, mk(1.f, 1.f, 1.f), -1.f);                          // 6 mirror
    b.material(z3, z3, 1.f, mk(1.f, 1.f, 1.f), 1.6f);                          // 7 glass
    b.material(mk(0.156863f, 0.172549f, 0.803922f), z3, 1.f, z3, -1.f);        // 8 blue wall

    // ---- geometry, src/scene.hxx:210-325 -------------------------------------------------
    const V3 cb[8] = {
        mk(-1.27029f,  1.30455f, -1.28002f), mk( 1.28975f,  1.30455f, -1.28002f),
        mk( 1.28975f,  1.30455f,  1.28002f), mk(-1.27029f,  1.30455f,  1.28002f),
        mk(-1.27029f, -1.25549f, -1.28002f), mk( 1.28975f, -1.25549f, -1.28002f),
        mk( 1.28975f, -1.25549f,  1.28002f), mk(-1.27029f, -1.25549f,  1.28002f) };

    const bool glossy = (mask & SVCM_BOX_GLOSSY_FLOOR) != 0;
    const int floorMat = glossy ? 2 : 5, backMat = glossy ? 8 : 5;
    b.triangle(cb[0], cb[4], cb[5], floorMat);                                 // floor
    b.triangle(cb[5], cb[1], cb[0], floorMat);
    b.triangle(cb[0], cb[1], cb[2], backMat);                                  // back wall
    b.triangle(cb[2], cb[3], cb[0], backMat);

    if(lightCeiling && !lightBox) { b.triangle(cb[2], cb[6], cb[7], 0); b.triangle(cb[7], cb[3], cb[2], 1); }
    else                          { b.triangle(cb[2], cb[6], cb[7], 5); b.triangle(cb[7], cb[3], cb[2], 5); }

    b.triangle(cb[3], cb[7], cb[4], 3);                                        // left wall
    b.triangle(cb[4], cb[0], cb[3], 3);
    b.triangle(cb[1], cb[5], cb[6], 4);                                        // right wall
    b.triangle(cb[6], cb[2], cb[1], 4);

    const float largeRadius = 0.8f;
    const V3 center = (cb[0] + cb[1] + cb[4] + cb[5]) * (1.f / 4.f) + mk(0, 0, largeRadius);
    if(mask & SVCM_BOX_LARGE_MIRROR_SPHERE) b.sphere(center, largeRadius, 6);
    if(mask & SVCM_BOX_LARGE_GLASS_SPHERE)  b.sphere(center, largeRadius, 7);

    const float smallRadius = 0.5f;
    const V3 leftWallCenter  = (cb[0] + cb[4]) * (1.f / 2.f) + mk(0, 0, smallRadius);
    const V3 rightWallCenter = (cb[1] + cb[5]) * (1.f / 2.f) + mk(0, 0, smallRadius);
    const float xlen = rightWallCenter.x - leftWallCenter.x;
    const V3 leftBallCenter  = leftWallCenter  + mk(2.f * xlen / 7.f, 0, 0);
    const V3 rightBallCenter = rightWallCenter - mk(2.f * xlen / 7.f, 0, 0);
    if(mask & SVCM_BOX_SMALL_MIRROR_SPHERE) b.sphere(leftBallCenter,  smallRadius, 6);
    if(mask & SVCM_BOX_SMALL_GLASS_SPHERE)  b.sphere(rightBallCenter, smallRadius, 7);

    const V3 lb[8] = {
        mk(-0.25f,  0.25f, 1.26002f), mk( 0.25f,  0.25f, 1.26002f), mk( 0.25f,  0.25f, 1.28002f),
        mk(-0.25f,  0.25f, 1.28002f), mk(-0.25f, -0.25f, 1.26002f), mk( 0.25f, -0.25f, 1.26002f),
        mk( 0.25f, -0.25f, 1.28002f), mk(-0.25f, -0.25f, 1.28002f) };
    if(lightBox)
    {
        b.triangle(lb[0], lb[2], lb[1], 5); b.triangle(lb[2], lb[0], lb[3], 5);   // back
        b.triangle(lb[3], lb[4], lb[7], 5); b.triangle(lb[4], lb[3], lb[0], 5);   // left
        b.triangle(lb[1], lb[6], lb[5], 5); b.triangle(lb[6], lb[1], lb[2], 5);   // right
        b.triangle(lb[4], lb[5], lb[6], 5); b.triangle(lb[6], lb[7], lb[4], 5);   // front
        if(lightCeiling) { b.triangle(lb[0], lb[5], lb[4], 0); b.triangle(lb[5], lb[0], lb[1], 1); }
        else             { b.triangle(lb[0], lb[5], lb[4], 5); b.triangle(lb[5], lb[0], lb[1], 5); }
    }

    // ---- lights, src/scene.hxx:327-384 ---------------------------------------------------
    b.mat2light.assign(b.mats.size(), -1);
    int background = -1;
    if(lightCeiling && !lightBox)
    {
        b.areaLight(cb[2], cb[6], cb[7], 0.95492965f); b.mat2light[0] = 0;
        b.areaLight(cb[7], cb[3], cb[2], 0.95492965f); b.mat2light[1] = 1;
    }
    else if(lightCeiling && lightBox)
    {
        b.areaLight(lb[0], lb[5], lb[4], 25.03329895614464f); b.mat2light[0] = 0;
        b.areaLight(lb[5], lb[0], lb[1], 25.03329895614464f); b.mat2light[1] = 1;
    }
    if(lightSun)
    {
        SvcmLight l; memset(&l, 0, sizeof(l));
        l.type = SVCM_LIGHT_DIRECTIONAL;
        V3 fx, fy, fz; frameFromZ(mk(-1.f, 1.5f, -1.f), fx, fy, fz);
        put(l.frameX, fx); put(l.frameY, fy); put(l.frameZ, fz);
        put(l.intensity, mk(0.5f, 0.2f, 0.f) * 20.f);
        b.lights.push_back(l);
    }
    if(lightPoint)
    {
        SvcmLight l; memset(&l, 0, sizeof(l));
        l.type = SVCM_LIGHT_POINT;
        put(l.p0, mk(0.0f, -0.5f, 1.0f));
        const float i = 70.f * (kInvPi * 0.25f);
        put(l.intensity, mk(i, i, i));
        b.lights.push_back(l);
    }
    if(lightBackground)
    {
        SvcmLight l; memset(&l, 0, sizeof(l));
        l.type = SVCM_LIGHT_BACKGROUND;
        put(l.intensity, mk(135.f / 255.f, 206.f / 255.f, 250.f / 255.f));     // src/lights.hxx:406
        l.scale = 1.f;
        background = int(b.lights.size());
        b.lights.push_back(l);
    }

    SvcmSceneDesc &d = b.desc;
    memset(&d, 0, sizeof(d));
    d.primitives = b.prims.data();       d.numPrimitives = int32_t(b.prims.size());
    d.materials = b.mats.data();         d.numMaterials = int32_t(b.mats.size());
    d.lights = b.lights.data();          d.numLights = int32_t(b.lights.size());
    d.materialToLight = b.mat2light.data();
    d.backgroundLight = background;

    // ---- camera: Camera::Setup, src/camera.hxx:36-80, called at src/scene.hxx:155-159 ----
    {
        const V3 aPosition = mk(-0.0439815f, -4.12529f, 0.222539f);
        const V3 aForward  = mk(0.00688625f, 0.998505f, -0.0542161f);
        const V3 aUp       = mk(3.73896e-4f, 0.0542148f, 0.998529f);
        const float resX = float(width), resY = float(height), fov = 45.f;

        const V3 forward = normalize(aForward);
        const V3 up      = normalize(cross(aUp, -forward));
        const V3 left    = cross(-forward, up);
        const V3 pos = mk(dot(up, aPosition), dot(left, aPosition), dot(-forward, aPosition));

        M4 worldToCamera = identity4();
        const V3 rows[3] = { up, left, -forward };
        const float negPos[3] = { -pos.x, -pos.y, -pos.z };
        for(int r = 0; r < 3; r++)
        {
            worldToCamera.at(r, 0) = rows[r].x; worldToCamera.at(r, 1) = rows[r].y;
            worldToCamera.at(r, 2) = rows[r].z; worldToCamera.at(r, 3) = negPos[r];
        }
        const M4 perspective    = perspective4(fov, 0.1f, 10000.f);
        const M4 worldToNScreen = mul4(perspective, worldToCamera);
        const M4 nscreenToWorld = invert4(worldToNScreen);

        const M4 worldToRaster = mul4(mul4(scale4(mk(resX * 0.5f, resY * 0.5f, 0)), translate4(mk(1.f, 1.f, 0))),
                                      worldToNScreen);
        const M4 rasterToWorld = mul4(mul4(nscreenToWorld, translate4(mk(-1.f, -1.f, 0))),
                                      scale4(mk(2.f / resX, 2.f / resY, 0)));

        put(d.camera.position, aPosition);
        put(d.camera.forward, forward);
        d.camera.resolution[0] = resX; d.camera.resolution[1] = resY;
        memcpy(d.camera.rasterToWorld, rasterToWorld.m, sizeof(float) * 16);
        memcpy(d.camera.worldToRaster, worldToRaster.m, sizeof(float) * 16);
        const float tanHalfAngle = std::tan(fov * kPi / 360.f);
        d.camera.imagePlaneDist = resX / (2.f * tanHalfAngle);
    }

    // ---- Scene::BuildSceneSphere, src/scene.hxx:387-398 (GrowBBox: src/geometry.hxx:158-170,239-259)
    {
        V3 mn = mk(1e36f, 1e36f, 1e36f), mx = mk(-1e36f, -1e36f, -1e36f);
        for(size_t i = 0; i < b.prims.size(); i++)
        {
            const SvcmPrimitive &p = b.prims[i];
            if(p.type == SVCM_PRIM_TRIANGLE)
            {
                for(int k = 0; k < 3; k++)
                {
                    mn.x = std::min(mn.x, p.v[3 * k]);     mx.x = std::max(mx.x, p.v[3 * k]);
                    mn.y = std::min(mn.y, p.v[3 * k + 1]); mx.y = std::max(mx.y, p.v[3 * k + 1]);
                    mn.z = std::min(mn.z, p.v[3 * k + 2]); mx.z = std::max(mx.z, p.v[3 * k + 2]);
                }
            }
            else
            {
                for(int k = 0; k < 8; k++)
                {
                    const float hx = (k & 1) ? -p.v[3] : p.v[3], hy = (k & 2) ? -p.v[3] : p.v[3],
                                hz = (k & 4) ? -p.v[3] : p.v[3];
                    const V3 c = mk(p.v[0] + hx, p.v[1] + hy, p.v[2] + hz);
                    mn.x = std::min(mn.x, c.x); mx.x = std::max(mx.x, c.x);
                    mn.y = std::min(mn.y, c.y); mx.y = std::max(mx.y, c.y);
                    mn.z = std::min(mn.z, c.z); mx.z = std::max(mx.z, c.z);
                }
            }
        }
        const V3 diag = mx - mn;
        const float radius2 = dot(diag, diag);
        put(d.sceneSphere.center, (mx + mn) * 0.5f);
        d.sceneSphere.radius = std::sqrt(radius2) * 0.5f;
        d.sceneSphere.invRadiusSqr = 1.f / (d.sceneSphere.radius * d.sceneSphere.radius);
    }

    *out = s;
    *descOut = &b.desc;
    return SVCM_OK;
}

void svcm_scene_free(svcm_scene *s) { delete s; }

} // extern "C"
