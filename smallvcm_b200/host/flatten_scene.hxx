// flatten_scene.hxx — host glue, reference side of the C-ABI.
//
// Turns the reference's pointer-graph `Scene` (reference src/scene.hxx:474-485) into the
// flat POD arrays of include/svcm.h.  Header-only and WITHOUT include guards on the
// reference side: it must be included AFTER the reference headers (scene.hxx, geometry.hxx,
// lights.hxx, camera.hxx, materials.hxx) inside the single translation unit those headers
// allow (their functions are non-inline).  Nothing here is compiled into
// libsvcm.so; the library never sees a reference type.
#ifndef SVCM_HOST_FLATTEN_SCENE_HXX_
#define SVCM_HOST_FLATTEN_SCENE_HXX_

#include <vector>
#include <cstring>
#include <cstdio>
#include "svcm.h"

namespace svcm_host {

struct FlatScene
{
    std::vector<SvcmPrimitive> primitives;
    std::vector<SvcmMaterial>  materials;
    std::vector<SvcmLight>     lights;
    std::vector<int32_t>       materialToLight;
    SvcmSceneDesc              desc;
};

inline void Put3(float *dst, const Vec3f &v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

// Returns false (after printing why) if the scene holds something the flat form cannot
// express — an unknown geometry or light class, or more items than the compiled-in limits.
inline bool FlattenScene(const Scene &aScene, FlatScene &oFlat)
{
    oFlat.primitives.clear();
    oFlat.materials.clear();
    oFlat.lights.clear();
    oFlat.materialToLight.clear();
    memset(&oFlat.desc, 0, sizeof(oFlat.desc));

    // Geometry: the reference's scenes hold one GeometryList (src/scene.hxx:221-222) of
    // Triangle / Sphere; the list order is kept because exact-tie hits resolve to the
    // first primitive in the list (src/geometry.hxx:146,226).
    const GeometryList *list = dynamic_cast<const GeometryList*>(aScene.mGeometry);
    if(list == NULL)
    {
        printf("svcm: scene geometry is not a GeometryList\n");
        return false;
    }

    for(size_t i = 0; i < list->mGeometry.size(); i++)
    {
        SvcmPrimitive prim;
        memset(&prim, 0, sizeof(prim));

        if(const Triangle *tri = dynamic_cast<const Triangle*>(list->mGeometry[i]))
        {
            prim.type  = SVCM_PRIM_TRIANGLE;
            prim.matID = tri->matID;
            Put3(prim.v + 0, tri->p[0]);
            Put3(prim.v + 3, tri->p[1]);
            Put3(prim.v + 6, tri->p[2]);
            Put3(prim.v + 9, tri->mNormal);
        }
        else if(const Sphere *sph = dynamic_cast<const Sphere*>(list->mGeometry[i]))
        {
            prim.type  = SVCM_PRIM_SPHERE;
            prim.matID = sph->matID;
            Put3(prim.v + 0, sph->center);
            prim.v[3] = sph->radius;
        }
        else
        {
            printf("svcm: primitive %d is neither Triangle nor Sphere\n", int(i));
            return false;
        }

        oFlat.primitives.push_back(prim);
    }

    // Materials: same fields, same order as src/materials.hxx:36-66.
    for(int i = 0; i < aScene.GetMaterialCount(); i++)
    {
        const Material &src = aScene.GetMaterial(i);
        SvcmMaterial mat;
        Put3(mat.diffuse, src.mDiffuseReflectance);
        Put3(mat.phong,   src.mPhongReflectance);
        mat.phongExponent = src.mPhongExponent;
        Put3(mat.mirror,  src.mMirrorReflectance);
        mat.ior = src.mIOR;
        oFlat.materials.push_back(mat);
    }

    // material -> light: densify the std::map of src/scene.hxx:61-66
    oFlat.materialToLight.assign(oFlat.materials.size(), -1);
    for(std::map<int, int>::const_iterator it = aScene.mMaterial2Light.begin();
        it != aScene.mMaterial2Light.end(); ++it)
    {
        if(it->first >= 0 && it->first < (int)oFlat.materialToLight.size())
            oFlat.materialToLight[it->first] = it->second;
    }

    // Lights: public fields of the four classes (src/lights.hxx:229-232,314-315,
    // 395-396,512-513)
    int background = -1;
    for(size_t i = 0; i < aScene.mLights.size(); i++)
    {
        SvcmLight light;
        memset(&light, 0, sizeof(light));
        const AbstractLight *src = aScene.mLights[i];

        if(const AreaLight *a = dynamic_cast<const AreaLight*>(src))
        {
            light.type = SVCM_LIGHT_AREA;
            Put3(light.p0, a->p0);
            Put3(light.e1, a->e1);
            Put3(light.e2, a->e2);
            Put3(light.frameX, a->mFrame.mX);
            Put3(light.frameY, a->mFrame.mY);
            Put3(light.frameZ, a->mFrame.mZ);
            Put3(light.intensity, a->mIntensity);
            light.invArea = a->mInvArea;
        }
        else if(const DirectionalLight *d = dynamic_cast<const DirectionalLight*>(src))
        {
            light.type = SVCM_LIGHT_DIRECTIONAL;
            Put3(light.frameX, d->mFrame.mX);
            Put3(light.frameY, d->mFrame.mY);
            Put3(light.frameZ, d->mFrame.mZ);
            Put3(light.intensity, d->mIntensity);
        }
        else if(const PointLight *p = dynamic_cast<const PointLight*>(src))
        {
            light.type = SVCM_LIGHT_POINT;
            Put3(light.p0, p->mPosition);
            Put3(light.intensity, p->mIntensity);
        }
        else if(const BackgroundLight *b = dynamic_cast<const BackgroundLight*>(src))
        {
            light.type = SVCM_LIGHT_BACKGROUND;
            Put3(light.intensity, b->mBackgroundColor);
            light.scale = b->mScale;
            if(aScene.GetBackground() == b)
                background = int(i);
        }
        else
        {
            printf("svcm: light %d has an unknown class\n", int(i));
            return false;
        }

        oFlat.lights.push_back(light);
    }

    if(oFlat.primitives.size() > SVCM_MAX_PRIMITIVES_BVH ||
       oFlat.materials.size()  > SVCM_MAX_MATERIALS  ||
       oFlat.lights.size()     > SVCM_MAX_LIGHTS)
    {
        printf("svcm: scene exceeds the compiled-in limits (%d primitives, %d materials, %d lights)\n",
            SVCM_MAX_PRIMITIVES_BVH, SVCM_MAX_MATERIALS, SVCM_MAX_LIGHTS);
        return false;
    }

    SvcmSceneDesc &d = oFlat.desc;
    d.primitives      = oFlat.primitives.data();
    d.numPrimitives   = int32_t(oFlat.primitives.size());
    d.materials       = oFlat.materials.data();
    d.numMaterials    = int32_t(oFlat.materials.size());
    d.lights          = oFlat.lights.data();
    d.numLights       = int32_t(oFlat.lights.size());
    d.materialToLight = oFlat.materialToLight.data();
    d.backgroundLight = background;

    // Camera (src/camera.hxx:121-126); Mat4f is 16 floats, column major (math.hxx:270-274)
    const Camera &cam = aScene.mCamera;
    Put3(d.camera.position, cam.mPosition);
    Put3(d.camera.forward,  cam.mForward);
    d.camera.resolution[0] = cam.mResolution.x;
    d.camera.resolution[1] = cam.mResolution.y;
    memcpy(d.camera.rasterToWorld, cam.mRasterToWorld.GetPtr(), sizeof(float) * 16);
    memcpy(d.camera.worldToRaster, cam.mWorldToRaster.GetPtr(), sizeof(float) * 16);
    d.camera.imagePlaneDist = cam.mImagePlaneDist;

    // Scene bounding sphere (src/lights.hxx:32-40, filled by src/scene.hxx:387-398)
    Put3(d.sceneSphere.center, aScene.mSceneSphere.mSceneCenter);
    d.sceneSphere.radius       = aScene.mSceneSphere.mSceneRadius;
    d.sceneSphere.invRadiusSqr = aScene.mSceneSphere.mInvSceneRadiusSqr;

    return true;
}

} // namespace svcm_host

#endif // SVCM_HOST_FLATTEN_SCENE_HXX_
