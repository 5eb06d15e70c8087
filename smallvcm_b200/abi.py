"""ctypes mirror of include/svcm.h (structs, enums and function prototypes).

Nothing here computes anything: it only describes the C-ABI of libsvcm.so so that the
Python host layer (renderer.py), the tests and bench.py can call it.  Field order and
types follow include/svcm.h one to one; tests/test_abi.py checks the struct sizes against
the sizes the library itself reports.
"""
import ctypes as C

# --- limits / enums (include/svcm.h) -------------------------------------------------------
PRIM_TRIANGLE, PRIM_SPHERE = 0, 1
LIGHT_AREA, LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_BACKGROUND = 0, 1, 2, 3
# VertexCM::AlgorithmType numbering (reference src/vertexcm.hxx:182-204)
IMAGE_BMP24, IMAGE_RGBE = 0, 1   # SVCM_IMAGE_*
MAX_PRIMITIVES, MAX_PRIMITIVES_BVH, MAX_MATERIALS, MAX_LIGHTS = 64, 1 << 24, 16, 8
MAX_RUNS = 8
ALG_LIGHT_TRACE, ALG_PPM, ALG_BPM, ALG_BPT, ALG_VCM, ALG_PATH_TRACE, ALG_EYE_LIGHT = 0, 1, 2, 3, 4, 5, 6
ALGORITHMS = {"lt": ALG_LIGHT_TRACE, "ppm": ALG_PPM, "bpm": ALG_BPM, "bpt": ALG_BPT, "vcm": ALG_VCM,
              "pt": ALG_PATH_TRACE, "el": ALG_EYE_LIGHT}
OPT_PROFILE, OPT_MERGE_GROUP, OPT_COUNTERS = 1, 2, 3
OPT_GRID_LAYOUT, OPT_MERGE_VARIANT, OPT_REF_CANDIDATES, OPT_OVERLAP = 4, 5, 6, 7
GRID_BUCKET, GRID_CELL = 0, 1
COMM_ID_BYTES = 128

OK, E_INVALID, E_NO_DEVICE, E_CUDA, E_NCCL, E_NOMEM, E_STATE = 0, -1, -2, -3, -4, -5, -6

f32, i32, u32, i64 = C.c_float, C.c_int32, C.c_uint32, C.c_int64


class Primitive(C.Structure):
    _fields_ = [("type", i32), ("matID", i32), ("v", f32 * 12)]


class Material(C.Structure):
    _fields_ = [("diffuse", f32 * 3), ("phong", f32 * 3), ("phongExponent", f32),
                ("mirror", f32 * 3), ("ior", f32)]


class Light(C.Structure):
    _fields_ = [("type", i32), ("p0", f32 * 3), ("e1", f32 * 3), ("e2", f32 * 3),
                ("frameX", f32 * 3), ("frameY", f32 * 3), ("frameZ", f32 * 3),
                ("intensity", f32 * 3), ("invArea", f32), ("scale", f32)]


class Camera(C.Structure):
    _fields_ = [("position", f32 * 3), ("forward", f32 * 3), ("resolution", f32 * 2),
                ("rasterToWorld", f32 * 16), ("worldToRaster", f32 * 16),
                ("imagePlaneDist", f32)]


class SceneSphere(C.Structure):
    _fields_ = [("center", f32 * 3), ("radius", f32), ("invRadiusSqr", f32)]


class BvhNode(C.Structure):
    _fields_ = [("lo", f32 * 3), ("first", i32), ("hi", f32 * 3), ("count", i32)]


class SceneDesc(C.Structure):
    _fields_ = [("primitives", C.POINTER(Primitive)), ("numPrimitives", i32),
                ("materials", C.POINTER(Material)), ("numMaterials", i32),
                ("lights", C.POINTER(Light)), ("numLights", i32),
                ("materialToLight", C.POINTER(i32)),
                ("backgroundLight", i32),
                ("camera", Camera),
                ("sceneSphere", SceneSphere)]


class Params(C.Structure):
    _fields_ = [("algorithm", i32), ("radiusFactor", f32), ("radiusAlpha", f32),
                ("seed", i32), ("device", i32), ("rank", i32), ("worldSize", i32),
                ("reserved", i32 * 5)]


class Stats(C.Structure):
    _fields_ = [("paths", i64), ("localPaths", i64), ("lightVertices", i64),
                ("localLightVertices", i64), ("mergeQueries", i64),
                ("mergeCandidates", i64), ("mergeAccepted", i64), ("cameraRays", i64),
                ("lightRays", i64), ("shadowRays", i64), ("radius", f32),
                ("msLightTrace", f32), ("msGather", f32), ("msGridBuild", f32),
                ("msCameraTrace", f32), ("msMerge", f32), ("msTotal", f32),
                ("kernelLaunches", i32), ("paramBytesPerLaunch", i32)]


class Profile(C.Structure):
    _fields_ = [("iterations", i32), ("timedIterations", i32), ("lightVertices", i64),
                ("mergeQueries", i64), ("mergeCandidates", i64), ("mergeAccepted", i64),
                ("lightRays", i64), ("cameraRays", i64), ("shadowRays", i64),
                ("kernelLaunches", i64), ("msLightTrace", C.c_double), ("msGather", C.c_double),
                ("msGridBuild", C.c_double), ("msCameraTrace", C.c_double),
                ("msMerge", C.c_double), ("msTotal", C.c_double)]


class MergeVertex(C.Structure):
    _fields_ = [("position", f32 * 3), ("throughput", f32 * 3), ("dirFix", f32 * 3),
                ("pathLength", u32), ("dVCM", f32), ("dVM", f32), ("contProb", f32)]


class MergeQuery(C.Structure):
    _fields_ = [("position", f32 * 3), ("normal", f32 * 3), ("dirFix", f32 * 3),
                ("matID", i32), ("pathLength", u32), ("dVCM", f32), ("dVM", f32)]


import numpy as np  # noqa: E402  (numpy views of the two record types)

MERGE_VERTEX_DTYPE = np.dtype([("position", "<f4", 3), ("throughput", "<f4", 3),
                               ("dirFix", "<f4", 3), ("pathLength", "<u4"),
                               ("dVCM", "<f4"), ("dVM", "<f4"), ("contProb", "<f4")])
MERGE_QUERY_DTYPE = np.dtype([("position", "<f4", 3), ("normal", "<f4", 3),
                              ("dirFix", "<f4", 3), ("matID", "<i4"),
                              ("pathLength", "<u4"), ("dVCM", "<f4"), ("dVM", "<f4")])
assert MERGE_VERTEX_DTYPE.itemsize == C.sizeof(MergeVertex) == 52
assert MERGE_QUERY_DTYPE.itemsize == C.sizeof(MergeQuery) == 52

VP = C.c_void_p
PSD = C.POINTER(SceneDesc)

# name -> (restype, argtypes): every symbol include/svcm.h declares
PROTOTYPES = {
    "svcm_create": (i32, [PSD, C.POINTER(Params), C.POINTER(VP)]),
    "svcm_destroy": (None, [VP]),
    "svcm_last_error": (C.c_char_p, [VP]),
    "svcm_set_option": (i32, [VP, i32, i64]),
    "svcm_run_iterations": (i32, [VP, i32, i32, u32, u32]),
    "svcm_synchronize": (i32, [VP]),
    "svcm_download_accum": (i32, [VP, VP, C.POINTER(i32)]),
    "svcm_clear_accum": (i32, [VP]),
    "svcm_image_bytes": (i64, [i32, i32, i32]),
    "svcm_image_header": (i32, [i32, i32, i32, VP, i32]),
    "svcm_download_image": (i32, [VP, i32, f32, VP, C.POINTER(i32)]),
    "svcm_save_image": (i32, [VP, C.c_char_p, f32]),
    "svcm_bvh_build": (i32, [VP, i32, VP, i32, VP, C.POINTER(i32)]),
    "svcm_primitive_runs": (i32, [VP, i32, VP, C.POINTER(u32), C.POINTER(i32)]),
    "svcm_image_convert_batch": (i32, [VP, i32, i32, i32, i32, f32, i32, VP]),
    "svcm_get_stats": (i32, [VP, C.POINTER(Stats)]),
    "svcm_get_profile": (i32, [VP, C.POINTER(Profile)]),
    "svcm_host_register": (i32, [VP, C.c_size_t]),
    "svcm_host_unregister": (i32, [VP]),
    "svcm_device_count": (i32, []),
    "svcm_event_record": (i32, [VP, i32]),
    "svcm_event_elapsed_ms": (i32, [VP, i32, i32, C.POINTER(f32)]),
    "svcm_shard_local_paths": (i64, [i32, i32, i32, i32]),
    "svcm_shard_owner": (i32, [i32, i32, i64]),
    "svcm_shard_local_index": (i64, [i32, i32, i64]),
    "svcm_shard_global_index": (i64, [i32, i32, i32, i64]),
    "svcm_comm_unique_id": (i32, [VP]),
    "svcm_comm_init": (i32, [VP, VP]),
    "svcm_reduce_accum": (i32, [VP, i32]),
    "svcm_hashgrid_build": (i32, [VP, i64, f32, i32, i32, C.POINTER(VP)]),
    "svcm_hashgrid_destroy": (None, [VP]),
    "svcm_hashgrid_download": (i32, [VP, VP, VP, VP]),
    "svcm_hashgrid_query": (i32, [VP, PSD, VP, i64, u32, u32, f32, i32, VP, VP]),
    "svcm_intersect_batch": (i32, [PSD, VP, i64, i32, VP, VP, VP, VP, VP]),
    "svcm_occluded_batch": (i32, [PSD, VP, i64, i32, VP]),
    "svcm_bsdf_evaluate_batch": (i32, [PSD, VP, i64, i32, i32, VP]),
    "svcm_bsdf_sample_batch": (i32, [PSD, VP, i64, i32, i32, VP, VP]),
    "svcm_rng_stream": (i32, [i32, i32, i32, u32, i32, i32, VP]),
    "svcm_rng_stream_device": (i32, [i32, i32, i32, u32, i32, i32, i32, VP]),
    "svcm_light_emit_batch": (i32, [PSD, i32, VP, i64, i32, VP]),
    "svcm_light_illuminate_batch": (i32, [PSD, i32, VP, i64, i32, VP]),
    "svcm_light_radiance_batch": (i32, [PSD, i32, VP, i64, i32, VP]),
    "svcm_camera_ray_batch": (i32, [PSD, VP, i64, i32, VP]),
    "svcm_scene_config_mask": (u32, [i32]),
    "svcm_scene_cornell_box": (i32, [u32, i32, i32, C.POINTER(VP), C.POINTER(PSD)]),
    "svcm_scene_free": (None, [VP]),
    "svcm_version": (C.c_char_p, []),
}


def bind(lib):
    """Attach restype/argtypes to every declared symbol; raises AttributeError if the
    library does not export one of them."""
    import os
    tolerant = os.environ.get("SVCM_ALLOW_OLD_LIBRARY") == "1"    # A/B runs against libraries of earlier commits
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if tolerant:
                continue
            raise
        fn.restype = res
        fn.argtypes = args
    return lib
