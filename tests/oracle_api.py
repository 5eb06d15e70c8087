"""ctypes access to the CHECKERS (test infrastructure): oracle/liboracle.so (our CPU
restatement) and, where they were built, oracle/_ref/libsvcm_ref.so (the unmodified
reference) and oracle/_ref/libsvcm_ref_ctr.so (reference arithmetic + per-path RNG).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from smallvcm_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsvcm_ref.so")
REF_CTR_SO = os.path.join(ORACLE_DIR, "_ref", "libsvcm_ref_ctr.so")

VP, i32, u32, f32, i64 = C.c_void_p, C.c_int, C.c_uint, C.c_float, C.c_longlong
PSD = C.POINTER(abi.SceneDesc)


def build_oracle():
    """(Re)build oracle/liboracle.so (and oracle/_ref where the reference checkout exists)."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(VP)


# functions shared (same signature after the handle) by liboracle.so and libsvcm_ref_ctr.so
_COMMON = {
    "destroy": (None, [VP]),
    "run_iteration": (None, [VP, i32]),
    "accum": (i32, [VP, VP]),
    "num_light_vertices": (i64, [VP]),
    "light_vertices": (None, [VP, VP, VP, VP, VP, VP]),
    "iteration_constants": (None, [VP, i32, VP, VP, VP, VP]),
    "grid": (i32, [VP, VP, VP, VP]),
    "intersect": (None, [VP, VP, i64, VP, VP, VP, VP, VP]),
    "occluded": (None, [VP, VP, i64, VP]),
    "bsdf_evaluate": (None, [VP, VP, i64, i32, VP]),
    "bsdf_sample": (None, [VP, VP, i64, i32, VP, VP]),
    "light": (None, [VP, i32, i32, VP, i64, VP]),
    "camera": (None, [VP, VP, i64, VP]),
}


class _Checker:
    """Common Python face of one checker handle (oracle or refctr)."""

    def __init__(self, lib, prefix, handle, width, height, keep=None):
        self._lib, self._p, self._h = lib, prefix, handle
        self.width, self.height = width, height
        self._keep = keep

    def _f(self, name):
        return getattr(self._lib, self._p + name)

    def close(self):
        if self._h:
            self._f("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_iteration(self, iteration, threads=None):
        """threads=None: the serial loop. threads=N (0: all hardware threads): the oracle's
        path-parallel iteration — per-path random streams only, same light vertices and counters,
        image equal up to the order of float additions (tests/test_oracle.py pins that)."""
        if threads is None or self._p != "oracle_":
            self._f("run_iteration")(self._h, iteration)
        else:
            oracle_lib().oracle_run_iteration_parallel(self._h, iteration, int(threads))

    def accum(self):
        out = np.zeros((self.height, self.width, 3), np.float32)
        n = self._f("accum")(self._h, _ptr(out))
        return out, n

    def num_light_vertices(self):
        return int(self._f("num_light_vertices")(self._h))

    def light_vertices(self):
        n = self.num_light_vertices()
        verts = np.zeros(n, abi.MERGE_VERTEX_DTYPE)
        dvc = np.zeros(n, np.float32)
        mat = np.zeros(n, np.int32)
        normal = np.zeros((n, 3), np.float32)
        ends = np.zeros(self.width * self.height, np.int32)
        self._f("light_vertices")(self._h, _ptr(verts), _ptr(dvc), _ptr(mat), _ptr(normal), _ptr(ends))
        return verts, dvc, mat, normal, ends

    def iteration_constants(self, iteration):
        vals = [f32() for _ in range(4)]
        self._f("iteration_constants")(self._h, iteration, *[C.byref(v) for v in vals])
        return tuple(v.value for v in vals)  # radius, vmNormalization, misVm, misVc

    def grid(self, num_cells):
        n = self.num_light_vertices()
        bbox = np.zeros(6, np.float32)
        ends = np.zeros(num_cells, np.int32)
        idx = np.zeros(n, np.int32)
        cells = self._f("grid")(self._h, _ptr(bbox), _ptr(ends), _ptr(idx))
        assert cells == num_cells, (cells, num_cells)
        return bbox, ends, idx

    def merge_queries(self, queries, want_counts=False):
        q = np.ascontiguousarray(queries, abi.MERGE_QUERY_DTYPE)
        contrib = np.zeros((len(q), 3), np.float32)
        if self._p == "oracle_":
            counts = np.zeros((len(q), 2), np.int32) if want_counts else None
            self._lib.oracle_merge_queries(self._h, _ptr(q), len(q), _ptr(contrib), _ptr(counts))
            return (contrib, counts) if want_counts else contrib
        assert not want_counts, "the reference build has no candidate counters"
        self._lib.refctr_merge_queries(self._h, _ptr(q), len(q), _ptr(contrib))
        return contrib

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        hit = np.zeros(n, np.int32); dist = np.zeros(n, np.float32)
        normal = np.zeros((n, 3), np.float32); mat = np.zeros(n, np.int32); light = np.zeros(n, np.int32)
        self._f("intersect")(self._h, _ptr(rays), n, _ptr(hit), _ptr(dist), _ptr(normal), _ptr(mat), _ptr(light))
        return hit, dist, normal, mat, light

    def occluded(self, segments):
        seg = np.ascontiguousarray(segments, np.float32).reshape(-1, 7)
        occ = np.zeros(len(seg), np.int32)
        self._f("occluded")(self._h, _ptr(seg), len(seg), _ptr(occ))
        return occ

    def bsdf_evaluate(self, items, fix_is_light):
        items = np.ascontiguousarray(items, np.float32).reshape(-1, 10)
        out = np.zeros((len(items), 8), np.float32)
        self._f("bsdf_evaluate")(self._h, _ptr(items), len(items), int(fix_is_light), _ptr(out))
        return out

    def bsdf_sample(self, items, fix_is_light):
        items = np.ascontiguousarray(items, np.float32).reshape(-1, 10)
        out = np.zeros((len(items), 8), np.float32)
        ev = np.zeros(len(items), np.uint32)
        self._f("bsdf_sample")(self._h, _ptr(items), len(items), int(fix_is_light), _ptr(out), _ptr(ev))
        return out, ev


_LIGHT_OPS = {"emit": 0, "illuminate": 1, "radiance": 2}


def _light(self, op, light_index, items):
    items = np.ascontiguousarray(items, np.float32).reshape(-1, 8)
    out = np.zeros((len(items), 12), np.float32)
    self._f("light")(self._h, _LIGHT_OPS[op], int(light_index), _ptr(items), len(items), _ptr(out))
    return out


def _camera(self, items):
    items = np.ascontiguousarray(items, np.float32).reshape(-1, 5)
    out = np.zeros((len(items), 6), np.float32)
    self._f("camera")(self._h, _ptr(items), len(items), _ptr(out))
    return out


_Checker.light = _light
_Checker.camera = _camera

_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = C.CDLL(ORACLE_SO)
        for name, (res, args) in _COMMON.items():
            fn = getattr(lib, "oracle_" + name)
            fn.restype, fn.argtypes = res, args
        lib.oracle_create.restype = VP
        lib.oracle_run_iteration_parallel.restype = None
        lib.oracle_run_iteration_parallel.argtypes = [VP, i32, i32]
        lib.oracle_create.argtypes = [PSD, i32, f32, f32, i32, i32, u32, u32]
        lib.oracle_merge_queries.restype = None
        lib.oracle_merge_queries.argtypes = [VP, VP, i64, VP, VP]
        lib.oracle_set_light_vertices.restype = None
        lib.oracle_set_light_vertices.argtypes = [VP, VP, i64, f32, i32, f32, i32, u32, u32]
        lib.oracle_rng_stream.restype = None
        lib.oracle_rng_stream.argtypes = [i32, i32, i32, u32, i32, i32, VP]
        lib.oracle_counters.restype = None
        lib.oracle_counters.argtypes = [VP, VP]
        lib.oracle_total_luminance.restype = f32
        lib.oracle_total_luminance.argtypes = [VP, i32, i32]
        lib.oracle_bmp_payload.restype = None
        lib.oracle_bmp_payload.argtypes = [VP, i32, i32, i32, f32, VP]
        lib.oracle_hdr_payload.restype = None
        lib.oracle_hdr_payload.argtypes = [VP, i32, i32, i32, VP]
        lib.oracle_render.restype = C.c_double
        lib.oracle_render.argtypes = [PSD, i32, i32, i32, i32, i32, u32, u32, f32, f32, VP]
        _oracle = lib
    return _oracle


class Oracle(_Checker):
    """oracle/vcm_oracle.cpp on a flat scene (abi.SceneDesc or smallvcm_b200.Scene)."""

    def __init__(self, scene, algorithm, radius_factor=0.003, radius_alpha=0.75, seed=1234,
                 rng_mode=1, max_path_length=10, min_path_length=0):
        desc = scene.desc if hasattr(scene, "desc") else scene
        lib = oracle_lib()
        h = lib.oracle_create(C.byref(desc), int(algorithm), radius_factor, radius_alpha, seed,
                              rng_mode, max_path_length, min_path_length)
        if not h:
            raise ValueError("oracle_create failed")
        super().__init__(lib, "oracle_", h, int(desc.camera.resolution[0]),
                         int(desc.camera.resolution[1]), keep=scene)

    def set_light_vertices(self, verts, radius, num_cells, mis_vc_weight_factor, ppm,
                           max_path_length, min_path_length):
        v = np.ascontiguousarray(verts, abi.MERGE_VERTEX_DTYPE)
        self._lib.oracle_set_light_vertices(self._h, _ptr(v), len(v), radius, num_cells,
                                            mis_vc_weight_factor, int(ppm), max_path_length,
                                            min_path_length)

    def counters(self):
        out = np.zeros(7, np.int64)
        self._lib.oracle_counters(self._h, _ptr(out))
        keys = ["lightRays", "cameraRays", "shadowRays", "mergeQueries", "mergeCandidates",
                "mergeAccepted", "connections"]
        return dict(zip(keys, (int(x) for x in out)))


def oracle_rng_stream(seed, iteration, phase, first_path, paths, draws):
    out = np.zeros((paths, draws), np.float32)
    oracle_lib().oracle_rng_stream(seed, iteration, phase, first_path, paths, draws, _ptr(out))
    return out


def oracle_total_luminance(img):
    img = np.ascontiguousarray(img, np.float32)
    return float(oracle_lib().oracle_total_luminance(_ptr(img), img.shape[1], img.shape[0]))


def oracle_bmp_payload(accum, iterations, gamma):
    accum = np.ascontiguousarray(accum, np.float32)
    out = np.empty(accum.shape[0] * accum.shape[1] * 3, np.uint8)
    oracle_lib().oracle_bmp_payload(_ptr(accum), accum.shape[1], accum.shape[0], iterations, gamma, _ptr(out))
    return out


def oracle_hdr_payload(accum, iterations):
    accum = np.ascontiguousarray(accum, np.float32)
    out = np.empty(accum.shape[0] * accum.shape[1] * 4, np.uint8)
    oracle_lib().oracle_hdr_payload(_ptr(accum), accum.shape[1], accum.shape[0], iterations, _ptr(out))
    return out


def oracle_render(scene, algorithm, iterations, threads, base_seed=1234, rng_mode=0,
                  max_path_length=10, min_path_length=0, radius_factor=0.003, radius_alpha=0.75,
                  want_image=True):
    desc = scene.desc if hasattr(scene, "desc") else scene
    w, h = int(desc.camera.resolution[0]), int(desc.camera.resolution[1])
    out = np.zeros((h, w, 3), np.float32) if want_image else None
    secs = oracle_lib().oracle_render(C.byref(desc), int(algorithm), iterations, threads, base_seed,
                                      rng_mode, max_path_length, min_path_length, radius_factor,
                                      radius_alpha, _ptr(out))
    return secs, out


# ------------------------------------------------------------------------------------------
# the reference itself (only where oracle/_ref was built)
# ------------------------------------------------------------------------------------------
def have_ref():
    return os.path.exists(REF_SO)


def have_ref_ctr():
    return os.path.exists(REF_CTR_SO)


_ref = None
_refctr = None

# Config::Algorithm numbering (reference src/config.hxx:54-64) for ref_render
REF_ALG = {"el": 0, "pt": 1, "lt": 2, "ppm": 3, "bpm": 4, "bpt": 5, "vcm": 6}


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.ref_render.restype = C.c_double
        lib.ref_render.argtypes = [i32] * 7 + [u32, u32, f32, f32, VP]
        lib.ref_total_luminance.restype = f32
        lib.ref_total_luminance.argtypes = [VP, i32, i32]
        lib.ref_num_procs.restype = i32
        lib.ref_scene_mask.restype = u32
        lib.ref_scene_mask.argtypes = [i32]
        lib.ref_save_image.restype = i32
        lib.ref_save_image.argtypes = [VP, i32, i32, i32, i32, f32, C.c_char_p]
        lib.ref_render_extra.restype = C.c_double
        lib.ref_render_extra.argtypes = [i32] * 7 + [u32, u32, f32, f32, VP, VP, i32, VP, VP, VP]
        lib.ref_flatten_scene.restype = i32
        lib.ref_flatten_scene.argtypes = [i32, i32, i32] + [VP] * 10
        _ref = lib
    return _ref


def ref_render(scene_id, algorithm, width, height, iterations, threads=1, base_seed=1234,
               max_path_length=10, min_path_length=0, radius_factor=0.003, radius_alpha=0.75,
               want_image=True):
    """The unmodified reference's render() (src/smallvcm.cxx:52). Returns (wall s, image)."""
    out = np.zeros((height, width, 3), np.float32) if want_image else None
    secs = ref_lib().ref_render(scene_id, REF_ALG[algorithm], width, height, iterations, threads,
                                base_seed, max_path_length, min_path_length, radius_factor,
                                radius_alpha, _ptr(out))
    if secs < 0:
        raise ValueError("ref_render failed")
    return secs, out


def ref_render_extra(scene_id, algorithm, width, height, iterations, vertices, materials, threads=1,
                     base_seed=1234, max_path_length=10, min_path_length=0, radius_factor=0.003,
                     radius_alpha=0.75):
    """The unmodified reference rendering box `scene_id` plus extra triangles appended to its
    GeometryList. Returns (image, the reference's flattened primitive bytes [n, 56], scene sphere)."""
    from smallvcm_b200 import abi
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 9)
    m = np.ascontiguousarray(materials, np.int32)
    out = np.zeros((height, width, 3), np.float32)
    prims = (abi.Primitive * (64 + len(v)))()
    sphere = abi.SceneSphere()
    secs = ref_lib().ref_render_extra(scene_id, REF_ALG[algorithm], width, height, iterations, threads, base_seed,
                                      max_path_length, min_path_length, radius_factor, radius_alpha, _ptr(v), _ptr(m),
                                      len(v), _ptr(out), C.cast(prims, VP), C.cast(C.byref(sphere), VP))
    if secs < 0:
        raise ValueError("ref_render_extra failed")
    return out, prims, sphere


def ref_save_image(accum, iterations, fmt, gamma, path):
    """The bytes of the file the reference's Framebuffer::SaveBMP (fmt 0) / SaveHDR (fmt 1) writes."""
    accum = np.ascontiguousarray(accum, np.float32)
    rc = ref_lib().ref_save_image(_ptr(accum), accum.shape[1], accum.shape[0], iterations, fmt, gamma,
                                  str(path).encode())
    assert rc == 0
    with open(path, "rb") as f:
        return f.read()


class FlatSceneCopy:
    """A SceneDesc plus the arrays it points to (keeps them alive)."""

    def __init__(self, prims, mats, lights, m2l, background, camera, sphere):
        self.prims, self.mats, self.lights, self.m2l = prims, mats, lights, m2l
        d = abi.SceneDesc()
        d.primitives = C.cast(prims, C.POINTER(abi.Primitive)); d.numPrimitives = len(prims)
        d.materials = C.cast(mats, C.POINTER(abi.Material)); d.numMaterials = len(mats)
        d.lights = C.cast(lights, C.POINTER(abi.Light)); d.numLights = len(lights)
        d.materialToLight = C.cast(m2l, C.POINTER(abi.i32))
        d.backgroundLight = background
        d.camera = camera
        d.sceneSphere = sphere
        self.desc = d


def ref_flatten_scene(scene_id, width, height):
    """The reference's own Scene for g_SceneConfigs[scene_id], flattened by the product's
    host glue (smallvcm_b200/host/flatten_scene.hxx)."""
    prims = (abi.Primitive * abi.MAX_PRIMITIVES)()
    mats = (abi.Material * abi.MAX_MATERIALS)()
    lights = (abi.Light * abi.MAX_LIGHTS)()
    m2l = (abi.i32 * abi.MAX_MATERIALS)()
    np_, nm, nl, bg = i32(), i32(), i32(), i32()
    cam, sph = abi.Camera(), abi.SceneSphere()
    rc = ref_lib().ref_flatten_scene(scene_id, width, height, prims, C.byref(np_), mats, C.byref(nm),
                                     lights, C.byref(nl), m2l, C.byref(bg), C.byref(cam), C.byref(sph))
    if rc != 0:
        raise ValueError("ref_flatten_scene failed: %d" % rc)
    p2 = (abi.Primitive * np_.value)(*prims[:np_.value])
    m2 = (abi.Material * nm.value)(*mats[:nm.value])
    l2 = (abi.Light * nl.value)(*lights[:nl.value])
    ml2 = (abi.i32 * nm.value)(*m2l[:nm.value])
    return FlatSceneCopy(p2, m2, l2, ml2, bg.value, cam, sph)


def refctr_lib():
    global _refctr
    if _refctr is None:
        lib = C.CDLL(REF_CTR_SO)
        for name, (res, args) in _COMMON.items():
            fn = getattr(lib, "refctr_" + name)
            fn.restype, fn.argtypes = res, args
        lib.refctr_create.restype = VP
        lib.refctr_create.argtypes = [i32] * 5 + [f32, f32, u32, u32]
        lib.refctr_merge_queries.restype = None
        lib.refctr_merge_queries.argtypes = [VP, VP, i64, VP]
        lib.refctr_rng_stream.restype = None
        lib.refctr_rng_stream.argtypes = [i32, i32, i32, u32, i32, i32, VP]
        _refctr = lib
    return _refctr


class RefCtr(_Checker):
    """The reference arithmetic with the per-path RNG (oracle/_ref/libsvcm_ref_ctr.so)."""

    def __init__(self, scene_id, algorithm, width, height, radius_factor=0.003, radius_alpha=0.75,
                 seed=1234, max_path_length=10, min_path_length=0):
        lib = refctr_lib()
        h = lib.refctr_create(scene_id, int(algorithm), width, height, seed, radius_factor,
                              radius_alpha, max_path_length, min_path_length)
        if not h:
            raise ValueError("refctr_create failed")
        super().__init__(lib, "refctr_", h, width, height)


def refctr_rng_stream(seed, iteration, phase, first_path, paths, draws):
    out = np.zeros((paths, draws), np.float32)
    refctr_lib().refctr_rng_stream(seed, iteration, phase, first_path, paths, draws, _ptr(out))
    return out
