"""ctypes host layer over libsvcm.so (include/svcm.h).

Mirrors, name for name, what a user of the reference touches on this path:

    reference (C++)                                   here
    ------------------------------------------------  -----------------------------------------
    Scene::LoadCornellBox + BuildSceneSphere          Scene.cornell_box(scene_id, w, h)
      (src/scene.hxx:132,387; g_SceneConfigs          (svcm_scene_cornell_box)
       src/config.hxx:146-151)
    new VertexCM(scene, alg, radiusFactor,            VertexCM(scene, algorithm, radius_factor,
                 radiusAlpha, seed)                            radius_alpha, seed)
      (src/vertexcm.hxx:208-214)
    renderer->mMaxPathLength / mMinPathLength         .max_path_length / .min_path_length
      (assigned after construction, smallvcm.cxx:70)
    renderer->RunIteration(i)   (vertexcm.hxx:284)    .run_iteration(i)
    renderer->GetFramebuffer(fb) (renderer.hxx:49)    .get_framebuffer()  (accum / iterations)
    renderer->WasUsed()          (renderer.hxx:58)    .was_used()

There is no CPU fallback: if the library is missing or no CUDA device is present every
computing call raises SvcmError.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SCENE_NAMES = {
    0: "glossy small spheres + sun (directional)",
    1: "glossy large mirror sphere + ceiling (area)",
    2: "glossy small spheres + point",
    3: "glossy small spheres + background (env. lighting)",
}


class SvcmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("svcm error %d: %s" % (code, message))
        self.code = code


def library_path():
    return os.environ.get("SVCM_LIBRARY", os.path.join(_HERE, "lib", "libsvcm.so"))


def load_library():
    """Loads libsvcm.so and binds every symbol of include/svcm.h. Raises if it is missing —
    build it with `python -c "import __graft_entry__ as g; g.build()"` or
    `make -C smallvcm_b200/csrc`."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise SvcmError(abi.E_STATE, "CUDA library not built: %s is missing (run make -C smallvcm_b200/csrc)" % path)
        _LIB = abi.bind(C.CDLL(path))
    return _LIB


def _check(rc, handle=None):
    if rc != 0:
        msg = load_library().svcm_last_error(handle)
        raise SvcmError(rc, msg.decode() if msg else "?")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Scene:
    """A flat scene (SvcmSceneDesc) and whatever owns its arrays."""

    def __init__(self, desc, owner=None, free=None):
        self.desc = desc
        self._owner = owner
        self._free = free

    @classmethod
    def cornell_box(cls, scene_id, width, height, mask=None):
        """g_SceneConfigs[scene_id] (or an explicit BoxMask) at width x height."""
        lib = load_library()
        if mask is None:
            mask = lib.svcm_scene_config_mask(scene_id)
            if mask == 0:
                raise SvcmError(abi.E_INVALID, "unknown scene id %r" % (scene_id,))
        handle = C.c_void_p()
        pdesc = abi.PSD()
        _check(lib.svcm_scene_cornell_box(mask, width, height, C.byref(handle), C.byref(pdesc)))
        return cls(pdesc.contents, owner=handle, free=lib.svcm_scene_free)

    def with_triangles(self, vertices, mat_ids):
        """A new scene: this one's camera, materials, lights and primitives plus the triangles
        `vertices` [n, 3, 3] (float32, counter-clockwise like src/geometry.hxx:99-107) with
        materials `mat_ids` [n], appended to the primitive list. The normal is the reference's
        Normalize(Cross(p1 - p0, p2 - p0)) evaluated in float32. The scene sphere is kept, so the
        triangles must lie inside the existing geometry's bounds. More than 64 primitives make
        svcm_create build a hierarchy (include/svcm.h, SVCM_MAX_PRIMITIVES)."""
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3, 3)
        mats = np.ascontiguousarray(mat_ids, np.int32).reshape(-1)
        assert len(v) == len(mats)
        e1, e2 = v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]
        n = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1], e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
                      e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], 1).astype(np.float32)
        length = np.sqrt((n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2]).astype(np.float32)).astype(np.float32)
        n = (n / length[:, None]).astype(np.float32)
        old = self.desc.numPrimitives
        prims = (abi.Primitive * (old + len(v)))()
        for i in range(old):
            prims[i] = self.desc.primitives[i]
        raw = np.frombuffer(prims, np.uint8).reshape(old + len(v), C.sizeof(abi.Primitive))
        rec = np.zeros((len(v), 14), np.float32)
        rec[:, 0] = np.full(len(v), abi.PRIM_TRIANGLE, np.int32).view(np.float32)
        rec[:, 1] = mats.view(np.float32)
        rec[:, 2:11] = v.reshape(-1, 9)
        rec[:, 11:14] = n
        raw[old:] = rec.view(np.uint8).reshape(len(v), -1)
        desc = abi.SceneDesc()
        C.memmove(C.byref(desc), C.byref(self.desc), C.sizeof(abi.SceneDesc))
        desc.primitives = C.cast(prims, C.POINTER(abi.Primitive))
        desc.numPrimitives = old + len(v)
        return Scene(desc, owner=(self, prims), free=None)

    def bvh(self):
        """(nodes structured array, order) of the hierarchy svcm_create would build (svcm_bvh_build)."""
        n = self.desc.numPrimitives
        nodes = (abi.BvhNode * (2 * n))()
        order = np.zeros(n, np.int32)
        count = C.c_int32()
        _check(load_library().svcm_bvh_build(self.desc.primitives, n, nodes, 2 * n, _ptr(order), C.byref(count)))
        dt = np.dtype([("lo", np.float32, 3), ("first", np.int32), ("hi", np.float32, 3), ("count", np.int32)])
        return np.frombuffer(nodes, dt)[:count.value].copy(), order

    def primitive_runs(self):
        """(runs structured array, skippable bit mask) of a constant-bank scene (svcm_primitive_runs)."""
        runs = (abi.BvhNode * abi.MAX_RUNS)()
        mask = C.c_uint32(); count = C.c_int32()
        _check(load_library().svcm_primitive_runs(self.desc.primitives, self.desc.numPrimitives, runs, C.byref(mask),
                                                  C.byref(count)))
        dt = np.dtype([("lo", np.float32, 3), ("first", np.int32), ("hi", np.float32, 3), ("count", np.int32)])
        return np.frombuffer(runs, dt)[:count.value].copy(), int(mask.value)

    @property
    def width(self):
        return int(self.desc.camera.resolution[0])

    @property
    def height(self):
        return int(self.desc.camera.resolution[1])

    def __del__(self):
        if self._free is not None and self._owner:
            try:
                self._free(self._owner)
            except Exception:
                pass
            self._owner = None


class VertexCM:
    """The reference's VertexCM renderer interface on top of the C-ABI."""

    def __init__(self, scene, algorithm, radius_factor=0.003, radius_alpha=0.75, seed=1234,
                 device=0, rank=0, world_size=1):
        self._lib = load_library()
        self.scene = scene
        if isinstance(algorithm, str):
            algorithm = abi.ALGORITHMS[algorithm]
        p = abi.Params()
        p.algorithm, p.radiusFactor, p.radiusAlpha, p.seed = int(algorithm), radius_factor, radius_alpha, seed
        p.device, p.rank, p.worldSize = device, rank, world_size
        self._h = C.c_void_p()
        _check(self._lib.svcm_create(C.byref(scene.desc), C.byref(p), C.byref(self._h)))
        self.width, self.height = scene.width, scene.height
        # AbstractRenderer's defaults (src/renderer.hxx:39-40); the driver overrides them
        self.max_path_length = 2
        self.min_path_length = 0
        self.algorithm = int(algorithm)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.svcm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- the hot path ---------------------------------------------------------------------
    def run_iteration(self, iteration):
        self.run_iterations(iteration, 1)

    def run_iterations(self, first, count):
        _check(self._lib.svcm_run_iterations(self._h, first, count, self.min_path_length,
                                             self.max_path_length), self._h)

    def synchronize(self):
        _check(self._lib.svcm_synchronize(self._h), self._h)

    def accum(self, out=None):
        """(accumulated image [H, W, 3] float32, iterations it holds)."""
        if out is None:
            out = np.empty((self.height, self.width, 3), np.float32)
        n = C.c_int32()
        _check(self._lib.svcm_download_accum(self._h, _ptr(out), C.byref(n)), self._h)
        return out, n.value

    def get_framebuffer(self):
        img, n = self.accum()
        if n > 0:
            img *= np.float32(1.0 / n)
        return img

    def image8(self, fmt="bmp", gamma=2.2):
        """The pixel payload of Framebuffer::SaveBMP ("bmp": [H, W, 3] uint8, rows bottom-up,
        B G R) or SaveHDR ("hdr": [H, W, 4] uint8 R G B E) of the averaged image, converted on
        the device (svcm_download_image)."""
        code = {"bmp": abi.IMAGE_BMP24, "hdr": abi.IMAGE_RGBE}[fmt]
        out = np.empty((self.height, self.width, 3 if code == abi.IMAGE_BMP24 else 4), np.uint8)
        _check(self._lib.svcm_download_image(self._h, code, gamma, _ptr(out), None), self._h)
        return out

    def save_image(self, path, gamma=2.2):
        """main()'s image output (src/smallvcm.cxx:311-320): *.bmp or *.hdr, chosen by extension."""
        _check(self._lib.svcm_save_image(self._h, os.fsencode(path), gamma), self._h)

    def was_used(self):
        return self.accum()[1] > 0

    def clear(self):
        _check(self._lib.svcm_clear_accum(self._h), self._h)

    def stats(self):
        s = abi.Stats()
        _check(self._lib.svcm_get_stats(self._h, C.byref(s)), self._h)
        return {name: getattr(s, name) for name, _ in abi.Stats._fields_}

    def profile(self):
        """Sums over the iterations run since the previous call (svcm_get_profile)."""
        p = abi.Profile()
        _check(self._lib.svcm_get_profile(self._h, C.byref(p)), self._h)
        return {name: getattr(p, name) for name, _ in abi.Profile._fields_}

    def event_record(self, slot):
        _check(self._lib.svcm_event_record(self._h, slot), self._h)

    def event_elapsed_ms(self, start_slot, stop_slot):
        ms = C.c_float()
        _check(self._lib.svcm_event_elapsed_ms(self._h, start_slot, stop_slot, C.byref(ms)), self._h)
        return ms.value

    def set_option(self, option, value):
        _check(self._lib.svcm_set_option(self._h, option, value), self._h)

    # --- multi-GPU ------------------------------------------------------------------------
    def comm_init(self, unique_id):
        buf = (C.c_uint8 * abi.COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        _check(self._lib.svcm_comm_init(self._h, buf), self._h)

    def reduce_accum(self, root=0):
        _check(self._lib.svcm_reduce_accum(self._h, root), self._h)


def comm_unique_id():
    buf = (C.c_uint8 * abi.COMM_ID_BYTES)()
    _check(load_library().svcm_comm_unique_id(buf))
    return bytes(buf)


# ------------------------------------------------------------------------------------------
# kernel-level entry points
# ------------------------------------------------------------------------------------------
class HashGrid:
    def __init__(self, handle, count, num_cells):
        self._h, self.count, self.num_cells = handle, count, num_cells
        self._lib = load_library()

    def download(self):
        bbox = np.zeros(6, np.float32)
        ends = np.zeros(self.num_cells, np.int32)
        idx = np.zeros(self.count, np.int32)
        _check(self._lib.svcm_hashgrid_download(self._h, _ptr(bbox), _ptr(ends), _ptr(idx)))
        return bbox, ends, idx

    def query(self, scene, queries, min_path_length, max_path_length, mis_vc_weight_factor, ppm,
              want_counts=True):
        q = np.ascontiguousarray(queries, abi.MERGE_QUERY_DTYPE)
        contrib = np.zeros((len(q), 3), np.float32)
        counts = np.zeros((len(q), 2), np.int32) if want_counts else None
        _check(self._lib.svcm_hashgrid_query(self._h, C.byref(scene.desc), _ptr(q), len(q), min_path_length,
                                             max_path_length, mis_vc_weight_factor, int(ppm), _ptr(contrib),
                                             _ptr(counts)))
        return (contrib, counts) if want_counts else contrib

    def close(self):
        if self._h:
            self._lib.svcm_hashgrid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hashgrid_build(vertices, radius, num_cells, device=0):
    v = np.ascontiguousarray(vertices, abi.MERGE_VERTEX_DTYPE)
    h = C.c_void_p()
    _check(load_library().svcm_hashgrid_build(_ptr(v), len(v), radius, num_cells, device, C.byref(h)))
    return HashGrid(h, len(v), num_cells)


def intersect_batch(scene, rays, device=0):
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    n = len(rays)
    hit = np.zeros(n, np.int32); dist = np.zeros(n, np.float32); normal = np.zeros((n, 3), np.float32)
    mat = np.zeros(n, np.int32); light = np.zeros(n, np.int32)
    _check(load_library().svcm_intersect_batch(C.byref(scene.desc), _ptr(rays), n, device, _ptr(hit), _ptr(dist),
                                               _ptr(normal), _ptr(mat), _ptr(light)))
    return hit, dist, normal, mat, light


def occluded_batch(scene, segments, device=0):
    seg = np.ascontiguousarray(segments, np.float32).reshape(-1, 7)
    occ = np.zeros(len(seg), np.int32)
    _check(load_library().svcm_occluded_batch(C.byref(scene.desc), _ptr(seg), len(seg), device, _ptr(occ)))
    return occ


def bsdf_evaluate_batch(scene, items, fix_is_light, device=0):
    items = np.ascontiguousarray(items, np.float32).reshape(-1, 10)
    out = np.zeros((len(items), 8), np.float32)
    _check(load_library().svcm_bsdf_evaluate_batch(C.byref(scene.desc), _ptr(items), len(items), int(fix_is_light),
                                                   device, _ptr(out)))
    return out


def bsdf_sample_batch(scene, items, fix_is_light, device=0):
    items = np.ascontiguousarray(items, np.float32).reshape(-1, 10)
    out = np.zeros((len(items), 8), np.float32)
    ev = np.zeros(len(items), np.uint32)
    _check(load_library().svcm_bsdf_sample_batch(C.byref(scene.desc), _ptr(items), len(items), int(fix_is_light),
                                                 device, _ptr(out), _ptr(ev)))
    return out, ev


def light_batch(scene, op, light_index, items, device=0):
    """op in ("emit", "illuminate", "radiance"): svcm_light_*_batch, items [n, 8] -> [n, 12]."""
    items = np.ascontiguousarray(items, np.float32).reshape(-1, 8)
    out = np.zeros((len(items), 12), np.float32)
    fn = getattr(load_library(), "svcm_light_%s_batch" % op)
    _check(fn(C.byref(scene.desc), int(light_index), _ptr(items), len(items), device, _ptr(out)))
    return out


def camera_ray_batch(scene, items, device=0):
    """svcm_camera_ray_batch: items [n, 5] (raster.xy, world point) -> [n, 6]."""
    items = np.ascontiguousarray(items, np.float32).reshape(-1, 5)
    out = np.zeros((len(items), 6), np.float32)
    _check(load_library().svcm_camera_ray_batch(C.byref(scene.desc), _ptr(items), len(items), device, _ptr(out)))
    return out


def rng_stream(seed, iteration, phase, first_path, paths, draws):
    out = np.zeros((paths, draws), np.float32)
    _check(load_library().svcm_rng_stream(seed, iteration, phase, first_path, paths, draws, _ptr(out)))
    return out


def image_convert_batch(accum, iterations, fmt="bmp", gamma=2.2, device=0):
    """svcm_image_convert_batch: SaveBMP / SaveHDR payload bytes of a host image [H, W, 3]."""
    accum = np.ascontiguousarray(accum, np.float32)
    h, w = accum.shape[:2]
    code = {"bmp": abi.IMAGE_BMP24, "hdr": abi.IMAGE_RGBE}[fmt]
    out = np.empty(h * w * (3 if code == abi.IMAGE_BMP24 else 4), np.uint8)
    _check(load_library().svcm_image_convert_batch(_ptr(accum), w, h, iterations, code, gamma, device, _ptr(out)))
    return out


def image_header(fmt, width, height):
    """The bytes the reference writes in front of the pixel payload (svcm_image_header)."""
    code = {"bmp": abi.IMAGE_BMP24, "hdr": abi.IMAGE_RGBE}[fmt]
    buf = np.zeros(128, np.uint8)
    n = load_library().svcm_image_header(code, width, height, _ptr(buf), len(buf))
    return buf[:n].tobytes()


def rng_stream_device(seed, iteration, phase, first_path, paths, draws, device=0):
    out = np.zeros((paths, draws), np.float32)
    _check(load_library().svcm_rng_stream_device(seed, iteration, phase, first_path, paths, draws, device, _ptr(out)))
    return out
