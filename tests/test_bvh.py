"""Scenes beyond the constant-bank limit: the bounding volume hierarchy must
return exactly what the reference's loops over the whole primitive list return
(GeometryList::Intersect / IntersectP, reference src/geometry.hxx:65-91).

CPU: the builder's structural invariants, and the traversal templates the kernels run
(smallvcm_b200/csrc/svcm_bvh.h) instantiated on the host against a list-order loop
(tests/native/bvh_check.cpp) — identical primitive, distance and any-hit answer for every ray.
GPU: the CUDA traversal against the ORACLE's loops on a 7 761-primitive scene (exact primitive /
material / occlusion, distances to the T1 tolerance), the reference's own scenes rendered through
the hierarchy vs through the constant-bank loop, and a render of the large scene vs the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_api as oa
import scene_gen
import smallvcm_b200 as sv
from smallvcm_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def prim_array(scene):
    n = scene.desc.numPrimitives
    raw = np.frombuffer(C.string_at(scene.desc.primitives, n * 56), np.uint8).reshape(n, 56)
    return raw[:, :4].copy().view(np.int32)[:, 0], raw[:, 4:8].copy().view(np.int32)[:, 0], raw[:, 8:].copy().view(np.float32)


def mixed_rays(scene, n, seed):
    """Random rays, axis-aligned rays, rays aimed at triangle vertices and edge midpoints, rays
    leaving surface points (what a path does), rays inside the planes of the walls."""
    rng = np.random.default_rng(seed)
    types, _, v = prim_array(scene)
    tri = v[types == 0][:, :9].reshape(-1, 3, 3)
    org = rng.uniform(-1.2, 1.2, (n, 3))
    d = rng.normal(size=(n, 3))
    k = n // 6
    d[:k] = np.eye(3)[rng.integers(0, 3, k)] * rng.choice([-1, 1], (k, 1))                  # axis aligned
    t = tri[rng.integers(0, len(tri), k)]
    d[k:2 * k] = t[np.arange(k), rng.integers(0, 3, k)] - org[k:2 * k]                       # at vertices
    t = tri[rng.integers(0, len(tri), k)]
    d[2 * k:3 * k] = 0.5 * (t[:, 0] + t[:, 1]) - org[2 * k:3 * k]                            # at edge midpoints
    t = tri[rng.integers(0, len(tri), k)]
    w = rng.dirichlet([1, 1, 1], k)
    org[3 * k:4 * k] = (t * w[:, :, None]).sum(1)                                            # from surface points
    org[4 * k:5 * k, 2] = -1.28                                                              # in the floor's plane
    d[4 * k:5 * k, 2] = 0.0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([org, d, rng.uniform(0.05, 3.0, (n, 1))], 1).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# CPU
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def big():
    scene, base = scene_gen.big_scene(sv, 1, 64, 64)
    return scene


def test_builder_invariants(big):
    nodes, order = big.bvh()
    n = big.desc.numPrimitives
    assert n == 7761 and np.array_equal(np.sort(order), np.arange(n)) and len(nodes) <= 2 * n - 1
    types, _, v = prim_array(big)
    lo = np.where(types[:, None] == 0, v[:, :9].reshape(n, 3, 3).min(1), v[:, :3] - np.abs(v[:, 3:4]))
    hi = np.where(types[:, None] == 0, v[:, :9].reshape(n, 3, 3).max(1), v[:, :3] + np.abs(v[:, 3:4]))
    seen = np.zeros(n, int)
    depth = np.zeros(len(nodes), int)
    for i, nd in enumerate(nodes):
        if nd["count"] > 0:                                     # leaf: contains its primitives, with padding
            sl = order[nd["first"]:nd["first"] + nd["count"]]
            seen[sl] += 1
            assert np.all(lo[sl] > nd["lo"]) and np.all(hi[sl] < nd["hi"])
        else:                                                   # inner: children adjacent, inside the parent
            f = nd["first"]
            assert 0 < f < len(nodes) - 1
            for c in (f, f + 1):
                assert np.all(nodes[c]["lo"] >= nd["lo"]) and np.all(nodes[c]["hi"] <= nd["hi"])
                depth[c] = depth[i] + 1
    assert np.all(seen == 1) and depth.max() < 56 and nodes["count"].max() <= 4


def test_primitive_runs_of_small_scenes():
    """svcm_primitive_runs: the constant-bank list in contiguous runs, list order kept; only runs of
    three or more small triangles are skippable, and their padded boxes contain them."""
    import ctypes as C
    for sid, nprim, skippable in [(0, 22, True), (1, 21, True), (2, 12, False), (3, 22, True)]:
        scene = sv.Scene.cornell_box(sid, 16, 16)
        runs, mask = scene.primitive_runs()
        types, _, v = prim_array(scene)
        assert scene.desc.numPrimitives == nprim and runs["first"][0] == 0
        assert np.array_equal(runs["first"][1:], (runs["first"] + runs["count"])[:-1]) and runs["count"].sum() == nprim
        if not skippable:
            assert mask == 0 and len(runs) == 1
            continue
        assert len(runs) == 2 and mask == 0b10 and runs["count"][1] == 10          # the light box: the last ten triangles
        assert np.all(runs["lo"][0] < -1e38) and np.all(runs["hi"][0] > 1e38)       # walls and spheres: always tested
        box = v[nprim - 10:, :9].reshape(-1, 3)
        assert np.all(types[nprim - 10:] == 0)
        assert np.all(box.min(0) > runs["lo"][1]) and np.all(box.max(0) < runs["hi"][1])
        assert np.all(box.min(0) - runs["lo"][1] < 1e-4) and np.all(runs["hi"][1] - box.max(0) < 1e-4)
    # many separate clusters of small triangles: never more than SVCM_MAX_RUNS runs, the tail is always tested
    base = sv.Scene.cornell_box(2, 16, 16)
    rng = np.random.default_rng(5)
    tri = []
    for c in range(12):
        centre = rng.uniform(-1, 1, 3)
        tri.append(centre + rng.uniform(-0.05, 0.05, (4, 3, 3)))
    scene = base.with_triangles(np.concatenate(tri).astype(np.float32), np.full(48, 5))
    runs, mask = scene.primitive_runs()
    assert len(runs) == abi.MAX_RUNS and runs["count"].sum() == 60 and not (mask >> (abi.MAX_RUNS - 1)) & 1
    assert np.array_equal(runs["first"][1:], (runs["first"] + runs["count"])[:-1])
    for r in range(len(runs)):
        if (mask >> r) & 1:
            _, _, v = prim_array(scene)
            pts = v[runs["first"][r]:runs["first"][r] + runs["count"][r], :9].reshape(-1, 3)
            assert runs["count"][r] >= 3 and np.all(pts.min(0) > runs["lo"][r]) and np.all(pts.max(0) < runs["hi"][r])
    lib = sv.load_library()
    out = (abi.BvhNode * abi.MAX_RUNS)(); m = C.c_uint32(); k = C.c_int32()
    assert lib.svcm_primitive_runs(None, 12, out, C.byref(m), C.byref(k)) == abi.E_INVALID
    assert lib.svcm_primitive_runs(base.desc.primitives, 65, out, C.byref(m), C.byref(k)) == abi.E_INVALID
    assert lib.svcm_primitive_runs(base.desc.primitives, 0, out, C.byref(m), C.byref(k)) == abi.E_INVALID


def test_degenerate_inputs_build():
    base = sv.Scene.cornell_box(0, 16, 16)
    same = np.tile(np.array([[[0, 0, 0], [0.1, 0, 0], [0, 0.1, 0]]], np.float32), (300, 1, 1))   # 300 identical triangles
    nodes, order = base.with_triangles(same, np.full(300, 5)).bvh()
    assert np.array_equal(np.sort(order), np.arange(322)) and nodes["count"].max() <= 4
    bad = same.copy(); bad[7, 1, 2] = np.nan
    with pytest.raises(sv.SvcmError):
        base.with_triangles(bad, np.full(300, 5)).bvh()


def test_traversal_equals_list_order_loop(big, tmp_path):
    so = str(tmp_path / "libbvh_check.so")
    subprocess.check_call(["g++", "-O2", "-std=c++11", "-ffp-contract=off", "-fPIC", "-shared",
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "smallvcm_b200", "csrc"),
                           os.path.join(HERE, "native", "bvh_check.cpp"), "-o", so])
    lib = C.CDLL(so)
    for scene, nrays in ((big, 60000), (sv.Scene.cornell_box(3, 16, 16), 200000)):
        nodes, order = scene.bvh()
        nodes = np.ascontiguousarray(nodes)
        rays = mixed_rays(scene, nrays, 3)
        out = [np.zeros(nrays, dt) for dt in (np.int32, np.float32, np.int32, np.float32, np.int32, np.int32)]
        bad = lib.bvh_check(scene.desc.primitives, scene.desc.numPrimitives, nodes.ctypes.data_as(C.c_void_p),
                            order.ctypes.data_as(C.c_void_p), rays.ctypes.data_as(C.c_void_p), nrays,
                            *[o.ctypes.data_as(C.c_void_p) for o in out])
        assert bad == 0
        assert np.array_equal(out[0], out[2]) and np.array_equal(out[1], out[3]) and np.array_equal(out[4], out[5])
        assert 0.3 < (out[0] >= 0).mean() and 0.1 < out[4].mean() < 0.95     # the rays do exercise both answers


@pytest.mark.parametrize("flags", [[], ["-Xcompiler", "-mfma,-ffp-contract=fast"]], ids=["unfused", "fma"])
def test_triangle_early_out_never_changes_the_answer(flags, tmp_path):
    """hitTriangle (svcm_device.cuh) rejects on the plane distance before the edge tests and without the
    division; tests/native/tri_early_out_check.cu runs it as host code against the reference's order of
    evaluation on 48 M random and adversarial ray / triangle / tmax combinations, with and without FMA
    contraction: same hit flag and bit-identical distance every time."""
    exe = str(tmp_path / "tri_early_out_check")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-x", "cu", "-Wno-deprecated-gpu-targets", *flags,
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "smallvcm_b200", "csrc"),
                           os.path.join(HERE, "native", "tri_early_out_check.cu"), "-o", exe])
    out = subprocess.run([exe, "200000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    fields = out.stdout.split()
    assert fields[-2:] == ["disagreements", "0"] and int(fields[3]) > 1000000 and int(fields[5]) > 1000000


def test_oracle_bit_exact_on_a_scene_that_is_not_a_hard_coded_box():
    """Pins the oracle beyond the reference's four boxes: box 1 (21 primitives) + 1 160 triangles appended to the
    reference's GeometryList, rendered by the UNMODIFIED reference (tests/golden/
    ref_extra_scene.npz, made by oracle/ref_glue ref_render_extra). Scene.with_triangles must build
    byte-identical primitives (the reference's Triangle constructor computes the normal), and the
    oracle with the reference's mt19937 stream must reproduce the images bit for bit."""
    gold = np.load(os.path.join(HERE, "golden", "ref_extra_scene.npz"))
    scene = sv.Scene.cornell_box(1, 64, 64).with_triangles(gold["vertices"], gold["materials"])
    n = scene.desc.numPrimitives
    assert n == 1181 and len(gold["primitives"]) == n
    mine = np.frombuffer(C.string_at(scene.desc.primitives, n * 56), np.uint8).reshape(n, 56)
    assert np.array_equal(mine, gold["primitives"])
    sph = scene.desc.sceneSphere
    assert np.array_equal(np.array(list(sph.center) + [sph.radius, sph.invRadiusSqr], np.float32), gold["sphere"])
    for alg in ("vcm", "bpt"):
        orc = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=1234, rng_mode=0)
        orc.run_iteration(0); orc.run_iteration(1)
        img, cnt = orc.accum()
        orc.close()
        assert np.array_equal((img * np.float32(1.0 / cnt)).view(np.uint32), gold["image_" + alg].view(np.uint32))


@pytest.mark.skipif(not oa.have_ref(), reason="oracle/_ref not built (no reference checkout)")
def test_with_triangles_matches_live_reference_constructor():
    rng = np.random.default_rng(8)
    verts = rng.uniform(-1, 1, (500, 3, 3)).astype(np.float32)
    mats = rng.integers(2, 9, 500).astype(np.int32)
    _, prims, _ = oa.ref_render_extra(0, "vcm", 16, 16, 0, verts, mats)
    scene = sv.Scene.cornell_box(0, 16, 16).with_triangles(verts, mats)
    n = scene.desc.numPrimitives
    assert np.array_equal(np.frombuffer(C.string_at(scene.desc.primitives, n * 56), np.uint8),
                          np.frombuffer(C.string_at(prims, n * 56), np.uint8))


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_device_traversal_matches_oracle_loops(big):
    n = 120000
    rays = mixed_rays(big, n, 4)
    k = n // 6
    rays[3 * k:4 * k, :3] += 1e-3 * rays[3 * k:4 * k, 3:6]     # paths leave surfaces with an offset (EPS_RAY)
    aimed = np.zeros(n, bool); aimed[k:3 * k] = True           # at vertices / edges: near-ties by construction
    orc = oa.Oracle(big, abi.ALG_VCM)
    o = orc.intersect(rays[:, :6])
    g = sv.intersect_batch(big, rays[:, :6])
    # closest hit: the same primitive (exact ties keep list order; FMA contraction in the device's
    # triangle test can swap near-ties, which is what the aimed rays provoke)
    mism = g[0] != o[0]
    assert mism[~aimed].mean() < 1e-4, mism[~aimed].sum()
    assert ((g[0] < 0) != (o[0] < 0))[aimed].mean() < 2e-2
    ok = ~mism
    assert np.array_equal(g[3][ok], o[3][ok]) and np.array_equal(g[4][ok], o[4][ok])
    hit = ok & (o[0] >= 0)
    assert hit.mean() > 0.5 and (o[0] >= 22).mean() > 0.1      # and many of the hits are on the added geometry
    d = np.abs(g[1][hit] - o[1][hit]) / np.maximum(o[1][hit], 1e-3)
    assert np.quantile(d, 0.999) < 1e-5 and d.max() < 1e-3
    assert np.abs(g[2][hit] - o[2][hit]).max() < 1e-5
    # (a ray aimed exactly at a vertex or an edge may slip through the crack on one side and
    # hit whatever lies behind — the same happens between CPU and GPU on the reference's scenes)
    go, oo = sv.occluded_batch(big, rays), orc.occluded(rays)
    assert (go != oo)[~aimed].mean() < 1e-4 and (go != oo)[aimed].mean() < 2e-2
    assert 0.1 < oo.mean() < 0.95
    orc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scene_id,alg", [(1, "vcm"), (3, "vcm"), (0, "bpt"), (2, "pt")])
def test_reference_scenes_through_the_hierarchy(monkeypatch, scene_id, alg):
    """The reference's own scenes (two spheres, coplanar triangle pairs sharing a diagonal) rendered
    with the hierarchy forced on vs with the constant-bank loop: the same paths, so the images
    differ only by the order of float atomics."""
    def render():
        scene = sv.Scene.cornell_box(scene_id, 128, 128)
        r = sv.VertexCM(scene, alg); r.max_path_length = 10
        r.run_iterations(0, 3)
        img, st = r.accum()[0], r.stats()
        r.close()
        return img, st
    ref, st0 = render()
    monkeypatch.setenv("SVCM_DEBUG_FORCE_BVH", "1")
    img, st1 = render()
    assert st1["lightVertices"] == st0["lightVertices"] and st1["shadowRays"] == st0["shadowRays"]
    assert rel_l2(img, ref) < 1e-5


@pytest.mark.gpu
def test_large_scene_render_matches_oracle():
    scene, _ = scene_gen.big_scene(sv, 1, 96, 96)
    r = sv.VertexCM(scene, "vcm"); r.max_path_length = 8
    r.run_iterations(0, 3)
    img, n = r.accum()
    st = r.stats()
    r.close()
    orc = oa.Oracle(scene, abi.ALG_VCM, rng_mode=1, max_path_length=8)
    for it in range(3):
        orc.run_iteration(it)
    want, _ = orc.accum()
    nv = orc.num_light_vertices()
    orc.close()
    assert abs(st["lightVertices"] - nv) <= max(2, 2e-3 * nv)
    assert rel_l2(img, want) < 6e-2
    assert abs(img.mean() - want.mean()) < 5e-3 * want.mean()


@pytest.mark.gpu
def test_large_scene_limits_and_throughput():
    scene, base = scene_gen.big_scene(sv, 1, 512, 512, n_lat=120, n_lon=240, n_confetti=20000)
    assert scene.desc.numPrimitives > 75000
    r = sv.VertexCM(scene, "vcm"); r.max_path_length = 10
    r.run_iterations(0, 2); r.synchronize()
    r.event_record(0); r.run_iterations(2, 4); r.event_record(1)
    ms = r.event_elapsed_ms(0, 1) / 4
    img = r.accum()[0]
    r.close()
    print("\\n77k-primitive scene, 512x512 vcm: %.2f ms / iteration, %.1f Mpaths/s" % (ms, 512 * 512 / ms / 1e3))
    assert np.isfinite(img).all() and img.mean() > 0.05
    # two handles with different scenes do not share geometry
    r1 = sv.VertexCM(base, "bpt"); r2 = sv.VertexCM(scene, "bpt")
    r1.max_path_length = r2.max_path_length = 6
    r1.run_iterations(0, 1); r2.run_iterations(0, 1)
    a, b = r1.accum()[0], r2.accum()[0]
    r1.close(); r2.close()
    assert rel_l2(a, b) > 0.05
