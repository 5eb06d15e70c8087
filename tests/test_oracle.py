"""Pins the parity oracle (oracle/vcm_oracle.cpp) to the reference.

* against committed golden outputs of the COMPILED reference (tests/golden/, regenerate with
  tests/golden/make_golden.py): bit-exact images for the reference's own mt19937_64 stream and
  for the per-path counter stream, and the known-answer scalars of BASELINE.md §3;
* where oracle/_ref exists: bit-exact light vertices, hash grid and merge sums against the
  reference's own classes.
"""
import json
import os

import numpy as np
import pytest

import oracle_api as oa
import smallvcm_b200 as sv
from smallvcm_b200 import abi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAIRS = [(1, "vcm"), (0, "bpt"), (3, "bpm"), (2, "vcm"), (1, "lt"), (1, "ppm"), (0, "vcm"), (3, "vcm"),
         (0, "pt"), (1, "pt"), (3, "pt"), (2, "el")]      # pt / el: SURVEY §8(f) rank 1


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def get_framebuffer(orc):
    """AbstractRenderer::GetFramebuffer (src/renderer.hxx:49-55): accumulated / iterations."""
    img, n = orc.accum()
    return img * np.float32(1.0 / n)


@pytest.mark.parametrize("scene_id,alg", PAIRS)
def test_reference_stream_bit_exact(scene_id, alg):
    """Oracle with the reference's sequential mt19937_64 stream == unmodified reference."""
    gold = np.load(os.path.join(GOLD, "ref_images.npz"))["s%d_%s" % (scene_id, alg)]
    scene = sv.Scene.cornell_box(scene_id, 64, 64)
    orc = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=1234, rng_mode=0)
    orc.run_iteration(0)
    orc.run_iteration(1)
    assert same_bits(get_framebuffer(orc), gold)


@pytest.mark.parametrize("scene_id,alg", PAIRS)
def test_counter_stream_bit_exact(scene_id, alg):
    """Oracle with the per-path stream == reference arithmetic with that generator swapped in."""
    gold = np.load(os.path.join(GOLD, "refctr_images.npz"))
    scene = sv.Scene.cornell_box(scene_id, 64, 64)
    orc = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=1234, rng_mode=1)
    orc.run_iteration(0)
    orc.run_iteration(1)
    img, n = orc.accum()
    assert n == 2
    assert same_bits(img, gold["s%d_%s" % (scene_id, alg)])
    assert orc.num_light_vertices() == int(gold["s%d_%s_nverts" % (scene_id, alg)][0])


def test_known_answer_scene1_vcm_512():
    """scene 1 vcm 512x512, 1 iteration, seed 1234 -> TotalLuminance 69769.8."""
    kat = [k for k in json.load(open(os.path.join(GOLD, "ref_kat.json")))
           if (k["scene"], k["alg"], k["width"], k["iterations"]) == (1, "vcm", 512, 1)][0]
    assert abs(kat["totalLuminance"] - 69769.8) < 0.05          # the survey's published value
    scene = sv.Scene.cornell_box(1, 512, 512)
    orc = oa.Oracle(scene, abi.ALG_VCM, seed=1234, rng_mode=0)
    orc.run_iteration(0)
    img = get_framebuffer(orc)
    assert oa.oracle_total_luminance(img) == kat["totalLuminance"]   # same float accumulation
    assert np.allclose(img.reshape(-1, 3).astype(np.float64).mean(0), kat["meanRGB"], rtol=0, atol=1e-12)


def test_known_answer_scene0_pt_512():
    """BASELINE config #1 / BASELINE.md §3: scene 0 pt 512x512 1 iteration -> 31074.3."""
    kat = [k for k in json.load(open(os.path.join(GOLD, "ref_kat.json")))
           if (k["scene"], k["alg"], k["width"]) == (0, "pt", 512)][0]
    assert abs(kat["totalLuminance"] - 31074.3) < 0.05
    scene = sv.Scene.cornell_box(0, 512, 512)
    orc = oa.Oracle(scene, abi.ALG_PATH_TRACE, seed=1234, rng_mode=0)
    orc.run_iteration(0)
    assert oa.oracle_total_luminance(get_framebuffer(orc)) == kat["totalLuminance"]


def test_known_answers_256():
    kats = json.load(open(os.path.join(GOLD, "ref_kat.json")))
    for k in kats:
        if k["width"] != 256:
            continue
        scene = sv.Scene.cornell_box(k["scene"], 256, 256)
        orc = oa.Oracle(scene, abi.ALGORITHMS[k["alg"]], seed=1234, rng_mode=0)
        for it in range(k["iterations"]):
            orc.run_iteration(it)
        assert oa.oracle_total_luminance(get_framebuffer(orc)) == k["totalLuminance"], k


def test_rng_stream_definition():
    """Three restatements of the per-path generator agree (product host code, oracle; the
    device code is checked in the gpu tests) and the stream looks uniform."""
    a = sv.rng_stream(1234, 3, 1, 1000, 64, 40)
    b = oa.oracle_rng_stream(1234, 3, 1, 1000, 64, 40)
    assert np.array_equal(a, b)
    assert a.min() > 0.0 and a.max() < 1.0            # strictly inside: no zero-pdf light samples
    big = sv.rng_stream(7, 0, 0, 0, 4096, 64)
    assert abs(big.mean() - 0.5) < 2e-3
    assert abs(big.var() - 1 / 12) < 2e-3
    # different (seed, iteration, phase, path) -> different streams
    assert not np.array_equal(sv.rng_stream(7, 0, 0, 0, 1, 16), sv.rng_stream(7, 0, 1, 0, 1, 16))
    assert not np.array_equal(sv.rng_stream(7, 0, 0, 0, 1, 16), sv.rng_stream(7, 1, 0, 0, 1, 16))
    assert not np.array_equal(sv.rng_stream(7, 0, 0, 0, 1, 16), sv.rng_stream(8, 0, 0, 0, 1, 16))
    # lag-1 correlation inside a path and across neighbouring paths
    x = big - 0.5
    assert abs((x[:, :-1] * x[:, 1:]).mean()) < 2e-3
    assert abs((x[:-1] * x[1:]).mean()) < 2e-3


def test_oracle_openmp_driver_matches_single_renderers():
    """oracle_render (the reference's OpenMP loop, src/smallvcm.cxx:52-151) with 1 thread equals
    running the renderer by hand; with T threads it averages T private renderers."""
    scene = sv.Scene.cornell_box(1, 32, 32)
    _, img = oa.oracle_render(scene, abi.ALG_VCM, 2, 1, base_seed=5, rng_mode=0)
    orc = oa.Oracle(scene, abi.ALG_VCM, seed=5, rng_mode=0)
    orc.run_iteration(0); orc.run_iteration(1)
    assert same_bits(img, get_framebuffer(orc))


def synthetic_queries(verts, normals, mats, n, seed, jitter):
    rng = np.random.default_rng(seed)
    q = np.zeros(n, abi.MERGE_QUERY_DTYPE)
    pick = rng.integers(0, len(verts), n)
    q["position"] = verts["position"][pick] + rng.normal(0, jitter, (n, 3)).astype(np.float32)
    q["normal"] = normals[pick]
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = np.where((d * q["normal"]).sum(1, keepdims=True) < 0, -d, d)
    q["dirFix"] = d
    q["matID"] = mats[pick]
    q["pathLength"] = rng.integers(1, 6, n)
    q["dVCM"] = rng.random(n) * 10
    q["dVM"] = rng.random(n) * 3
    return q


@pytest.mark.ref
@pytest.mark.skipif(not oa.have_ref_ctr(), reason="oracle/_ref not built (no reference checkout)")
@pytest.mark.parametrize("scene_id,alg", [(1, "vcm"), (3, "bpm"), (2, "vcm"), (1, "ppm"), (0, "bpt"), (1, "pt"),
                                          (3, "pt"), (0, "el")])
def test_internals_against_reference_classes(scene_id, alg):
    w = 80
    scene = sv.Scene.cornell_box(scene_id, w, w)
    ref = oa.RefCtr(scene_id, abi.ALGORITHMS[alg], w, w, seed=77)
    orc = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=77, rng_mode=1)
    for it in range(2):
        ref.run_iteration(it)
        orc.run_iteration(it)
    assert same_bits(ref.accum()[0], orc.accum()[0])
    if alg in ("pt", "el"):
        return
    va, vb = ref.light_vertices(), orc.light_vertices()
    for x, y in zip(va, vb):
        assert same_bits(x, y)
    assert ref.iteration_constants(1) == orc.iteration_constants(1)
    if alg in ("vcm", "bpm", "ppm"):
        for x, y in zip(ref.grid(w * w), orc.grid(w * w)):
            assert same_bits(x, y)
        q = synthetic_queries(va[0], va[3], va[2], 3000, 1, 0.004)
        assert same_bits(ref.merge_queries(q), orc.merge_queries(q))


@pytest.mark.ref
@pytest.mark.skipif(not oa.have_ref_ctr(), reason="oracle/_ref not built (no reference checkout)")
def test_device_functions_against_reference_classes():
    """Scene::Intersect/Occluded and BSDF Evaluate/Sample restatements vs the reference's."""
    rng = np.random.default_rng(3)
    for scene_id in range(4):
        scene = sv.Scene.cornell_box(scene_id, 64, 64)
        ref = oa.RefCtr(scene_id, abi.ALG_VCM, 64, 64)
        orc = oa.Oracle(scene, abi.ALG_VCM)
        n = 4000
        org = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays = np.concatenate([org, d], 1)
        for x, y in zip(ref.intersect(rays), orc.intersect(rays)):
            assert same_bits(x, y)
        seg = np.concatenate([org, d, rng.uniform(0.01, 3, (n, 1)).astype(np.float32)], 1)
        assert same_bits(ref.occluded(seg), orc.occluded(seg))
        nmat = scene.desc.numMaterials
        nrm = rng.normal(size=(n, 3)).astype(np.float32); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        fix = rng.normal(size=(n, 3)).astype(np.float32); fix /= np.linalg.norm(fix, axis=1, keepdims=True)
        gen = rng.normal(size=(n, 3)).astype(np.float32); gen /= np.linalg.norm(gen, axis=1, keepdims=True)
        mat = rng.integers(0, nmat, n).astype(np.int32).view(np.float32)[:, None]
        for light in (0, 1):
            items = np.concatenate([nrm, fix, gen, mat], 1)
            assert same_bits(ref.bsdf_evaluate(items, light), orc.bsdf_evaluate(items, light))
            items = np.concatenate([nrm, fix, rng.random((n, 3)).astype(np.float32), mat], 1)
            (a, ea), (b, eb) = ref.bsdf_sample(items, light), orc.bsdf_sample(items, light)
            assert same_bits(a, b) and same_bits(ea, eb)


@pytest.mark.parametrize("scene_id,alg", [(1, "vcm"), (3, "bpm"), (0, "bpt"), (1, "lt"), (2, "ppm"), (0, "pt"), (1, "el")])
def test_parallel_iteration_equals_serial(scene_id, alg):
    """The oracle's path-parallel iteration (what the full-size GPU parity tests call): same light
    vertices in the same order, same path ends, grid and counters as the serial loop that is pinned
    bit-exact to the reference; the image equal up to the order of float additions. A thread count
    that does not divide the path count exercises the ragged blocks."""
    import smallvcm_b200 as sv
    from smallvcm_b200 import abi
    scene = sv.Scene.cornell_box(scene_id, 80, 56)
    ser = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=77, rng_mode=1)
    par = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=77, rng_mode=1)
    for it in range(2):
        ser.run_iteration(it)
        par.run_iteration(it, threads=7)
    a, na = ser.accum(); b, nb = par.accum()
    assert na == nb == 2
    assert np.abs(a.astype(np.float64) - b).max() <= 1e-5 * np.abs(a).max()
    assert ser.counters() == par.counters()
    if alg not in ("pt", "el"):
        va, vb = ser.light_vertices(), par.light_vertices()
        assert all(np.array_equal(x, y) for x, y in zip(va, vb))
        if alg in ("vcm", "bpm", "ppm"):
            ga, gb = ser.grid(80 * 56), par.grid(80 * 56)
            assert all(np.array_equal(x, y) for x, y in zip(ga, gb))
    # the reference stream cannot be split: the call falls back to the serial loop, bit for bit
    r0 = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=77, rng_mode=0)
    r1 = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=77, rng_mode=0)
    r0.run_iteration(0); r1.run_iteration(0, threads=5)
    assert np.array_equal(r0.accum()[0], r1.accum()[0])


def light_camera_inputs(rng, scene, n):
    """Random inputs of the light / camera entry points (shared with tests/test_gpu_kernels.py)."""
    emit = np.zeros((n, 8), np.float32); emit[:, :4] = rng.random((n, 4))
    emit[: n // 50, :4] = rng.integers(0, 2, (n // 50, 4))             # corners of the unit squares
    emit[:, :4] = np.maximum(emit[:, :4], np.float32(2.0 ** -24))      # the path never draws exactly 0 (DESIGN: RNG)
    illum = np.zeros((n, 8), np.float32)
    illum[:, :3] = rng.uniform(-1.3, 1.3, (n, 3)); illum[:, 3:5] = np.maximum(rng.random((n, 2)), np.float32(2.0 ** -24))
    rad = np.zeros((n, 8), np.float32)
    d = rng.normal(size=(n, 3)); rad[:, :3] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rad[:, 3:6] = rng.uniform(-1.3, 1.3, (n, 3))
    cam = np.zeros((n, 5), np.float32)
    cam[:, 0] = rng.uniform(-2, scene.width + 2, n); cam[:, 1] = rng.uniform(-2, scene.height + 2, n)
    cam[:, 2:] = rng.uniform(-1.3, 1.3, (n, 3))
    return emit, illum, rad, cam


@pytest.mark.ref
@pytest.mark.parametrize("scene_id", [0, 1, 2, 3])
def test_lights_and_camera_against_reference_classes(scene_id):
    """AbstractLight::Emit / Illuminate / GetRadiance of all four light types (scene 0 directional,
    1 area x2, 2 point, 3 background) and Camera::GenerateRay / WorldToRaster: the oracle's
    restatements against the reference's classes, bit for bit."""
    if not oa.have_ref_ctr():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(40 + scene_id)
    scene = sv.Scene.cornell_box(scene_id, 96, 64)
    ref = oa.RefCtr(scene_id, abi.ALG_VCM, 96, 64)
    orc = oa.Oracle(scene, abi.ALG_VCM)
    emit, illum, rad, cam = light_camera_inputs(rng, scene, 5000)
    for light in range(scene.desc.numLights):
        for op, items in (("emit", emit), ("illuminate", illum), ("radiance", rad)):
            assert same_bits(ref.light(op, light, items), orc.light(op, light, items)), (light, op)
    assert same_bits(ref.camera(cam), orc.camera(cam))
