"""Kernel-level parity (SURVEY §8c tier T1): each stage of the CUDA path, called on its own
through the C-ABI, against the oracle on the same seeded inputs.

Integer / index / decision results must be EXACT (primitive ids, occlusion flags, sampled
events, bounding box, per-cell membership, candidate and accepted counts); float results
within 1e-5 relative (CUDA's sinf/cosf/powf differ from glibc's in the last ulps and the
device code contracts a*b+c into FMAs)."""
import numpy as np
import pytest

import oracle_api as oa
import smallvcm_b200 as sv
from smallvcm_b200 import abi

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def unit(rng, n):
    v = rng.normal(size=(n, 3)).astype(np.float32)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def close(a, b, rtol=RTOL, atol=1e-6):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b))


def test_rng_device_matches_host_and_oracle():
    for (seed, it, phase, first) in [(1234, 0, 0, 0), (1234, 7, 1, 1 << 20), (-5, 3, 0, 4000000000)]:
        dev = sv.rng_stream_device(seed, it, phase, first, 257, 33)
        assert np.array_equal(dev, sv.rng_stream(seed, it, phase, first, 257, 33))
        assert np.array_equal(dev, oa.oracle_rng_stream(seed, it, phase, first, 257, 33))
    assert sv.rng_stream_device(1, 0, 0, 0, 0, 4).shape == (0, 4)       # empty input


@pytest.mark.parametrize("scene_id", [0, 1, 2, 3])
def test_intersect_and_occluded(scene_id):
    rng = np.random.default_rng(10 + scene_id)
    scene = sv.Scene.cornell_box(scene_id, 64, 64)
    orc = oa.Oracle(scene, abi.ALG_VCM)
    n = 200000
    org = rng.uniform(-1.25, 1.25, (n, 3)).astype(np.float32)
    org[: n // 4] = np.array(scene.desc.camera.position, np.float32)      # camera rays
    rays = np.concatenate([org, unit(rng, n)], 1)
    g = sv.intersect_batch(scene, rays)
    o = orc.intersect(rays)
    # closest hit: same primitive, material and light; distance and normal to float accuracy
    mism = g[0] != o[0]
    assert mism.mean() < 2e-5, mism.sum()               # exact ties broken by ulps only
    ok = ~mism
    assert np.array_equal(g[3][ok], o[3][ok]) and np.array_equal(g[4][ok], o[4][ok])
    hit = ok & (o[0] >= 0)
    # distances: FMA contraction in the triangle test moves them by ulps; allow 1e-5 relative
    # for all but a sliver of badly conditioned (near-parallel) rays, 1e-3 for every ray
    dist_ok = close(g[1][hit], o[1][hit], rtol=1e-5)
    assert dist_ok.mean() > 0.999 and close(g[1][hit], o[1][hit], rtol=1e-3).all()
    assert close(g[2][hit], o[2][hit], rtol=1e-5, atol=2e-6).mean() > 0.999
    # sphere hits are evaluated unfused like the reference: bit-identical distance and normal
    prim_type = np.array([scene.desc.primitives[i].type for i in range(scene.desc.numPrimitives)])
    sph = hit & (prim_type[np.maximum(o[0], 0)] == abi.PRIM_SPHERE)
    if sph.any():
        assert np.array_equal(g[1][sph], o[1][sph]) and np.array_equal(g[2][sph], o[2][sph])
    assert (g[1][ok & (o[0] < 0)] == np.float32(1e36)).all()

    seg = np.concatenate([org, unit(rng, n), rng.uniform(0.002, 3.0, (n, 1)).astype(np.float32)], 1)
    go, oo = sv.occluded_batch(scene, seg), orc.occluded(seg)
    assert (go != oo).mean() < 2e-5
    assert 0.05 < go.mean() < 0.95


@pytest.mark.parametrize("scene_id", [0, 1, 3])
def test_primitive_runs_never_change_a_hit(scene_id, monkeypatch):
    """The constant-bank list is cut into runs and a ray skips a run of small triangles whose padded
    box it misses (DevScene::runs). Skipping must never change a result: closest hits and occlusion
    flags with the runs and with the whole list in one always-tested run (SVCM_DEBUG_NO_RUNS=1) are
    BIT-identical, on random rays and on rays built to graze, start on, end on and run along the
    faces of the light box — the run that is skipped in the reference's scenes."""
    rng = np.random.default_rng(77 + scene_id)
    scene = sv.Scene.cornell_box(scene_id, 64, 64)
    n = 100000
    lo = np.array([-0.25, -0.25, 1.26002], np.float32); hi = np.array([0.25, 0.25, 1.28002], np.float32)
    inside = rng.uniform(-1.25, 1.25, (n, 3)).astype(np.float32)
    # points on the faces of the light box (one coordinate snapped to a face), and just off them
    on_box = (lo + (hi - lo) * rng.uniform(0, 1, (n, 3))).astype(np.float32)
    axis = rng.integers(0, 3, n); side = rng.integers(0, 2, n)
    on_box[np.arange(n), axis] = np.where(side == 1, hi[axis], lo[axis])
    near_box = (on_box + rng.normal(scale=2e-3, size=(n, 3))).astype(np.float32)
    rays, segs = [], []

    def toward(a, b):
        d = (b - a).astype(np.float32); ln = np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
        return d / np.maximum(ln, 1e-20), ln
    for org, dst in [(inside, rng.uniform(-1.25, 1.25, (n, 3)).astype(np.float32)),      # anywhere to anywhere
                     (inside, on_box), (inside, near_box),                              # ending on / near the box
                     (on_box, inside), (near_box, inside),                              # starting on / near the box
                     (on_box, np.roll(on_box, 1, 0)), (near_box, np.roll(on_box, 7, 0))]:   # along the faces
        d, ln = toward(org, dst)
        rays.append(np.concatenate([org, d], 1))
        for scale in (1.0, 0.5, 1.5):
            segs.append(np.concatenate([org, d, (ln * scale).astype(np.float32)], 1))
    # axis-parallel rays in the planes of the box faces (zero direction components: infinite slab reciprocals)
    d = np.zeros((n, 3), np.float32); d[np.arange(n), (axis + 1) % 3] = np.where(rng.integers(0, 2, n) == 1, 1.0, -1.0)
    start = on_box - d * rng.uniform(0, 2, (n, 1)).astype(np.float32)
    rays.append(np.concatenate([start, d], 1)); segs.append(np.concatenate([start, d, rng.uniform(0.01, 3, (n, 1)).astype(np.float32)], 1))
    rays = np.concatenate(rays); segs = np.concatenate(segs)

    with_runs = sv.intersect_batch(scene, rays), sv.occluded_batch(scene, segs)
    monkeypatch.setenv("SVCM_DEBUG_NO_RUNS", "1")
    without = sv.intersect_batch(scene, rays), sv.occluded_batch(scene, segs)
    for a, b in zip(with_runs[0], without[0]):
        assert np.array_equal(a, b)
    assert np.array_equal(with_runs[1], without[1])
    hits_box = np.isin(with_runs[0][0], np.arange(scene.desc.numPrimitives - 10, scene.desc.numPrimitives))
    assert hits_box.mean() > 0.05 and 0.02 < with_runs[1].mean() < 0.98          # the box itself is hit often enough


@pytest.mark.parametrize("scene_id", [0, 1])
@pytest.mark.parametrize("fix_is_light", [0, 1])
def test_bsdf_evaluate_and_sample(scene_id, fix_is_light):
    rng = np.random.default_rng(20 + scene_id)
    scene = sv.Scene.cornell_box(scene_id, 64, 64)
    orc = oa.Oracle(scene, abi.ALG_VCM)
    n = 100000
    nrm, fix, gen = unit(rng, n), unit(rng, n), unit(rng, n)
    # make a third of the generated directions cluster around the mirror direction (Phong lobe)
    refl = 2 * (fix * nrm).sum(1, keepdims=True) * nrm - fix
    k = n // 3
    gen[:k] = refl[:k] + 0.05 * unit(rng, k)
    gen /= np.linalg.norm(gen, axis=1, keepdims=True)
    mat = rng.integers(0, scene.desc.numMaterials, n).astype(np.int32).view(np.float32)[:, None]

    items = np.concatenate([nrm, fix, gen, mat], 1).astype(np.float32)
    g, o = sv.bsdf_evaluate_batch(scene, items, fix_is_light), orc.bsdf_evaluate(items, fix_is_light)
    assert np.array_equal(g[:, 7], o[:, 7])                        # valid / delta flags
    zero_mismatch = (np.abs(g[:, :3]).sum(1) == 0) != (np.abs(o[:, :3]).sum(1) == 0)
    assert zero_mismatch.mean() < 1e-4                             # epsilon-branch flips only
    ok = ~zero_mismatch
    # Lambertian-only materials: 1e-5 relative. The glossy floor (material 2, Phong exponent
    # 90) amplifies the ulp-level difference of dot(R, w) by the exponent: d(x^90)/x^90 =
    # 90 dx/x, so 1e-7 absolute on the dot product is 2e-5 .. 1e-3 relative on the lobe.
    phong = items[:, 9].view(np.int32) == 2
    assert close(g[ok & ~phong][:, :7], o[ok & ~phong][:, :7], rtol=1e-5).all()
    assert close(g[ok & phong][:, :7], o[ok & phong][:, :7], rtol=2e-3).all()
    assert close(g[ok & phong][:, :7], o[ok & phong][:, :7], rtol=1e-4).mean() > 0.99

    items = np.concatenate([nrm, fix, rng.random((n, 3)).astype(np.float32), mat], 1).astype(np.float32)
    (g, ge), (o, oe) = sv.bsdf_sample_batch(scene, items, fix_is_light), orc.bsdf_sample(items, fix_is_light)
    flips = ge != oe
    assert flips.mean() < 1e-4
    ok = ~flips
    phong = items[:, 9].view(np.int32) == 2
    # Conditioning: the specular lobes return coeff / |cos(theta)| (src/bsdf.hxx:331-332,
    # 379-381), so at grazing incidence an ulp on cos is a large relative error of the value;
    # judge the well-conditioned samples strictly and the grazing tail loosely.
    cos_fix = np.abs((nrm * fix).sum(1))
    cos_gen = np.maximum(o[:, 7], 1e-30)
    well = ok & (cos_fix > 0.2) & ((cos_gen > 0.2) | (oe == 0))
    assert well.mean() > 0.5
    assert close(g[well & ~phong], o[well & ~phong], rtol=1e-4, atol=2e-6).all()
    assert close(g[well & phong], o[well & phong], rtol=2e-3, atol=2e-6).all()
    assert close(g[ok], o[ok], rtol=2e-2, atol=2e-6).all()
    assert close(g[ok], o[ok], rtol=1e-4, atol=2e-6).mean() > 0.99
    assert (oe > 0).mean() > 0.3


def oracle_iteration(scene_id, alg, w, iteration=0, seed=1234):
    scene = sv.Scene.cornell_box(scene_id, w, w)
    orc = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=seed, rng_mode=1)
    for it in range(iteration + 1):
        orc.run_iteration(it)
    return scene, orc


def cells_as_sets(ends, idx):
    starts = np.concatenate([[0], ends[:-1]])
    return [frozenset(idx[s:e].tolist()) for s, e in zip(starts, ends) if e > s]


LAYOUTS = ["bucket", "cell"]


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("scene_id,alg,w", [(1, "vcm", 96), (3, "bpm", 128), (2, "vcm", 64)])
def test_hashgrid_build_exact(scene_id, alg, w, layout, monkeypatch):
    """HashGrid::Build: bbox, mCellEnds and per-cell membership identical to the oracle — in the
    reference's bucket order, and restated from the cell-ordered layout (the device supplies the
    bucket of every sorted particle and the per-bucket histogram)."""
    monkeypatch.setenv("SVCM_GRID_LAYOUT", layout)
    scene, orc = oracle_iteration(scene_id, alg, w)
    verts = orc.light_vertices()[0]
    radius = orc.iteration_constants(0)[0]
    obbox, oends, oidx = orc.grid(w * w)
    grid = sv.hashgrid_build(verts, radius, w * w)
    bbox, ends, idx = grid.download()
    assert np.array_equal(bbox, obbox)
    assert np.array_equal(ends, oends)                       # histogram + scan, exact
    assert sorted(idx.tolist()) == list(range(len(verts)))   # a permutation
    assert cells_as_sets(ends, idx) == cells_as_sets(oends, oidx)
    # size-independent properties
    assert ends[-1] == len(verts) and (np.diff(ends) >= 0).all()


@pytest.mark.parametrize("layout,variant", [("bucket", "1"), ("bucket", "2"), ("bucket", "3"), ("bucket", "4"),
                                            ("cell", "0"), ("cell", "1"), ("cell", "3")])
def test_hashgrid_edge_cases(layout, variant, monkeypatch):
    monkeypatch.setenv("SVCM_GRID_LAYOUT", layout)
    monkeypatch.setenv("SVCM_MERGE_VARIANT", variant)
    monkeypatch.setenv("SVCM_MERGE_GROUP", variant)
    scene = sv.Scene.cornell_box(1, 32, 32)
    # empty vertex set: bbox stays (+1e36, -1e36), every query is rejected
    grid = sv.hashgrid_build(np.zeros(0, abi.MERGE_VERTEX_DTYPE), 0.01, 64)
    bbox, ends, idx = grid.download()
    assert (bbox[:3] == np.float32(1e36)).all() and (bbox[3:] == np.float32(-1e36)).all()
    assert (ends == 0).all() and len(idx) == 0
    q = np.zeros(3, abi.MERGE_QUERY_DTYPE); q["normal"] = [0, 0, 1]; q["dirFix"] = [0, 0, 1]; q["matID"] = 5
    contrib, counts = grid.query(scene, q, 0, 10, 0.1, False)
    assert (contrib == 0).all() and (counts == 0).all()
    # one cell only (numCells = 1): everything collides, the 8 visited "cells" are the same
    # bucket -> every particle is visited 8 times, like the reference
    v = np.zeros(5, abi.MERGE_VERTEX_DTYPE)
    v["position"] = np.array([[0, 0, 0], [0.001, 0, 0], [0.5, 0.5, 0.5], [0, 0.002, 0], [1, 1, 1]], np.float32)
    v["throughput"] = 1; v["dirFix"] = [0, 0, 1]; v["pathLength"] = 1; v["dVCM"] = 1; v["dVM"] = 1; v["contProb"] = 0.8
    grid = sv.hashgrid_build(v, 0.01, 1)
    orc = oa.Oracle(scene, abi.ALG_VCM)
    orc.set_light_vertices(v, 0.01, 1, 0.1, False, 10, 0)
    q = np.zeros(2, abi.MERGE_QUERY_DTYPE)
    q["position"] = [[0.0005, 0.0005, 0.0], [0.9, 0.9, 0.9]]
    q["normal"] = [0, 0, 1]; q["dirFix"] = [0.1, 0.1, 0.98]; q["matID"] = 5; q["pathLength"] = 2; q["dVCM"] = 1; q["dVM"] = 1
    gc, gn = grid.query(scene, q, 0, 10, 0.1, False)
    oc, on = orc.merge_queries(q, True)
    assert np.array_equal(gn, on) and gn[0, 0] == 40 and gn[0, 1] == 24
    assert close(gc, oc).all()


@pytest.mark.parametrize("merge_kernel", ["1", "2", "3", "4", "5", "8", "32", "cell0", "cell1", "cell3"])
@pytest.mark.parametrize("scene_id,alg,w,ppm", [(1, "vcm", 128, False), (3, "bpm", 128, False), (1, "ppm", 96, True),
                                                (0, "vcm", 96, False)])
def test_hashgrid_query_against_oracle(scene_id, alg, w, ppm, merge_kernel, monkeypatch):
    """HashGrid::Process + RangeQuery::Process on the oracle's light vertices: visited and
    accepted counts exact, merged radiance within 1e-5 relative."""
    from test_oracle import synthetic_queries
    # every merge kernel is product code: bucket layout (lane-per-query with compaction, cooperative
    # groups) and cell layout (lane-per-query, staged with bulk copies / plain loads / small stages)
    if merge_kernel.startswith("cell"):
        monkeypatch.setenv("SVCM_GRID_LAYOUT", "cell")
        monkeypatch.setenv("SVCM_MERGE_VARIANT", merge_kernel[4:])
    else:
        monkeypatch.setenv("SVCM_GRID_LAYOUT", "bucket")
        monkeypatch.setenv("SVCM_MERGE_GROUP", merge_kernel)
    scene, orc = oracle_iteration(scene_id, alg, w)
    verts, _, mats, normals, _ = orc.light_vertices()
    radius, _, _, mis_vc = orc.iteration_constants(0)
    n = 60000
    q = synthetic_queries(verts, normals, mats, n, 5, radius * 0.6)
    q["position"][: n // 20] += 10.0                        # some far outside the bbox
    q["matID"][n // 20: n // 10] = 6                        # some on a purely specular material
    orc.set_light_vertices(verts, radius, w * w, mis_vc, ppm, 10, 0)
    oc, on = orc.merge_queries(q, True)
    grid = sv.hashgrid_build(verts, radius, w * w)
    gc, gn = grid.query(scene, q, 0, 10, mis_vc, ppm)
    assert np.array_equal(gn, on)                           # N_cand and N_acc per query, exact
    assert on[:, 1].sum() > n // 4
    scale = np.abs(oc).max(1, keepdims=True) + 1e-30
    assert (np.abs(gc - oc) / scale).max() < 2e-5
    # path-length window (min/max) is honoured
    oc2 = None
    orc.set_light_vertices(verts, radius, w * w, mis_vc, ppm, 4, 3)
    oc2, _ = orc.merge_queries(q, True)
    gc2, _ = grid.query(scene, q, 3, 4, mis_vc, ppm)
    assert (np.abs(gc2 - oc2) / (np.abs(oc2).max(1, keepdims=True) + 1e-30)).max() < 2e-5
    assert np.abs(oc2).sum() < np.abs(oc).sum()


@pytest.mark.parametrize("layout,variant", [("bucket", "1"), ("bucket", "2"), ("bucket", "3"), ("cell", "0"), ("cell", "1")])
@pytest.mark.parametrize("num_cells", [2, 7, 64, 1000])
def test_hashgrid_bucket_collisions(num_cells, layout, variant, monkeypatch):
    """With few buckets most queries see two or more of their 8 cells share a bucket; the reference
    visits such a bucket once per cell (src/hashgrid.hxx:142-155) and counts its in-range particles
    that often. Visited and accepted counts must stay exact in every layout."""
    from test_oracle import synthetic_queries
    monkeypatch.setenv("SVCM_GRID_LAYOUT", layout)
    monkeypatch.setenv("SVCM_MERGE_VARIANT", variant)
    monkeypatch.setenv("SVCM_MERGE_GROUP", variant)
    scene, orc = oracle_iteration(1, "vcm", 64)
    verts, _, mats, normals, _ = orc.light_vertices()
    radius, _, _, mis_vc = orc.iteration_constants(0)
    radius *= 2.0
    q = synthetic_queries(verts, normals, mats, 20000, 7, radius * 0.6)
    orc.set_light_vertices(verts, radius, num_cells, mis_vc, False, 10, 0)
    oc, on = orc.merge_queries(q, True)
    grid = sv.hashgrid_build(verts, radius, num_cells)
    bbox, ends, idx = grid.download()
    obbox, oends, oidx = orc.grid(num_cells)
    assert np.array_equal(ends, oends) and cells_as_sets(ends, idx) == cells_as_sets(oends, oidx)
    gc, gn = grid.query(scene, q, 0, 10, mis_vc, False)
    assert np.array_equal(gn, on)
    # collisions really happen: more accepted visits than distinct in-range particles
    orc.set_light_vertices(verts, radius, 1 << 24, mis_vc, False, 10, 0)
    _, on_sparse = orc.merge_queries(q, True)
    assert on[:, 1].sum() > on_sparse[:, 1].sum()
    scale = np.abs(oc).max(1, keepdims=True) + 1e-30
    assert (np.abs(gc - oc) / scale).max() < 2e-5


@pytest.mark.parametrize("scene_id", [0, 1, 2, 3])
def test_lights_and_camera(scene_id):
    """AbstractLight::Emit / Illuminate / GetRadiance for every light of the four scenes
    (directional, area, point, background) and Camera::GenerateRay / WorldToRaster on random
    inputs: device functions vs the oracle (itself bit-exact to the reference classes,
    tests/test_oracle.py). Decisions (zero radiance, inside-the-image) exact, values 1e-5."""
    from test_oracle import light_camera_inputs
    rng = np.random.default_rng(70 + scene_id)
    scene = sv.Scene.cornell_box(scene_id, 96, 64)
    orc = oa.Oracle(scene, abi.ALG_VCM)
    emit, illum, rad, cam = light_camera_inputs(rng, scene, 100000)
    for light in range(scene.desc.numLights):
        for op, items in (("emit", emit), ("illuminate", illum), ("radiance", rad)):
            g, o = sv.light_batch(scene, op, light, items), orc.light(op, light, items)
            assert np.array_equal((g[:, :3] == 0).all(1), (o[:, :3] == 0).all(1)), (light, op)   # same zero / non-zero decision
            ok = close(g, o, rtol=1e-5, atol=2e-6)
            # sin/cos of the sphere and disc samplers: a handful of arguments next to a multiple of
            # pi/2 lose relative accuracy in a component that is ~0 there
            assert ok.all(1).mean() > 0.9999 and close(g, o, rtol=1e-3, atol=2e-5).all(), (light, op, (~ok).sum())
    g, o = sv.camera_ray_batch(scene, cam), orc.camera(cam)
    border = np.minimum(np.abs(o[:, 3:5] - np.round(o[:, 3:5])).min(1), 1.0) < 1e-3     # raster within 1e-3 of a pixel edge
    assert np.array_equal(g[~border, 5], o[~border, 5])
    assert close(g[:, :3], o[:, :3], rtol=1e-5, atol=1e-6).all()
    assert close(g[:, 3:5], o[:, 3:5], rtol=2e-5, atol=1e-3).all()
    assert sv.light_batch(scene, "emit", 0, np.zeros((0, 8), np.float32)).shape == (0, 12)     # empty input
    with pytest.raises(sv.SvcmError):
        sv.light_batch(scene, "emit", 7, emit[:4])
