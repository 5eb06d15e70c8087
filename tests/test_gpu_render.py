"""Path-level and statistical parity of the whole iteration (SURVEY §8c tiers T2 and T3) and
size-independent properties at BASELINE.json's full sizes — all through the C-ABI.

Tolerances (stated here, used below):
  T2  GPU vs oracle, SAME per-path random stream (seed, iteration, phase, path): the images
      agree up to rare epsilon-branch flips and libm ulps that send single paths elsewhere:
        relative L2  ||gpu - oracle|| / ||oracle||  <= 6e-2   (measured 1e-3 .. 3.5e-2; the
                     norm is dominated by the handful of flipped paths, each a firefly)
        fraction of pixels off by > 0.1 % of the pixel's value   <= 1 %
        mean image value within 5e-3 relative; stored light vertices within 1e-4 relative
  T3  GPU vs the UNMODIFIED reference at the same mBaseSeed (different generator, so only
      statistics can agree), N = 16 iterations at 128x128, reference seeds 1234 and 1235 (a, b):
        per-pixel L1 distance  |gpu - a|_1 and |gpu - b|_1   <= 1.2  x |a - b|_1
        8x8-block means        |blk(gpu) - blk((a+b)/2)|_1   <= 1.15 x |blk(a) - blk(b)|_1
                               (0.87 expected for three iid renders; a missing or doubled
                               transport component shows up here first)
        mean luminance within 2 % (scenes 0, 1) / 10 % (scene 3, firefly-dominated: the
        reference's own seeds spread 0.89 .. 0.97 at this sample count)
      L1 instead of the survey's L2 because at 16 samples/pixel the L2 norm is carried by a few
      fireflies: the unmodified reference's OWN other seeds score 0.72 .. 1.25 x on L2 against
      (a, b) but 0.95 .. 1.11 x on L1 (measured with the oracle in reference-stream mode).
"""
import os

import numpy as np
import pytest

import oracle_api as oa
import smallvcm_b200 as sv
from smallvcm_b200 import abi

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LUM = np.array([0.212671, 0.715160, 0.072169])


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) /
                 np.linalg.norm(b.astype(np.float64)))


def render(scene, alg, iterations, seed=1234, max_len=10, min_len=0, first=0, layout=None, variant=None, **kw):
    r = sv.VertexCM(scene, alg, seed=seed, **kw)
    r.max_path_length, r.min_path_length = max_len, min_len
    # mergeCandidates in the reference's convention (buckets visited) whatever the particle order
    r.set_option(abi.OPT_REF_CANDIDATES, 1)
    if layout is not None:
        r.set_option(abi.OPT_GRID_LAYOUT, layout)
    if variant is not None:
        r.set_option(abi.OPT_MERGE_VARIANT, variant)
    r.run_iterations(first, iterations)
    img, n = r.accum()
    st = r.stats()
    r.close()
    return img, n, st


@pytest.mark.parametrize("scene_id,alg", [(1, "vcm"), (0, "bpt"), (3, "bpm"), (2, "vcm"), (1, "lt"), (1, "ppm"),
                                          (0, "vcm"), (3, "vcm"), (2, "bpm"), (0, "lt"),
                                          (0, "pt"), (1, "pt"), (2, "pt"), (3, "pt"), (1, "el"), (3, "el")])
def test_t2_same_stream_parity(scene_id, alg):
    w, iters = 128, 3
    scene = sv.Scene.cornell_box(scene_id, w, w)
    img, n, st = render(scene, alg, iters)
    orc = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=1234, rng_mode=1)
    for it in range(iters):
        orc.run_iteration(it)
    ref, _ = orc.accum()
    cnt = orc.counters()
    assert n == iters
    pix = np.abs(img - ref).max(2) / (np.abs(ref).max(2) + 1e-3)
    _log_measured("t2_128", scene=scene_id, alg=alg, rel_l2=rel_l2(img, ref), pixels_off=float((pix > 1e-3).mean()))
    assert rel_l2(img, ref) <= 6e-2
    assert (pix > 1e-3).mean() <= 1e-2
    assert abs(img.mean() - ref.mean()) <= 5e-3 * ref.mean()
    # workload counters of the last iteration: vertices exact, the rest to within branch flips
    assert abs(st["lightVertices"] - orc.num_light_vertices()) <= 1e-4 * orc.num_light_vertices() + 2
    for key, okey in [("lightRays", "lightRays"), ("cameraRays", "cameraRays"), ("shadowRays", "shadowRays"),
                      ("mergeQueries", "mergeQueries"), ("mergeCandidates", "mergeCandidates"),
                      ("mergeAccepted", "mergeAccepted")]:
        assert abs(st[key] - cnt[okey]) <= 2e-3 * cnt[okey] + 8, (key, st[key], cnt[okey])


@pytest.mark.parametrize("variant", [0, 1, 3])
@pytest.mark.parametrize("scene_id,alg", [(1, "vcm"), (3, "bpm"), (1, "ppm")])
def test_t2_cell_layout_parity(scene_id, alg, variant):
    """T2 with the particles in true cell order (every merge variant) and, for comparison, in the
    reference's bucket order: same bars, and the two layouts agree with each other to summation order."""
    w, iters = 128, 3
    scene = sv.Scene.cornell_box(scene_id, w, w)
    img, n, st = render(scene, alg, iters, layout=abi.GRID_CELL, variant=variant)
    bucket, _, stb = render(scene, alg, iters, layout=abi.GRID_BUCKET)
    orc = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=1234, rng_mode=1)
    for it in range(iters):
        orc.run_iteration(it)
    ref, _ = orc.accum()
    cnt = orc.counters()
    assert rel_l2(img, ref) <= 6e-2
    assert rel_l2(img, bucket) < 1e-5
    pix = np.abs(img - ref).max(2) / (np.abs(ref).max(2) + 1e-3)
    assert (pix > 1e-3).mean() <= 1e-2
    for key in ("lightVertices", "mergeQueries", "mergeCandidates", "mergeAccepted"):
        assert st[key] == stb[key], (key, st[key], stb[key])
    for key in ("mergeQueries", "mergeCandidates", "mergeAccepted"):
        assert abs(st[key] - cnt[key]) <= 2e-3 * cnt[key] + 8, (key, st[key], cnt[key])


@pytest.mark.parametrize("scene_id,alg", [(1, "vcm"), (0, "vcm"), (3, "bpt")])
def test_primitive_runs_do_not_change_a_single_path(scene_id, alg, monkeypatch):
    """Whole iterations with the skippable primitive runs (the default) and with the plain list
    (SVCM_DEBUG_NO_RUNS=1): every workload counter is IDENTICAL — no path took a different branch —
    and the images agree to the order of the float atomics."""
    scene = sv.Scene.cornell_box(scene_id, 256, 256)
    a_img, _, a_st = render(scene, alg, 4)
    monkeypatch.setenv("SVCM_DEBUG_NO_RUNS", "1")
    b_img, _, b_st = render(scene, alg, 4)
    for key in ("lightVertices", "mergeQueries", "mergeCandidates", "mergeAccepted", "cameraRays", "lightRays", "shadowRays"):
        assert a_st[key] == b_st[key], key
    assert rel_l2(a_img, b_img) <= 1e-5


def test_t2_path_length_limits():
    """min/max path length windows (src/vertexcm.hxx:133-135,382,387,438,472,482,489,512-521)."""
    scene = sv.Scene.cornell_box(1, 96, 96)
    for (mn, mx) in [(0, 2), (0, 1), (3, 5), (2, 2), (0, 20)]:
        img, _, _ = render(scene, "vcm", 2, max_len=mx, min_len=mn)
        orc = oa.Oracle(scene, abi.ALG_VCM, seed=1234, rng_mode=1, max_path_length=mx, min_path_length=mn)
        orc.run_iteration(0); orc.run_iteration(1)
        ref, _ = orc.accum()
        assert rel_l2(img, ref) <= 6e-2, (mn, mx)
        assert abs(img.mean() - ref.mean()) <= 3e-3 * ref.mean() + 1e-6


def test_t2_nonsquare_and_iteration_offset():
    """Radius schedule uses the GLOBAL iteration index (src/vertexcm.hxx:295-298); W != H."""
    scene = sv.Scene.cornell_box(1, 160, 72)
    r = sv.VertexCM(scene, "vcm", seed=99)
    r.max_path_length = 8
    r.run_iterations(5, 2)
    img, n = r.accum()
    st = r.stats()
    orc = oa.Oracle(scene, abi.ALG_VCM, seed=99, rng_mode=1, max_path_length=8)
    orc.run_iteration(5); orc.run_iteration(6)
    ref, _ = orc.accum()
    assert n == 2 and img.shape == (72, 160, 3)
    assert abs(st["radius"] - orc.iteration_constants(6)[0]) < 1e-9
    assert rel_l2(img, ref) <= 6e-2
    r.close()


def test_accumulate_clear_and_linearity():
    """mFramebuffer semantics: iterations add up; Clear() resets; running iterations one by one
    equals running them in a batch (float atomics allow ulp-level differences)."""
    scene = sv.Scene.cornell_box(1, 96, 96)
    r = sv.VertexCM(scene, "vcm"); r.max_path_length = 10
    r.run_iterations(0, 3)
    batch, n = r.accum()
    assert n == 3
    r.clear()
    parts = np.zeros_like(batch, np.float64)
    for it in range(3):
        r.run_iteration(it)
        img, k = r.accum()
        assert k == 1
        parts += img
        r.clear()
    assert rel_l2(parts, batch) < 1e-5
    assert r.accum()[1] == 0 and float(np.abs(r.accum()[0]).max()) == 0.0
    fb = sv.VertexCM(scene, "vcm"); fb.max_path_length = 10
    assert not fb.was_used()
    fb.run_iterations(0, 3)
    assert fb.was_used()
    assert rel_l2(fb.get_framebuffer() * 3, batch) < 1e-5
    r.close(); fb.close()


@pytest.mark.parametrize("scene_id,alg,w", [(1, "vcm", 160), (3, "bpm", 128), (0, "ppm", 96), (0, "bpt", 96), (1, "lt", 96)])
@pytest.mark.parametrize("single_buffer", [0, 1])
def test_stream_schedules_agree(scene_id, alg, w, single_buffer, monkeypatch):
    """SVCM_OPT_OVERLAP: every schedule of the stages — one after the other (0), the three overlaps
    inside an iteration (7), the merge of an iteration left running under the next iteration's light
    pass (15, the default), iterations enqueued in one call, one call each, or with an entry point that
    joins in between — renders the same image and reports the same per-iteration counters."""
    # the camera vertices exist twice when memory allows (the next walk writes one set while the pending
    # merge reads the other); SVCM_SINGLE_QUERY_BUFFER=1 is the memory-bounded arrangement
    monkeypatch.setenv("SVCM_SINGLE_QUERY_BUFFER", str(single_buffer))
    scene = sv.Scene.cornell_box(scene_id, w, w)
    results = []
    for mask, mode in [(0, "batch"), (7, "batch"), (15, "batch"), (15, "single"), (15, "joined"), (8, "batch")]:
        r = sv.VertexCM(scene, alg); r.max_path_length = 8
        r.set_option(abi.OPT_OVERLAP, mask)
        r.set_option(abi.OPT_PROFILE, 1)
        if mode == "batch":
            r.run_iterations(0, 7)
        else:
            for it in range(7):
                r.run_iterations(it, 1)
                if mode == "joined":
                    if it % 2: r.event_record(0)
                    else: r.synchronize()
        prof = r.profile()
        img, n = r.accum()
        st = r.stats()
        r.close()
        assert n == 7 and prof["iterations"] == 7
        results.append((img, prof, st))
    img0, prof0, st0 = results[0]
    for img, prof, st in results[1:]:
        assert rel_l2(img, img0) < 1e-5
        for key in ("lightVertices", "mergeQueries", "mergeCandidates", "mergeAccepted", "lightRays", "cameraRays", "shadowRays"):
            assert prof[key] == prof0[key], key
            if key in st: assert st[key] == st0[key], key


def test_merge_group_sizes_agree():
    scene = sv.Scene.cornell_box(3, 128, 128)
    imgs = []
    for group in (1, 2, 3, 4, 5, 8, 16, 32):
        r = sv.VertexCM(scene, "bpm"); r.max_path_length = 10
        r.set_option(abi.OPT_MERGE_GROUP, group)
        r.run_iterations(0, 2)
        imgs.append(r.accum()[0])
        r.close()
    assert max(rel_l2(imgs[k], imgs[0]) for k in range(1, len(imgs))) < 1e-5


@pytest.mark.parametrize("scene_id,alg", [(1, "vcm"), (3, "bpm"), (0, "bpt"), (1, "pt")])
def test_t3_statistical_parity_with_unmodified_reference(scene_id, alg):
    """Golden: the unmodified reference at 128x128, 16 iterations, seeds 1234 and 1235."""
    gold = np.load(os.path.join(GOLD, "ref_noise.npz"))
    a = gold["s%d_%s_seed1234" % (scene_id, alg)].astype(np.float32)
    b = gold["s%d_%s_seed1235" % (scene_id, alg)].astype(np.float32)
    scene = sv.Scene.cornell_box(scene_id, 128, 128)
    r = sv.VertexCM(scene, alg, seed=1234); r.max_path_length = 10
    r.run_iterations(0, 16)
    img = r.get_framebuffer()
    r.close()
    l1 = lambda x, y: float(np.abs(x.astype(np.float64) - y).sum())
    assert l1(img, a) <= 1.2 * l1(a, b)
    assert l1(img, b) <= 1.2 * l1(a, b)

    def blocks(x, k=8):
        h, w, c = x.shape
        return x.reshape(h // k, k, w // k, k, c).mean((1, 3)).astype(np.float64)
    A, B, G = blocks(a), blocks(b), blocks(img)
    assert np.abs(G - 0.5 * (A + B)).sum() <= 1.15 * np.abs(A - B).sum()

    lum = lambda x: float((x.astype(np.float64) @ LUM).mean())
    ref_lum = 0.5 * (lum(a) + lum(b))
    assert abs(lum(img) - ref_lum) <= (0.10 if scene_id == 3 else 0.02) * ref_lum


def _log_measured(name, **values):
    """Measured parity figures, kept next to the run (gpurun_out/ travels back from the GPU box)."""
    try:
        import json
        d = os.path.join(os.path.dirname(GOLD), "..", "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **values)) + "\n")
    except Exception:
        pass


# (scene, algorithm, width, maxLen, iterations, bound on relative L2): BASELINE.json configs #2, the
# metric configuration, #3 and #4 at FULL size against the oracle on the same per-path random
# streams. The oracle runs its path-parallel iteration (tests/test_oracle.py pins it to the serial
# loop, which is pinned bit-exact to the reference). Bounds are ~3x what was measured on B200
# (profiles/r2_parity_measured.jsonl: rel. L2 3.4e-4 / 3.1e-4 / 1.6e-2 / 1.7e-2, mean radiance equal
# to 1e-5 .. 1e-6). The L2 norm is carried by the few paths an epsilon branch sends elsewhere; a
# flipped LIGHT path moves every pixel whose query merges with it by a little, so the share of
# pixels off by more than 0.1 % grows with the resolution (0.9 % .. 2.1 % measured) while the share
# off by more than 1 % stays small.
# (BPM at 2048^2 has one merge query per pixel with ~14 accepted photons: a flipped photon moves its
# pixels by several per cent, 1.3 % of the pixels are off by more than 1 % there)
FULL_SIZE = [(1, "vcm", 512, 10, 2, 1e-3, 1e-2), (1, "vcm", 1024, 10, 1, 1e-3, 1e-2), (0, "bpt", 1024, 10, 1, 5e-2, 1e-2),
             (3, "bpm", 2048, 10, 1, 5e-2, 4e-2)]


@pytest.mark.parametrize("scene_id,alg,w,max_len,iters,l2_bound,off1_bound", FULL_SIZE,
                         ids=["config2_s1_vcm_512", "metric_s1_vcm_1024", "config3_s0_bpt_1024", "config4_s3_bpm_2048"])
def test_t2_full_size_same_stream(scene_id, alg, w, max_len, iters, l2_bound, off1_bound):
    scene = sv.Scene.cornell_box(scene_id, w, w)
    img, n, st = render(scene, alg, iters, max_len=max_len)
    orc = oa.Oracle(scene, abi.ALGORITHMS[alg], seed=1234, rng_mode=1, max_path_length=max_len)
    for it in range(iters):
        orc.run_iteration(it, threads=0)
    ref, _ = orc.accum()
    cnt = orc.counters()
    l2 = rel_l2(img, ref)
    pix = np.abs(img - ref).max(2) / (np.abs(ref).max(2) + 1e-3)
    off = float((pix > 1e-3).mean())
    off1 = float((pix > 1e-2).mean())
    _log_measured("t2_full_size", scene=scene_id, alg=alg, w=w, rel_l2=l2, pixels_off=off, pixels_off_1pct=off1,
                  mean_rel=float(abs(img.mean() - ref.mean()) / ref.mean()))
    assert n == iters
    assert l2 <= l2_bound
    assert off <= 5e-2                                             # pixels off by more than 0.1 %
    assert off1 <= off1_bound                                      # pixels off by more than 1 %
    assert abs(img.mean() - ref.mean()) <= 1e-4 * ref.mean()
    assert abs(st["lightVertices"] - orc.num_light_vertices()) <= 1e-4 * orc.num_light_vertices() + 2
    for key in ("lightRays", "cameraRays", "shadowRays", "mergeQueries", "mergeCandidates", "mergeAccepted"):
        assert abs(st[key] - cnt[key]) <= 1e-3 * cnt[key] + 8, (key, st[key], cnt[key])


def test_full_size_properties_scene1_vcm_1024():
    """BASELINE north-star config (scene 1, vcm, 1024x1024, maxLen 10): properties that do not
    need a full CPU render — counters against the reference's measured workload (BASELINE.md
    §2.2: N_v 2 232 535, N_q 2 700 569 at iteration 0), energy against the reference's
    TotalLuminance 277 959, determinism across runs."""
    scene = sv.Scene.cornell_box(1, 1024, 1024)
    img, n, st = render(scene, "vcm", 1)
    assert st["paths"] == 1024 * 1024
    assert abs(st["lightVertices"] - 2232535) < 0.005 * 2232535
    assert abs(st["mergeQueries"] - 2700569) < 0.005 * 2700569
    assert abs(st["mergeCandidates"] - 176523356) < 0.01 * 176523356
    assert abs(st["mergeAccepted"] - 25559886) < 0.01 * 25559886
    total_lum = float((img.astype(np.float64) @ LUM).sum())
    assert abs(total_lum - 277959) < 0.005 * 277959
    assert np.isfinite(img).all() and img.min() >= 0
    img2, _, st2 = render(scene, "vcm", 1)
    for key in ("lightVertices", "mergeQueries", "mergeCandidates", "mergeAccepted", "lightRays", "cameraRays",
                "shadowRays"):
        assert st[key] == st2[key], key                    # same paths, same work
    assert rel_l2(img2, img) < 1e-5                        # float-atomic order only


def test_full_size_properties_scene3_bpm_2048():
    """BASELINE config #4 (scene 3, bpm, 2048x2048, radiusFactor 0.003): N_v 3 358 921,
    N_q 11 079 677, TotalLuminance 527 477."""
    scene = sv.Scene.cornell_box(3, 2048, 2048)
    img, n, st = render(scene, "bpm", 1)
    assert abs(st["lightVertices"] - 3358921) < 0.005 * 3358921
    assert abs(st["mergeQueries"] - 11079677) < 0.005 * 11079677
    assert abs(st["mergeAccepted"] - 157258615) < 0.01 * 157258615
    total_lum = float((img.astype(np.float64) @ LUM).sum())
    assert abs(total_lum - 527477) < 0.01 * 527477


def test_full_size_properties_scene1_vcm_4096_maxlen20():
    """BASELINE config #5's image on ONE GPU (scene 1, vcm, 4096x4096, maxPathLength 20; the
    8-GPU sharded run of it is tools/config5.py): the reference's instrumented light phase stores
    37 817 883 vertices and issues ~45.6 M merge queries; one iteration is
    ~0.5 s here and over an hour on a CPU core, so only these counters and the sanity of the
    image are checked."""
    scene = sv.Scene.cornell_box(1, 4096, 4096)
    img, n, st = render(scene, "vcm", 1, max_len=20)
    assert st["paths"] == 4096 * 4096
    assert abs(st["lightVertices"] - 37817883) < 0.005 * 37817883
    assert abs(st["mergeQueries"] - 45.6e6) < 0.01 * 45.6e6
    assert st["mergeAccepted"] > 5e9 and st["mergeCandidates"] > 3e10
    assert np.isfinite(img).all() and img.min() >= 0
    # same scene, same estimator: the mean radiance agrees with the 1024^2 image to its noise
    small, _, _ = render(sv.Scene.cornell_box(1, 1024, 1024), "vcm", 4, max_len=20)
    assert np.isfinite(small).all()
    assert abs(img.mean() - small.mean() / 4) < 0.02 * img.mean()


def test_full_size_properties_scene0_bpt_1024():
    """BASELINE config #3 (scene 0, bpt, 1024x1024): N_v 1 268 565, TotalLuminance 148 822."""
    scene = sv.Scene.cornell_box(0, 1024, 1024)
    img, n, st = render(scene, "bpt", 1)
    assert abs(st["lightVertices"] - 1268565) < 0.005 * 1268565
    assert st["mergeQueries"] == 0
    total_lum = float((img.astype(np.float64) @ LUM).sum())
    assert abs(total_lum - 148822) < 0.01 * 148822


def test_full_size_properties_scene0_pt_512():
    """BASELINE config #1 on the GPU (scene 0, pt, 512x512, 1 iteration; the reference runs it on
    the CPU): energy against the reference's TotalLuminance 31074.3 — that is
    ONE iteration of a different random stream and scene 0's sun makes firefly paths, so the
    total is only comparable to its sampling noise (3 %; measured 1.0 % at 4 iterations)."""
    scene = sv.Scene.cornell_box(0, 512, 512)
    img, n, st = render(scene, "pt", 4)
    total_lum = float((img.astype(np.float64) @ LUM).sum()) / n
    assert abs(total_lum - 31074.3) < 0.03 * 31074.3
    assert st["cameraRays"] > 512 * 512 and st["lightVertices"] == 0


def test_grid_of_full_iteration_is_consistent():
    """The in-pipeline grid: every stored vertex is counted by the merge (candidates >= accepted,
    accepted > 0), and a second handle fed the same vertices through the kernel-level entry
    point reproduces the per-query sums (ties the pipeline to the T1-tested kernels)."""
    scene = sv.Scene.cornell_box(1, 256, 256)
    img, _, st = render(scene, "bpm", 1)
    assert st["mergeCandidates"] >= st["mergeAccepted"] > 0
    assert st["mergeQueries"] > 0 and st["lightVertices"] > 0
    assert st["kernelLaunches"] >= 20


def test_multi_handle_isolation():
    """Two handles with different scenes / algorithms interleaved give the same images as run
    alone (the scene travels as a kernel parameter, not as shared __constant__ state)."""
    s1, s3 = sv.Scene.cornell_box(1, 96, 96), sv.Scene.cornell_box(3, 64, 64)
    a = sv.VertexCM(s1, "vcm"); b = sv.VertexCM(s3, "bpm")
    a.max_path_length = b.max_path_length = 10
    for it in range(2):
        a.run_iteration(it); b.run_iteration(it)
    ia, ib = a.accum()[0], b.accum()[0]
    a.close(); b.close()
    assert rel_l2(ia, render(s1, "vcm", 2)[0]) < 1e-5
    assert rel_l2(ib, render(s3, "bpm", 2)[0]) < 1e-5


def test_full_shadow_queue_falls_back_to_inline_tracing(monkeypatch):
    """A visibility queue that is far too small (64 segments for ~80 000) makes the producers trace
    inline; the image must not change beyond float-atomic order."""
    scene = sv.Scene.cornell_box(1, 96, 96)
    ref, _, st = render(scene, "vcm", 2)
    monkeypatch.setenv("SVCM_DEBUG_SHADOW_CAP", "64")
    img, _, st2 = render(scene, "vcm", 2)
    assert st2["shadowRays"] == st["shadowRays"] > 10000
    assert rel_l2(img, ref) < 1e-5


def test_pair_list_overflow_is_an_error_not_a_dim_image(monkeypatch):
    """Records that do not fit are never dropped silently: the next synchronisation fails."""
    scene = sv.Scene.cornell_box(1, 96, 96)
    monkeypatch.setenv("SVCM_DEBUG_PAIR_CAP", "1000")
    r = sv.VertexCM(scene, "vcm"); r.max_path_length = 10
    r.run_iterations(0, 1)
    with pytest.raises(sv.SvcmError) as e:
        r.synchronize()
    assert e.value.code == abi.E_NOMEM and "overflow" in str(e.value)
    r.close()


def test_handles_driven_from_concurrent_threads():
    """render() calls RunIteration concurrently on DIFFERENT renderers from OpenMP threads
    (src/smallvcm.cxx:82-108): two host threads, two handles (seed, seed+1), same device. ctypes
    releases the GIL, so the library calls really overlap."""
    import threading
    scene = sv.Scene.cornell_box(1, 128, 128)
    out = {}

    def work(k):
        r = sv.VertexCM(scene, "vcm", seed=1234 + k); r.max_path_length = 10
        for it in range(6):
            r.run_iteration(it)
            r.accum()
        out[k] = r.accum()[0]
        r.close()
    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in threads: t.start()
    for t in threads: t.join()
    for k in range(2):
        assert rel_l2(out[k], render(scene, "vcm", 6, seed=1234 + k)[0]) < 1e-5


def test_drop_in_driver_binary(tmp_path):
    """The reference's own main()/render() with VertexCMGpu injected (smallvcm_b200/bin/
    smallvcm_gpu, built where the reference checkout exists): `-s 1 -a vcm -i 2` writes an HDR
    whose mean luminance matches the library driven directly."""
    import subprocess
    exe = os.path.join(os.path.dirname(GOLD), "..", "smallvcm_b200", "bin", "smallvcm_gpu")
    exe = os.path.abspath(exe)
    if not os.path.exists(exe):
        pytest.skip("drop-in driver not built (needs the reference checkout)")
    out = str(tmp_path / "out.hdr")
    env = dict(os.environ, OMP_NUM_THREADS="1", SVCM_DEVICE="0")
    res = subprocess.run([exe, "-s", "1", "-a", "vcm", "-i", "2", "-o", out], capture_output=True, text=True,
                         env=env, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "vertex connection and merging" in res.stdout
    assert os.path.getsize(out) > 512 * 512


def test_drop_in_driver_time_budget(tmp_path):
    """`-t <sec>` of the reference's render() (src/smallvcm.cxx:79-95) with the GPU renderer: the
    driver build substitutes a wall clock for clock(), so two seconds are two seconds of
    rendering, not CPU-seconds. main() sizes the thread pool with omp_get_num_procs()
    (src/smallvcm.cxx:278-280; OMP_NUM_THREADS is ignored), so the process is pinned to two CPUs:
    two host threads, two handles on one device."""
    import re
    import subprocess
    import time
    exe = os.path.abspath(os.path.join(os.path.dirname(GOLD), "..", "smallvcm_b200", "bin", "smallvcm_gpu"))
    if not os.path.exists(exe):
        pytest.skip("drop-in driver not built (needs the reference checkout)")
    out = str(tmp_path / "timed.bmp")
    env = dict(os.environ, SVCM_DEVICE="0")
    cpus = sorted(os.sched_getaffinity(0))[:2]
    t0 = time.time()
    res = subprocess.run([exe, "-s", "1", "-a", "vcm", "-t", "2", "-o", out], capture_output=True, text=True,
                         env=env, timeout=600, preexec_fn=lambda: os.sched_setaffinity(0, cpus))
    wall = time.time() - t0
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Target:  2 seconds render time" in res.stdout
    reported = float(re.search(r"done in ([0-9.]+) s", res.stdout).group(1))
    # the first iteration of each thread pays CUDA start-up inside the loop (seconds on a cold box),
    # so only the lower bound is sharp: never less than the budget, never more than elapsed time
    assert 2.0 <= reported <= wall + 0.05 and wall < 60.0, (reported, wall)
    assert os.path.getsize(out) == 54 + 512 * 512 * 3
    pixels = np.fromfile(out, np.uint8, offset=54)
    assert 40 < pixels.mean() < 200         # a converged-looking picture, many iterations averaged
