#!/usr/bin/env python
"""bench.py — the headline benchmark: VCM Mpaths/s on the built-in Cornell box (BASELINE.json).

Workload (config.workload): scene 1 (glossy floor, large mirror sphere, ceiling area light),
algorithm vcm, 1024 x 1024, maxPathLength 10, radiusFactor 0.003, radiusAlpha 0.75, seed 1234 —
the configuration BASELINE.json's metric is quoted on. One "step" = one VertexCM::RunIteration
= W*H light sub-paths + W*H camera sub-paths; one "path" = one such pair (the reference's own
unit, src/vertexcm.hxx:290-292). Mpaths/s = W*H*steps / seconds / 1e6.

Our arm (default):
  value     K iterations enqueued back to back on the renderer's stream, timed with CUDA events
            on THAT stream (svcm_event_record), state resident in HBM, max over ranks.
  e2e       the same metric through the reference-facing call sequence of
            VertexCMGpu::RunIteration: svcm_run_iterations(1) + svcm_download_accum into a host
            buffer every step (the non-virtual GetFramebuffer forces that copy), wall clock.
  roofline  k_merge (the hash-grid range query + merge): algorithmic bytes
            N_q*140 + N_cand*12 + N_acc*40 per launch over its CUDA-event time.
  cpu_baseline  the reference's own OpenMP path (oracle/_ref/libsvcm_ref.so -> render(),
            src/smallvcm.cxx:52) timed on this host's cores on a bounded sample; rank 0, N=1.
N > 1 (torchrun): headline = iteration-parallel replicas, one private renderer per GPU with
  seed + rank (the reference's own parallel model, src/smallvcm.cxx:61-142; no data-path
  collective; "weak": every GPU renders one full iteration per step). The same line also
  carries "sharded": every iteration pixel-sharded over the ranks (rows round-robin) with one
  NCCL all-gather of the light vertices before the grid build, N iterations per step.

--impl reference: times the UNMODIFIED reference's CPU implementation of the same path on the
  same workload with all host threads (one private renderer per thread, one iteration each per
  step — its own parallel model), or the oracle port if oracle/_ref was not built.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SCENE, ALG, WIDTH, HEIGHT, MAXLEN, MINLEN = 1, "vcm", 1024, 1024, 10, 0
RADIUS_FACTOR, RADIUS_ALPHA, SEED = 0.003, 0.75, 1234
WORKLOAD = "scene 1 (glossy large mirror sphere + ceiling area light) -a vcm %dx%d maxPathLength %d" % (WIDTH, HEIGHT, MAXLEN)
METRIC, UNIT = "VCM Mpaths/s at 1024^2 scene 1", "Mpaths/s"


# ------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe) sampled while the timed region runs
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock / throttle reasons of the samples that arrived in [t_begin, t_end]
        (perf_counter times; the sampler itself is started well before, nvidia-smi needs
        ~0.3 s to deliver its first row)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.12)]
        if not rows:
            rows = [r for (_, r) in self.rows[-3:]]
        for row in rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def merge_algorithmic_bytes(nq, ncand, nacc):
    """query record 64 B + 8 cell ranges x 8 B + result 12 B per query, 12 B per
    distance test, 40 B of merge payload per accepted particle."""
    return nq * 140 + ncand * 12 + nacc * 40


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own path on the host cores
# ------------------------------------------------------------------------------------------------
def host_cpu_info():
    """What the host really offers this process: the CPU model, the CPUs it may be scheduled on
    (sched_getaffinity) and the cgroup CPU quota (cpu.max) — os.cpu_count() alone says none of it."""
    info = {"model": None, "os_cpu_count": os.cpu_count(), "affinity": None, "cgroup_cpu_max": None, "quota_cpus": None}
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                info["model"] = line.split(":", 1)[1].strip(); break
    except Exception:
        pass
    try:
        info["affinity"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            text = open(path).read().split()
            if path.endswith("cpu.max"):
                info["cgroup_cpu_max"] = " ".join(text)
                if text[0] != "max":
                    info["quota_cpus"] = float(text[0]) / float(text[1])
            else:
                quota = float(text[0]); period = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                info["cgroup_cpu_max"] = "%d %d" % (quota, period)
                if quota > 0:
                    info["quota_cpus"] = quota / period
            break
        except Exception:
            continue
    try:
        info["loadavg_1m"] = float(open("/proc/loadavg").read().split()[0])
    except Exception:
        pass
    return info


def host_threads():
    """Threads the CPU arm uses: the CPUs this process may run on (affinity, capped by the cgroup
    quota), and no more renderers than memory holds."""
    info = host_cpu_info()
    cores = info["affinity"] or os.cpu_count() or 1
    if info["quota_cpus"]:
        cores = max(1, min(cores, int(info["quota_cpus"] + 0.5)))
    try:
        avail_kb = [int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        # one private renderer per thread: ~0.6 GB at 1024^2 (vertices 120 B x 2.2 M + vectors + framebuffer)
        cores = max(1, min(cores, int(avail_kb / 1024 / 1024 * 0.5 / 0.6)))
    except Exception:
        pass
    return cores


def cpu_reference_step(threads, iterations):
    """One bounded sample: `iterations` iterations of the workload on `threads` threads.
    Returns (Mpaths/s, seconds, kind)."""
    import oracle_api as oa
    if oa.have_ref():
        secs, _ = oa.ref_render(SCENE, ALG, WIDTH, HEIGHT, iterations, threads=threads, base_seed=SEED,
                                max_path_length=MAXLEN, min_path_length=MINLEN, radius_factor=RADIUS_FACTOR,
                                radius_alpha=RADIUS_ALPHA, want_image=False)
        kind = "reference"
    else:
        import smallvcm_b200 as sv
        scene = sv.Scene.cornell_box(SCENE, WIDTH, HEIGHT)
        secs, _ = oa.oracle_render(scene, sv.abi.ALGORITHMS[ALG], iterations, threads, base_seed=SEED, rng_mode=0,
                                   max_path_length=MAXLEN, min_path_length=MINLEN, radius_factor=RADIUS_FACTOR,
                                   radius_alpha=RADIUS_ALPHA, want_image=False)
        kind = "port"
    return WIDTH * HEIGHT * iterations / secs / 1e6, secs, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    threads = host_threads()
    # the reference parallelises over ITERATIONS (src/smallvcm.cxx:99-108): one step = one
    # render() call with one iteration per thread — the smallest sample of this workload that
    # uses every host thread (one 1024^2 iteration is 15-20 s on an idle core, ~2 minutes when all
    # 128 threads share the memory system). The arm must end within a few minutes whatever
    # --steps / --warmup say, so steps beyond the time budget are not run and "steps" reports
    # what was timed; a CPU has no clocks to warm, warm-up is skipped when it does not fit.
    iters = threads
    budget = float(os.environ.get("SVCM_REF_BUDGET_S", "240"))
    t_start = time.perf_counter()
    step_s = None
    warm = 0
    while warm < args.warmup and step_s is not None and (time.perf_counter() - t_start) + 2 * step_s <= budget:
        cpu_reference_step(threads, iters); warm += 1
    kind, done, secs = "reference", 0, 0.0
    while done < args.steps:
        t0 = time.perf_counter()
        _, _, kind = cpu_reference_step(threads, iters)
        step_s = time.perf_counter() - t0
        secs += step_s; done += 1
        if (time.perf_counter() - t_start) + step_s > budget:
            break
    args.steps, args.warmup = done, warm
    value = WIDTH * HEIGHT * iters * args.steps / secs / 1e6
    sample = "%d iterations of the full %dx%d workload per step, one private renderer per thread (%s)" % (
        iters, WIDTH, HEIGHT, "reference render(), src/smallvcm.cxx:52" if kind == "reference" else "oracle port")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "paths_per_step": WIDTH * HEIGHT * iters, "seed": SEED,
                       "radiusFactor": RADIUS_FACTOR, "radiusAlpha": RADIUS_ALPHA,
                       "iteration_window": [0, iters],
                       "iteration_window_note": "render() numbers its iterations 0..N-1 (src/smallvcm.cxx:99-108); the "
                                                "GPU arm times global iterations 0..steps-1, the same radius schedule"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                             "host": host_cpu_info()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def measure(sv, D, args, rank, world, device, tdev, sharded, want_e2e, clocks=None):
    """Times `args.steps` steps in one of the two multi-GPU modes and returns the numbers.

    replicas (sharded=False): every GPU owns a private renderer (seed + rank) and renders one
        full iteration per step — the reference's own iteration-parallel model
        (src/smallvcm.cxx:61-142), no data-path collective. At N = 1 this is simply "the renderer".
    sharded  (sharded=True) : every iteration is split over the GPUs by image rows; the light
        vertices are all-gathered (NCCL) before each grid build. A step is N iterations, so each
        GPU still traces W*H paths per step.
    """
    import numpy as np
    from smallvcm_b200 import abi
    scene = sv.Scene.cornell_box(SCENE, WIDTH, HEIGHT)
    if sharded:
        r = sv.VertexCM(scene, ALG, radius_factor=RADIUS_FACTOR, radius_alpha=RADIUS_ALPHA, seed=SEED, device=device,
                        rank=rank, world_size=world)
        uid = sv.comm_unique_id() if rank == 0 else None
        r.comm_init(D.broadcast_bytes(uid, abi.COMM_ID_BYTES, 0, tdev))
        iters_per_step, first = world, 0
    else:
        r = sv.VertexCM(scene, ALG, radius_factor=RADIUS_FACTOR, radius_alpha=RADIUS_ALPHA, seed=SEED + rank,
                        device=device, rank=0, world_size=1)
        # replica g renders the global iterations g, g + N, g + 2N, ... (radius schedule, :295-298)
        iters_per_step, first = 1, rank
    r.max_path_length, r.min_path_length = MAXLEN, MINLEN
    r.set_option(abi.OPT_PROFILE, 1)
    stride = 1 if sharded else world
    state = {"next": first}
    window_first = first

    def rewind():
        """Every timed region covers the SAME global iterations (the radius, hence the merge work,
        shrinks with the iteration index, src/vertexcm.hxx:295-298): 0 .. steps*N - 1."""
        state["next"] = window_first

    def run_steps(n):
        if sharded:
            r.run_iterations(state["next"], n * iters_per_step)
            state["next"] += n * iters_per_step
        else:
            for _ in range(n):
                r.run_iterations(state["next"], 1)
                state["next"] += stride

    def sync_all():
        r.synchronize()
        if world > 1:
            D.barrier(tdev)

    run_steps(args.warmup)
    sync_all()
    r.profile()
    if clocks is not None:
        # keep the GPU busy until the clock sampler has delivered its first rows (~0.3 s)
        t_wait = time.perf_counter()
        while clocks.proc is not None and not clocks.rows and time.perf_counter() - t_wait < 2.0:
            run_steps(1)
            r.synchronize()
    sync_all()
    r.profile()
    rewind()

    # timed: K steps, device time on the launching stream, max over ranks
    t_begin = time.perf_counter()
    r.event_record(0)
    run_steps(args.steps)
    r.event_record(1)
    ms = r.event_elapsed_ms(0, 1)
    sync_all()
    t_end = time.perf_counter()
    prof = r.profile()
    if world > 1:
        ms = D.max_over_ranks(ms, tdev)
    paths_per_step = WIDTH * HEIGHT * world          # both modes: W*H paths per GPU per step
    out = {"value": paths_per_step * args.steps / (ms * 1e-3) / 1e6, "ms_per_step": ms / args.steps,
           "paths_per_step": paths_per_step, "prof": prof, "window": (t_begin, t_end),
           "iterations_per_step": world, "iteration_window": [0, args.steps * world]}

    if want_e2e:
        # the plugin's call sequence with a HOST framebuffer every iteration
        host_fb = np.empty((HEIGHT, WIDTH, 3), np.float32)
        lib = sv.load_library()
        lib.svcm_host_register(host_fb.ctypes.data, host_fb.nbytes)     # what VertexCMGpu does once
        st = r.stats()
        rewind()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for _ in range(iters_per_step):
                r.run_iterations(state["next"], 1)
                state["next"] += stride
                r.accum(host_fb)                                      # D2H of the accumulated image
        e2e_s = time.perf_counter() - t0
        if world > 1:
            e2e_s = D.max_over_ranks(e2e_s, tdev)
        lib.svcm_host_unregister(host_fb.ctypes.data)
        out["e2e"] = {"value": paths_per_step * args.steps / e2e_s / 1e6, "unit": UNIT,
                      "h2d_bytes_per_step": int(st["kernelLaunches"] * st["paramBytesPerLaunch"] * world),
                      "d2h_bytes_per_step": int((host_fb.nbytes + 64) * world),
                      "note": "svcm_run_iterations(1) + svcm_download_accum per iteration, as "
                              "VertexCMGpu::RunIteration does (the non-virtual GetFramebuffer forces the copy); the "
                              "inputs of a step are the iteration index and the kernel parameter blocks"}
    if want_e2e:
        # per-stage times of the stages ALONE: the default schedule runs the camera walk with the light
        # pass, the grid build with the camera walk and the merge with vertex connection, so its stage
        # times overlap and are inflated by sharing the SMs. The roofline is stated for the kernel alone.
        r.set_option(abi.OPT_OVERLAP, 0)
        rewind()
        r.run_iterations(state["next"], 2); r.synchronize(); r.profile()
        rewind()
        n_iso = min(args.steps, 20)
        for _ in range(n_iso):
            r.run_iterations(state["next"], 1)
            state["next"] += stride
        r.synchronize()
        out["prof_isolated"] = r.profile()
        r.set_option(abi.OPT_OVERLAP, 15)
    r.close()
    return out


def sharded_parity(sv, D, rank, world, device, tdev, iterations=2, width=512):
    """rel. L2 between the pixel-sharded render (all ranks, light vertices all-gathered, images
    summed with svcm_reduce_accum) and ONE GPU rendering the same iterations: the only place the
    driver's records can show that the N-rank image is the 1-rank image (its test box has one GPU)."""
    import numpy as np
    from smallvcm_b200 import abi
    scene = sv.Scene.cornell_box(SCENE, width, width)
    r = sv.VertexCM(scene, ALG, radius_factor=RADIUS_FACTOR, radius_alpha=RADIUS_ALPHA, seed=SEED, device=device,
                    rank=rank, world_size=world)
    uid = sv.comm_unique_id() if rank == 0 else None
    r.comm_init(D.broadcast_bytes(uid, abi.COMM_ID_BYTES, 0, tdev))
    r.max_path_length, r.min_path_length = MAXLEN, MINLEN
    r.run_iterations(0, iterations)
    r.reduce_accum(0)
    img, _ = r.accum()
    r.close()
    out = None
    if rank == 0:
        one = sv.VertexCM(scene, ALG, radius_factor=RADIUS_FACTOR, radius_alpha=RADIUS_ALPHA, seed=SEED, device=device)
        one.max_path_length, one.min_path_length = MAXLEN, MINLEN
        one.run_iterations(0, iterations)
        ref, _ = one.accum()
        one.close()
        out = float(np.linalg.norm(img.astype(np.float64) - ref) / np.linalg.norm(ref.astype(np.float64)))
    D.barrier(tdev)
    return out


def config5_record(sv, D, rank, world, device, tdev, iterations=2):
    """BASELINE config #5 (scene 1 vcm 4096^2 maxPathLength 20), ONE iteration strong-scaled over
    the ranks by the pixel-sharded mode; a bounded side measurement (1 warm-up + `iterations`)."""
    from smallvcm_b200 import abi
    W, L = 4096, 20
    scene = sv.Scene.cornell_box(SCENE, W, W)
    r = sv.VertexCM(scene, ALG, radius_factor=RADIUS_FACTOR, radius_alpha=RADIUS_ALPHA, seed=SEED, device=device,
                    rank=rank, world_size=world)
    if world > 1:
        uid = sv.comm_unique_id() if rank == 0 else None
        r.comm_init(D.broadcast_bytes(uid, abi.COMM_ID_BYTES, 0, tdev))
    r.max_path_length = L
    r.set_option(abi.OPT_PROFILE, 1)
    r.run_iterations(0, 1); r.synchronize()
    if world > 1:
        D.barrier(tdev)
    r.profile()
    r.event_record(0); r.run_iterations(1, iterations); r.event_record(1)
    ms = r.event_elapsed_ms(0, 1)
    if world > 1:
        ms = D.max_over_ranks(ms, tdev)
    p = r.profile()
    r.close()
    n = max(1, p["timedIterations"])
    return {"workload": "scene 1 -a vcm %dx%d maxPathLength %d, iterations 1..%d" % (W, W, L, iterations), "n_gpus": world,
            "ms_per_iteration": ms / iterations, "mpaths_s": W * W * iterations / (ms * 1e-3) / 1e6,
            "stage_ms": {k: p[k] / n for k in ("msLightTrace", "msGather", "msGridBuild", "msCameraTrace", "msMerge")},
            "light_vertices": p["lightVertices"] / n, "merge_queries": p["mergeQueries"] / n}


def run_gpu_arm(args):
    import smallvcm_b200 as sv
    from smallvcm_b200 import dist as D

    rank, local_rank, world = D.env_rank()
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus %d needs one process per GPU: launch with torch.distributed.run" % args.gpus)
    device = local_rank
    tdev = None
    if world > 1:
        import torch
        torch.cuda.set_device(device)
        tdev = "cuda:%d" % device
        D.init_process_group("nccl")

    clocks = ClockSampler(device) if rank == 0 else None
    if clocks is not None:
        clocks.start()

    # headline: one renderer per GPU (at N > 1: iteration-parallel replicas, no collective)
    head = measure(sv, D, args, rank, world, device, tdev, sharded=False, want_e2e=True, clocks=clocks)
    clk = clocks.stop(*head["window"]) if clocks is not None else None
    # N > 1: also the pixel-sharded mode with the NCCL all-gather of light vertices
    shard = measure(sv, D, args, rank, world, device, tdev, sharded=True, want_e2e=False) if world > 1 else None
    parity = sharded_parity(sv, D, rank, world, device, tdev) if world > 1 else None
    config5 = None
    if not args.no_config5:
        try:
            config5 = config5_record(sv, D, rank, world, device, tdev)
        except Exception as exc:                      # a side record must never cost the headline
            config5 = {"error": str(exc)[:200]}

    # roofline of the merge stage (one launch set per iteration), timed ALONE (see measure())
    prof = head["prof"]
    iso = head.get("prof_isolated") or prof
    n = max(1, iso["timedIterations"])
    nq, ncand, nacc = iso["mergeQueries"] / n, iso["mergeCandidates"] / n, iso["mergeAccepted"] / n
    merge_ms = iso["msMerge"] / n
    peak, peak_kind = measured_peak_gbs()
    achieved = merge_algorithmic_bytes(nq, ncand, nacc) / (merge_ms * 1e-3) / 1e9 if merge_ms > 0 else 0.0
    # DRAM bytes of the merge kernel cannot be read without a profiler: the figure is the committed
    # ncu capture of this workload on ONE GPU (profiles/merge_traffic.json names kernel, iteration,
    # commit); it is not reported for N > 1, where every rank runs a different share
    traffic, traffic_source = None, None
    if world == 1:
        try:
            rec = json.load(open(os.path.join(ROOT, "profiles", "merge_traffic.json")))
            traffic = rec["dram_bytes_per_launch"]
            traffic_source = {k: rec.get(k) for k in ("kernel", "iteration", "n_gpus", "captured", "how")}
        except Exception:
            pass
    stage = lambda p: {k: p[k] / max(1, p["timedIterations"]) for k in
                       ("msLightTrace", "msGather", "msGridBuild", "msCameraTrace", "msMerge", "msTotal")}
    st_ms = stage(prof)
    iso_ms = stage(iso)
    build_bytes = iso["lightVertices"] / n * 136 + WIDTH * HEIGHT * 8
    trace_s = (iso_ms["msLightTrace"] + iso_ms["msCameraTrace"]) * 1e-3
    rays = {k: prof[k] / max(1, prof["timedIterations"]) for k in ("lightRays", "cameraRays", "shadowRays")}
    trace = {"closest_hit_rays_per_iteration": rays["lightRays"] + rays["cameraRays"],
             "shadow_rays_per_iteration": rays["shadowRays"],
             "grays_per_s": (rays["lightRays"] + rays["cameraRays"] + rays["shadowRays"]) / trace_s / 1e9 if trace_s > 0 else None,
             "ms_light_plus_camera": trace_s * 1e3,
             "ms_start_to_camera_pass_done_in_schedule": st_ms["msLightTrace"] + st_ms["msCameraTrace"],
             "note": "ms_light_plus_camera: light + camera passes (closest-hit rays of both walks, deferred any-hit "
                     "segments, vertex connection) timed alone, one after the other (SVCM_OPT_OVERLAP = 0); "
                     "instruction-bound on 22 primitives in the constant bank, reported as rays/s, not against the "
                     "HBM roofline (SURVEY 8d). In the default schedule the two walks run concurrently with each "
                     "other and the light pass shares the SMs with the previous iteration's merge: "
                     "ms_start_to_camera_pass_done_in_schedule is the wall time from the first kernel of an "
                     "iteration to the end of its vertex-connection visibility tests under that sharing"}
    roofline = {"kernel": "merge stage: query sort + merge kernel (HashGrid::Process + RangeQuery::Process)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_source,
                "peak_source": peak_kind, "ms_per_launch": merge_ms, "queries": nq, "candidates": ncand,
                "accepted": nacc, "algorithmic_bytes_per_launch": merge_algorithmic_bytes(nq, ncand, nacc),
                "timed": "alone: %d iterations of the same window with SVCM_OPT_OVERLAP = 0 (no other kernel on the "
                         "device); in the default schedule (SVCM_OPT_OVERLAP = 15) the stage runs concurrently with "
                         "vertex connection and with the next iteration's light pass" % n,
                "stage_ms_per_iteration": st_ms,
                "stage_ms_per_iteration_alone": iso_ms,
                "grid_build": {"algorithmic_bytes": build_bytes, "ms": iso_ms["msGridBuild"],
                               "ms_in_schedule": st_ms["msGridBuild"],
                               "achieved": build_bytes / (iso_ms["msGridBuild"] * 1e-3) / 1e9 if iso_ms["msGridBuild"] > 0 else None,
                               "frac": build_bytes / (iso_ms["msGridBuild"] * 1e-3) / 1e9 / peak if iso_ms["msGridBuild"] > 0 else None,
                               "formula": "N_v*136 + cells*8 (SURVEY 8d); timed alone like the merge"}}

    # CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        v, secs, kind = cpu_reference_step(threads, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "host": host_cpu_info(),
               "sample": "%d iterations of the full %dx%d workload, one private renderer and one iteration per "
                         "thread (the reference's own OpenMP model, src/smallvcm.cxx:61-108), %.1f s wall"
                         % (threads, WIDTH, HEIGHT, secs)}

    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "paths_per_step": head["paths_per_step"],
                           "iterations_per_step": head["iterations_per_step"], "seed": SEED,
                           "radiusFactor": RADIUS_FACTOR, "radiusAlpha": RADIUS_ALPHA,
                           "iteration_window": head["iteration_window"],
                           "iteration_window_note": "value and e2e both time global iterations 0..steps*N-1 (warm-up "
                                                    "runs first, then the index rewinds); the reference arm's render() "
                                                    "numbers its iterations from 0 as well",
                           "parallelism": "1 GPU" if world == 1 else
                           "%d iteration-parallel replicas (one renderer per GPU, seed + rank, global iteration "
                           "indices interleaved: the reference's own model, src/smallvcm.cxx:61-142); no data-path "
                           "collective. The pixel-sharded mode with the NCCL all-gather is under \"sharded\"" % world,
                           "l2": "per-iteration working set (path state, light vertices, camera vertices, pair and "
                                 "shadow queues, grid: ~0.8 GB) exceeds the 126 MB L2; no flush needed"},
                "clocks": clk, "e2e": head["e2e"], "gpu_launches": int(prof["kernelLaunches"]), "roofline": roofline,
                "trace": trace, "cpu_baseline": cpu, "config5": config5}
        if shard is not None:
            line["sharded"] = {"value": shard["value"], "unit": UNIT, "ms_per_step": shard["ms_per_step"],
                               "iterations_per_step": world, "stage_ms_per_iteration": stage(shard["prof"]),
                               "gpu_launches": int(shard["prof"]["kernelLaunches"]),
                               "parity_rel_l2": parity,
                               "parity_note": "||reduce_accum(N-rank render) - 1-GPU render|| / ||1-GPU render||, scene 1 "
                                              "vcm 512^2, iterations 0..1 (differs only by the order of float atomics)",
                               "note": "every iteration split over the GPUs by image rows (round robin); per "
                                       "iteration ONE in-place ncclAllGather of fixed-size blocks [256-byte header with the "
                                       "rank's count and bounding box | its 52-byte light-vertex merge records] before "
                                       "the grid build, no host round trip; "
                                       "a step is N such iterations. Fixed per-iteration costs (replicated grid "
                                       "build, gather, launch floors of 1/N-sized kernels) bound this mode at 1024^2; "
                                       "it exists for iterations too large for one GPU (BASELINE config #5)"}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    ap.add_argument("--no-config5", action="store_true", help="skip the BASELINE config #5 side record")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
