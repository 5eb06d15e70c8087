"""Quick GPU-vs-oracle sweep (developer tool; the real parity tests live in tests/)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import smallvcm_b200 as sv
import oracle_api as oa
from smallvcm_b200 import abi

def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))

def main():
    W = int(os.environ.get("W", 128)); iters = int(os.environ.get("ITERS", 2))
    res = []
    for sid, alg in [(1, "vcm"), (0, "bpt"), (3, "bpm"), (2, "vcm"), (1, "lt"), (1, "ppm"), (0, "vcm"), (3, "vcm")]:
        sc = sv.Scene.cornell_box(sid, W, W)
        o = oa.Oracle(sc, abi.ALGORITHMS[alg], seed=1234, rng_mode=1)
        g = sv.VertexCM(sc, alg, seed=1234); g.max_path_length = 10
        g.set_option(abi.OPT_PROFILE, 1)
        t0 = time.time()
        for it in range(iters): o.run_iteration(it)
        t1 = time.time()
        g.run_iterations(0, iters)
        gi, n = g.accum(); oi, _ = o.accum()
        st = g.stats(); oc = o.counters()
        d = np.abs(gi - oi).reshape(-1, 3).max(1); scale = np.abs(oi).reshape(-1,3).max(1) + 1e-3
        r = dict(scene=sid, alg=alg, W=W, rel_l2=rel_l2(gi, oi), frac_pix_gt_1e3=float(((d/scale) > 1e-3).mean()),
                 max_abs=float(d.max()), mean_gpu=float(gi.mean()), mean_orc=float(oi.mean()),
                 gpu=dict((k, st[k]) for k in ("lightVertices","mergeQueries","mergeCandidates","mergeAccepted","lightRays","cameraRays","shadowRays","msTotal","msLightTrace","msGridBuild","msCameraTrace","msMerge","kernelLaunches")),
                 oracle=oc, oracle_verts=o.num_light_vertices(), oracle_s=t1-t0)
        print(json.dumps(r)); res.append(r)
        g.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
