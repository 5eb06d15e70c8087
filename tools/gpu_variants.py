"""A/B stage timings of library variants and option sets on one configuration (developer tool).

  python tools/gpu_variants.py [name=lib:ENV=V,ENV=V ...]

Each variant runs in its own process (the library is loaded once per process): iterations
FIRST..FIRST+ITERS-1 after 3 warm-up iterations, per-stage CUDA-event times from svcm_get_profile
and the whole-window time from user events on the library's stream. One JSON line per variant on
stdout and appended to gpurun_out/variants.jsonl.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
import smallvcm_b200 as sv
from smallvcm_b200 import abi
sid = int(os.environ.get("SCENE", 1)); alg = os.environ.get("ALG", "vcm"); W = int(os.environ.get("W", 1024))
H = int(os.environ.get("H", W))
L = int(os.environ.get("MAXLEN", 10)); iters = int(os.environ.get("ITERS", 20)); first = int(os.environ.get("FIRST", 5))
sc = sv.Scene.cornell_box(sid, W, H)
g = sv.VertexCM(sc, alg); g.max_path_length = L
g.set_option(abi.OPT_PROFILE, 1)
for opt in os.environ.get("OPTS", "").split(";"):
    if opt:
        k, v = opt.split("="); g.set_option(int(k), int(v))
g.run_iterations(max(first - 3, 0), 3); g.synchronize(); g.profile()
g.event_record(0); g.run_iterations(first, iters); g.event_record(1)
ms = g.event_elapsed_ms(0, 1)
p = g.profile(); st = g.stats()
n = max(p["timedIterations"], 1)
out = dict(name=os.environ.get("VARIANT", "?"), scene=sid, alg=alg, W=W, H=H, L=L, iters=iters, first=first,
           ms_per_iter=ms / iters, mpaths_s=W * H * iters / ms / 1e3,
           light=p["msLightTrace"] / n, gather=p["msGather"] / n, build=p["msGridBuild"] / n,
           camera=p["msCameraTrace"] / n, merge=p["msMerge"] / n, launches=p["kernelLaunches"] / max(p["iterations"], 1),
           queries=p["mergeQueries"] / n, candidates=p["mergeCandidates"] / n, accepted=p["mergeAccepted"] / n,
           lightVertices=p["lightVertices"] / n)
img, k = g.accum()
out["mean"] = float(img.mean() / max(k, 1))
print("VARIANT_JSON " + json.dumps(out))
'''


def main():
    specs = sys.argv[1:]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for spec in specs:
        name, _, rest = spec.partition("=")
        lib, _, envs = rest.partition(":")
        env = dict(os.environ)
        env["VARIANT"] = name
        if lib:
            env["SVCM_LIBRARY"] = os.path.join(ROOT, lib)
        for kv in envs.split(","):
            if kv:
                k, _, v = kv.partition("=")
                env[k] = v
        try:
            res = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT)], env=env, capture_output=True, text=True,
                                 timeout=300)
            line = [l for l in res.stdout.splitlines() if l.startswith("VARIANT_JSON ")]
            rec = json.loads(line[0][13:]) if line else dict(name=name, error=(res.stderr or res.stdout)[-600:])
        except subprocess.TimeoutExpired:
            rec = dict(name=name, error="timeout")
        print(json.dumps(rec), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "variants.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
