"""Bitwise A/B of the ray queries between two builds of the library (developer tool).

  python tools/ab_intersect.py libA.so libB.so

Each library runs in its own process on the same seeded rays and segments of the four reference
scenes (svcm_intersect_batch, svcm_occluded_batch) and dumps every output array; the parent
compares them byte for byte. Used to show that a change of the intersection code (evaluation order,
skipped primitive runs) does not move a single hit.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r)
import smallvcm_b200 as sv
out = {}
for sid in range(4):
    rng = np.random.default_rng(100 + sid)
    scene = sv.Scene.cornell_box(sid, 64, 64)
    n = 400000
    org = rng.uniform(-1.27, 1.28, (n, 3)).astype(np.float32)
    snap = rng.integers(0, 4, n)                     # a quarter of the origins ON a wall plane, a quarter at the camera
    org[snap == 1, 2] = np.float32(-1.28002); org[snap == 2, 0] = np.float32(1.28975)
    org[snap == 3] = np.array(scene.desc.camera.position, np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[rng.integers(0, n, n // 50), rng.integers(0, 3, n // 50)] = 0.0          # zero components
    rays = np.concatenate([org, d], 1)
    for k, a in enumerate(sv.intersect_batch(scene, rays)):
        out["s%%d_hit%%d" %% (sid, k)] = a
    dst = rng.uniform(-1.27, 1.28, (n, 3)).astype(np.float32)
    dst[snap == 1, 2] = np.float32(1.26002)           # towards the light box's lower face
    v = dst - org; ln = np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
    seg = np.concatenate([org, v / np.maximum(ln, 1e-20), ln], 1).astype(np.float32)
    out["s%%d_occ" %% sid] = sv.occluded_batch(scene, seg)
np.savez(sys.argv[1], **out)
'''


def main():
    libs = sys.argv[1:3]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    files = []
    for i, lib in enumerate(libs):
        env = dict(os.environ, SVCM_LIBRARY=os.path.join(ROOT, lib), SVCM_ALLOW_OLD_LIBRARY="1")
        f = "/tmp/ab_intersect_%d.npz" % i
        subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT), f], env=env, check=True, timeout=300)
        files.append(f)
    import numpy as np
    a, b = np.load(files[0]), np.load(files[1])
    rec = {"libs": libs, "arrays": len(a.files), "different": {}}
    for k in a.files:
        same = a[k].tobytes() == b[k].tobytes()
        if not same:
            rec["different"][k] = int((a[k] != b[k]).sum())
    rec["hits"] = float((a["s1_hit0"] >= 0).mean()); rec["occluded"] = float(a["s1_occ"].mean())
    rec["identical"] = not rec["different"]
    print(json.dumps(rec))
    with open(os.path.join(ROOT, "gpurun_out", "ab_intersect.json"), "a") as f:
        f.write(json.dumps(rec) + "\n")
    return 0 if rec["identical"] else 1


if __name__ == "__main__":
    sys.exit(main())
