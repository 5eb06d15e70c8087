"""Scan iterations for non-finite pixels (developer tool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smallvcm_b200 as sv
W = int(os.environ.get("W", 1024)); L = int(os.environ.get("MAXLEN", 20)); N = int(os.environ.get("ITERS", 16))
alg = os.environ.get("ALG", "vcm"); sid = int(os.environ.get("SCENE", 1))
sc = sv.Scene.cornell_box(sid, W, W)
r = sv.VertexCM(sc, alg); r.max_path_length = L
for it in range(N):
    r.clear(); r.run_iteration(it)
    img, _ = r.accum()
    bad = np.argwhere(~np.isfinite(img).all(2))
    if len(bad):
        print("iteration", it, "non-finite pixels:", [(int(y), int(x), img[y, x].tolist()) for y, x in bad[:8]])
print("done", alg, sid, W, L, N)
