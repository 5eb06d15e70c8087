"""Per-stage device timings of one configuration (developer tool)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smallvcm_b200 as sv
from smallvcm_b200 import abi
sid = int(os.environ.get("SCENE", 1)); alg = os.environ.get("ALG", "vcm"); W = int(os.environ.get("W", 1024))
L = int(os.environ.get("MAXLEN", 10)); iters = int(os.environ.get("ITERS", 8)); grp = int(os.environ.get("GROUP", 8))
sc = sv.Scene.cornell_box(sid, W, W)
g = sv.VertexCM(sc, alg); g.max_path_length = L
g.set_option(abi.OPT_PROFILE, 1); g.set_option(abi.OPT_MERGE_GROUP, grp)
g.run_iterations(0, 2); g.synchronize()
t0 = time.time(); g.run_iterations(2, iters); g.synchronize(); t1 = time.time()
st = g.stats()
print(json.dumps(dict(scene=sid, alg=alg, W=W, L=L, group=grp, ms_per_iter=(t1 - t0) * 1e3 / iters,
                      mpaths_s=W * W * iters / (t1 - t0) / 1e6, **st)))
