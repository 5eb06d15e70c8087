import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smallvcm_b200 as sv
from smallvcm_b200 import abi
sid = int(os.environ.get("SCENE", 1)); alg = os.environ.get("ALG", "vcm"); W = int(os.environ.get("W", 1024))
sc = sv.Scene.cornell_box(sid, W, W)
g = sv.VertexCM(sc, alg); g.max_path_length = 10
g.set_option(abi.OPT_PROFILE, 1)
g.run_iterations(0, 3); g.synchronize()
fb = np.empty((W, W, 3), np.float32)
lib = sv.load_library(); print('register', lib.svcm_host_register(fb.ctypes.data, fb.nbytes))
rows = []
it = 3
for k in range(12):
    t0 = time.perf_counter(); g.run_iterations(it, 1); it += 1
    t1 = time.perf_counter(); g.synchronize()
    t2 = time.perf_counter(); g.accum(fb)
    t3 = time.perf_counter()
    st = g.stats()
    rows.append((1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), st['msTotal']))
for r in rows: print('enqueue %.2f ms  sync %.2f ms  download %.2f ms   events total %.2f ms' % r)
# batch of 8 without sync
t0 = time.perf_counter(); g.run_iterations(it, 8); t1 = time.perf_counter(); g.synchronize(); t2 = time.perf_counter()
print('batch8: enqueue %.2f ms, sync %.2f ms -> %.2f ms/iter' % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t2-t0)/8))
