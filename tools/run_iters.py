"""Runs a few iterations of one configuration (for ncu captures): SCENE, ALG, W, MAXLEN, ITERS, FIRST."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smallvcm_b200 as sv
sid = int(os.environ.get("SCENE", 1)); alg = os.environ.get("ALG", "vcm"); W = int(os.environ.get("W", 1024))
L = int(os.environ.get("MAXLEN", 10)); iters = int(os.environ.get("ITERS", 5)); first = int(os.environ.get("FIRST", 0))
g = sv.VertexCM(sv.Scene.cornell_box(sid, W, W), alg); g.max_path_length = L
g.run_iterations(first, iters); g.synchronize()
print("done", g.stats()["mergeQueries"])
