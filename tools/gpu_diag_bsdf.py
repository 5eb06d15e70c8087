import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import smallvcm_b200 as sv, oracle_api as oa
from smallvcm_b200 import abi
def unit(rng, n):
    v = rng.normal(size=(n, 3)).astype(np.float32); return v / np.linalg.norm(v, axis=1, keepdims=True)
rng = np.random.default_rng(20)
scene = sv.Scene.cornell_box(0, 64, 64); orc = oa.Oracle(scene, abi.ALG_VCM)
n = 100000
nrm, fix, gen = unit(rng, n), unit(rng, n), unit(rng, n)
mat = rng.integers(0, scene.desc.numMaterials, n).astype(np.int32)
items = np.concatenate([nrm, fix, gen, mat.view(np.float32)[:, None]], 1).astype(np.float32)
g, o = sv.bsdf_evaluate_batch(scene, items, 0), orc.bsdf_evaluate(items, 0)
err = np.abs(g - o) / (np.maximum(np.abs(g), np.abs(o)) + 1e-30)
err[np.abs(g - o) < 1e-6] = 0
for c in range(7):
    i = err[:, c].argmax()
    print('col', c, 'max rel err', err[i, c], 'mat', mat[i], 'g', g[i], 'o', o[i])
for m in range(9):
    sel = mat == m
    print('mat', m, 'max err', err[sel][:, :7].max())
print("---- sample")
for light in (0, 1):
    items = np.concatenate([nrm, fix, rng.random((n, 3)).astype(np.float32), mat.view(np.float32)[:, None]], 1).astype(np.float32)
    (g, ge), (o, oe) = sv.bsdf_sample_batch(scene, items, light), orc.bsdf_sample(items, light)
    flips = ge != oe
    print('flips', flips.sum())
    err = np.abs(g - o) / (np.maximum(np.abs(g), np.abs(o)) + 1e-30)
    err[np.abs(g - o) < 2e-6] = 0
    err[flips] = 0
    for m in range(9):
        sel = mat == m
        i = np.where(sel)[0][err[sel].max(1).argmax()]
        print('light', light, 'mat', m, 'max err', err[i].max(), 'ev', ge[i], 'g', g[i], 'o', o[i], 'in', items[i, :9])
