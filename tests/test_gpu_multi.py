"""Pixel-sharded multi-GPU run (needs >= 2 GPUs): two processes, one GPU each, rows dealt
round-robin, light vertices all-gathered with NCCL before the grid build, images summed with
ncclReduce. Because every sub-path's random stream is keyed by its GLOBAL path index, the
sharded render traces exactly the single-GPU paths: the images must agree up to float-atomic
summation order."""
import os
import socket

import numpy as np
import pytest

import smallvcm_b200 as sv
from smallvcm_b200 import abi

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, cfg, out):
    import torch
    from smallvcm_b200 import dist as D
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    D.init_process_group("gloo")
    scene_id, alg, w, h, iters = cfg
    scene = sv.Scene.cornell_box(scene_id, w, h)
    r = sv.VertexCM(scene, alg, seed=1234, device=rank, rank=rank, world_size=world)
    r.max_path_length = 10
    uid = sv.comm_unique_id() if rank == 0 else None
    r.comm_init(D.broadcast_bytes(uid, abi.COMM_ID_BYTES, 0))
    if os.environ.get("SVCM_TEST_LAYOUT"):
        r.set_option(abi.OPT_GRID_LAYOUT, int(os.environ["SVCM_TEST_LAYOUT"]))
    # two calls with a synchronisation in between: the second sizes its all-gather blocks from the
    # counts the first one saw (a different block stride, hence a different buffer layout)
    r.run_iterations(0, 1)
    st0 = r.stats()
    r.reduce_accum(0)
    first, n1 = r.accum()
    r.run_iterations(1, iters - 1)
    st = r.stats()
    r.reduce_accum(0)                    # reducing again must not count the first iteration twice
    img, n = r.accum()
    out[rank] = (img if rank == 0 else None, n, st["lightVertices"], st["localLightVertices"], st["localPaths"],
                 st["mergeQueries"], first if rank == 0 else None, n1)
    r.close()
    D.barrier()


@pytest.mark.parametrize("layout", ["0", "1"])
@pytest.mark.parametrize("cfg", [(1, "vcm", 128, 96, 3), (3, "bpm", 96, 96, 2), (0, "bpt", 96, 64, 2)])
def test_two_rank_render_equals_single_gpu(cfg, layout, monkeypatch):
    lib = sv.load_library()
    if lib.svcm_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("SVCM_TEST_LAYOUT", layout)
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), cfg, out), nprocs=world, join=True)
        res = dict(out)
    scene_id, alg, w, h, iters = cfg
    scene = sv.Scene.cornell_box(scene_id, w, h)
    r = sv.VertexCM(scene, alg, seed=1234)
    r.max_path_length = 10
    r.set_option(abi.OPT_GRID_LAYOUT, int(layout))
    r.run_iterations(0, 1)
    ref_first, _ = r.accum()
    r.run_iterations(1, iters - 1)
    ref, n = r.accum()
    st = r.stats()
    r.close()
    img = res[0][0]
    assert res[0][1] == iters and n == iters
    assert res[0][4] + res[1][4] == w * h                                   # every path owned once
    assert res[0][3] + res[1][3] == st["lightVertices"]                     # same light sub-paths
    if alg != "bpt":
        assert res[0][2] == res[1][2] == st["lightVertices"]                # both ranks saw the full set
        assert abs(res[0][5] + res[1][5] - st["mergeQueries"]) <= 2
    rel = np.linalg.norm(img.astype(np.float64) - ref) / np.linalg.norm(ref.astype(np.float64))
    assert rel < 1e-5, rel
    first = res[0][6]
    assert res[0][7] == 1
    assert np.linalg.norm(first.astype(np.float64) - ref_first) / np.linalg.norm(ref_first.astype(np.float64)) < 1e-5
