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
    r.run_iterations(0, iters)
    st = r.stats()
    r.reduce_accum(0)
    img, n = r.accum()
    out[rank] = (img if rank == 0 else None, n, st["lightVertices"], st["localLightVertices"], st["localPaths"],
                 st["mergeQueries"])
    r.close()
    D.barrier()


@pytest.mark.parametrize("cfg", [(1, "vcm", 128, 96, 2), (3, "bpm", 96, 96, 2), (0, "bpt", 96, 64, 2)])
def test_two_rank_render_equals_single_gpu(cfg):
    lib = sv.load_library()
    if lib.svcm_device_count() < 2:
        pytest.skip("needs 2 GPUs")
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
    r.run_iterations(0, iters)
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
