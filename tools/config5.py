"""BASELINE config #5: scene 1, vcm, 4096x4096, maxPathLength 20 — one GPU, or pixel-sharded over
the ranks of a torchrun launch with the NCCL all-gather of light vertices (strong scaling of ONE
large iteration, the case the sharded mode exists for).

  python tools/config5.py                                   # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config5.py
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smallvcm_b200 as sv
from smallvcm_b200 import abi, dist as D

W = int(os.environ.get("W", 4096)); L = int(os.environ.get("MAXLEN", 20)); iters = int(os.environ.get("ITERS", 2))
rank, local_rank, world = D.env_rank()
tdev = None
if world > 1:
    import torch
    torch.cuda.set_device(local_rank); tdev = "cuda:%d" % local_rank
    D.init_process_group("nccl")
scene = sv.Scene.cornell_box(1, W, W)
r = sv.VertexCM(scene, "vcm", seed=1234, device=local_rank, rank=rank, world_size=world)
r.max_path_length = L
r.set_option(abi.OPT_PROFILE, 1)
if os.environ.get("MERGE_GROUP"):
    r.set_option(abi.OPT_MERGE_GROUP, int(os.environ["MERGE_GROUP"]))
if world > 1:
    uid = sv.comm_unique_id() if rank == 0 else None
    r.comm_init(D.broadcast_bytes(uid, abi.COMM_ID_BYTES, 0, tdev))
r.run_iterations(0, 1); r.synchronize()
if world > 1: D.barrier(tdev)
r.profile()
r.event_record(0); r.run_iterations(1, iters); r.event_record(1)
ms = r.event_elapsed_ms(0, 1)
if world > 1: ms = D.max_over_ranks(ms, tdev)
st = r.stats(); p = r.profile()
if rank == 0:
    n = max(1, p["timedIterations"])
    print(json.dumps(dict(config="scene 1 vcm %dx%d maxLen %d" % (W, W, L), n_gpus=world, iterations=iters,
                          ms_per_iteration=ms / iters, mpaths_s=W * W * iters / (ms * 1e-3) / 1e6,
                          lightVertices=st["lightVertices"], localLightVertices=st["localLightVertices"],
                          mergeQueries=p["mergeQueries"] / n, mergeCandidates=p["mergeCandidates"] / n,
                          mergeAccepted=p["mergeAccepted"] / n,
                          stage_ms={k: p[k] / n for k in ("msLightTrace", "msGather", "msGridBuild", "msCameraTrace", "msMerge", "msTotal")})))
r.close()
if world > 1:
    import torch.distributed as dist
    dist.destroy_process_group()
