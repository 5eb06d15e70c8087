"""One-process-per-GPU plumbing for bench.py (torch.distributed is plumbing here, not product):
rank discovery from the torchrun environment, broadcast of the NCCL unique id that
svcm_comm_init needs, and the max-over-ranks reduction of device timings.  Works with the
`nccl` backend on GPUs and with `gloo` on CPU (tests/test_dist.py, world_size 2)."""
import os


def env_rank():
    """(rank, local_rank, world_size) from RANK / LOCAL_RANK / WORLD_SIZE (1 process if unset)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def init_process_group(backend):
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend)
    return dist


def broadcast_bytes(payload, nbytes, src=0, device=None):
    """Every rank returns rank `src`'s `payload` (bytes of length nbytes)."""
    import torch
    import torch.distributed as dist
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8)
        assert t.numel() == nbytes
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def max_over_ranks(value, device=None):
    """MAX all-reduce of one float (device timings: the slowest rank defines the step)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(device=None):
    import torch
    import torch.distributed as dist
    if device is not None:
        dist.barrier(device_ids=[torch.device(device).index])
    else:
        dist.barrier()


def shard_rows(height, rank, world_size):
    """Rows owned by `rank` under the library's round-robin row map (svcm_shard_owner)."""
    return list(range(rank, height, world_size))
