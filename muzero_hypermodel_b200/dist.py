"""Multi-GPU plumbing for self-play: one process per GPU, games sharded, no data-path collective.

The reference runs `num_workers` independent SelfPlay actors (muzero.py:170-186, seed = config.seed + worker
index :175).  Here rank r of N owns the global game slots [r*G, (r+1)*G): the slot is the RNG counter, so a
game's trajectory does not depend on N or on which GPU plays it.  The only collectives are the barrier and
the reductions that aggregate timing / counters (torch.distributed: NCCL on GPUs, gloo on CPU for tests).
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None, device=None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (no-op when N = 1)."""
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def first_slot(rank, games_per_rank):
    """Global id of a rank's game 0."""
    return rank * games_per_rank


def owner_of(slot, games_per_rank):
    """(rank, local game index) that plays global slot `slot`."""
    return slot // games_per_rank, slot % games_per_rank


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def _reduce(value, op, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=op)
    return float(t[0])


def max_over_ranks(value, device="cpu"):
    """Device-timed milliseconds -> the slowest rank's (what throughput is computed from)."""
    return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value, device="cpu"):
    return _reduce(value, dist.ReduceOp.SUM, device)


def aggregate_throughput(units_this_rank, ms_this_rank, device="cpu"):
    """Whole-job units/s = all ranks' units / the slowest rank's time."""
    units = sum_over_ranks(units_this_rank, device)
    ms = max_over_ranks(ms_this_rank, device)
    return units / (ms * 1e-3), ms
