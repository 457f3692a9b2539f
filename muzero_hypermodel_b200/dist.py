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


# ---- the two collectives the learner side needs; neither is on the self-play data path
def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def broadcast_weights(state_dict, src=0, device=None):
    """Weight refresh (replaces `shared_storage.get_info.remote("weights")`, self_play.py:37): rank `src`'s state
    dict reaches every rank as ONE flat float32 bucket (launch latency, not link count, is what a 1.5 K - 5.5 M
    parameter model pays on NVSwitch).  Returns a state dict of tensors on `device` (same keys and shapes)."""
    keys = list(state_dict.keys())
    shapes = [tuple(state_dict[k].shape) for k in keys]
    sizes = [int(torch.as_tensor(state_dict[k]).numel()) for k in keys]
    dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() and dist.is_initialized()
                                                          and dist.get_backend() == "nccl" else "cpu"))
    flat = torch.cat([torch.as_tensor(state_dict[k]).reshape(-1).to(dev, torch.float32) for k in keys]) if keys else \
        torch.zeros(0, device=dev)
    if _world() > 1:
        dist.broadcast(flat, src=src)
    out, off = {}, 0
    for k, shp, n in zip(keys, shapes, sizes):
        out[k] = flat[off:off + n].reshape(shp)
        off += n
    return out


def allreduce_gradients(grads, average=True):
    """The trainer's gradient all-reduce (the only collective north_star names): the gradients are packed into one
    flat float32 bucket, summed over the ranks (NCCL over NVLink / NVSwitch, in-switch reduction when NVLS is up) and
    written back in place; `average` divides by the world size (data-parallel mean)."""
    grads = [g for g in grads if g is not None]
    if not grads:
        return
    world = _world()
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat /= world
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].reshape(g.shape))
        off += n


# ---- gradient all-reduce fused into the optimiser step (csrc/mzb_optim.cu), over NVLink peer memory
class PeerGradientBuckets:
    """Every rank's flat gradient bucket in IPC-shared device memory, mapped by all ranks of the node, so that ONE
    kernel per rank can sum the ranks' gradients over NVLink / NVSwitch and apply the optimiser update
    (`mzb_adam_step_allreduce` / `mzb_sgd_step_allreduce`) - no NCCL call on the step.  torch.distributed only carries
    the 64-byte IPC handles once, at construction.  Two buckets per rank (step parity), one flag word per peer."""

    def __init__(self, n_floats, device):
        import ctypes as C
        from . import _lib
        from ._lib import check
        vp = C.c_void_p
        for name, res, args in (("mzb_p2p_alloc", C.c_int, [C.POINTER(vp), C.c_size_t]), ("mzb_p2p_free", C.c_int, [vp]),
                                ("mzb_p2p_export", C.c_int, [vp, C.c_char_p]), ("mzb_p2p_import", C.c_int, [C.c_char_p, C.POINTER(vp)]),
                                ("mzb_p2p_close", C.c_int, [vp])):
            _lib.bind(name, res, args)
        self._C, self._lib, self._check = C, _lib, check
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > 8:
            raise NotImplementedError("fused gradient all-reduce: one node, at most 8 ranks")
        self.n, self.device = int(n_floats), torch.device(device)
        self._own, handles = [], []
        with torch.cuda.device(self.device):
            for nbytes in (4 * self.n, 4 * self.n, 4 * 64):            # bucket 0, bucket 1, flags
                p = vp()
                check(_lib.lib.mzb_p2p_alloc(C.byref(p), nbytes))
                h = C.create_string_buffer(64)
                check(_lib.lib.mzb_p2p_export(p, h))
                self._own.append(p)
                handles.append(bytes(h.raw))
            everyone = [None] * self.world
            dist.all_gather_object(everyone, handles)
            self._peer = []                                             # [rank][bucket0, bucket1, flags] device pointers
            for r, hs in enumerate(everyone):
                if r == self.rank:
                    self._peer.append([p.value for p in self._own])
                    continue
                ptrs = []
                for h in hs:
                    p = vp()
                    check(_lib.lib.mzb_p2p_import(C.create_string_buffer(h, 64), C.byref(p)))
                    ptrs.append(p.value)
                self._peer.append(ptrs)
        dist.barrier()

    def bucket(self, seq):
        """This rank's bucket of parity seq & 1 as a float32 tensor (to copy the step's flat gradients into)."""
        ptr, n = self._own[seq & 1].value, self.n

        class _Raw:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}

        return torch.as_tensor(_Raw(), device=self.device)

    def pointers(self, seq):
        """(host array of every rank's bucket pointer for this parity, host array of every rank's flag pointer)."""
        C = self._C
        grads = (C.c_void_p * self.world)(*[self._peer[r][seq & 1] for r in range(self.world)])
        flags = (C.c_void_p * self.world)(*[self._peer[r][2] for r in range(self.world)])
        return grads, flags

    def close(self):
        for r, ptrs in enumerate(getattr(self, "_peer", [])):
            if r != self.rank:
                for p in ptrs:
                    self._lib.lib.mzb_p2p_close(self._C.c_void_p(p))
        for p in getattr(self, "_own", []):
            self._lib.lib.mzb_p2p_free(p)
        self._peer, self._own = [], []
