"""Device-side target construction with the reference's names (replay_buffer.py:222-295).

`make_target_batch` builds the value / reward / policy / action targets of a batch of sampled
positions on the GPU (K11, csrc/mzb_targets.cu) from games stored in the export layout of
`envs.VectorEnv.drain_raw()`.  The replay store, PER sampling and Reanalyse are learner-side
("next" rows) and are not part of this package yet.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

_vp, _i32 = C.c_void_p, C.c_int32
_lib.bind("mzb_make_target", C.c_int, [_vp] * 8 + [_i32] + [_vp] * 4 + [_i32, _i32, _i32, _vp, C.c_uint64] + [_vp] * 5)


class DeviceGames:
    """Finished games on the device in entry-array form (what the export ring holds)."""

    def __init__(self, reward, to_play, root_value, visits, action, game_start, game_len, reanalysed=None, device="cuda"):
        dev = torch.device(device)
        t = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x)).to(dev, dt).contiguous()
        self.reward, self.to_play = t(reward, torch.float32), t(to_play, torch.int8)
        self.root_value = t(root_value, torch.float64)
        self.visits = torch.from_numpy(np.ascontiguousarray(visits).astype(np.uint16)).to(dev).contiguous()
        self.action, self.game_start, self.game_len = t(action, torch.int32), t(game_start, torch.int32), t(game_len, torch.int32)
        self.reanalysed = None if reanalysed is None else t(reanalysed, torch.float64)
        self.A = self.visits.shape[1]
        self.device = dev


def make_target_batch(games, batch_game, batch_index, config, seed=0, batch_slot=None, batch_step=None):
    """Returns (target_values [B,K+1] f64, target_rewards [B,K+1] f64, target_policies [B,K+1,A] f64, actions [B,K+1] i32)."""
    dev = games.device
    B = len(batch_game)
    K, td = int(config.num_unroll_steps), int(config.td_steps)
    bg = torch.as_tensor(np.asarray(batch_game, dtype=np.int32)).to(dev)
    bi = torch.as_tensor(np.asarray(batch_index, dtype=np.int32)).to(dev)
    bs = None if batch_slot is None else torch.as_tensor(np.asarray(batch_slot, dtype=np.int64)).to(dev).to(torch.int32)
    bt = None if batch_step is None else torch.as_tensor(np.asarray(batch_step, dtype=np.int64)).to(dev).to(torch.int32)
    # Python's own pow, so discount ** i matches replay_buffer.py:240, 253 bit-for-bit
    dp = torch.tensor([config.discount ** i for i in range(td + 1)], dtype=torch.float64, device=dev)
    tv = torch.empty((B, K + 1), dtype=torch.float64, device=dev)
    tr = torch.empty((B, K + 1), dtype=torch.float64, device=dev)
    tp = torch.empty((B, K + 1, games.A), dtype=torch.float64, device=dev)
    ta = torch.empty((B, K + 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib.mzb_make_target(ptr(games.reward), ptr(games.to_play), ptr(games.root_value), ptr(games.reanalysed),
                                       ptr(games.visits), ptr(games.action), ptr(games.game_start), ptr(games.game_len),
                                       games.A, ptr(bg), ptr(bi), ptr(bs), ptr(bt), B, K, td, ptr(dp),
                                       int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(tv), ptr(tr), ptr(tp), ptr(ta),
                                       _lib.current_stream()))
    return tv, tr, tp, ta
