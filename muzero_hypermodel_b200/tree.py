"""Batched MCTS tree store: thin host wrapper over the mzb_tree_* C ABI (include/mzb200.h).

Replaces the reference's Python Node graph + MinMaxStats (self_play.py:434-477, 551-568) for G
games at once.  All tensors live on the tree's CUDA device; every call is asynchronous on the
current torch stream.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


class _BlockedHidden:
    """Index adaptor over the blocked hidden-state slots: h[game, slot] / h[game] without copying the store."""

    def __init__(self, blocked, G):
        self._b, self.G = blocked, G
        self.shape = (G, blocked.shape[1], blocked.shape[3])

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            g, rest = idx[0], idx[1:]
        else:
            g, rest = idx, ()
        if not isinstance(g, int):
            raise TypeError("index the blocked hidden-state store by an integer game first")
        v = self._b[g >> 5, :, g & 31]
        return v[rest] if rest else v

    def dense(self):
        """A [G, S+1, H] copy in game-major order."""
        return self._b.permute(0, 2, 1, 3).reshape(-1, self._b.shape[1], self._b.shape[3])[:self.G].contiguous()


class BatchedTree:
    def __init__(self, n_games, n_actions, num_simulations, n_players, discount, pb_c_base, pb_c_init,
                 hidden_floats=0, seed=0, device=None):
        if n_players > 2:
            raise NotImplementedError("More than two player mode not implemented.")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.G, self.A, self.S, self.P, self.H = n_games, n_actions, num_simulations, n_players, hidden_floats
        self.cfg = _lib.TreeConfig(n_games, n_actions, num_simulations, n_players, float(discount), float(pb_c_base),
                                   float(pb_c_init), hidden_floats, 0, seed & 0xFFFFFFFFFFFFFFFF)
        nbytes = _lib.lib.mzb_tree_workspace_bytes(C.byref(self.cfg))
        if nbytes == 0:
            check(-1)
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = self.workspace.data_ptr()
            self._ws_ptr = (base + 255) // 256 * 256
            # the caller's own math.log, so pb_c matches ucb_score bit-for-bit (self_play.py:385-390)
            lut = (C.c_double * (num_simulations + 1))(
                *[math.log((n + pb_c_base + 1) / pb_c_base) + pb_c_init for n in range(num_simulations + 1)])
            self._h = C.c_void_p()
            check(_lib.lib.mzb_tree_create(C.byref(self._h), C.byref(self.cfg), C.c_void_p(self._ws_ptr), nbytes, lut))
        self.nbytes = nbytes

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib.mzb_tree_destroy(h)
            except (AttributeError, TypeError):      # interpreter shutdown
                pass
            self._h = None

    # -- search steps
    def root_init(self, reward, policy, policy_is_logits=True, legal=None, to_play=None, noise=None, alpha=0.25,
                  frac=0.0, slot=None, step=None):
        check(_lib.lib.mzb_tree_root_init(self._h, ptr(reward), ptr(policy), int(policy_is_logits), ptr(legal),
                                          ptr(to_play), ptr(noise), float(alpha), float(frac), ptr(slot), ptr(step),
                                          _lib.current_stream()))

    def select(self, parent_slot=None, action=None, depth=None):
        check(_lib.lib.mzb_tree_select(self._h, ptr(parent_slot), ptr(action), ptr(depth), _lib.current_stream()))

    def expand_backup(self, value, reward, policy, policy_is_logits=True):
        check(_lib.lib.mzb_tree_expand_backup(self._h, ptr(value), ptr(reward), ptr(policy), int(policy_is_logits),
                                              _lib.current_stream()))

    def root_stats(self, full=False):
        dev = self.device
        visits = torch.empty((self.G, self.A), dtype=torch.int32, device=dev)
        root_value = torch.empty(self.G, dtype=torch.float64, device=dev)
        max_depth = torch.empty(self.G, dtype=torch.int32, device=dev)
        out = {"visits": visits, "root_value": root_value, "max_depth": max_depth}
        cvs = crw = cpr = mm = None
        if full:
            cvs = torch.empty((self.G, self.A), dtype=torch.float64, device=dev)
            crw = torch.empty((self.G, self.A), dtype=torch.float32, device=dev)
            cpr = torch.empty((self.G, self.A), dtype=torch.float64, device=dev)
            mm = torch.empty((self.G, 2), dtype=torch.float64, device=dev)
            out.update(child_value_sum=cvs, child_reward=crw, child_prior=cpr, minmax=mm)
        check(_lib.lib.mzb_tree_root_stats(self._h, ptr(visits), ptr(root_value), ptr(max_depth), ptr(cvs), ptr(crw),
                                           ptr(cpr), ptr(mm), _lib.current_stream()))
        return out

    def counters(self, reset=False):
        c = (C.c_uint64 * 2)()
        check(_lib.lib.mzb_tree_counters_sync(self._h, c, int(reset), _lib.current_stream()))
        return {"path_length_sum": int(c[0]), "simulations": int(c[1])}

    def hidden(self):
        """fp32 [G, S+1, H] view of the hidden-state slots.  The store keeps them blocked by 32 games and node-major
        inside a block ([G/32][S+1][32][H], csrc/mzb_tree.cuh); the returned adaptor indexes that memory as
        h[game, slot] without copying (h.dense() makes a game-major copy)."""
        p = _lib.lib.mzb_tree_hidden_ptr(self._h)
        if not p:
            return None
        off = (p - self.workspace.data_ptr())
        nb = (self.G + 31) // 32
        n = nb * (self.S + 1) * 32 * self.H
        blocked = self.workspace[off:off + 4 * n].view(torch.float32).view(nb, self.S + 1, 32, self.H)
        return _BlockedHidden(blocked, self.G)

    def export_game(self, game):
        S1, A = self.S + 1, self.A
        vs = np.empty((S1, A), dtype=np.float64)
        pr = np.empty((S1, A), dtype=np.float32)
        vi = np.empty((S1, A), dtype=np.int32)
        rw = np.empty((S1, A), dtype=np.float32)
        ch = np.empty((S1, A), dtype=np.int32)
        rp = np.empty(A, dtype=np.float64)
        sc = np.empty(4, dtype=np.float64)
        check(_lib.lib.mzb_tree_export_game_sync(self._h, int(game), ptr(vs), ptr(pr), ptr(vi), ptr(rw), ptr(ch),
                                                 ptr(rp), ptr(sc), _lib.current_stream()))
        n = int(sc[3])
        return {"value_sum": vs[:n], "prior": pr[:n], "visit": vi[:n], "reward": rw[:n], "child": ch[:n],
                "root_prior": rp, "root_visit": int(sc[0]), "root_value_sum": float(sc[1]),
                "root_reward": float(sc[2])}
