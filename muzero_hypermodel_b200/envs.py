"""Vectorised device environments: host wrapper over the mzb_env_* C ABI (include/mzb200.h).

G independent games of one kind live on the GPU; `observe()` yields what `play_game` hands to
`MCTS.run` (self_play.py:138-150), `act_step()` is select_action + Game.step + GameHistory appends
(:152-182), `harvest()` exports finished games and restarts them, `drain()` returns them to the host
as reference-format GameHistory objects (self_play.py:480-495).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

_vp, _i32 = C.c_void_p, C.c_int32

KINDS = {"cartpole": 0, "tictactoe": 1, "connect4": 2, "gomoku": 3, "breakout": 4}
SHAPES = {"cartpole": (1, 1, 4), "tictactoe": (3, 3, 3), "connect4": (3, 6, 7), "gomoku": (3, 11, 11),
          "breakout": (3, 96, 96)}


class EnvConfig(C.Structure):
    _fields_ = [("kind", _i32), ("n_games", _i32), ("max_moves", _i32), ("export_entries", _i32),
                ("export_games", _i32), ("first_slot", C.c_uint32), ("seed", C.c_uint64)]


_lib.bind("mzb_env_workspace_bytes", C.c_size_t, [C.POINTER(EnvConfig)])
_lib.bind("mzb_env_create", C.c_int, [C.POINTER(_vp), C.POINTER(EnvConfig), _vp, C.c_size_t, _vp])
_lib.bind("mzb_env_destroy", C.c_int, [_vp])
_lib.bind("mzb_env_info", C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)])
_lib.bind("mzb_env_reset", C.c_int, [_vp, _vp])
_lib.bind("mzb_env_observe", C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp])
_lib.bind("mzb_env_observe_stacked", C.c_int, [_vp, _i32, _vp, _vp])
_lib.bind("mzb_env_act_step", C.c_int, [_vp, _vp, _vp, _vp, C.c_double, _i32, _vp, _vp, _vp, _vp, _vp, _vp])
_lib.bind("mzb_env_harvest", C.c_int, [_vp, C.c_int, _vp])
_lib.bind("mzb_env_counters_sync", C.c_int, [_vp, C.POINTER(C.c_uint64), _vp])
_lib.bind("mzb_env_state_ptrs", C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)])
_lib.bind("mzb_env_export_drain_sync", C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32)] + [_vp] * 9 + [_vp])


def game_kind(config_or_name):
    name = config_or_name if isinstance(config_or_name, str) else getattr(config_or_name, "GAME", None)
    if name not in KINDS:
        shape = tuple(getattr(config_or_name, "observation_shape", ()))
        name = {v: k for k, v in SHAPES.items()}.get(shape)
    if name not in KINDS:
        raise NotImplementedError(f"no device environment for {config_or_name!r}")
    return name


class VectorEnv:
    def __init__(self, kind, n_games, max_moves, seed=0, first_slot=0, export_entries=None, export_games=None,
                 device=None):
        self.kind = kind
        self.G = int(n_games)
        self.max_moves = int(max_moves)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if export_games is None:
            export_games = max(1024, self.G)
        if export_entries is None:
            # worst case: every game ends in the same step with a full-length history (cartpole: 262,144 x 501 entries, 5 GB)
            export_entries = int(min(2 ** 30, max(4 * (self.max_moves + 1), self.G * (self.max_moves + 1))))
        self.cfg = EnvConfig(KINDS[kind], self.G, self.max_moves, int(export_entries), int(export_games),
                             int(first_slot), int(seed) & 0xFFFFFFFFFFFFFFFF)
        nbytes = _lib.lib.mzb_env_workspace_bytes(C.byref(self.cfg))
        if nbytes == 0:
            check(-1)
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = (self.workspace.data_ptr() + 255) // 256 * 256
            self._h = _vp()
            check(_lib.lib.mzb_env_create(C.byref(self._h), C.byref(self.cfg), _vp(base), nbytes, _lib.current_stream()))
        a, o, r = _i32(), _i32(), _i32()
        check(_lib.lib.mzb_env_info(self._h, C.byref(a), C.byref(o), C.byref(r)))
        self.A, self.obs_dim, self.rec_floats = a.value, o.value, r.value
        self.obs_shape = SHAPES[kind]
        dev = self.device
        self.obs = torch.empty((self.G, self.obs_dim), dtype=torch.float32, device=dev)
        self.legal = torch.empty((self.G, self.A), dtype=torch.uint8, device=dev)
        self.to_play = torch.empty(self.G, dtype=torch.int8, device=dev)
        self.slot = torch.empty(self.G, dtype=torch.int32, device=dev)
        self.step_count = torch.empty(self.G, dtype=torch.int32, device=dev)
        self.nbytes = nbytes

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                _lib.lib.mzb_env_destroy(self._h)
            except (AttributeError, TypeError):      # interpreter shutdown
                pass
            self._h = None

    def reset(self):
        check(_lib.lib.mzb_env_reset(self._h, _lib.current_stream()))

    def observe(self):
        check(_lib.lib.mzb_env_observe(self._h, ptr(self.obs), ptr(self.legal), ptr(self.to_play), ptr(self.slot),
                                       ptr(self.step_count), _lib.current_stream()))
        return self.obs, self.legal, self.to_play

    def observe_stacked(self, stacked_observations):
        """observe() with GameHistory.get_stacked_observations(-1, S) (self_play.py:514-548) as the observation:
        [G, C (S + 1) + S, H, W] float32, assembled on the device from the running histories."""
        S = int(stacked_observations)
        _, legal, to_play = self.observe()
        if S == 0:
            return self.obs, legal, to_play
        c, h, w = self.obs_shape
        buf = getattr(self, "_stacked", None)
        if buf is None or buf.shape[1] != c * (S + 1) + S:
            buf = self._stacked = torch.empty((self.G, c * (S + 1) + S, h, w), dtype=torch.float32, device=self.device)
        check(_lib.lib.mzb_env_observe_stacked(self._h, S, ptr(buf), _lib.current_stream()))
        return buf, legal, to_play

    def act_step(self, visits, root_value, legal=None, temperature=1.0, temperature_threshold=None, uniforms=None,
                 forced_action=None, want_outputs=False):
        out = (None, None, None)
        if want_outputs:
            out = (torch.empty(self.G, dtype=torch.int32, device=self.device),
                   torch.empty(self.G, dtype=torch.float32, device=self.device),
                   torch.empty(self.G, dtype=torch.uint8, device=self.device))
        T = float(temperature)
        check(_lib.lib.mzb_env_act_step(self._h, ptr(visits), ptr(root_value), ptr(self.legal if legal is None else legal),
                                        T, int(temperature_threshold or 0), ptr(uniforms), ptr(forced_action),
                                        ptr(out[0]), ptr(out[1]), ptr(out[2]), _lib.current_stream()))
        return out

    def harvest(self, export=True):
        check(_lib.lib.mzb_env_harvest(self._h, int(bool(export)), _lib.current_stream()))

    def counters(self):
        c = (C.c_uint64 * 4)()
        check(_lib.lib.mzb_env_counters_sync(self._h, c, _lib.current_stream()))
        return {"games": int(c[0]), "finished_moves": int(c[1]), "env_steps": int(c[2]), "dropped_games": int(c[3])}

    def _state_view(self):
        p = [_vp() for _ in range(5)]
        check(_lib.lib.mzb_env_state_ptrs(self._h, *[C.byref(x) for x in p]))
        base = self.workspace.data_ptr()

        def view(pp, nbytes, dtype, shape):
            off = pp.value - base
            return self.workspace[off:off + nbytes].view(dtype).view(shape)

        return {"cartpole": view(p[0], self.G * 32, torch.float64, (self.G, 4)),
                "board": view(p[1], self.G * max(1, int(np.prod(self.obs_shape[1:])) if self.kind != "cartpole" else 1),
                              torch.int8, (self.G, -1)),
                "player": view(p[2], self.G, torch.int8, (self.G,)),
                "hist_len": view(p[3], self.G * 4, torch.int32, (self.G,)),
                "finished": view(p[4], self.G, torch.uint8, (self.G,))}

    def drain_raw(self):
        """Finished games since the last drain as flat numpy arrays (see mzb_env_export_drain_sync)."""
        E, XG, A, R = self.cfg.export_entries, self.cfg.export_games, self.A, self.rec_floats
        if not hasattr(self, "_host"):
            self._host = {"obs": np.empty((E, R), np.float32), "action": np.empty(E, np.int32),
                          "reward": np.empty(E, np.float32), "to_play": np.empty(E, np.int8),
                          "visits": np.empty((E, A), np.uint16), "root_value": np.empty(E, np.float64),
                          "start": np.empty(XG, np.int32), "len": np.empty(XG, np.int32), "slot": np.empty(XG, np.uint32)}
        h = self._host
        ne, ng = _i32(), _i32()
        check(_lib.lib.mzb_env_export_drain_sync(self._h, C.byref(ne), C.byref(ng), ptr(h["obs"]), ptr(h["action"]),
                                                 ptr(h["reward"]), ptr(h["to_play"]), ptr(h["visits"]),
                                                 ptr(h["root_value"]), ptr(h["start"]), ptr(h["len"]), ptr(h["slot"]),
                                                 _lib.current_stream()))
        return ne.value, ng.value, h

    def decode_observation(self, rec):
        """History record -> the observation array the reference's Game returns."""
        if self.kind == "cartpole":
            return np.array([[rec[:4]]], dtype=np.float32)
        c, hh, ww = self.obs_shape
        raw = rec.view(np.int8)
        board = raw[:hh * ww].reshape(hh, ww)
        player = int(raw[hh * ww])
        dtype = np.int32 if self.kind == "tictactoe" else np.float64
        return np.array([board == 1, board == -1, np.full((hh, ww), player)], dtype=dtype)
