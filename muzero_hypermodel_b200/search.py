"""Batched MCTS: G independent `MCTS.run` calls (self_play.py:261-362) in lock-step on one GPU.

`BatchedMCTS(config, n_games).run(model, observations, legal_mask, to_play, add_exploration_noise)`
takes what G calls of the reference's `MCTS(config).run(model, observation, legal_actions, to_play,
add_exploration_noise)` take, stacked along a leading game axis, and returns the root statistics the
reference's callers read from the returned root Node (visit counts, root.value(), max_tree_depth,
root_predicted_value).  `self_play.MCTS` wraps it with G=1 and materialises the Node graph.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr
from .tree import BatchedTree

_vp = C.c_void_p
_lib.bind("mzb_search_fc", C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp, C.c_int32, C.c_int,
                                     _vp, _vp, _vp, _vp, _vp])
_lib.bind("mzb_search_fc_is_fused", C.c_int, [_vp])
_lib.bind("mzb_u8_to_unit_float", C.c_int, [_vp, C.c_int64, _vp, _vp])
_lib.bind("mzb_search_resnet", C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp, C.c_int32, _vp, _vp,
                                         C.c_size_t, _vp, _vp, _vp, _vp, _vp])


class BatchedMCTS:
    def __init__(self, config, n_games, device=None, seed=None, hidden_floats=None):
        self.config = config
        self.G = int(n_games)
        self.A = len(config.action_space)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if len(config.players) > 2:
            raise NotImplementedError("More than two player mode not implemented.")
        if hidden_floats is None:
            # fully-connected: slots inside the tree store; residual: a separate NHWC pool (see run)
            hidden_floats = config.encoding_size if config.network == "fullyconnected" else 0
        self.tree = BatchedTree(self.G, self.A, config.num_simulations, len(config.players), config.discount,
                                config.pb_c_base, config.pb_c_init, hidden_floats=hidden_floats,
                                seed=config.seed if seed is None else seed, device=self.device)

    def run(self, model, observations, legal_mask=None, to_play=None, add_exploration_noise=True, noise=None,
            slot=None, step=None, allow_fused=True, num_simulations=None, out=None):
        """observations [G, ...] float tensor on the device; legal_mask [G, A] bool/uint8 or None (all legal);
        to_play [G] int8 or None; noise [G, A] float64 injected Dirichlet sample (by action) or None = generated
        on the device; slot/step [G] int32 RNG counters.  Returns dict of device tensors."""
        cfg = self.config
        dev = self.device
        G, A = self.G, self.A
        S = cfg.num_simulations if num_simulations is None else num_simulations
        if observations.dtype == torch.uint8:
            # emulator frames: normalised on the device exactly as the reference's wrapper does on the host
            # (games/breakout.py:141-159, float32(frame) / 255), so uint8 crosses PCIe instead of float32
            src = observations.to(device=dev).contiguous()
            buf = getattr(self, "_frames_f32", None)
            if buf is None or buf.shape != src.shape:
                buf = self._frames_f32 = torch.empty(src.shape, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                check(_lib.lib.mzb_u8_to_unit_float(ptr(src), src.numel(), ptr(buf), _lib.current_stream()))
            observations = buf
        obs = observations.to(device=dev, dtype=torch.float32).reshape(G, -1).contiguous()
        lg = None if legal_mask is None else legal_mask.to(device=dev, dtype=torch.uint8).contiguous()
        tp = None if to_play is None else to_play.to(device=dev, dtype=torch.int8).contiguous()
        nz = None if noise is None else noise.to(device=dev, dtype=torch.float64).contiguous()
        sl = None if slot is None else slot.to(device=dev, dtype=torch.int32).contiguous()
        st = None if step is None else step.to(device=dev, dtype=torch.int32).contiguous()
        if out is None:
            out = {"visits": torch.empty((G, A), dtype=torch.int32, device=dev),
                   "root_value": torch.empty(G, dtype=torch.float64, device=dev),
                   "root_predicted_value": torch.empty(G, dtype=torch.float32, device=dev),
                   "max_depth": torch.empty(G, dtype=torch.int32, device=dev)}
        frac = float(cfg.root_exploration_fraction) if add_exploration_noise else 0.0
        if hasattr(model, "handle") and cfg.network == "fullyconnected":
            with torch.cuda.device(dev):
                check(_lib.lib.mzb_search_fc(self.tree._h, model.handle(), ptr(obs), ptr(lg), ptr(tp), ptr(nz),
                                             float(cfg.root_dirichlet_alpha), frac, ptr(sl), ptr(st), int(S),
                                             int(allow_fused), ptr(out["visits"]), ptr(out["root_value"]),
                                             ptr(out["root_predicted_value"]), ptr(out["max_depth"]),
                                             _lib.current_stream()))
            return out
        if hasattr(model, "handle") and cfg.network == "resnet":
            h = model.handle()
            state = int(model.latent_shape[0] * model.latent_shape[1] * model.latent_shape[2])
            dt = torch.bfloat16 if model.precision == "bf16" else torch.float32
            pool = getattr(self, "_pool", None)
            if pool is None or pool.dtype != dt:
                pool = self._pool = torch.empty((G, cfg.num_simulations + 1, state), dtype=dt, device=dev)
            # [G, C(S+1)+S, H, W]: the stacked planes of get_stacked_observations ride in the channel axis
            obs4 = observations.to(device=dev, dtype=torch.float32).reshape((G, -1) + tuple(model.observation_shape[1:])).contiguous()
            with torch.cuda.device(dev):
                ws = model._workspace(G, dev)
                check(_lib.lib.mzb_search_resnet(self.tree._h, h, ptr(obs4), ptr(lg), ptr(tp), ptr(nz),
                                                 float(cfg.root_dirichlet_alpha), frac, ptr(sl), ptr(st), int(S), ptr(pool),
                                                 ptr(ws), ws.numel(), ptr(out["visits"]), ptr(out["root_value"]),
                                                 ptr(out["root_predicted_value"]), ptr(out["max_depth"]),
                                                 _lib.current_stream()))
            return out
        return self._run_generic(model, observations, lg, tp, nz, frac, sl, st, S, out)

    def _run_generic(self, model, observations, lg, tp, nz, frac, sl, st, S, out):
        """Any model exposing the reference's initial_inference / recurrent_inference on CUDA tensors
        (e.g. the residual networks): tree kernels + batched network calls per simulation."""
        from . import models
        cfg, tree, dev, G = self.config, self.tree, self.device, self.G
        obs = observations.to(device=dev, dtype=torch.float32)
        v, r, pl, hs = model.initial_inference(obs)
        out["root_predicted_value"].copy_(models.support_to_scalar(v, cfg.support_size).reshape(G))
        reward = models.support_to_scalar(r, cfg.support_size).reshape(G).contiguous()
        hidden = getattr(self, "_generic_hidden", None)
        if hidden is None or hidden.shape[2] != hs[0].numel():
            hidden = self._generic_hidden = torch.empty((G, cfg.num_simulations + 1, hs[0].numel()), device=dev)
        hidden[:, 0] = hs.reshape(G, -1)
        tree.root_init(reward, pl.contiguous(), True, lg, tp, nz, cfg.root_dirichlet_alpha, frac, sl, st)
        parent = torch.empty(G, dtype=torch.int32, device=dev)
        action = torch.empty(G, dtype=torch.int32, device=dev)
        ar = torch.arange(G, device=dev)
        for sim in range(S):
            tree.select(parent, action)
            state = hidden[ar, parent.long()].reshape(hs.shape)
            v, r, pl, ns = model.recurrent_inference(state, action.reshape(G, 1))
            hidden[:, sim + 1] = ns.reshape(G, -1)
            tree.expand_backup(models.support_to_scalar(v, cfg.support_size).reshape(G).contiguous(),
                               models.support_to_scalar(r, cfg.support_size).reshape(G).contiguous(),
                               pl.contiguous(), True)
        stats = tree.root_stats()
        out["visits"].copy_(stats["visits"])
        out["root_value"].copy_(stats["root_value"])
        out["max_depth"].copy_(stats["max_depth"])
        return out
