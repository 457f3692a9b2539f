"""Networks of the self-play hot path: same names, constructor arguments and state-dict keys as the
reference's models.py, inference on hand-written sm_100a kernels (libmzb200.so).

  MuZeroNetwork(config)                 models.py:7-41
  AbstractNetwork.get_weights/set_weights  :56-73
  MuZeroFullyConnectedNetwork           :80-195   -> csrc/mzb_fc.cu (K4/K5)
  MuZeroResidualNetwork                 :432-619  -> csrc/mzb_resnet.cu (K6-K8)
  support_to_scalar / scalar_to_support :641-685

The modules only HOLD the parameters (so `state_dict()` / `load_state_dict()` / `.parameters()` /
`.to()` behave as in the reference and reference checkpoints load unchanged); `initial_inference`
and `recurrent_inference` run the CUDA kernels and raise if the model is not on a CUDA device -
there is no CPU fallback on the hot path.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64


class FcConfig(C.Structure):
    _fields_ = [("obs_dim", _i32), ("encoding_size", _i32), ("n_actions", _i32), ("support_size", _i32),
                ("n_rep", _i32), ("rep", _i32 * 3), ("n_dyn", _i32), ("dyn", _i32 * 3),
                ("n_rew", _i32), ("rew", _i32 * 3), ("n_val", _i32), ("val", _i32 * 3),
                ("n_pol", _i32), ("pol", _i32 * 3)]


_lib.bind("mzb_fc_create", C.c_int, [C.POINTER(_vp), C.POINTER(FcConfig)])
_lib.bind("mzb_fc_destroy", C.c_int, [_vp])
_lib.bind("mzb_fc_num_tensors", C.c_int, [_vp])
_lib.bind("mzb_fc_set_weights", C.c_int, [_vp, C.POINTER(_vp), C.c_int, _vp])
_lib.bind("mzb_fc_initial", C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp])
_lib.bind("mzb_fc_recurrent", C.c_int,
          [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp])
_lib.bind("mzb_support_to_scalar", C.c_int, [_vp, _i64, C.c_int, _vp, _vp])
_lib.bind("mzb_scalar_to_support", C.c_int, [_vp, _i64, C.c_int, _vp, _vp])


class MuZeroNetwork:
    def __new__(cls, config):
        if config.network == "fullyconnected":
            return MuZeroFullyConnectedNetwork(
                config.observation_shape, config.stacked_observations, len(config.action_space),
                config.encoding_size, config.fc_reward_layers, config.fc_value_layers, config.fc_policy_layers,
                config.fc_representation_layers, config.fc_dynamics_layers, config.support_size)
        if config.network == "resnet":
            from .resnet import MuZeroResidualNetwork
            return MuZeroResidualNetwork(
                config.observation_shape, config.stacked_observations, len(config.action_space), config.blocks,
                config.channels, config.reduced_channels_reward, config.reduced_channels_value,
                config.reduced_channels_policy, config.resnet_fc_reward_layers, config.resnet_fc_value_layers,
                config.resnet_fc_policy_layers, config.support_size, config.downsample,
                precision=getattr(config, "resnet_precision", "fp32"))
        raise NotImplementedError('The network parameter should be "fullyconnected" or "resnet".')


def dict_to_cpu(dictionary):
    out = {}
    for key, value in dictionary.items():
        if isinstance(value, torch.Tensor):
            out[key] = value.cpu()
        elif isinstance(value, dict):
            out[key] = dict_to_cpu(value)
        else:
            out[key] = value
    return out


class AbstractNetwork(torch.nn.Module):
    def initial_inference(self, observation):
        raise NotImplementedError

    def recurrent_inference(self, encoded_state, action):
        raise NotImplementedError

    def get_weights(self):
        return dict_to_cpu(self.state_dict())

    def set_weights(self, weights):
        self.load_state_dict(weights)


def mlp(input_size, layer_sizes, output_size, output_activation=torch.nn.Identity, activation=torch.nn.ELU):
    sizes = [input_size] + list(layer_sizes) + [output_size]
    layers = []
    for i in range(len(sizes) - 1):
        layers += [torch.nn.Linear(sizes[i], sizes[i + 1]), (activation if i < len(sizes) - 2 else output_activation)()]
    return torch.nn.Sequential(*layers)


class Replicated(torch.nn.Module):
    """Parameter holder that keeps the reference's `<net>.module.<i>` key names (it wraps every
    sub-network in torch.nn.DataParallel, models.py:98-126; the B200 path shards GAMES over GPUs instead)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: the B200 hot path has no CPU fallback - move the model and inputs to a CUDA device")


class MuZeroFullyConnectedNetwork(AbstractNetwork):
    def __init__(self, observation_shape, stacked_observations, action_space_size, encoding_size, fc_reward_layers,
                 fc_value_layers, fc_policy_layers, fc_representation_layers, fc_dynamics_layers, support_size):
        super().__init__()
        self.action_space_size = action_space_size
        self.support_size = support_size
        self.full_support_size = 2 * support_size + 1
        self.encoding_size = encoding_size
        self.obs_dim = (observation_shape[0] * observation_shape[1] * observation_shape[2] * (stacked_observations + 1)
                        + stacked_observations * observation_shape[1] * observation_shape[2])
        self.representation_network = Replicated(mlp(self.obs_dim, fc_representation_layers, encoding_size))
        self.dynamics_encoded_state_network = Replicated(
            mlp(encoding_size + action_space_size, fc_dynamics_layers, encoding_size))
        self.dynamics_reward_network = Replicated(mlp(encoding_size, fc_reward_layers, self.full_support_size))
        self.prediction_policy_network = Replicated(mlp(encoding_size, fc_policy_layers, action_space_size))
        self.prediction_value_network = Replicated(mlp(encoding_size, fc_value_layers, self.full_support_size))
        cfg = FcConfig()
        cfg.obs_dim, cfg.encoding_size, cfg.n_actions, cfg.support_size = self.obs_dim, encoding_size, action_space_size, support_size
        for name, layers in (("rep", fc_representation_layers), ("dyn", fc_dynamics_layers), ("rew", fc_reward_layers),
                             ("val", fc_value_layers), ("pol", fc_policy_layers)):
            if len(layers) > 3:
                raise NotImplementedError("fully-connected mlp() with more than 3 hidden layers")
            setattr(cfg, "n_" + name, len(layers))
            arr = getattr(cfg, name)
            for i, w in enumerate(layers):
                arr[i] = int(w)
        self._cfg = cfg
        self._h = None
        self._h_device = None
        self._synced = None

    # ---- device handle / weight sync
    def _params_in_order(self):
        nets = (self.representation_network, self.dynamics_encoded_state_network, self.dynamics_reward_network,
                self.prediction_policy_network, self.prediction_value_network)
        out = []
        for net in nets:
            for layer in net.module:
                if isinstance(layer, torch.nn.Linear):
                    out += [layer.weight, layer.bias]
        return out

    def handle(self):
        """mzb_fc_model* for the device the parameters live on, weights re-packed if they changed."""
        params = self._params_in_order()
        dev = params[0].device
        _require_cuda(params[0], "MuZeroFullyConnectedNetwork")
        with torch.cuda.device(dev):
            if self._h is None or self._h_device != dev:
                self._free()
                h = _vp()
                check(_lib.lib.mzb_fc_create(C.byref(h), C.byref(self._cfg)))
                self._h, self._h_device, self._synced = h, dev, None
            stamp = tuple((p.data_ptr(), p._version) for p in params)
            if stamp != self._synced:
                host = [p.detach().to("cpu", torch.float32).contiguous() for p in params]
                arr = (_vp * len(host))(*[t.data_ptr() for t in host])
                check(_lib.lib.mzb_fc_set_weights(self._h, arr, len(host), _lib.current_stream()))
                self._synced = stamp
        return self._h

    def _free(self):
        if getattr(self, "_h", None):
            try:
                _lib.lib.mzb_fc_destroy(self._h)
            except (AttributeError, TypeError):      # interpreter shutdown
                pass
            object.__setattr__(self, "_h", None)     # not Module.__setattr__: torch may be half torn down at exit

    def __del__(self):
        self._free()

    # ---- reference API (models.py:172-195)
    def initial_inference(self, observation):
        _require_cuda(observation, "initial_inference")
        out = self.initial_inference_fused(observation, want_logits=True)
        return out["value_logits"], out["reward_logits"], out["policy_logits"], out["state"]

    def recurrent_inference(self, encoded_state, action):
        _require_cuda(encoded_state, "recurrent_inference")
        out = self.recurrent_inference_fused(encoded_state, action, want_logits=True)
        return out["value_logits"], out["reward_logits"], out["policy_logits"], out["state"]

    # ---- batched entry points used by the device search (scalars + priors straight out of the kernel)
    def initial_inference_fused(self, observation, legal=None, want_logits=False, state_out=None, out_row_stride=None,
                                out_offset=0):
        h = self.handle()
        obs = observation.to(torch.float32).reshape(observation.shape[0], -1).contiguous()
        B, dev = obs.shape[0], obs.device
        assert obs.shape[1] == self.obs_dim, f"observation has {obs.shape[1]} features, network expects {self.obs_dim}"
        A, F = self.action_space_size, self.full_support_size
        res = {"value": torch.empty(B, device=dev), "reward": torch.empty(B, device=dev),
               "priors": torch.empty((B, A), device=dev)}
        if state_out is None:
            res["state"] = torch.empty((B, self.encoding_size), device=dev)
            state_ptr, stride = res["state"], self.encoding_size
        else:
            state_ptr, stride = state_out, out_row_stride
        vl = rl = pl = None
        if want_logits:
            vl = res["value_logits"] = torch.empty((B, F), device=dev)
            rl = res["reward_logits"] = torch.empty((B, F), device=dev)
            pl = res["policy_logits"] = torch.empty((B, A), device=dev)
        lg = None if legal is None else legal.to(torch.uint8).contiguous()
        with torch.cuda.device(dev):
            check(_lib.lib.mzb_fc_initial(h, B, ptr(obs), ptr(lg), ptr(state_ptr), stride, out_offset, ptr(vl), ptr(rl),
                                          ptr(pl), ptr(res["value"]), ptr(res["reward"]), ptr(res["priors"]),
                                          _lib.current_stream()))
        return res

    def recurrent_inference_fused(self, encoded_state, action, want_logits=False, in_slot=None, in_row_stride=None,
                                  slot_stride=0, state_out=None, out_row_stride=None, out_offset=0):
        h = self.handle()
        dev = encoded_state.device
        act = action.reshape(-1).to(torch.int32).contiguous()
        B = act.shape[0]
        A, F = self.action_space_size, self.full_support_size
        if in_row_stride is None:
            encoded_state = encoded_state.to(torch.float32).reshape(B, -1).contiguous()
            in_row_stride = self.encoding_size
        res = {"value": torch.empty(B, device=dev), "reward": torch.empty(B, device=dev),
               "priors": torch.empty((B, A), device=dev)}
        if state_out is None:
            res["state"] = torch.empty((B, self.encoding_size), device=dev)
            state_ptr, stride = res["state"], self.encoding_size
        else:
            state_ptr, stride = state_out, out_row_stride
        vl = rl = pl = None
        if want_logits:
            vl = res["value_logits"] = torch.empty((B, F), device=dev)
            rl = res["reward_logits"] = torch.empty((B, F), device=dev)
            pl = res["policy_logits"] = torch.empty((B, A), device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib.mzb_fc_recurrent(h, B, ptr(encoded_state), in_row_stride, ptr(in_slot), slot_stride, ptr(act),
                                            ptr(state_ptr), stride, out_offset, ptr(vl), ptr(rl), ptr(pl),
                                            ptr(res["value"]), ptr(res["reward"]), ptr(res["priors"]),
                                            _lib.current_stream()))
        return res


def support_to_scalar(logits, support_size):
    """[B, 2S+1] logits -> [B, 1] scalars (models.py:641-662) on the device kernel."""
    _require_cuda(logits, "support_to_scalar")
    x = logits.to(torch.float32).contiguous()
    out = torch.empty((x.shape[0], 1), device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib.mzb_support_to_scalar(ptr(x), x.shape[0], int(support_size), ptr(out), _lib.current_stream()))
    return out


def scalar_to_support(x, support_size):
    """[B, T] scalars -> [B, T, 2S+1] two-hot categorical (models.py:665-685) on the device kernel."""
    _require_cuda(x, "scalar_to_support")
    xf = x.to(torch.float32).contiguous()
    out = torch.empty(tuple(xf.shape) + (2 * support_size + 1,), device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib.mzb_scalar_to_support(ptr(xf), xf.numel(), int(support_size), ptr(out), _lib.current_stream()))
    return out
