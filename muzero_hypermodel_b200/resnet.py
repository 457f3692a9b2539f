"""MuZeroResidualNetwork with the reference's constructor and state-dict keys (models.py:432-619),
inference on csrc/mzb_resnet.cu (+ mzb_conv_tc.cu for the bf16 tensor-core path).

The torch modules below only HOLD the parameters/buffers under the reference's names
(`representation_network.module.resblocks.0.conv1.weight`, ...); they are never called on the hot path.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr
from .models import AbstractNetwork, Replicated, _require_cuda, mlp

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64


class ResnetConfig(C.Structure):
    _fields_ = [("obs_channels", _i32), ("height", _i32), ("width", _i32), ("n_actions", _i32), ("blocks", _i32),
                ("channels", _i32), ("reduced_channels_reward", _i32), ("reduced_channels_value", _i32),
                ("reduced_channels_policy", _i32), ("n_fc_reward", _i32), ("fc_reward", _i32 * 3),
                ("n_fc_value", _i32), ("fc_value", _i32 * 3), ("n_fc_policy", _i32), ("fc_policy", _i32 * 3),
                ("support_size", _i32), ("downsample", _i32), ("precision", _i32)]


_lib.bind("mzb_resnet_create", C.c_int, [C.POINTER(_vp), C.POINTER(ResnetConfig)])
_lib.bind("mzb_resnet_destroy", C.c_int, [_vp])
_lib.bind("mzb_resnet_num_tensors", C.c_int, [_vp])
_lib.bind("mzb_resnet_set_weights", C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_i64), C.c_int, _vp])
_lib.bind("mzb_resnet_workspace_bytes", C.c_size_t, [_vp, _i64])
_lib.bind("mzb_conv_tc_enable", None, [C.c_int])
_lib.bind("mzb_tower16_enable", None, [C.c_int])
_lib.bind("mzb_stem16_enable", None, [C.c_int])
_lib.bind("mzb_resnet_workspace_init", C.c_int, [_vp, _vp, C.c_size_t, _vp])
_lib.bind("mzb_resnet_initial", C.c_int, [_vp, _i64, _vp, _vp, _vp, C.c_size_t, _vp, C.c_int, _i64, _i64] + [_vp] * 7)
_lib.bind("mzb_resnet_conv_probe", C.c_int, [_vp, _i64, _vp, C.c_size_t, C.c_int32, _vp])
_lib.bind("mzb_resnet_recurrent", C.c_int, [_vp, _i64, _vp, C.c_int, _i64, _vp, _i64, _vp, _vp, C.c_size_t, _vp, C.c_int,
                                            _i64, _i64] + [_vp] * 7)


def conv3x3(cin, cout, stride=1):
    return torch.nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=False)


class ResidualBlock(torch.nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv1 = conv3x3(channels, channels)
        self.bn1 = torch.nn.BatchNorm2d(channels)
        self.conv2 = conv3x3(channels, channels)
        self.bn2 = torch.nn.BatchNorm2d(channels)


class DownSample(torch.nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = conv3x3(cin, cout // 2, stride=2)
        self.resblocks1 = torch.nn.ModuleList([ResidualBlock(cout // 2) for _ in range(2)])
        self.conv2 = conv3x3(cout // 2, cout, stride=2)
        self.resblocks2 = torch.nn.ModuleList([ResidualBlock(cout) for _ in range(3)])
        self.resblocks3 = torch.nn.ModuleList([ResidualBlock(cout) for _ in range(3)])


class RepresentationNetwork(torch.nn.Module):
    def __init__(self, observation_shape, stacked_observations, num_blocks, num_channels, downsample):
        super().__init__()
        cin = observation_shape[0] * (stacked_observations + 1) + stacked_observations
        if downsample:
            if downsample != "resnet":
                raise NotImplementedError('device path implements downsample=False or "resnet"')
            self.downsample_net = DownSample(cin, num_channels)
        self.conv = conv3x3(cin, num_channels)
        self.bn = torch.nn.BatchNorm2d(num_channels)
        self.resblocks = torch.nn.ModuleList([ResidualBlock(num_channels) for _ in range(num_blocks)])


class DynamicsNetwork(torch.nn.Module):
    def __init__(self, num_blocks, num_channels, reduced_channels_reward, fc_reward_layers, full_support_size,
                 block_output_size_reward):
        super().__init__()
        self.conv = conv3x3(num_channels, num_channels - 1)
        self.bn = torch.nn.BatchNorm2d(num_channels - 1)
        self.resblocks = torch.nn.ModuleList([ResidualBlock(num_channels - 1) for _ in range(num_blocks)])
        self.conv1x1_reward = torch.nn.Conv2d(num_channels - 1, reduced_channels_reward, 1)
        self.fc = mlp(block_output_size_reward, fc_reward_layers, full_support_size)


class PredictionNetwork(torch.nn.Module):
    def __init__(self, action_space_size, num_blocks, num_channels, reduced_channels_value, reduced_channels_policy,
                 fc_value_layers, fc_policy_layers, full_support_size, block_output_size_value, block_output_size_policy):
        super().__init__()
        self.resblocks = torch.nn.ModuleList([ResidualBlock(num_channels) for _ in range(num_blocks)])
        self.conv1x1_value = torch.nn.Conv2d(num_channels, reduced_channels_value, 1)
        self.conv1x1_policy = torch.nn.Conv2d(num_channels, reduced_channels_policy, 1)
        self.fc_value = mlp(block_output_size_value, fc_value_layers, full_support_size)
        self.fc_policy = mlp(block_output_size_policy, fc_policy_layers, action_space_size)


class MuZeroResidualNetwork(AbstractNetwork):
    PRECISION = {"fp32": 0, "bf16": 1}

    def __init__(self, observation_shape, stacked_observations, action_space_size, num_blocks, num_channels,
                 reduced_channels_reward, reduced_channels_value, reduced_channels_policy, fc_reward_layers,
                 fc_value_layers, fc_policy_layers, support_size, downsample, precision="fp32"):
        super().__init__()
        self.action_space_size = action_space_size
        self.support_size = support_size
        self.full_support_size = 2 * support_size + 1
        self.observation_shape = tuple(observation_shape)
        self.num_channels = num_channels
        self.precision = precision
        h = -(-observation_shape[1] // 16) if downsample else observation_shape[1]
        w = -(-observation_shape[2] // 16) if downsample else observation_shape[2]
        self.latent_shape = (num_channels, h, w)
        self.representation_network = Replicated(
            RepresentationNetwork(observation_shape, stacked_observations, num_blocks, num_channels, downsample))
        self.dynamics_network = Replicated(
            DynamicsNetwork(num_blocks, num_channels + 1, reduced_channels_reward, fc_reward_layers,
                            self.full_support_size, reduced_channels_reward * h * w))
        self.prediction_network = Replicated(
            PredictionNetwork(action_space_size, num_blocks, num_channels, reduced_channels_value,
                              reduced_channels_policy, fc_value_layers, fc_policy_layers, self.full_support_size,
                              reduced_channels_value * h * w, reduced_channels_policy * h * w))
        cfg = ResnetConfig()
        cfg.obs_channels = observation_shape[0] * (stacked_observations + 1) + stacked_observations
        cfg.height, cfg.width, cfg.n_actions = observation_shape[1], observation_shape[2], action_space_size
        cfg.blocks, cfg.channels = num_blocks, num_channels
        cfg.reduced_channels_reward, cfg.reduced_channels_value = reduced_channels_reward, reduced_channels_value
        cfg.reduced_channels_policy = reduced_channels_policy
        for name, layers in (("reward", fc_reward_layers), ("value", fc_value_layers), ("policy", fc_policy_layers)):
            if len(layers) > 3:
                raise NotImplementedError("head mlp() with more than 3 hidden layers")
            setattr(cfg, "n_fc_" + name, len(layers))
            arr = getattr(cfg, "fc_" + name)
            for i, wd in enumerate(layers):
                arr[i] = int(wd)
        cfg.support_size = support_size
        cfg.downsample = 1 if downsample else 0
        cfg.precision = self.PRECISION[precision]
        self._cfg = cfg
        self._h = None
        self._h_device = None
        self._synced = None
        self._ws = None

    def set_precision(self, precision):
        """"fp32" (exact CUDA-core path) or "bf16" (tcgen05 tensor-core path, fp32 accumulation)."""
        if precision != self.precision:
            self._free()
            self.precision = precision
            self._cfg.precision = self.PRECISION[precision]
        return self

    def _tensors_in_order(self):
        return [v for k, v in self.state_dict().items() if not k.endswith("num_batches_tracked")]

    def handle(self):
        params = self._tensors_in_order()
        dev = params[0].device
        _require_cuda(params[0], "MuZeroResidualNetwork")
        with torch.cuda.device(dev):
            if self._h is None or self._h_device != dev:
                self._free()
                h = _vp()
                check(_lib.lib.mzb_resnet_create(C.byref(h), C.byref(self._cfg)))
                self._h, self._h_device, self._synced = h, dev, None
            stamp = tuple((p.data_ptr(), p._version) for p in params)
            if stamp != self._synced:
                host = [p.detach().to("cpu", torch.float32).contiguous() for p in params]
                arr = (_vp * len(host))(*[t.data_ptr() for t in host])
                numel = (_i64 * len(host))(*[t.numel() for t in host])
                check(_lib.lib.mzb_resnet_set_weights(self._h, arr, numel, len(host), _lib.current_stream()))
                self._synced = stamp
        return self._h

    def _free(self):
        if getattr(self, "_h", None):
            try:
                _lib.lib.mzb_resnet_destroy(self._h)
            except (AttributeError, TypeError):      # interpreter shutdown
                pass
            object.__setattr__(self, "_h", None)     # not Module.__setattr__: torch may be half torn down at exit

    def __del__(self):
        self._free()

    def _workspace(self, B, dev):
        need = _lib.lib.mzb_resnet_workspace_bytes(self._h, B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            check(_lib.lib.mzb_resnet_workspace_init(self._h, ptr(self._ws), need, _lib.current_stream()))
        return self._ws

    # ---- reference API
    def initial_inference(self, observation):
        _require_cuda(observation, "initial_inference")
        o = self.initial_inference_fused(observation, want_logits=True)
        return o["value_logits"], o["reward_logits"], o["policy_logits"], o["state"]

    def recurrent_inference(self, encoded_state, action):
        _require_cuda(encoded_state, "recurrent_inference")
        o = self.recurrent_inference_fused(encoded_state, action, want_logits=True)
        return o["value_logits"], o["reward_logits"], o["policy_logits"], o["state"]

    def _outputs(self, B, dev, want_logits):
        A, F = self.action_space_size, self.full_support_size
        res = {"value": torch.empty(B, device=dev), "reward": torch.empty(B, device=dev),
               "priors": torch.empty((B, A), device=dev)}
        vl = rl = pl = None
        if want_logits:
            vl = res["value_logits"] = torch.empty((B, F), device=dev)
            rl = res["reward_logits"] = torch.empty((B, F), device=dev)
            pl = res["policy_logits"] = torch.empty((B, A), device=dev)
        return res, vl, rl, pl

    def initial_inference_fused(self, observation, legal=None, want_logits=False, state_out=None, state_layout=0,
                                out_row_stride=None, out_offset=0):
        h = self.handle()
        obs = observation.to(torch.float32).contiguous()
        B, dev = obs.shape[0], obs.device
        res, vl, rl, pl = self._outputs(B, dev, want_logits)
        if state_out is None:
            res["state"] = torch.empty((B,) + self.latent_shape, device=dev)
            state_out, state_layout, out_row_stride = res["state"], 0, res["state"][0].numel()
        lg = None if legal is None else legal.to(torch.uint8).contiguous()
        with torch.cuda.device(dev):
            ws = self._workspace(B, dev)
            check(_lib.lib.mzb_resnet_initial(h, B, ptr(obs), ptr(lg), ptr(ws), ws.numel(), ptr(state_out), state_layout,
                                              out_row_stride, out_offset, ptr(vl), ptr(rl), ptr(pl), ptr(res["value"]),
                                              ptr(res["reward"]), ptr(res["priors"]), _lib.current_stream()))
        return res

    def recurrent_inference_fused(self, encoded_state, action, want_logits=False, in_layout=0, in_slot=None,
                                  in_row_stride=None, slot_stride=0, state_out=None, out_layout=0, out_row_stride=None,
                                  out_offset=0):
        h = self.handle()
        dev = encoded_state.device
        act = action.reshape(-1).to(torch.int32).contiguous()
        B = act.shape[0]
        if in_row_stride is None:
            encoded_state = encoded_state.to(torch.float32).contiguous()
            in_row_stride = encoded_state[0].numel()
        res, vl, rl, pl = self._outputs(B, dev, want_logits)
        if state_out is None:
            res["state"] = torch.empty((B,) + self.latent_shape, device=dev)
            state_out, out_layout, out_row_stride = res["state"], 0, res["state"][0].numel()
        with torch.cuda.device(dev):
            ws = self._workspace(B, dev)
            check(_lib.lib.mzb_resnet_recurrent(h, B, ptr(encoded_state), in_layout, in_row_stride, ptr(in_slot),
                                                slot_stride, ptr(act), ptr(ws), ws.numel(), ptr(state_out), out_layout,
                                                out_row_stride, out_offset, ptr(vl), ptr(rl), ptr(pl), ptr(res["value"]),
                                                ptr(res["reward"]), ptr(res["priors"]), _lib.current_stream()))
        return res
