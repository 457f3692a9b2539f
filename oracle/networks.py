"""CPU restatement of the reference networks and support codec (test infrastructure).

Follows /root/reference/models.py:
  mlp :626-638, MuZeroFullyConnectedNetwork :80-195, ResidualBlock :213-229,
  DownSample :233-275, RepresentationNetwork :300-349, DynamicsNetwork :352-387,
  PredictionNetwork :390-429, MuZeroResidualNetwork :432-619,
  support_to_scalar :641-662, scalar_to_support :665-685.

The reference builds torch.nn.Module graphs wrapped in DataParallel; this restatement is
purely functional and runs straight off a reference-format state-dict (keys keep the
`.module.` infix): numpy float32 for the fully-connected family,
torch.nn.functional primitives on CPU for the residual family (conv/batch-norm arithmetic is
PyTorch's in the reference too).  Pinned to tolerance by tests/golden/net_*.npz (outputs of
the reference models on the shipped cartpole checkpoint and on seeded random weights).
"""
import math

import numpy as np

F32 = np.float32


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def _linear_stack(sd, prefix):
    """[(W, b), ...] of the Linear layers of an mlp() stored under `prefix.module.<2i>`."""
    layers, i = [], 0
    while f"{prefix}.module.{2 * i}.weight" in sd or f"{prefix}.{2 * i}.weight" in sd:
        key = f"{prefix}.module.{2 * i}" if f"{prefix}.module.{2 * i}.weight" in sd else f"{prefix}.{2 * i}"
        layers.append((_np(sd[key + ".weight"]).astype(F32), _np(sd[key + ".bias"]).astype(F32)))
        i += 1
    return layers


def _elu(x):
    # ATen's ELU kernel evaluates exp(x) - 1 (not expm1) for x <= 0.
    return np.where(x > 0, x, np.exp(np.minimum(x, 0), dtype=F32) - F32(1)).astype(F32)


def _mlp(layers, x):
    for i, (w, b) in enumerate(layers):
        x = (x @ w.T + b).astype(F32)
        if i < len(layers) - 1:
            x = _elu(x)
    return x


def _minmax_rows(x):
    """Per-sample min-max scaling with the `scale < 1e-5 -> += 1e-5` guard (models.py:138-145)."""
    lo = x.min(axis=1, keepdims=True)
    hi = x.max(axis=1, keepdims=True)
    scale = (hi - lo).astype(F32)
    scale = np.where(scale < F32(1e-5), scale + F32(1e-5), scale).astype(F32)
    return ((x - lo) / scale).astype(F32)


def zero_reward_logits(batch, support_size):
    """log(one-hot(centre)) = -inf / 0 rows (models.py:176-183)."""
    n = 2 * support_size + 1
    r = np.full((batch, n), -np.inf, dtype=F32)
    r[:, n // 2] = 0.0
    return r


class FullyConnected:
    def __init__(self, state_dict, n_actions, support_size):
        self.rep = _linear_stack(state_dict, "representation_network")
        self.dyn = _linear_stack(state_dict, "dynamics_encoded_state_network")
        self.rew = _linear_stack(state_dict, "dynamics_reward_network")
        self.pol = _linear_stack(state_dict, "prediction_policy_network")
        self.val = _linear_stack(state_dict, "prediction_value_network")
        self.A = n_actions
        self.S = support_size

    def initial_inference(self, observation):
        obs = np.asarray(observation, dtype=F32)
        x = obs.reshape(obs.shape[0], -1)
        state = _minmax_rows(_mlp(self.rep, x))                      # :133-145
        return (_mlp(self.val, state), zero_reward_logits(len(x), self.S), _mlp(self.pol, state), state)

    def recurrent_inference(self, state, action):
        state = np.asarray(state, dtype=F32)
        a = np.asarray(action).reshape(-1).astype(np.int64)
        onehot = np.zeros((len(a), self.A), dtype=F32)              # :149-155
        onehot[np.arange(len(a)), a] = 1.0
        nxt = _mlp(self.dyn, np.concatenate([state, onehot], axis=1))
        reward = _mlp(self.rew, nxt)                                 # un-normalised state :159
        nxt = _minmax_rows(nxt)
        return _mlp(self.val, nxt), reward, _mlp(self.pol, nxt), nxt


class Residual:
    """Residual family, eval-mode batch-norm, torch CPU functional ops."""

    def __init__(self, state_dict, observation_shape, n_actions, blocks, support_size, downsample=False):
        import torch
        self.t = torch
        self.sd = {k: (v.detach().float().cpu() if hasattr(v, "detach") else torch.as_tensor(v).float())
                   for k, v in state_dict.items()}
        self.A = n_actions
        self.S = support_size
        self.blocks = blocks
        self.downsample = downsample

    # -- building blocks
    def _bn(self, x, p):
        sd = self.sd
        return self.t.nn.functional.batch_norm(
            x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
            training=False, eps=1e-5)

    def _conv(self, x, p, stride=1, padding=1):
        return self.t.nn.functional.conv2d(x, self.sd[p + ".weight"], self.sd.get(p + ".bias"),
                                           stride=stride, padding=padding)

    def _block(self, x, p):                                          # :213-229
        relu = self.t.nn.functional.relu
        out = relu(self._bn(self._conv(x, p + ".conv1"), p + ".bn1"))
        out = self._bn(self._conv(out, p + ".conv2"), p + ".bn2")
        return relu(out + x)

    def _fc(self, x, p):
        i, n = 0, 0
        while f"{p}.{2 * n}.weight" in self.sd:
            n += 1
        for i in range(n):
            x = self.t.nn.functional.linear(x, self.sd[f"{p}.{2 * i}.weight"], self.sd[f"{p}.{2 * i}.bias"])
            if i < n - 1:
                x = self.t.nn.functional.elu(x)
        return x

    def _minmax(self, s):                                            # per channel over HxW :525-549
        b, c, h, w = s.shape
        flat = s.reshape(b, c, h * w)
        lo = flat.min(2, keepdim=True)[0].unsqueeze(-1)
        hi = flat.max(2, keepdim=True)[0].unsqueeze(-1)
        scale = hi - lo
        scale = self.t.where(scale < 1e-5, scale + 1e-5, scale)
        return (s - lo) / scale

    def _down(self, x, p):                                           # :233-275
        x = self._conv(x, p + ".conv1", stride=2)
        for i in range(2):
            x = self._block(x, f"{p}.resblocks1.{i}")
        x = self._conv(x, p + ".conv2", stride=2)
        for i in range(3):
            x = self._block(x, f"{p}.resblocks2.{i}")
        x = self.t.nn.functional.avg_pool2d(x, 3, 2, 1)
        for i in range(3):
            x = self._block(x, f"{p}.resblocks3.{i}")
        return self.t.nn.functional.avg_pool2d(x, 3, 2, 1)

    def _prediction(self, s):                                        # :420-429
        p = "prediction_network.module"
        x = s
        for i in range(self.blocks):
            x = self._block(x, f"{p}.resblocks.{i}")
        v = self._conv(x, p + ".conv1x1_value", padding=0).reshape(len(x), -1)
        pi = self._conv(x, p + ".conv1x1_policy", padding=0).reshape(len(x), -1)
        return self._fc(pi, p + ".fc_policy"), self._fc(v, p + ".fc_value")

    def initial_inference(self, observation):
        t = self.t
        with t.no_grad():
            x = t.as_tensor(np.asarray(observation, dtype=F32))
            p = "representation_network.module"
            if self.downsample == "resnet":
                x = self._down(x, p + ".downsample_net")
            elif self.downsample:
                raise NotImplementedError("CNN downsample is not on the hot path of any BASELINE config")
            else:
                x = t.nn.functional.relu(self._bn(self._conv(x, p + ".conv"), p + ".bn"))
            for i in range(self.blocks):
                x = self._block(x, f"{p}.resblocks.{i}")
            s = self._minmax(x)
            pol, val = self._prediction(s)
            rew = t.as_tensor(zero_reward_logits(len(x), self.S))
            return val.numpy(), rew.numpy(), pol.numpy(), s.numpy()

    def recurrent_inference(self, state, action):
        t = self.t
        with t.no_grad():
            s = t.as_tensor(np.asarray(state, dtype=F32))
            a = t.as_tensor(np.asarray(action)).reshape(-1, 1).float()
            plane = a[:, :, None, None] * t.ones((s.shape[0], 1, s.shape[2], s.shape[3])) / self.A   # :553-568
            x = t.cat((s, plane), dim=1)
            p = "dynamics_network.module"
            x = t.nn.functional.relu(self._bn(self._conv(x, p + ".conv"), p + ".bn"))
            for i in range(self.blocks):
                x = self._block(x, f"{p}.resblocks.{i}")
            r = self._conv(x, p + ".conv1x1_reward", padding=0).reshape(len(x), -1)
            reward = self._fc(r, p + ".fc")
            nxt = self._minmax(x)
            pol, val = self._prediction(nxt)
            return val.numpy(), reward.numpy(), pol.numpy(), nxt.numpy()


def support_to_scalar(logits, support_size):
    """models.py:641-662 in numpy float32: softmax expectation then inverse h-transform."""
    x = np.asarray(logits, dtype=F32)
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m, dtype=F32)
    p = (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)
    support = np.arange(-support_size, support_size + 1, dtype=F32)
    v = (p * support).sum(axis=1, keepdims=True, dtype=F32)
    eps = F32(0.001)
    inner = np.sqrt(F32(1) + F32(4) * eps * (np.abs(v) + F32(1) + eps), dtype=F32)
    return (np.sign(v) * (((inner - F32(1)) / (F32(2) * eps)) ** 2 - F32(1))).astype(F32)


def scalar_to_support(x, support_size):
    """models.py:665-685 in numpy float32: h-transform, clamp, two-hot."""
    x = np.asarray(x, dtype=F32)
    x = (np.sign(x) * (np.sqrt(np.abs(x) + F32(1), dtype=F32) - F32(1)) + F32(0.001) * x).astype(F32)
    x = np.clip(x, -support_size, support_size).astype(F32)
    floor = np.floor(x)
    prob = (x - floor).astype(F32)
    out = np.zeros(x.shape + (2 * support_size + 1,), dtype=F32)
    lo = (floor + support_size).astype(np.int64)
    b, t = np.indices(x.shape)
    out[b, t, lo] = F32(1) - prob
    hi = lo + 1
    ok = hi <= 2 * support_size
    out[b[ok], t[ok], hi[ok]] = prob[ok]
    # reference quirk (:682-684): an out-of-range upper index is redirected to slot 0 with prob 0,
    # and scatter_ then WRITES that 0 into slot 0.
    bad = ~ok
    out[b[bad], t[bad], 0] = np.where(lo[bad] == 0, out[b[bad], t[bad], 0], F32(0))
    return out
