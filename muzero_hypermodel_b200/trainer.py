"""Trainer: the reference's class (trainer.py:11-298) on one GPU per rank.

Same constructor, `update_lr`, `update_weights(batch)`, `continuous_update_weights(replay_buffer, shared_storage)`,
loss definition (cross-entropy on the categorical support, value loss weight, PER importance weights, 0.5 gradient
scaling into the dynamics function, per-sample 1 / gradient_scale) and return values.

What runs where - stated plainly, because this row is NOT yet on hand-written kernels end to end:
  * the unrolled forward / backward is PyTorch autograd over the parameter-holder modules of `models` / `resnet`
    (library kernels: cuBLAS / cuDNN), in the reference's own operation order;
  * `scalar_to_support` / `support_to_scalar` are this package's device kernels;
  * every parameter and every gradient is a VIEW into one flat float32 bucket each, so the data-parallel step is one
    NCCL all-reduce of the gradient bucket (`dist.allreduce_gradients`' collective, here without the pack / unpack
    copies) followed by ONE launch of the hand-written flat optimiser kernel (csrc/mzb_optim.cu) instead of
    torch.optim's per-tensor launches;
  * batches come from the device replay store as device tensors (no host hop) and priorities go back the same way.
"""
import ctypes as C
import time

import numpy
import torch
import torch.distributed as tdist

from . import _lib, models
from ._lib import check, ptr

_vp, _i64, _f64 = C.c_void_p, C.c_int64, C.c_double
_lib.bind("mzb_adam_step", C.c_int, [_vp, _vp, _vp, _vp, _i64, _f64, _f64, _f64, _f64, _f64, _i64, _f64, _vp])
_lib.bind("mzb_sgd_step", C.c_int, [_vp, _vp, _vp, _i64, _f64, _f64, _f64, _i64, _f64, _vp])



class FcTrainDesc(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("encoding_size", C.c_int32), ("n_actions", C.c_int32), ("support_size", C.c_int32),
                ("n_layers", C.c_int32 * 5), ("in_", (C.c_int32 * 4) * 5), ("out", (C.c_int32 * 4) * 5),
                ("w_off", (C.c_int64 * 4) * 5), ("b_off", (C.c_int64 * 4) * 5)]


_lib.bind("mzb_fc_train_workspace_bytes", C.c_int64, [C.POINTER(FcTrainDesc), C.c_int32, C.c_int32])
_lib.bind("mzb_fc_train_fits", C.c_int, [C.POINTER(FcTrainDesc), C.c_int32, C.c_int32])
_lib.bind("mzb_fc_train_grad", C.c_int, [C.POINTER(FcTrainDesc), _vp, _i64, C.c_int32, C.c_int32] + [_vp] * 8 + [_f64, _f64] + [_vp] * 5 + [_i64, _vp])

F = torch.nn.functional


# ------------------------------------------------------------------------------------------ differentiable forward
def _minmax(x):
    """Per-sample min-max scaling over everything but the batch dimension with the +1e-5 guard
    (models.py:138-145 for vectors, :525-549 for [B, C, H, W]: per sample AND channel)."""
    if x.dim() == 2:
        lo, hi = x.min(1, keepdim=True)[0], x.max(1, keepdim=True)[0]
    else:
        flat = x.reshape(x.shape[0], x.shape[1], -1)
        lo, hi = flat.min(2, keepdim=True)[0].unsqueeze(-1), flat.max(2, keepdim=True)[0].unsqueeze(-1)
    scale = hi - lo
    scale = torch.where(scale < 1e-5, scale + 1e-5, scale)
    return (x - lo) / scale


def _block(b, x):
    out = F.relu(b.bn1(b.conv1(x)))
    out = b.bn2(b.conv2(out))
    return F.relu(out + x)


class _FcGraph:
    """MuZeroFullyConnectedNetwork.forward pieces (models.py:128-195)."""

    def __init__(self, net):
        self.n = net

    def initial(self, obs):
        n = self.n
        state = _minmax(n.representation_network(obs.reshape(obs.shape[0], -1).float()))
        policy, value = n.prediction_policy_network(state), n.prediction_value_network(state)
        reward = torch.full((obs.shape[0], n.full_support_size), -float("inf"), device=obs.device)
        reward[:, n.full_support_size // 2] = 0.0            # log of the one-hot centre bin (:176-183)
        return value, reward, policy, state

    def recurrent(self, state, action):
        n = self.n
        onehot = torch.zeros((action.shape[0], n.action_space_size), device=action.device)
        onehot.scatter_(1, action.long(), 1.0)
        nxt = n.dynamics_encoded_state_network(torch.cat((state, onehot), dim=1))
        reward = n.dynamics_reward_network(nxt)              # on the un-normalised state (:157-159)
        nxt = _minmax(nxt)
        return n.prediction_value_network(nxt), reward, n.prediction_policy_network(nxt), nxt


class _ResnetGraph:
    """MuZeroResidualNetwork.forward pieces (models.py:206-619)."""

    def __init__(self, net):
        self.n = net

    def _represent(self, obs):
        r = self.n.representation_network.module
        if hasattr(r, "downsample_net"):
            d = r.downsample_net
            x = d.conv1(obs)
            for b in d.resblocks1:
                x = _block(b, x)
            x = d.conv2(x)
            for b in d.resblocks2:
                x = _block(b, x)
            x = F.avg_pool2d(x, kernel_size=3, stride=2, padding=1)
            for b in d.resblocks3:
                x = _block(b, x)
            x = F.avg_pool2d(x, kernel_size=3, stride=2, padding=1)
        else:
            x = F.relu(r.bn(r.conv(obs)))
        for b in r.resblocks:
            x = _block(b, x)
        return _minmax(x)

    def _predict(self, state):
        p = self.n.prediction_network.module
        x = state
        for b in p.resblocks:
            x = _block(b, x)
        value = p.fc_value(p.conv1x1_value(x).reshape(x.shape[0], -1))
        policy = p.fc_policy(p.conv1x1_policy(x).reshape(x.shape[0], -1))
        return policy, value

    def initial(self, obs):
        n = self.n
        state = self._represent(obs.float())
        policy, value = self._predict(state)
        reward = torch.full((obs.shape[0], n.full_support_size), -float("inf"), device=obs.device)
        reward[:, n.full_support_size // 2] = 0.0
        return value, reward, policy, state

    def recurrent(self, state, action):
        n, d = self.n, self.n.dynamics_network.module
        plane = (action.float() / n.action_space_size).reshape(-1, 1, 1, 1).expand(-1, 1, state.shape[2], state.shape[3])
        x = F.relu(d.bn(d.conv(torch.cat((state, plane), dim=1))))
        for b in d.resblocks:
            x = _block(b, x)
        reward = d.fc(d.conv1x1_reward(x).reshape(x.shape[0], -1))
        nxt = _minmax(x)
        policy, value = self._predict(nxt)
        return value, reward, policy, nxt


# ------------------------------------------------------------------------------------------ loss of one batch
def loss_function(value, reward, policy_logits, target_value, target_reward, target_policy):
    """Cross-entropy against the categorical targets (trainer.py:267-284)."""
    value_loss = (-target_value * F.log_softmax(value, dim=1)).sum(1)
    reward_loss = (-target_reward * F.log_softmax(reward, dim=1)).sum(1)
    policy_loss = (-target_policy * F.log_softmax(policy_logits, dim=1)).sum(1)
    return value_loss, reward_loss, policy_loss


def unrolled_loss(graph, cfg, tensors, scalar_to_support=None, support_to_scalar=None):
    """The training objective of one batch (trainer.py:140-243): initial inference + K recurrent steps, per-step
    cross-entropies with the 0.5 gradient scale into the dynamics function and 1 / gradient_scale per sample, value loss
    weight, PER importance weights, batch mean.  `tensors` = (observation [B,...], action [B,K+1] i64, target value /
    reward [B,K+1], target policy [B,K+1,A], weight [B] | None, gradient scale [B,K+1]) on one device.
    Returns (loss, value_loss [B], reward_loss [B], policy_loss [B], priorities [B,K+1]).
    The codec callables default to this package's device kernels."""
    scalar_to_support = scalar_to_support or models.scalar_to_support
    support_to_scalar = support_to_scalar or models.support_to_scalar
    observation_batch, action_batch, target_value_scalar, target_reward, target_policy, weight_batch, gradient_scale_batch = tensors
    action_batch = action_batch.unsqueeze(-1)
    target_value = scalar_to_support(target_value_scalar, cfg.support_size)
    target_reward = scalar_to_support(target_reward, cfg.support_size)
    priorities = torch.zeros_like(target_value_scalar)

    value, reward, policy_logits, hidden_state = graph.initial(observation_batch)
    predictions = [(value, reward, policy_logits)]
    for i in range(1, action_batch.shape[1]):
        value, reward, policy_logits, hidden_state = graph.recurrent(hidden_state, action_batch[:, i])
        hidden_state.register_hook(lambda grad: grad * 0.5)        # gradient scale into the dynamics function
        predictions.append((value, reward, policy_logits))

    # the reference's per-step hooks all divide by the LAST column of gradient_scale_batch (late-binding closure,
    # trainer.py:207-215); the columns of a row are equal by construction (replay_buffer.py:103-110)
    gscale = gradient_scale_batch[:, action_batch.shape[1] - 1]
    value, reward, policy_logits = predictions[0]
    value_loss, _, policy_loss = loss_function(value, reward, policy_logits, target_value[:, 0], target_reward[:, 0], target_policy[:, 0])
    reward_loss = torch.zeros_like(value_loss)
    priorities[:, 0] = (support_to_scalar(value.detach(), cfg.support_size).reshape(-1) - target_value_scalar[:, 0]).abs() ** cfg.PER_alpha
    for i in range(1, len(predictions)):
        value, reward, policy_logits = predictions[i]
        cv, cr, cp = loss_function(value, reward, policy_logits, target_value[:, i], target_reward[:, i], target_policy[:, i])
        for x in (cv, cr, cp):
            x.register_hook(lambda grad: grad / gscale)
        value_loss = value_loss + cv
        reward_loss = reward_loss + cr
        policy_loss = policy_loss + cp
        priorities[:, i] = (support_to_scalar(value.detach(), cfg.support_size).reshape(-1) - target_value_scalar[:, i]).abs() ** cfg.PER_alpha
    loss = value_loss * cfg.value_loss_weight + reward_loss + policy_loss
    if cfg.PER:
        loss = loss * weight_batch
    return loss.mean(), value_loss, reward_loss, policy_loss, priorities


def training_graph(model, cfg):
    """The differentiable forward of a `models.MuZeroNetwork` (autograd over its parameter-holder modules)."""
    return _FcGraph(model) if cfg.network == "fullyconnected" else _ResnetGraph(model)


# ------------------------------------------------------------------------------------------ trainer
class Trainer:
    def __init__(self, initial_checkpoint, config, device=None):
        self.config = config
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("Trainer: the B200 path has no CPU fallback")
        numpy.random.seed(config.seed)
        torch.manual_seed(config.seed)
        self.model = models.MuZeroNetwork(config)
        if initial_checkpoint.get("weights") is not None:
            self.model.set_weights(initial_checkpoint["weights"])
        self.model.to(self.device)
        self.model.train()
        self.training_step = int(initial_checkpoint.get("training_step", 0))
        if config.optimizer not in ("Adam", "SGD"):
            raise NotImplementedError(f"{config.optimizer} is not implemented. You can change the optimizer manually in trainer.py.")
        self.graph = training_graph(self.model, config)
        # flat buckets: parameters and gradients become views (the DDP bucket idea, for the whole model at once)
        unused = set()
        if config.network == "resnet" and config.downsample:        # never reached by forward (models.py:338-345):
            r = self.model.representation_network.module             # torch.optim skips parameters without a gradient
            unused = {id(p) for p in list(r.conv.parameters()) + list(r.bn.parameters())}
        self.params = [p for p in self.model.parameters() if p.requires_grad and id(p) not in unused]
        # torch.optim's state_dict numbers ALL model.parameters() (the unused ones keep their index and simply have no
        # state): index in that numbering of every parameter the flat bucket holds
        self.n_all_params = sum(1 for _ in self.model.parameters())
        self.param_index = [i for i, p in enumerate(self.model.parameters()) if p.requires_grad and id(p) not in unused]
        total = sum(p.numel() for p in self.params)
        self.flat_param = torch.empty(total, dtype=torch.float32, device=self.device)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device)
        off = 0
        for p in self.params:
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view_as(p)
            p.grad = self.flat_grad[off:off + n].view_as(p)
            off += n
        self.state1 = torch.zeros(total, dtype=torch.float32, device=self.device)        # Adam exp_avg / SGD momentum buffer
        self.state2 = torch.zeros(total, dtype=torch.float32, device=self.device)        # Adam exp_avg_sq
        self.opt_step = 0
        self.lr = float(config.lr_init)
        self.use_cuda_graph = True
        self._graphs = {}
        st = initial_checkpoint.get("optimizer_state")
        if st is not None:
            self.load_optimizer_state(st)

    # -- optimiser state in torch.optim's state_dict format (shared_storage / model.checkpoint compatibility)
    def optimizer_state(self):
        state, off = {}, 0
        for i, p in zip(self.param_index, self.params):
            n = p.numel()
            if self.opt_step > 0:
                if self.config.optimizer == "Adam":
                    state[i] = {"step": torch.tensor(float(self.opt_step)), "exp_avg": self.state1[off:off + n].view_as(p).cpu().clone(),
                                "exp_avg_sq": self.state2[off:off + n].view_as(p).cpu().clone()}
                else:
                    state[i] = {"momentum_buffer": self.state1[off:off + n].view_as(p).cpu().clone()}
            off += n
        group = {"lr": self.lr, "weight_decay": self.config.weight_decay, "params": list(range(self.n_all_params))}
        if self.config.optimizer == "Adam":
            group.update(betas=(0.9, 0.999), eps=1e-8, amsgrad=False)
        else:
            group.update(momentum=self.config.momentum, dampening=0, nesterov=False)
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state(self, st):
        off = 0
        for i, p in zip(self.param_index, self.params):
            n = p.numel()
            s = st["state"].get(i)
            if s:
                if "exp_avg" in s:
                    self.state1[off:off + n].copy_(s["exp_avg"].reshape(-1))
                    self.state2[off:off + n].copy_(s["exp_avg_sq"].reshape(-1))
                    self.opt_step = int(float(s["step"]))
                elif s.get("momentum_buffer") is not None:
                    self.state1[off:off + n].copy_(s["momentum_buffer"].reshape(-1))
                    self.opt_step = max(self.opt_step, 1)
            off += n

    # -- trainer.py:257-265
    def update_lr(self):
        self.lr = self.config.lr_init * self.config.lr_decay_rate ** (self.training_step / self.config.lr_decay_steps)

    loss_function = staticmethod(loss_function)

    # -- trainer.py:124-255
    def update_weights(self, batch):
        dev = self.device
        t = lambda x, dt: None if x is None else torch.as_tensor(numpy.asarray(x) if not torch.is_tensor(x) else x).to(dev, dt)
        observation_batch, action_batch, target_value, target_reward, target_policy, weight_batch, gradient_scale_batch = batch
        tensors = (t(observation_batch, torch.float32), t(action_batch, torch.int64), t(target_value, torch.float32),
                   t(target_reward, torch.float32), t(target_policy, torch.float32),
                   t(weight_batch, torch.float32) if self.config.PER else None, t(gradient_scale_batch, torch.float32))
        loss, value_loss, reward_loss, policy_loss, priorities = self._forward_backward(tensors)
        self._step()
        self.training_step += 1
        # a fresh tensor like the reference's numpy array: `priorities` aliases the CUDA graph's static output, which
        # the next replay overwrites
        priorities = priorities.clone()
        return priorities, loss.item(), value_loss.mean().item(), reward_loss.mean().item(), policy_loss.mean().item()

    # -- one-kernel training step of the fully-connected family (csrc/mzb_fc_train.cu)
    def _fc_desc(self):
        """Layer table of the FC network over the flat parameter bucket, or None (residual family / frozen parameters)."""
        if getattr(self, "_fc_desc_cache", 0) != 0:
            return self._fc_desc_cache
        self._fc_desc_cache = None
        m = self.model
        if self.config.network != "fullyconnected" or len(self.params) != sum(1 for _ in m.parameters()):
            return None
        offset, off = {}, 0
        for p in self.params:
            offset[id(p)] = off
            off += p.numel()
        d = FcTrainDesc()
        d.obs_dim, d.encoding_size, d.n_actions, d.support_size = m.obs_dim, m.encoding_size, m.action_space_size, m.support_size
        nets = (m.representation_network, m.dynamics_encoded_state_network, m.dynamics_reward_network,
                m.prediction_policy_network, m.prediction_value_network)
        for k, net in enumerate(nets):
            layers = [x for x in net.module if isinstance(x, torch.nn.Linear)]
            if not 1 <= len(layers) <= 4:
                return None
            d.n_layers[k] = len(layers)
            for l, lin in enumerate(layers):
                d.in_[k][l], d.out[k][l] = lin.in_features, lin.out_features
                d.w_off[k][l], d.b_off[k][l] = offset[id(lin.weight)], offset[id(lin.bias)]
        self._fc_desc_cache = d
        return d

    def _fc_kernel_step(self, tensors):
        """Forward + backward of one batch in ONE launch; returns None when the shape is outside the kernel's reach."""
        import os
        if os.environ.get("MZB_TRAIN_KERNEL", "1") == "0":
            return None
        d = self._fc_desc()
        obs, action, tv, tr_, tp, weight, gscale = tensors
        if d is None:
            return None
        B, K1 = int(action.shape[0]), int(action.shape[1])
        if not _lib.lib.mzb_fc_train_fits(C.byref(d), B, K1):
            return None
        cfg, S = self.config, self.config.support_size
        key = (B, K1)
        ws = getattr(self, "_fc_ws", {}).get(key)
        if ws is None:
            nbytes = int(_lib.lib.mzb_fc_train_workspace_bytes(C.byref(d), B, K1))
            ws = (torch.zeros(nbytes, dtype=torch.uint8, device=self.device), torch.empty(3, B, device=self.device),
                  torch.empty(B, K1, device=self.device), torch.empty(1, device=self.device))
            self._fc_ws = dict(getattr(self, "_fc_ws", {}))
            self._fc_ws[key] = ws
        work, losses, priorities, loss = ws
        tv_sup = models.scalar_to_support(tv, S).contiguous()
        tr_sup = models.scalar_to_support(tr_, S).contiguous()
        obs2 = obs.reshape(B, -1).contiguous()
        with torch.cuda.device(self.device):
            check(_lib.lib.mzb_fc_train_grad(C.byref(d), ptr(self.flat_param), self.flat_param.numel(), B, K1, ptr(obs2),
                                             ptr(action.contiguous()), ptr(tv_sup), ptr(tr_sup), ptr(tp.contiguous()), ptr(tv.contiguous()),
                                             ptr(weight.contiguous()) if weight is not None else None, ptr(gscale.contiguous()),
                                             float(cfg.value_loss_weight), float(cfg.PER_alpha), ptr(self.flat_grad), ptr(losses),
                                             ptr(priorities), ptr(loss), ptr(work), work.numel(), _lib.current_stream()))
        return loss[0], losses[0], losses[1], losses[2], priorities

    def _forward_backward(self, tensors):
        out = self._fc_kernel_step(tensors)
        if out is not None:
            return out
        return self._forward_backward_autograd(tensors)

    def _forward_backward_autograd(self, tensors):
        """Gradients of one batch into the flat bucket.  The unrolled graph is hundreds of small kernels (launch-bound:
        13 ms per cartpole step eagerly), so from the second batch of a given shape on it is replayed from a CUDA graph
        captured over static input buffers; the optimiser launch stays outside (its step count and learning rate
        are host scalars)."""
        sig = tuple(None if t is None else (tuple(t.shape), t.dtype) for t in tensors)
        st = self._graphs.get(sig) if self.use_cuda_graph else None
        if st is None:
            self.flat_grad.zero_()
            out = unrolled_loss(self.graph, self.config, tensors)
            out[0].backward()
            if self.use_cuda_graph:
                self._graphs[sig] = "seen"
            return out
        if st == "seen":
            static = [None if t is None else t.clone() for t in tensors]
            # capture needs a few warm-up passes on a side stream (cuDNN plans, allocator pools); they must not count
            # as training passes: batch-norm running statistics are restored afterwards
            buffers = [b.clone() for b in self.model.buffers()]
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.flat_grad.zero_()
                    unrolled_loss(self.graph, self.config, static)[0].backward()
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.flat_grad.zero_()
                out = unrolled_loss(self.graph, self.config, static)
                out[0].backward()
            for b, saved in zip(self.model.buffers(), buffers):
                b.copy_(saved)
            st = self._graphs[sig] = (g, static, out)
        g, static, out = st
        for dst, src in zip(static, tensors):
            if dst is not None:
                dst.copy_(src)
        g.replay()
        return out

    def _step(self):
        """Gradient all-reduce over the ranks + optimiser update of the flat bucket.  With more than one rank on the
        node both happen in ONE kernel per rank: the ranks' gradient buckets are peer-mapped over NVLink and summed in
        rank order inside the optimiser kernel (dist.PeerGradientBuckets, csrc/mzb_optim.cu); MZB_FUSED_ALLREDUCE=0
        selects an NCCL all-reduce followed by the plain optimiser launch."""
        import os
        world = tdist.get_world_size() if tdist.is_available() and tdist.is_initialized() else 1
        self.opt_step += 1
        n, cfg = self.flat_param.numel(), self.config
        fused = world > 1 and os.environ.get("MZB_FUSED_ALLREDUCE", "1") != "0"
        with torch.cuda.device(self.device):
            if fused:
                from .dist import PeerGradientBuckets
                if getattr(self, "_peers", None) is None:
                    self._peers = PeerGradientBuckets(n, self.device)
                    _lib.bind("mzb_adam_step_allreduce", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint32,
                                                                  C.c_void_p, C.c_void_p, C.c_int64] + [C.c_double] * 5 + [C.c_int64, C.c_void_p])
                    _lib.bind("mzb_sgd_step_allreduce", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint32,
                                                                 C.c_void_p, C.c_int64] + [C.c_double] * 3 + [C.c_int64, C.c_void_p])
                    self._seq = 0
                self._seq += 1
                pb = self._peers
                pb.bucket(self._seq).copy_(self.flat_grad)                 # this step's gradients into the shared bucket
                grads, flags = pb.pointers(self._seq)
                if cfg.optimizer == "Adam":
                    check(_lib.lib.mzb_adam_step_allreduce(ptr(self.flat_param), grads, flags, pb.rank, pb.world, self._seq,
                                                           ptr(self.state1), ptr(self.state2), n, self.lr, 0.9, 0.999, 1e-8,
                                                           float(cfg.weight_decay), self.opt_step, _lib.current_stream()))
                else:
                    check(_lib.lib.mzb_sgd_step_allreduce(ptr(self.flat_param), grads, flags, pb.rank, pb.world, self._seq,
                                                          ptr(self.state1), n, self.lr, float(cfg.momentum),
                                                          float(cfg.weight_decay), self.opt_step, _lib.current_stream()))
            else:
                if world > 1:
                    tdist.all_reduce(self.flat_grad, op=tdist.ReduceOp.SUM)
                if cfg.optimizer == "Adam":
                    check(_lib.lib.mzb_adam_step(ptr(self.flat_param), ptr(self.flat_grad), ptr(self.state1), ptr(self.state2), n,
                                                 self.lr, 0.9, 0.999, 1e-8, float(cfg.weight_decay), self.opt_step, 1.0 / world,
                                                 _lib.current_stream()))
                else:
                    check(_lib.lib.mzb_sgd_step(ptr(self.flat_param), ptr(self.flat_grad), ptr(self.state1), n, self.lr,
                                                float(cfg.momentum), float(cfg.weight_decay), self.opt_step, 1.0 / world,
                                                _lib.current_stream()))
        self.model._synced = None          # the inference kernels re-pack the weights on their next call

    # -- trainer.py:55-122 over plain get_info / set_info objects
    def continuous_update_weights(self, replay_buffer, shared_storage, max_steps=None):
        while shared_storage.get_info("num_played_games") < 1:
            time.sleep(0.1)
        done = 0
        while self.training_step < self.config.training_steps and not shared_storage.get_info("terminate"):
            index_batch, batch = replay_buffer.get_batch()
            self.update_lr()
            priorities, total_loss, value_loss, reward_loss, policy_loss = self.update_weights(batch)
            if self.config.PER:
                replay_buffer.update_priorities(priorities, index_batch)
            if self.training_step % self.config.checkpoint_interval == 0:
                shared_storage.set_info("weights", self.model.get_weights())
                shared_storage.set_info("optimizer_state", self.optimizer_state())
            for k, v in (("training_step", self.training_step), ("lr", self.lr), ("total_loss", total_loss),
                         ("value_loss", value_loss), ("reward_loss", reward_loss), ("policy_loss", policy_loss)):
                shared_storage.set_info(k, v)
            done += 1
            if max_steps is not None and done >= max_steps:
                break
            if self.config.training_delay:
                time.sleep(self.config.training_delay)
            if self.config.ratio:
                while (self.training_step / max(1, shared_storage.get_info("num_played_steps")) > self.config.ratio
                       and self.training_step < self.config.training_steps and not shared_storage.get_info("terminate")):
                    time.sleep(0.5)
