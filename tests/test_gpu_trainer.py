"""Trainer (muzero_hypermodel_b200.trainer) vs two steps of the UNMODIFIED reference Trainer.update_weights on CPU
(tests/golden/trainer.npz): losses, PER priorities, learning rate and the updated weights, for the FC family (Adam),
the residual family (Adam, SGD with momentum) and the DownSample network.  Floating point: the reference ran PyTorch on
the CPU, this runs autograd on the GPU + the flat optimiser kernel - tolerances are stated per quantity."""
import ast
import importlib

import numpy as np
import pytest
import torch

import _tables as T

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
Z = T.load("trainer")


def setup_case(ci):
    from muzero_hypermodel_b200.trainer import Trainer
    pre = f"{ci}/"
    name, over = str(Z[pre + "game"]), ast.literal_eval(str(Z[pre + "over"]))
    cfg = importlib.import_module(f"muzero_hypermodel_b200.games.{name}").MuZeroConfig()
    for k, v in over.items():
        setattr(cfg, k, v)
    w0 = {k[len(pre + "w0/"):]: torch.tensor(Z[k]) for k in Z.files if k.startswith(pre + "w0/")}
    names = ["observation", "action", "value", "reward", "policy", "weight", "gradient_scale"]
    batch = [Z[pre + "batch/" + n] if pre + "batch/" + n in Z.files else None for n in names]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    tr = Trainer({"weights": w0, "training_step": 0, "optimizer_state": None}, cfg, device=DEV)
    return tr, cfg, batch, pre


@pytest.mark.parametrize("ci", range(int(Z["n"])))
def test_two_steps_match_reference_trainer(ci):
    tr, cfg, batch, pre = setup_case(ci)
    for step in range(2):
        tr.update_lr()
        pr, total, vl, rl, pl = tr.update_weights(batch)
        ref = Z[pre + f"step{step}/losses"]
        # FC: tight.  Residual nets: batch-statistics BatchNorm over 4-12 samples followed by per-channel min-max
        # rescaling amplifies the CPU (oneDNN) vs GPU (cuDNN) float32 convolution differences to ~1e-3 of a loss.
        rtol = 2e-4 if cfg.network == "fullyconnected" else 3e-3
        np.testing.assert_allclose([total, vl, rl, pl], ref[:4], rtol=rtol, atol=1e-4, err_msg=f"losses step {step}")
        assert abs(tr.lr - ref[4]) <= 1e-12 * max(1.0, abs(ref[4]))
        # priorities = |support_to_scalar(value) - target| ** alpha: compared as the value error itself (alpha = 0.5
        # would amplify small differences near zero).  The decode is ill-conditioned and, for the
        # residual nets, CPU-vs-GPU differences of the batch-statistics BatchNorm grow along the unrolled steps.
        alpha = cfg.PER_alpha
        d = np.abs(pr.cpu().numpy().astype(np.float64) ** (1 / alpha) - Z[pre + f"step{step}/priorities"].astype(np.float64) ** (1 / alpha))
        if cfg.network == "fullyconnected":
            assert np.quantile(d, 0.98) <= 2e-2 and d.max() <= 0.15, (float(np.quantile(d, 0.98)), float(d.max()))
        elif step == 0:     # chaotic along the unroll (see above): only the initial inference's column; the
            assert d[:, 0].max() <= 5e-2, float(d[:, 0].max())   # same-arithmetic comparison is tests/test_trainer_graph.py
    assert tr.training_step == 2
    got = tr.model.state_dict()
    checked = 0
    for k in Z.files:
        if not k.startswith(pre + "w2/") or k.endswith("num_batches_tracked"):
            continue
        a, b = got[k[len(pre + "w2/"):]].detach().cpu().numpy().astype(np.float64), Z[k].astype(np.float64)
        err = np.abs(a - b)
        if cfg.network == "fullyconnected":
            # Adam's first steps move every weight by ~lr * sign(gradient): an element whose gradient is at
            # rounding-noise level can land on the other side (2 * lr away) - at most 0.5 % of a tensor's elements
            assert (err <= 2e-4 + 2e-3 * np.abs(b)).mean() >= 0.995, (k, float(err.max()))
        # residual nets: the CPU-vs-GPU gradient differences are not small relative to the early layers' gradients
        # (batch statistics of 4-12 samples, see above), so only the size of the optimiser steps bounds the distance;
        # the same-arithmetic comparison of the objective and its gradients is tests/test_trainer_graph.py
        bound = 2.5 * 2 * cfg.lr_init + 1e-3 if cfg.optimizer == "Adam" else 0.1
        assert err.max() <= bound, (k, float(err.max()))
        checked += 1
    assert checked > 8


def test_optimizer_state_round_trip_and_kernel_weight_sync():
    """optimizer_state() is torch.optim's state_dict layout (what shared_storage / model.checkpoint hold), a trainer
    resumed from it takes the same next step, and the inference kernels see the updated weights."""
    from muzero_hypermodel_b200.trainer import Trainer
    tr, cfg, batch, pre = setup_case(0)
    obs = torch.tensor(batch[0], device=DEV)
    v_before = tr.model.initial_inference(obs)[0].clone()
    tr.update_lr(); tr.update_weights(batch)
    v_after = tr.model.initial_inference(obs)[0]
    assert not torch.equal(v_before, v_after)                 # kernels re-packed the trained weights
    st, w = tr.optimizer_state(), tr.model.get_weights()
    assert set(st) == {"state", "param_groups"} and "exp_avg_sq" in st["state"][0]
    tr2 = Trainer({"weights": w, "training_step": tr.training_step, "optimizer_state": st}, cfg, device=DEV)
    tr.update_lr(); tr2.update_lr()
    a = tr.update_weights(batch)
    b = tr2.update_weights(batch)
    assert a[1:] == b[1:]
    for (k, x), (_, y) in zip(tr.model.state_dict().items(), tr2.model.state_dict().items()):
        assert torch.equal(x, y), k


@pytest.mark.parametrize("kind", ["Adam", "SGD"])
def test_flat_optimiser_kernel_equals_torch_optim(kind):
    """csrc/mzb_optim.cu on one flat bucket vs torch.optim.{Adam,SGD} (trainer.py:35-52 settings) fed the same
    gradients for five steps with a changing learning rate."""
    import ctypes as C
    from muzero_hypermodel_b200 import _lib, trainer  # noqa: F401  (binds the entry points)
    torch.manual_seed(3)
    n, wd, mom = 70001, 1e-4, 0.9
    p_ref = torch.nn.Parameter(torch.randn(n, device=DEV))
    p = p_ref.detach().clone()
    s1, s2 = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    opt = torch.optim.Adam([p_ref], lr=0.02, weight_decay=wd) if kind == "Adam" else \
        torch.optim.SGD([p_ref], lr=0.02, momentum=mom, weight_decay=wd)
    for step in range(1, 6):
        g = torch.randn(n, device=DEV) * (10.0 ** torch.randint(-6, 2, (n,), device=DEV).float())
        lr = 0.02 * 0.9 ** step
        for grp in opt.param_groups:
            grp["lr"] = lr
        p_ref.grad = g.clone()
        opt.step()
        if kind == "Adam":
            _lib.check(_lib.lib.mzb_adam_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(s1), _lib.ptr(s2), n, lr, 0.9, 0.999, 1e-8, wd,
                                              step, 1.0, _lib.current_stream()))
        else:
            _lib.check(_lib.lib.mzb_sgd_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(s1), n, lr, mom, wd, step, 1.0, _lib.current_stream()))
        torch.testing.assert_close(p, p_ref.detach(), rtol=1e-5, atol=1e-6)


def test_self_play_replay_train_loop():
    """The whole loop on one GPU with plain-object shared storage: batched self-play feeds the device replay store, the
    trainer draws batches from it, writes priorities back and publishes weights, self-play picks them up."""
    from muzero_hypermodel_b200.games.cartpole import MuZeroConfig
    from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
    from muzero_hypermodel_b200.self_play import SelfPlay
    from muzero_hypermodel_b200.trainer import Trainer
    cfg = MuZeroConfig()
    cfg.num_simulations, cfg.max_moves, cfg.batch_size, cfg.replay_buffer_size = 10, 40, 64, 512
    cfg.checkpoint_interval, cfg.self_play_delay, cfg.training_delay, cfg.ratio = 2, 0, 0, None

    class Storage:
        def __init__(self):
            self.info = {"training_step": 0, "terminate": False, "weights": None, "optimizer_state": None,
                         "num_played_games": 0, "num_played_steps": 0}

        def get_info(self, key):
            return self.info[key]

        def set_info(self, key, value=None):
            self.info[key] = value

    st = Storage()
    sp = SelfPlay({"weights": None}, None, cfg, 1, n_games=128, device=DEV)
    env, _ = sp._setup()
    rb = ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg, device=DEV, record_env=env)
    tr = Trainer({"weights": sp.model.get_weights(), "training_step": 0, "optimizer_state": None}, cfg, device=DEV)
    w_before = {k: v.clone() for k, v in sp.model.get_weights().items()}
    sp.continuous_self_play(st, rb, max_moves=45)
    assert st.info["num_played_games"] >= 128 and len(rb) >= 128
    tr.continuous_update_weights(rb, st, max_steps=6)
    assert tr.training_step == 6 == st.info["training_step"] and np.isfinite(st.info["total_loss"])
    assert st.info["weights"] is not None and st.info["optimizer_state"]["state"]
    sp2 = SelfPlay({"weights": None}, None, cfg, 1, n_games=128, device=DEV)
    sp2.continuous_self_play(st, rb, max_moves=2)           # refreshes the weights from the storage (self_play.py:37)
    assert any(not torch.equal(w_before[k], v) for k, v in sp2.model.get_weights().items())


def test_optimizer_state_indices_follow_torch_optim_on_downsample_net():
    """Breakout's DownSample representation never reaches conv/bn (models.py:338-345): torch.optim still NUMBERS those
    parameters (they just have no state).  optimizer_state() must use the same numbering so a reference checkpoint's
    optimizer_state loads into the right moments."""
    from muzero_hypermodel_b200.games.breakout import MuZeroConfig
    from muzero_hypermodel_b200.trainer import Trainer
    cfg = MuZeroConfig()
    cfg.batch_size, cfg.num_unroll_steps = 2, 2
    torch.manual_seed(0)
    tr = Trainer({"weights": None, "training_step": 0, "optimizer_state": None}, cfg, device=DEV)
    B, K, A = 2, cfg.num_unroll_steps, len(cfg.action_space)
    rs = np.random.RandomState(0)
    batch = [rs.rand(B, 3, 96, 96).astype(np.float32), rs.randint(A, size=(B, K + 1)), rs.randn(B, K + 1), rs.randn(B, K + 1),
             np.full((B, K + 1, A), 1.0 / A), np.ones(B, np.float32), np.ones((B, K + 1), np.float32)]
    tr.use_cuda_graph = False
    tr.update_lr(); tr.update_weights(batch)
    ours = tr.optimizer_state()
    # torch's own numbering: Adam over ALL model.parameters(), stepping with the same gradients present / absent
    allp = list(tr.model.parameters())
    opt = torch.optim.Adam(allp, lr=cfg.lr_init, weight_decay=cfg.weight_decay)
    used = {id(p) for p in tr.params}
    for p in allp:
        p.grad = torch.zeros_like(p) if id(p) in used else None
    opt.step()
    ref = opt.state_dict()
    assert ours["param_groups"][0]["params"] == ref["param_groups"][0]["params"]
    assert sorted(ours["state"]) == sorted(ref["state"]) and len(ref["state"]) < len(allp)
    for i in ref["state"]:
        assert tuple(ours["state"][i]["exp_avg"].shape) == tuple(ref["state"][i]["exp_avg"].shape), i
    # and a state in torch's numbering loads back into the same flat offsets
    tr2 = Trainer({"weights": tr.model.get_weights(), "training_step": 1, "optimizer_state": ours}, cfg, device=DEV)
    assert torch.equal(tr2.state1, tr.state1) and torch.equal(tr2.state2, tr.state2) and tr2.opt_step == tr.opt_step


def test_update_weights_returns_fresh_priorities():
    """The returned priorities must not alias the CUDA graph's static output (the reference returns a new array)."""
    tr, cfg, batch, pre = setup_case(0)
    outs = []
    for _ in range(3):
        tr.update_lr()
        outs.append(tr.update_weights(batch)[0])
    assert outs[1].data_ptr() != outs[2].data_ptr()
    assert not torch.equal(outs[1], outs[2])          # weights moved between the steps; an alias would compare equal


@pytest.mark.parametrize("game,B,per", [("cartpole", 128, True), ("cartpole", 7, False), ("tictactoe", 33, True)])
def test_one_kernel_fc_training_step_equals_autograd(game, B, per):
    """The one-kernel forward / backward of the fully-connected family (csrc/mzb_fc_train.cu) against PyTorch autograd
    over the same parameters (the path pinned to the reference by tests/golden/trainer.npz and tests/test_trainer_graph.py):
    every gradient, the per-sample losses, the batch objective and the priorities; twice, to show it is deterministic."""
    from muzero_hypermodel_b200.trainer import Trainer
    cfg = importlib.import_module(f"muzero_hypermodel_b200.games.{game}").MuZeroConfig()
    cfg.network = "fullyconnected"
    cfg.PER = per                                          # without PER the batch carries no importance weights
    torch.backends.cuda.matmul.allow_tf32 = False
    tr = Trainer({"weights": None, "training_step": 0, "optimizer_state": None}, cfg, device=DEV)
    rs = np.random.RandomState(5)
    K1, A = cfg.num_unroll_steps + 1, len(cfg.action_space)
    obs = torch.tensor(rs.uniform(-1, 1, (B,) + tuple(cfg.observation_shape)).astype(np.float32), device=DEV)
    action = torch.tensor(rs.randint(0, A, (B, K1)), device=DEV)
    tv = torch.tensor(rs.uniform(-3, 30, (B, K1)).astype(np.float32), device=DEV)
    trw = torch.tensor(rs.uniform(-1, 1, (B, K1)).astype(np.float32), device=DEV)
    tp = torch.tensor(rs.dirichlet([0.5] * A, (B, K1)).astype(np.float32), device=DEV)
    tp[:, -1] = 0.0                                        # an absorbing tail with an all-zero policy target
    w = torch.tensor(rs.uniform(0.2, 1.0, B).astype(np.float32), device=DEV)
    gs = torch.tensor(np.repeat(rs.randint(1, K1 + 1, (B, 1)), K1, 1).astype(np.float32), device=DEV)
    tensors = (obs, action, tv, trw, tp, w if cfg.PER else None, gs)
    assert tr._fc_desc() is not None
    out = tr._fc_kernel_step(tensors)
    assert out is not None, "the one-kernel step must take this shape"
    g1 = tr.flat_grad.clone()
    k = [x.clone() for x in out]
    out2 = tr._fc_kernel_step(tensors)
    assert torch.equal(tr.flat_grad, g1) and all(torch.equal(a, b) for a, b in zip(out2, k)), "not deterministic"
    tr.use_cuda_graph = False
    ref = tr._forward_backward_autograd(tensors)
    g0 = tr.flat_grad.clone()
    scale = float(g0.abs().max())
    assert scale > 0
    err = float((g1 - g0).abs().max())
    assert err <= 2e-5 * scale + 1e-7, (err, scale)
    np.testing.assert_allclose(float(k[0]), float(ref[0]), rtol=2e-5)
    for a, b in zip(k[1:4], ref[1:4]):
        np.testing.assert_allclose(a.cpu().numpy(), b.detach().cpu().numpy(), rtol=2e-5, atol=1e-5)
    # priorities through the ill-conditioned decode: compare the value errors, not their roots
    al = cfg.PER_alpha
    d = np.abs(k[4].cpu().numpy().astype(np.float64) ** (1 / al) - ref[4].detach().cpu().numpy().astype(np.float64) ** (1 / al))
    assert d.max() <= 5e-3, float(d.max())


def test_one_kernel_fc_training_step_argument_checks():
    """mzb_fc_train_grad refuses what it cannot run instead of computing garbage: a layer table that does not cover the
    parameter bucket, a workspace that is too small, an unroll too long for the kernel's shared memory (the trainer then
    falls back to autograd)."""
    import ctypes as C
    from muzero_hypermodel_b200 import _lib
    from muzero_hypermodel_b200.trainer import Trainer
    cfg = importlib.import_module("muzero_hypermodel_b200.games.cartpole").MuZeroConfig()
    tr = Trainer({"weights": None, "training_step": 0, "optimizer_state": None}, cfg, device=DEV)
    d = tr._fc_desc()
    B, K1 = 8, cfg.num_unroll_steps + 1
    assert _lib.lib.mzb_fc_train_fits(C.byref(d), B, K1) == 1
    assert _lib.lib.mzb_fc_train_fits(C.byref(d), B, 4000) == 0                     # activations of 4,000 steps do not fit
    need = int(_lib.lib.mzb_fc_train_workspace_bytes(C.byref(d), B, K1))
    assert need > 0
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=DEV)
    full, A = 2 * cfg.support_size + 1, len(cfg.action_space)
    # the tensors stay referenced for the whole test: a temporary's block would be handed to the next allocation
    t = dict(obs=z(B, 4), action=z(B, K1, dt=torch.int64), tv=z(B, K1, full), tr=z(B, K1, full), tp=z(B, K1, A), tvs=z(B, K1),
             gs=torch.ones(B, K1, device=DEV), losses=z(3, B), prio=z(B, K1), loss=z(1))
    args = lambda n_params, ws: (C.byref(d), _lib.ptr(tr.flat_param), n_params, B, K1, _lib.ptr(t["obs"]), _lib.ptr(t["action"]),
                                 _lib.ptr(t["tv"]), _lib.ptr(t["tr"]), _lib.ptr(t["tp"]), _lib.ptr(t["tvs"]), None,
                                 _lib.ptr(t["gs"]), 0.25, 0.5, _lib.ptr(tr.flat_grad), _lib.ptr(t["losses"]),
                                 _lib.ptr(t["prio"]), _lib.ptr(t["loss"]), _lib.ptr(ws), ws.numel(), _lib.current_stream())
    ws = torch.zeros(need, dtype=torch.uint8, device=DEV)
    assert _lib.lib.mzb_fc_train_grad(*args(tr.flat_param.numel(), ws)) == 0
    torch.cuda.synchronize()
    assert _lib.lib.mzb_fc_train_grad(*args(tr.flat_param.numel() - 1, ws)) != 0      # table / bucket mismatch
    assert b"parameters" in _lib.lib.mzb_last_error()
    small = torch.zeros(16, dtype=torch.uint8, device=DEV)
    assert _lib.lib.mzb_fc_train_grad(*args(tr.flat_param.numel(), small)) != 0       # workspace too small
    assert b"workspace" in _lib.lib.mzb_last_error()
