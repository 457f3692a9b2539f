"""The training objective (muzero_hypermodel_b200.trainer.unrolled_loss over the differentiable graph of the
parameter-holder modules) against two steps of the UNMODIFIED reference Trainer.update_weights (tests/golden/trainer.npz),
on the same arithmetic the reference used (PyTorch on the CPU): losses, priorities and - through torch.optim with the
reference's settings - the weights after two steps.  CPU-only: the graph is plain autograd; the Trainer class itself
(flat buckets, optimiser kernel, device codec) is GPU-only and tested in test_gpu_trainer.py."""
import ast
import importlib

import numpy as np
import pytest
import torch

import _tables as T
from oracle import networks as onet

Z = T.load("trainer")


def _codec(fn):
    return lambda x, s: torch.tensor(fn(x.detach().cpu().numpy(), s))


@pytest.mark.parametrize("ci", range(int(Z["n"])))
def test_objective_and_gradients_match_reference(ci):
    from muzero_hypermodel_b200 import models
    from muzero_hypermodel_b200.trainer import training_graph, unrolled_loss
    torch.set_num_threads(1)
    pre = f"{ci}/"
    name, over = str(Z[pre + "game"]), ast.literal_eval(str(Z[pre + "over"]))
    cfg = importlib.import_module(f"muzero_hypermodel_b200.games.{name}").MuZeroConfig()
    for k, v in over.items():
        setattr(cfg, k, v)
    model = models.MuZeroNetwork(cfg)
    model.set_weights({k[len(pre + "w0/"):]: torch.tensor(Z[k]) for k in Z.files if k.startswith(pre + "w0/")})
    model.train()
    graph = training_graph(model, cfg)
    if cfg.optimizer == "Adam":
        opt = torch.optim.Adam(model.parameters(), lr=cfg.lr_init, weight_decay=cfg.weight_decay)
    else:
        opt = torch.optim.SGD(model.parameters(), lr=cfg.lr_init, momentum=cfg.momentum, weight_decay=cfg.weight_decay)
    names = ["observation", "action", "value", "reward", "policy", "weight", "gradient_scale"]
    b = {n: (torch.tensor(Z[pre + "batch/" + n]) if pre + "batch/" + n in Z.files else None) for n in names}
    tensors = (b["observation"], b["action"].long(), b["value"], b["reward"], b["policy"], b["weight"], b["gradient_scale"])
    for step in range(2):
        lr = cfg.lr_init * cfg.lr_decay_rate ** (step / cfg.lr_decay_steps)
        for g in opt.param_groups:
            g["lr"] = lr
        loss, vl, rl, pl, pr = unrolled_loss(graph, cfg, tensors, _codec(onet.scalar_to_support), _codec(onet.support_to_scalar))
        ref = Z[pre + f"step{step}/losses"]
        # step 0 is a pure forward pass; step 1 follows one optimiser step, where Adam turns rounding-level gradient
        # differences (thread count / summation order of the convolutions) into +-lr weight differences
        np.testing.assert_allclose([loss.item(), vl.mean().item(), rl.mean().item(), pl.mean().item()], ref[:4],
                                   rtol=2e-5 if step == 0 else 1e-3, atol=1e-5)
        assert lr == ref[4]
        alpha = cfg.PER_alpha
        d = np.abs(pr.numpy().astype(np.float64) ** (1 / alpha) - Z[pre + f"step{step}/priorities"].astype(np.float64) ** (1 / alpha))
        if step == 0:
            assert d.max() <= 2e-3, float(d.max())
        else:       # 20 unrolled steps of batch-norm + min-max rescaling amplify the post-Adam weight differences
            assert np.quantile(d, 0.95) <= 5e-2 and d.max() <= 0.5, (float(np.quantile(d, 0.95)), float(d.max()))
        opt.zero_grad()
        loss.backward()
        opt.step()
    got = model.state_dict()
    worst = 0.0
    for k in Z.files:
        if k.startswith(pre + "w2/") and not k.endswith("num_batches_tracked"):
            a, w = got[k[len(pre + "w2/"):]].detach().numpy(), Z[k]
            err = np.abs(a.astype(np.float64) - w)
            assert err.mean() <= 5e-4 + 5e-3 * np.abs(w).mean(), (k, float(err.mean()), float(err.max()))
            worst = max(worst, float(err.max()))
    assert worst <= 4 * cfg.lr_init + 1e-3
