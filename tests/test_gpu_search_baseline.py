"""Search-level parity of the residual family AT THE BASELINE CONFIGS:

* `mzb_search_resnet` for gomoku (A = 121, 400 simulations), breakout (DownSample stem + 30 simulations) and connect4 at
  the full 200 simulations, against `oracle.mcts.search` (pinned to the reference's MCTS.run by tests/golden/tree.npz) fed
  the SAME kernels' network outputs one row at a time: fp32 path bit-exact; bf16 tcgen05 path within a stated bound;
* the bf16 path against the fp32 path over the same roots, noise and tie-breaks: measured visit-count agreement and
  root-value error per config, asserted against floors and written to gpurun_out/r2_parity_search.json (bench.py
  prints the same metrics as `parity`)."""
import json
import os

import numpy as np
import pytest
import torch

import _tables as T
from _configs import product_config
from _weights import seeded_state_dict
from oracle import mcts as omcts
from oracle import rng

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _net(tag, precision):
    from muzero_hypermodel_b200 import models
    cfg = product_config(tag)
    z = T.load("net")
    if tag == "gomoku":
        keys = str(z["gomoku/keys"]).split("\n")
        shapes = [[int(d) for d in s.split("x")] if s else [] for s in z["gomoku/shapes"]]
        sd = {k: torch.tensor(v) for k, v in seeded_state_dict(keys, shapes).items()}
    else:
        pre = tag + "/w/"
        sd = {k[len(pre):]: torch.tensor(z[k]) for k in z.files if k.startswith(pre)}
    net = models.MuZeroNetwork(cfg)
    net.set_weights(sd)
    net.set_precision(precision)
    return net.to(DEV).eval(), cfg


def _inputs(cfg, G, seed):
    rs = np.random.RandomState(seed)
    C, H, W = cfg.observation_shape
    A = len(cfg.action_space)
    if len(cfg.players) == 2:                      # a random mid-game position; legal = empty cells / open columns
        n_stones = rs.randint(0, H * W // 2, size=G)
        stones = np.zeros((G, H * W), dtype=np.int64)
        for g in range(G):
            idx = rs.permutation(H * W)[:n_stones[g]]
            stones[g, idx] = rs.choice([-1, 1], size=n_stones[g])
        stones = stones.reshape(G, H, W)
        tp = rs.choice([-1, 1], size=(G, 1, 1))
        obs = np.stack([(stones == 1), (stones == -1), np.broadcast_to(tp, stones.shape)], axis=1).astype(np.float32)
        if A == H * W:
            legal = stones.reshape(G, -1) == 0
        else:                                      # connect4: a column is open while its top cell is empty
            legal = stones[:, H - 1, :] == 0
        legal[np.arange(G), rs.randint(A, size=G)] |= ~legal.any(1)
        to_play = (tp.reshape(G) == -1).astype(np.int8)
    else:                                          # synthetic frames (BASELINE configs[4])
        obs = rs.uniform(0, 1, size=(G, C, H, W)).astype(np.float32)
        legal = np.ones((G, A), dtype=bool)
        to_play = np.zeros(G, dtype=np.int8)
    noise = np.zeros((G, A))
    for g in range(G):
        noise[g, legal[g]] = rs.dirichlet([cfg.root_dirichlet_alpha] * int(legal[g].sum()))
    slot = rs.randint(1 << 20, size=G).astype(np.int32)
    step = rs.randint(100, size=G).astype(np.int32)
    return obs, legal, to_play, noise, slot, step


def _search(net, cfg, G, inp):
    from muzero_hypermodel_b200.search import BatchedMCTS
    obs, legal, to_play, noise, slot, step = inp
    eng = BatchedMCTS(cfg, G, device=DEV, seed=T.SEED)
    out = eng.run(net, torch.tensor(obs, device=DEV), torch.tensor(legal, device=DEV), torch.tensor(to_play, device=DEV),
                  True, noise=torch.tensor(noise, device=DEV), slot=torch.tensor(slot, device=DEV),
                  step=torch.tensor(step, device=DEV))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def _oracle(net, cfg, inp, g):
    obs, legal, to_play, noise, slot, step = inp
    A = len(cfg.action_space)
    la = np.nonzero(legal[g])[0].tolist()
    o = net.initial_inference_fused(torch.tensor(obs[g:g + 1], device=DEV), legal=torch.tensor(legal[g:g + 1], device=DEV))
    root = (float(o["value"][0]), float(o["reward"][0]), [float(o["priors"][0, a]) for a in la], o["state"])

    def rec(hidden, action):
        r = net.recurrent_inference_fused(hidden, torch.tensor([[action]], device=DEV))
        return float(r["value"][0]), float(r["reward"][0]), [float(x) for x in r["priors"][0]], r["state"]

    res = omcts.search(rec, root, la, int(to_play[g]), n_actions=A, n_players=len(cfg.players),
                       num_simulations=cfg.num_simulations, discount=cfg.discount, pb_c_base=cfg.pb_c_base,
                       pb_c_init=cfg.pb_c_init, noise=[float(noise[g, a]) for a in la],
                       exploration_fraction=cfg.root_exploration_fraction,
                       tie=lambda n, sim, depth: rng.tie_index(T.SEED, int(slot[g]), int(step[g]), sim, depth, n))
    return la, res


def _dump(key, value):
    path = os.path.join(ROOT, "gpurun_out", "r2_parity_search.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    try:
        d = json.load(open(path))
    except (OSError, ValueError):
        d = {}
    d[key] = value
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)


# (tag, games searched, games checked against the oracle): BASELINE simulation counts (200 / 400 / 30)
CASES = [("connect4", 64, 64), ("gomoku", 8, 3), ("breakout", 24, 6)]


@pytest.mark.parametrize("tag,G,n_check", CASES)
def test_fp32_search_equals_oracle_at_baseline_config(tag, G, n_check):
    """fp32 kernels: the device search must rebuild the oracle's tree bit-for-bit - visit counts, float64 root value,
    maximum depth - at the configuration's own simulation count and action-space width."""
    net, cfg = _net(tag, "fp32")
    assert cfg.num_simulations == {"connect4": 200, "gomoku": 400, "breakout": 30}[tag]
    inp = _inputs(cfg, G, 7)
    out = _search(net, cfg, G, inp)
    assert (out["visits"].sum(1) == cfg.num_simulations).all() and (out["visits"][~inp[1]] == 0).all()
    for g in range(n_check):
        la, res = _oracle(net, cfg, inp, g)
        np.testing.assert_array_equal(out["visits"][g][la], res.visits, err_msg=f"game {g}")
        assert np.float64(out["root_value"][g]).tobytes() == np.float64(res.root_value()).tobytes(), g
        assert out["max_depth"][g] == res.max_tree_depth, g


@pytest.mark.parametrize("tag,G,n_check", CASES)
def test_bf16_search_vs_oracle_at_baseline_config(tag, G, n_check):
    """bf16 tcgen05 kernels (the path every resnet bench number is quoted on) against the oracle fed the same kernels'
    row-at-a-time outputs.  The convolution's accumulation order per output row does not depend on the batch, so the
    batched search must rebuild the oracle's tree EXACTLY for the C >= 64 towers (connect4: 64 games x 200 simulations,
    gomoku: 400 simulations over 121 actions - measured 64/64 and 3/3 identical, root values bit-equal).  Breakout's
    16-channel heads run a different kernel per batch shape (warp-per-image vs the fused projection): visit counts must
    still coincide, root values within 1e-2 (measured 2.6e-3 mean)."""
    from muzero_hypermodel_b200.parity import visit_agreement
    net, cfg = _net(tag, "bf16")
    inp = _inputs(cfg, G, 7)
    out = _search(net, cfg, G, inp)
    A = len(cfg.action_space)
    ov, orv, dv, drv, depth_same = [], [], [], [], 0
    for g in range(n_check):
        la, res = _oracle(net, cfg, inp, g)
        v = np.zeros(A, dtype=np.int64)
        v[la] = res.visits
        ov.append(v); orv.append(res.root_value()); dv.append(out["visits"][g]); drv.append(out["root_value"][g])
        depth_same += int(out["max_depth"][g] == res.max_tree_depth)
    m = visit_agreement(np.array(dv), np.array(ov), np.array(drv), np.array(orv))
    m["same_max_depth"] = depth_same / n_check
    _dump(f"bf16_vs_oracle/{tag}", m)
    assert m["identical_visit_counts"] == 1.0 and m["same_max_depth"] == 1.0, m
    if tag == "breakout":
        assert m["root_value_max_err"] <= 1e-2, m
    else:
        assert m["root_value_max_err"] == 0.0, m


# floors = what the bf16 convolutions may cost against the fp32 path at the BASELINE simulation counts, set from the
# measured figures (gpurun_out/r2_parity_search.json -> DESIGN.md §9): (visit agreement >=, root value MAE <=)
FLOORS = {"tictactoe": (0.95, 0.01), "connect4": (0.96, 0.01), "breakout": (0.99, 0.01), "gomoku": (0.50, 0.35)}


@pytest.mark.parametrize("tag,G", [("tictactoe", 512), ("connect4", 256), ("gomoku", 24), ("breakout", 128)])
def test_bf16_vs_fp32_visit_agreement(tag, G):
    """The bf16 search against the fp32 search over the same roots, injected noise and tie-break counters at the
    BASELINE simulation counts: how far the training targets (child_visits, root values) move when the convolutions run
    in bf16.  Seeded random weights make this the hard case (near-uniform priors, values near zero: ties everywhere)."""
    from muzero_hypermodel_b200.parity import visit_agreement
    nf, cfg = _net(tag, "fp32")
    nb, _ = _net(tag, "bf16")
    inp = _inputs(cfg, G, 11)
    a = _search(nf, cfg, G, inp)
    b = _search(nb, cfg, G, inp)
    m = visit_agreement(b["visits"], a["visits"], b["root_value"], a["root_value"])
    m["num_simulations"] = int(cfg.num_simulations)
    m["root_predicted_value_mae"] = float(np.abs(a["root_predicted_value"] - b["root_predicted_value"]).mean())
    _dump(f"bf16_vs_fp32/{tag}", m)
    # gomoku is the stated exception: 6 blocks x 128 channels of SEEDED RANDOM weights give near-uniform priors over 121
    # actions (max prior ~ 1 %) and values near zero, so 400 simulations amplify the 3e-2 bf16 value error (sqrt(L)
    # growth over 25-51 bf16 layers, tests/test_gpu_resnet.py) into different explored moves: measured 0.66 / 0.19.
    # The bound is reported, not hidden: bench.py prints the same metrics per run as `parity`.
    floor, mae = FLOORS[tag]
    assert m["visit_agreement"] >= floor, m
    assert m["root_value_mae"] <= mae, m
