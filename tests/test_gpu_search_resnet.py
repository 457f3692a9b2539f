"""Batched search with residual networks (mzb_search_resnet) vs the oracle MCTS fed by the same kernels."""
import numpy as np
import pytest
import torch

import _tables as T
from _configs import product_config
from oracle import mcts as omcts
from oracle import rng

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net(tag, precision):
    from muzero_hypermodel_b200 import models
    cfg = product_config(tag)
    z = T.load("net")
    pre = tag + "/w/"
    net = models.MuZeroNetwork(cfg)
    net.set_weights({k[len(pre):]: torch.tensor(z[k]) for k in z.files if k.startswith(pre)})
    net.set_precision(precision)
    return net.to(DEV).eval(), cfg


@pytest.mark.parametrize("tag,precision,sims", [("tictactoe", "fp32", 25), ("tictactoe", "bf16", 25),
                                                 ("connect4", "fp32", 24), ("connect4", "bf16", 40)])
def test_resnet_search_equals_oracle_given_kernel_outputs(tag, precision, sims):
    from muzero_hypermodel_b200.search import BatchedMCTS
    net, cfg = _net(tag, precision)
    cfg.num_simulations = sims
    G = 7
    rs = np.random.RandomState(2)
    C, H, W = cfg.observation_shape
    A = len(cfg.action_space)
    stones = rs.randint(-1, 2, (G, H, W))
    tp = rs.choice([-1, 1], size=(G, 1, 1))
    obs = np.stack([(stones == 1), (stones == -1), np.broadcast_to(tp, stones.shape)], axis=1).astype(np.float32)
    legal = rs.uniform(size=(G, A)) < 0.7
    legal[np.arange(G), rs.randint(A, size=G)] = True
    to_play = rs.randint(2, size=G).astype(np.int8)
    noise = np.zeros((G, A))
    for g in range(G):
        noise[g, legal[g]] = rs.dirichlet([cfg.root_dirichlet_alpha] * int(legal[g].sum()))
    slot = rs.randint(1 << 20, size=G).astype(np.int32)
    step = rs.randint(100, size=G).astype(np.int32)
    eng = BatchedMCTS(cfg, G, device=DEV, seed=T.SEED)
    out = eng.run(net, torch.tensor(obs, device=DEV), torch.tensor(legal, device=DEV), torch.tensor(to_play, device=DEV),
                  True, noise=torch.tensor(noise, device=DEV), slot=torch.tensor(slot, device=DEV),
                  step=torch.tensor(step, device=DEV))
    out = {k: v.cpu().numpy() for k, v in out.items()}
    assert (out["visits"].sum(1) == sims).all() and (out["visits"][~legal] == 0).all()
    for g in range(3):
        la = np.nonzero(legal[g])[0].tolist()
        o = net.initial_inference_fused(torch.tensor(obs[g:g + 1], device=DEV), legal=torch.tensor(legal[g:g + 1], device=DEV))
        root = (float(o["value"][0]), float(o["reward"][0]), [float(o["priors"][0, a]) for a in la], o["state"])

        def rec(hidden, action):
            r = net.recurrent_inference_fused(hidden, torch.tensor([[action]], device=DEV))
            return float(r["value"][0]), float(r["reward"][0]), [float(x) for x in r["priors"][0]], r["state"]

        res = omcts.search(rec, root, la, int(to_play[g]), n_actions=A, n_players=2, num_simulations=sims,
                           discount=cfg.discount, pb_c_base=cfg.pb_c_base, pb_c_init=cfg.pb_c_init,
                           noise=[float(noise[g, a]) for a in la], exploration_fraction=cfg.root_exploration_fraction,
                           tie=lambda n, sim, depth: rng.tie_index(T.SEED, int(slot[g]), int(step[g]), sim, depth, n))
        if precision == "fp32":
            np.testing.assert_array_equal(out["visits"][g][la], res.visits)
            assert np.float64(out["root_value"][g]).tobytes() == np.float64(res.root_value()).tobytes()
            assert out["max_depth"][g] == res.max_tree_depth
        else:
            # the oracle run passes hidden states through fp32 NCHW tensors (exact for bf16 values), so the trees
            # coincide unless a bf16 rounding tie flips; require near-identical statistics
            assert np.abs(out["visits"][g][la] - np.array(res.visits)).sum() <= max(2, sims // 10)
            assert abs(out["root_value"][g] - res.root_value()) < 0.05 * max(1.0, abs(res.root_value()))


def test_mcts_run_dropin_materialises_node_graph():
    """self_play.MCTS(config).run(...) with G = 1: Node graph read API of the reference (resnet + fc)."""
    from muzero_hypermodel_b200 import self_play
    net, cfg = _net("tictactoe", "fp32")
    obs = np.zeros((3, 3, 3), dtype=np.int32)
    obs[2] = 1
    root, info = self_play.MCTS(cfg).run(net, obs, [0, 1, 2, 3, 4, 5, 6, 7, 8], 0, True)
    assert root.visit_count == cfg.num_simulations and sum(c.visit_count for c in root.children.values()) == cfg.num_simulations
    assert list(root.children.keys()) == list(range(9)) and root.expanded() and root.to_play == 0
    assert isinstance(info["max_tree_depth"], int) and isinstance(info["root_predicted_value"], float)
    assert tuple(root.hidden_state.shape) == (1, 16, 3, 3)
    best = max(root.children.values(), key=lambda c: c.visit_count)
    assert best.expanded() and best.to_play == 1 and abs(best.value()) < 1e3
    a = self_play.SelfPlay.select_action(root, 0)
    assert root.children[a].visit_count == max(c.visit_count for c in root.children.values())
    gh = self_play.GameHistory()
    gh.store_search_statistics(root, cfg.action_space)
    assert abs(sum(gh.child_visits[0]) - 1) < 1e-12 and gh.root_values[0] == root.value()
    with pytest.raises(AssertionError):
        self_play.MCTS(cfg).run(net, obs, [], 0, True)
    # override_root_with (diagnose_model.py:57-69): caller-expanded root, root_predicted_value None
    v, r, pl, hs = net.recurrent_inference(root.hidden_state, torch.tensor([[a]], device=DEV))
    new_root = self_play.Node(0)
    new_root.expand(cfg.action_space, 1, 0.0, pl, hs)
    root2, info2 = self_play.MCTS(cfg).run(net, None, cfg.action_space, 1, True, new_root)
    assert info2["root_predicted_value"] is None and root2.visit_count == cfg.num_simulations


def test_reference_style_calls_draw_fresh_randomness():
    """MCTS.run without RNG counters and repeated play_game calls must not replay the same noise / tie-breaks /
    action samples (the reference draws from numpy's advancing global stream)."""
    from muzero_hypermodel_b200 import models, self_play
    from muzero_hypermodel_b200.games.cartpole import MuZeroConfig, Game
    cfg = MuZeroConfig()
    cfg.max_moves = 40
    torch.manual_seed(1)
    net = models.MuZeroNetwork(cfg).to(DEV).eval()
    obs = np.zeros((1, 1, 4), dtype=np.float32)
    priors = []
    for _ in range(3):
        root, _ = self_play.MCTS(cfg).run(net, obs, [0, 1], 0, True)
        priors.append(tuple(root.children[a].prior for a in (0, 1)))
    assert len(set(priors)) == 3, priors                  # three different Dirichlet samples mixed into the priors
    sp = self_play.SelfPlay({"weights": net.get_weights()}, Game, cfg, 0, device=DEV)
    games = [sp.play_game(1.0, None, False, "self", 0) for _ in range(3)]
    assert len({(tuple(g.action_history), tuple(np.round(g.root_values, 12))) for g in games}) == 3
