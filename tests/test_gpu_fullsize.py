"""Size-independent properties at BASELINE.json's full batch sizes (where the oracle cannot follow): every search
spends exactly num_simulations visits on legal root actions, root values stay inside the value support, a batch made
of one position replicated gives one answer (no cross-game leakage at scale), and a whole self-play move keeps the
integer boards consistent with the recorded actions."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _selfplay(workload, G, sims=None):
    from muzero_hypermodel_b200.self_play import SelfPlay
    cfg = bench.make_config(workload)
    if sims:
        cfg.num_simulations = sims
    w = {k: torch.tensor(v) for k, v in bench.load_weights(bench.WORKLOADS[workload][0]).items()}
    return SelfPlay({"weights": w}, None, cfg, cfg.seed, n_games=G, device=DEV), cfg


@pytest.mark.parametrize("workload,G,sims", [("cartpole", 262144, None), ("tictactoe", 65536, None),
                                              ("connect4", 16384, None), ("gomoku", 1024, 100), ("breakout", 16384, None)])
def test_full_batch_search_invariants(workload, G, sims):
    sp, cfg = _selfplay(workload, G, sims)
    env, mcts = sp._setup()
    S, A = cfg.num_simulations, len(cfg.action_space)
    for move in range(3):                           # the first call launches kernel by kernel, later ones replay the graph
        obs, legal, to_play = env.observe()
        out = mcts.run(sp.model, obs, legal, to_play, True, slot=env.slot, step=env.step_count)
        visits, rv = out["visits"], out["root_value"]
        assert visits.shape == (G, A) and int(visits.min()) >= 0
        assert torch.all(visits.sum(1) == S), "every search spends exactly num_simulations visits"
        assert int((visits * (legal == 0)).sum()) == 0, "no visit on an illegal root action"
        assert torch.isfinite(rv).all()
        bound = 400.0                               # |value| <= inverse transform of the support edge (S = 10 -> ~121) x rewards
        assert float(rv.abs().max()) < bound
        assert int(out["max_depth"].min()) >= 1 and int(out["max_depth"].max()) <= S
        env.act_step(visits, rv, legal, 1.0, None)
        env.harvest(True)
    c = env.counters()
    assert c["env_steps"] == 3 * G and c["dropped_games"] == 0


@pytest.mark.parametrize("workload,G", [("cartpole", 65536), ("connect4", 4096)])
def test_replicated_position_gives_one_answer(workload, G):
    """All G games share one observation, slot and step: every row of the result must be identical."""
    sp, cfg = _selfplay(workload, G)
    env, mcts = sp._setup()
    obs, legal, to_play = env.observe()
    obs1 = obs[:1].expand(G, -1).contiguous()
    legal1 = legal[:1].expand(G, -1).contiguous()
    tp1 = to_play[:1].expand(G).contiguous()
    slot = torch.full((G,), 7, dtype=torch.int32, device=DEV)
    step = torch.full((G,), 3, dtype=torch.int32, device=DEV)
    out = mcts.run(sp.model, obs1, legal1, tp1, True, slot=slot, step=step)
    v = out["visits"]
    assert torch.all(v == v[0:1]), "visit counts differ between replicas of one position"
    assert out["root_value"].cpu().numpy().tobytes() == out["root_value"][0:1].expand(G).contiguous().cpu().numpy().tobytes()


def test_boards_follow_recorded_actions_at_scale():
    """connect4, 16,384 games, 6 moves: replaying every game's recorded actions on the oracle's integer board rules
    reproduces the device boards and to-play flags exactly (checked on a sample of games)."""
    from oracle import games as ogames
    sp, cfg = _selfplay("connect4", 16384, 8)
    env, mcts = sp._setup()
    acts = []
    for move in range(6):
        obs, legal, to_play = env.observe()
        out = mcts.run(sp.model, obs, legal, to_play, True, slot=env.slot, step=env.step_count)
        a, _, _ = env.act_step(out["visits"], out["root_value"], legal, 1.0, None, want_outputs=True)
        acts.append(a.cpu().numpy())
        env.harvest(True)
    obs, legal, to_play = env.observe()
    planes = obs.cpu().numpy().reshape(-1, 3, 6, 7)
    sample = np.random.RandomState(0).choice(16384, 256, replace=False)
    for g in sample:
        o = ogames.Connect4(1)
        for move in range(6):
            _, _, done = o.step(np.array([acts[move][g]]))
            assert not done[0]                      # a connect4 game cannot end within 6 plies
        np.testing.assert_array_equal(o.observation()[0], planes[g])
        assert int(o.to_play()[0]) == int(to_play[g])
