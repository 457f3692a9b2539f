"""G7 on the device: whole games of the UNMODIFIED reference SelfPlay.play_game (tests/golden/episode.npz) replayed by
the device environment kernels (observe / select_action + Game.step + history record / harvest + export ring) and the
CUDA tree kernels, fed the same table-driven network outputs and Dirichlet noise: actions, rewards, to-play,
observations, child-visit distributions and root values of every move are bit-exact."""
import zlib

import numpy as np
import pytest
import torch

import _tables as T
from oracle import rng
from test_gpu_tree import child_row_torch

pytestmark = pytest.mark.gpu

Z = T.load("episode")
N = int(Z["n"])
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("i", range(N))
def test_device_episode_equals_reference_play_game(i):
    from muzero_hypermodel_b200.envs import VectorEnv
    from muzero_hypermodel_b200.self_play import decode_export
    from muzero_hypermodel_b200.tree import BatchedTree
    pre = f"{i}/"
    A, players, sims, max_moves, discount, alpha, temp, thr, slot, seed_env = Z[pre + "meta"]
    A, players, S, max_moves, thr, slot, seed_env = int(A), int(players), int(sims), int(max_moves), int(thr), int(slot), int(seed_env)
    discount = int(discount) if float(discount).is_integer() else float(discount)
    game, tab = str(Z[pre + "game"]), str(Z[pre + "table"])
    V = torch.tensor(Z[f"{tab}/V"], device=DEV); Rw = torch.tensor(Z[f"{tab}/Rw"], device=DEV); P = torch.tensor(Z[f"{tab}/P"], device=DEV)
    cart = game == "cartpole"
    # cartpole: the golden game drew its reset state with (seed_env, slot 0) and its action uniforms with (SEED, slot):
    # the environment gets the former, the uniforms are injected; board games draw their own uniforms on the device
    env = VectorEnv(game, 1, max_moves, seed=seed_env if cart else T.SEED, first_slot=0 if cart else slot, device=DEV)
    tree = BatchedTree(1, A, S, players, discount, T.PB_C_BASE, T.PB_C_INIT, seed=T.SEED, device=DEV)
    n_moves = len(Z[pre + "actions"]) - 1
    parent = torch.empty(1, dtype=torch.int32, device=DEV)
    action = torch.empty(1, dtype=torch.int32, device=DEV)
    games = []
    for mi in range(n_moves):
        obs, legal, to_play = env.observe()
        o = obs.cpu().numpy().reshape(env.obs_shape)
        row = int(Z[pre + "root_rows"][mi])
        if not cart:     # integer boards: the observation (hence its hash) must be the reference's exactly
            assert row == zlib.crc32(np.ascontiguousarray(o, dtype=np.float32).tobytes()) % T.N_ROWS, f"move {mi}"
        lg = np.nonzero(legal.cpu().numpy()[0])[0]
        pri = np.zeros((1, A), dtype=np.float32); pri[0, lg] = Z[pre + f"root_priors/{mi}"]
        nz = np.zeros((1, A), dtype=np.float64); nz[0, lg] = Z[pre + f"noise/{mi}"]
        rows = torch.zeros((1, S + 1), dtype=torch.int64, device=DEV)
        rows[0, 0] = row
        tree.root_init(Rw[rows[:, 0]].contiguous(), torch.tensor(pri, device=DEV), policy_is_logits=False, legal=legal,
                       to_play=to_play, noise=torch.tensor(nz, device=DEV), alpha=float(alpha), frac=T.FRAC,
                       slot=torch.tensor([slot], dtype=torch.int32, device=DEV),
                       step=torch.tensor([mi], dtype=torch.int32, device=DEV))
        for sim in range(S):
            tree.select(parent, action)
            crow = child_row_torch(rows[0, parent.long()], action.long())
            rows[0, sim + 1] = crow
            tree.expand_backup(V[crow].contiguous(), Rw[crow].contiguous(), P[crow].contiguous(), policy_is_logits=False)
        st = tree.root_stats()
        u = torch.tensor([rng.action_uniform(T.SEED, slot, mi)], dtype=torch.float64, device=DEV) if cart else None
        env.act_step(st["visits"], st["root_value"], legal, float(temp), thr or None, uniforms=u)
        env.harvest(True)
        games += decode_export(env)
    assert len(games) == 1, "the device game must end exactly where the reference's did"
    gh = games[0]
    assert gh.action_history == Z[pre + "actions"].tolist()
    assert gh.to_play_history == Z[pre + "to_play"].tolist()
    assert np.array(gh.reward_history, dtype=np.float64).tobytes() == Z[pre + "rewards"].tobytes()
    assert np.array(gh.root_values, dtype=np.float64).tobytes() == Z[pre + "root_values"].tobytes()
    assert np.array(gh.child_visits, dtype=np.float64).tobytes() == Z[pre + "child_visits"].tobytes()
    got = np.array([np.asarray(o, dtype=np.float32) for o in gh.observation_history])
    assert got.shape == Z[pre + "observations"].shape
    if cart:     # float64 physics on both sides, observation cast to float32: stated 1e-5 relative bound (sin/cos)
        np.testing.assert_allclose(got, Z[pre + "observations"], rtol=1e-5, atol=1e-7)
    else:
        assert got.tobytes() == Z[pre + "observations"].tobytes()
