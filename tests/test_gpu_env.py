"""Device environments, action selection, recording and self-play loop vs golden vectors and the oracle."""
import numpy as np
import pytest
import torch

import _tables as T
from _configs import product_config
from oracle import games as ogames
from oracle import mcts as omcts
from oracle import rng, selfplay as oselfplay

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", ["tictactoe", "connect4", "gomoku"])
def test_board_envs_bit_exact_vs_reference(name):
    from muzero_hypermodel_b200.envs import VectorEnv
    z = T.load("env")
    n = int(z[f"{name}/n"])
    seqs = [z[f"{name}/{g}/actions"] for g in range(n)]
    env = VectorEnv(name, n, 200, device=DEV)
    sv = env._state_view()
    H, W = env.obs_shape[1:]
    alive = np.ones(n, dtype=bool)
    t = 0
    while alive.any():
        obs, legal, tp = [x.cpu().numpy() for x in env.observe()]
        board = sv["board"].cpu().numpy().reshape(n, H, W)
        for g in np.nonzero(alive)[0]:
            pre = f"{name}/{g}/"
            np.testing.assert_array_equal(board[g], z[pre + "boards"][t])
            np.testing.assert_array_equal(obs[g].reshape(3, H, W), z[pre + "obs"][t])
            np.testing.assert_array_equal(legal[g], z[pre + "legal"][t])
            assert tp[g] == z[pre + "to_play"][t]
        acts = torch.tensor([int(seqs[g][t]) if t < len(seqs[g]) else 0 for g in range(n)], dtype=torch.int32, device=DEV)
        a, r, d = env.act_step(None, None, forced_action=acts, want_outputs=True)
        r, d = r.cpu().numpy(), d.cpu().numpy()
        for g in np.nonzero(alive)[0]:
            pre = f"{name}/{g}/"
            assert r[g] == z[pre + "rewards"][t]
            assert bool(d[g]) == bool(z[pre + "dones"][t])
            if t + 1 >= len(seqs[g]):
                alive[g] = False
        t += 1


def test_cartpole_physics_vs_golden():
    from muzero_hypermodel_b200.envs import VectorEnv
    z = T.load("env")
    n = int(z["cartpole/n"])
    env = VectorEnv("cartpole", n, 500, seed=0, device=DEV)
    sv = env._state_view()
    # reset state = Philox(seed, slot, step 0): the golden episodes were generated with seed = game index, slot 0
    for g in range(n):
        e1 = VectorEnv("cartpole", 1, 500, seed=g, device=DEV)
        o = e1.observe()[0].cpu().numpy()[0]
        np.testing.assert_array_equal(o, z[f"cartpole/{g}/obs"][0].reshape(4))        # bit-exact initial state
        sv["cartpole"][g] = e1._state_view()["cartpole"][0]
    # free run on the golden action sequences: float64 states, libm vs CUDA sin/cos -> tolerance, first 60 steps
    L = min(len(z[f"cartpole/{g}/actions"]) for g in range(n))
    for t in range(min(L, 60)):
        acts = torch.tensor([int(z[f"cartpole/{g}/actions"][t]) for g in range(n)], dtype=torch.int32, device=DEV)
        env.act_step(None, None, forced_action=acts)
        o = env.observe()[0].cpu().numpy()
        want = np.stack([z[f"cartpole/{g}/obs"][t + 1].reshape(4) for g in range(n)])
        np.testing.assert_allclose(o, want, rtol=1e-5, atol=1e-7)
    # teacher-forced single steps along whole episodes (incl. the 500-step balanced ones) in float64
    for g in (0, n - 1):
        oenv = ogames.CartPoleV1(seed=g)
        oenv.reset()
        acts = z[f"cartpole/{g}/actions"]
        states = [oenv.state]
        for a in acts:
            oenv.step(int(a))
            states.append(oenv.state)
        states = np.array(states)
        k = len(acts)
        envk = VectorEnv("cartpole", k, 1000, device=DEV)
        envk._state_view()["cartpole"].copy_(torch.tensor(states[:-1], device=DEV))
        _, _, d = envk.act_step(None, None, forced_action=torch.tensor(acts, dtype=torch.int32, device=DEV), want_outputs=True)
        got = envk._state_view()["cartpole"].cpu().numpy()
        np.testing.assert_allclose(got, states[1:], rtol=1e-12, atol=1e-15)
        out_of_bounds = (np.abs(states[1:, 0]) > 2.4) | (np.abs(states[1:, 2]) > ogames.THETA_THRESHOLD)
        np.testing.assert_array_equal(d.cpu().numpy().astype(bool), out_of_bounds)


def test_select_action_kernel_vs_reference():
    """act_step's action choice against the reference select_action golden cases (injected uniform)."""
    from muzero_hypermodel_b200.envs import VectorEnv
    z = T.load("action")
    by_A = {}
    for i in range(int(z["n"])):
        by_A.setdefault(int(z[f"{i}/A"]), []).append(i)
    kinds = {2: "cartpole", 9: "tictactoe", 7: "connect4", 121: "gomoku"}
    checked = 0
    for A, idx in by_A.items():
        if A not in kinds:
            continue
        for T_ in sorted(set(float(z[f"{i}/T"]) for i in idx)):
            cases = [i for i in idx if float(z[f"{i}/T"]) == T_]
            G = len(cases)
            env = VectorEnv(kinds[A], G, 300, device=DEV)
            visits = np.zeros((G, A), dtype=np.int32)
            legal = np.zeros((G, A), dtype=np.uint8)
            u = np.zeros(G)
            for g, i in enumerate(cases):
                acts = z[f"{i}/actions"]
                visits[g, acts] = z[f"{i}/visits"]
                legal[g, acts] = 1
                u[g] = float(z[f"{i}/u"])
            a, _, _ = env.act_step(torch.tensor(visits, device=DEV), torch.zeros(G, dtype=torch.float64, device=DEV),
                                   legal=torch.tensor(legal, device=DEV), temperature=T_,
                                   uniforms=torch.tensor(u, device=DEV), want_outputs=True)
            want = np.array([int(z[f"{i}/action"]) for i in cases])
            np.testing.assert_array_equal(a.cpu().numpy(), want)
            checked += G
    assert checked > 250


def _fc_net(tag):
    from muzero_hypermodel_b200 import models
    cfg = product_config(tag)
    z = T.load("net")
    pre = tag + "/w/"
    net = models.MuZeroNetwork(cfg)
    sd = {k[len(pre):]: torch.tensor(z[k]) for k in z.files if k.startswith(pre)}
    net.set_weights(sd)
    return net, cfg, sd


@pytest.mark.parametrize("tag,kind", [("tictactoe_fc", "tictactoe"), ("cartpole", "cartpole")])
def test_self_play_games_equal_oracle_play_game(tag, kind):
    """Full episodes: device self-play (fused search, device env, device recording, export ring) vs the oracle's
    play_game fed by the same network kernels, with injected Dirichlet noise and the shared counter RNG."""
    from muzero_hypermodel_b200.self_play import SelfPlay
    net, cfg, sd = _fc_net(tag)
    if kind == "cartpole":
        cfg.max_moves = 40                          # keep the oracle side short
    G, SEED = 5, 77
    sp = SelfPlay({"weights": sd}, None, cfg, SEED, n_games=G, device=DEV)
    env, mcts = sp._setup()
    A = env.A
    rs = np.random.RandomState(3)
    noise_log = {}
    finished = []
    for move in range(cfg.max_moves + 1):
        obs, legal, to_play = env.observe()
        lg = legal.cpu().numpy().astype(bool)
        steps = env.step_count.cpu().numpy()
        nz = np.zeros((G, A))
        for g in range(G):
            nz[g, lg[g]] = rs.dirichlet([cfg.root_dirichlet_alpha] * int(lg[g].sum()))
            noise_log[(g, int(steps[g]))] = nz[g, lg[g]].copy()
        out = mcts.run(sp.model, obs, legal, to_play, True, noise=torch.tensor(nz, device=DEV), slot=env.slot,
                       step=env.step_count)
        env.act_step(out["visits"], out["root_value"], legal, 1.0, None)
        env.harvest(True)
        finished += sp.drain()
        if len(finished) >= G:
            break
    assert len(finished) >= G
    dev = sp.model
    first = {}
    for gh in finished:
        first.setdefault(gh.slot, gh)               # first episode of each slot
    for g in range(G):
        gh = first[g]
        oenv = {"tictactoe": ogames.TicTacToe, "cartpole": None}[kind]
        if kind == "cartpole":
            oenv = ogames.CartPole(1, seed=SEED, slot0=g)
        else:
            oenv = oenv(1)

        def initial(obs, legal_actions):
            lm = torch.zeros((1, A), dtype=torch.uint8, device=DEV)
            lm[0, legal_actions] = 1
            o = dev.initial_inference_fused(torch.tensor(obs, device=DEV).reshape(1, -1), legal=lm)
            return (float(o["value"][0]), float(o["reward"][0]), [float(o["priors"][0, a]) for a in legal_actions], o["state"])

        def recurrent(hidden, action):
            r = dev.recurrent_inference_fused(hidden, torch.tensor([[action]], device=DEV))
            return float(r["value"][0]), float(r["reward"][0]), [float(x) for x in r["priors"][0]], r["state"]

        oh = oselfplay.play_game(oenv, initial, recurrent, cfg, 1.0, None, SEED, g,
                                 lambda step, legal_actions, _g=g: [float(x) for x in noise_log[(_g, step)]])
        assert gh.action_history == oh.action_history
        assert gh.to_play_history == oh.to_play_history
        assert [float(x) for x in gh.reward_history] == [float(x) for x in oh.reward_history]
        assert np.array(gh.root_values).tobytes() == np.array(oh.root_values).tobytes()
        assert np.array(gh.child_visits, dtype=np.float64).tobytes() == np.array(oh.child_visits, dtype=np.float64).tobytes()
        for a, b in zip(gh.observation_history, oh.observation_history):
            if kind == "cartpole":
                np.testing.assert_allclose(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1), rtol=1e-5, atol=1e-7)
            else:
                np.testing.assert_array_equal(np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32))
    c = env.counters()
    assert c["games"] >= G and c["dropped_games"] == 0


def test_auto_reset_and_counters():
    from muzero_hypermodel_b200.self_play import SelfPlay
    net, cfg, sd = _fc_net("tictactoe_fc")
    G = 512
    sp = SelfPlay({"weights": sd}, None, cfg, 1, n_games=G, device=DEV)
    games = sp.play_games(30, drain_every=3)
    c = sp._env.counters()
    assert c["env_steps"] == 30 * G                      # every game moves every step (auto-reset keeps the batch full)
    assert c["games"] == len(games) and c["dropped_games"] == 0
    assert c["finished_moves"] == sum(len(g.root_values) for g in games)
    for gh in games[:50]:
        n = len(gh.root_values)
        assert 5 <= n <= 9 and len(gh.action_history) == n + 1 and len(gh.observation_history) == n + 1
        assert all(abs(sum(cv) - 1.0) < 1e-12 for cv in gh.child_visits)
        # replay the recorded actions through the oracle environment: same observations / rewards / turns
        env = ogames.TicTacToe(1)
        obs = env.observation()
        for i in range(n):
            np.testing.assert_array_equal(np.asarray(gh.observation_history[i], dtype=np.float32), obs[0])
            assert gh.child_visits[i][gh.action_history[i + 1]] > 0
            obs, r, d = env.step(np.array([gh.action_history[i + 1]]))
            assert float(r[0]) == gh.reward_history[i + 1] and int(env.to_play()[0]) == gh.to_play_history[i + 1]
        assert bool(d[0])


def test_game_plugin_contract():
    """AbstractGame contract: reset / legal_actions / to_play / step return what the reference wrappers return."""
    from muzero_hypermodel_b200.games import connect4, tictactoe
    z = T.load("env")
    for mod, name in ((tictactoe, "tictactoe"), (connect4, "connect4")):
        game = mod.Game(seed=0)
        obs = game.reset()
        pre = f"{name}/0/"
        acts = z[pre + "actions"]
        for t, a in enumerate(acts):
            np.testing.assert_array_equal(np.asarray(obs, dtype=np.float32), z[pre + "obs"][t])
            assert game.legal_actions() == np.nonzero(z[pre + "legal"][t])[0].tolist()
            assert game.to_play() == z[pre + "to_play"][t]
            obs, reward, done = game.step(int(a))
            assert reward == z[pre + "rewards"][t] and done == bool(z[pre + "dones"][t])
        assert obs.dtype == (np.int32 if name == "tictactoe" else np.float64)


def test_synthetic_frames_env_matches_oracle():
    from muzero_hypermodel_b200.envs import VectorEnv
    env = VectorEnv("breakout", 3, 50, seed=5, first_slot=7, device=DEV)
    o = ogames.SyntheticFrames(3, seed=5, slot0=7)
    for t in range(3):
        obs, legal, tp = env.observe()
        np.testing.assert_array_equal(obs.cpu().numpy().reshape(3, 3, 96, 96), o.observation())
        assert legal.all() and not tp.any()
        _, r, d = env.act_step(None, None, forced_action=torch.zeros(3, dtype=torch.int32, device=DEV), want_outputs=True)
        o.step(np.zeros(3))
        assert not r.any() and not d.any()
    assert 0.45 < float(obs.mean()) < 0.55


@pytest.mark.parametrize("kind,S", [("tictactoe", 2), ("connect4", 4), ("cartpole", 3)])
def test_device_stacked_observations_equal_game_history(kind, S):
    """mzb_env_observe_stacked against GameHistory.get_stacked_observations(-1, S) (pinned to the reference by
    tests/golden/stacked.npz) applied to the same running games: at every move of random play, incl. the zero planes
    before the start of a game and across an auto-reset."""
    from muzero_hypermodel_b200.envs import VectorEnv
    from muzero_hypermodel_b200.self_play import GameHistory
    G, moves = 6, 14
    env = VectorEnv(kind, G, 500 if kind == "cartpole" else 60, seed=3, device=DEV)
    rs = np.random.RandomState(1)
    hist = [GameHistory() for _ in range(G)]
    obs, legal, _ = env.observe()
    for g in range(G):
        hist[g].observation_history.append(obs[g].cpu().numpy().reshape(env.obs_shape).copy())
        hist[g].action_history.append(0)
    for mv in range(moves):
        st, legal, _ = env.observe_stacked(S)
        st = st.cpu().numpy()
        for g in range(G):
            want = np.asarray(hist[g].get_stacked_observations(-1, S), dtype=np.float32)
            assert st[g].shape == want.shape and np.array_equal(st[g], want), (mv, g)
        lg = legal.cpu().numpy().astype(bool)
        act = np.array([rs.choice(np.nonzero(lg[g])[0]) for g in range(G)], dtype=np.int32)
        _, _, done = env.act_step(None, None, forced_action=torch.tensor(act, device=DEV), want_outputs=True)
        done = done.cpu().numpy().astype(bool)
        env.harvest(False)
        obs, _, _ = env.observe()
        for g in range(G):
            if done[g]:
                hist[g] = GameHistory()
                hist[g].action_history.append(0)
            else:
                hist[g].action_history.append(int(act[g]))
            hist[g].observation_history.append(obs[g].cpu().numpy().reshape(env.obs_shape).copy())


def test_caller_supplied_game_class_and_opponents():
    """SelfPlay honours the `Game` argument (self_play.py:16-19): a host AbstractGame subclass is played through
    step / legal_actions / to_play with every search on the device kernels, incl. the evaluation opponents
    ("random", "expert") and test mode of continuous_self_play (:54-88)."""
    from muzero_hypermodel_b200.games.abstract_game import AbstractGame
    from muzero_hypermodel_b200.games.tictactoe import MuZeroConfig
    from muzero_hypermodel_b200.self_play import SelfPlay
    from muzero_hypermodel_b200.shared_storage import SharedStorage, new_checkpoint

    class Nim(AbstractGame):
        """3x3x3-shaped toy: players alternately take 1-3 of 9 stones; taking the last stone wins."""
        def __init__(self, seed=None): self.reset()
        def reset(self):
            self.left, self.player = 9, 0
            return self._obs()
        def _obs(self):
            o = np.zeros((3, 3, 3), dtype=np.int32)
            o[0].flat[:self.left] = 1
            o[2][:] = 1 if self.player == 0 else -1
            return o
        def step(self, action):
            take = action % 3 + 1
            self.left -= min(take, self.left)
            done = self.left == 0
            reward = 1 if done else 0
            self.player = 1 - self.player
            return self._obs(), reward, done
        def to_play(self): return self.player
        def legal_actions(self): return [a for a in range(9) if a % 3 + 1 <= self.left]
        def expert_agent(self): return (self.left % 4 or 1) - 1
        def close(self): pass
        def render(self): pass
        def action_to_string(self, a): return f"take {a % 3 + 1}"

    cfg = MuZeroConfig()
    cfg.network, cfg.num_simulations, cfg.opponent, cfg.muzero_player = "fullyconnected", 12, "expert", 0
    torch.manual_seed(0)
    sp = SelfPlay({"weights": None}, Nim, cfg, 0, device=DEV)
    assert sp.host_game
    gh = sp.play_game(1.0, None, False, "self", 0)
    assert len(gh.root_values) == len(gh.action_history) - 1 >= 3 and all(v is not None for v in gh.root_values)
    assert sum(gh.reward_history) == 1 and all(len(cv) == 9 and abs(sum(cv) - 1) < 1e-12 for cv in gh.child_visits)
    for opp in ("random", "expert"):
        g = sp.play_game(0, None, False, opp, 0)
        assert any(v is None for v in g.root_values) and any(v is not None for v in g.root_values)      # opponent moves carry no search
    st = SharedStorage(new_checkpoint(sp.model.get_weights()), cfg)
    sp.continuous_self_play(st, None, test_mode=True, max_moves=2)
    assert st.get_info("episode_length") >= 3 and st.get_info("muzero_reward") + st.get_info("opponent_reward") == 1
    with pytest.raises(NotImplementedError):
        SelfPlay({"weights": None}, Nim, cfg, 0, n_games=64, device=DEV)
