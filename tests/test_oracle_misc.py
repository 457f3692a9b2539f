"""CPU oracle vs the committed reference outputs: environments, codec, networks, targets, action choice."""
import numpy as np
import pytest

import _tables as T
from _weights import seeded_state_dict
from oracle import games, mcts, networks, rng, targets


@pytest.mark.parametrize("name,cls", [("tictactoe", games.TicTacToe), ("connect4", games.Connect4), ("gomoku", games.Gomoku)])
def test_board_games_bit_exact(name, cls):
    z = T.load("env")
    n = int(z[f"{name}/n"])
    seqs = [z[f"{name}/{g}/actions"] for g in range(n)]
    env = cls(n)
    obs = env.observation()
    t = 0
    alive = np.ones(n, dtype=bool)
    while alive.any():
        for g in range(n):
            if alive[g]:
                pre = f"{name}/{g}/"
                np.testing.assert_array_equal(env.board[g], z[pre + "boards"][t])
                np.testing.assert_array_equal(obs[g], z[pre + "obs"][t])
                np.testing.assert_array_equal(env.legal_mask()[g].astype(np.uint8), z[pre + "legal"][t])
                assert env.to_play()[g] == z[pre + "to_play"][t]
        acts = np.array([seqs[g][t] if t < len(seqs[g]) else 0 for g in range(n)])
        obs, rew, done = env.step(acts)
        for g in range(n):
            if alive[g]:
                pre = f"{name}/{g}/"
                assert rew[g] == z[pre + "rewards"][t]
                assert bool(done[g]) == bool(z[pre + "dones"][t])
                if t + 1 >= len(seqs[g]):
                    np.testing.assert_array_equal(env.board[g], z[pre + "boards"][t + 1])
                    alive[g] = False
        t += 1


def test_cartpole_matches_reference_wrapper():
    z = T.load("env")
    for g in range(int(z["cartpole/n"])):
        env = games.CartPoleV1(seed=g)
        obs = [env.reset()]
        for a in z[f"cartpole/{g}/actions"]:
            o, r, d = env.step(int(a))
            obs.append(o)
        assert d
        np.testing.assert_array_equal(np.array(obs, dtype=np.float32), z[f"cartpole/{g}/obs"].reshape(-1, 4))


def test_codec():
    z = T.load("codec")
    with np.errstate(divide="ignore", invalid="ignore"):
        s = networks.support_to_scalar(z["logits"], 10)
    # The inverse h-transform subtracts 1 from sqrt(1 + 0.004(|x|+1.001)) in float32: a 1-ulp change of the
    # softmax expectation moves the result by ~1e-4 relative (DESIGN.md "support codec conditioning"), so
    # across different exp() implementations only ~3e-4 holds; with an exact expectation it is bit-exact.
    np.testing.assert_allclose(s, z["scalars"], rtol=3e-4, atol=2e-4)
    np.testing.assert_array_equal(s[256:277], z["scalars"][256:277])        # one-hot rows: exact integers in
    assert np.signbit(s[-2, 0]) and s[-2, 0] == 0          # log one-hot(centre) decodes to -0.0
    sup = networks.scalar_to_support(z["x"], 10)
    np.testing.assert_allclose(sup, z["support"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(sup.sum(-1), 1.0, atol=1e-6)


def _weights(z, tag):
    pre = tag + "/w/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


@pytest.mark.parametrize("tag,A", [("cartpole_shipped", 2), ("cartpole", 2), ("tictactoe_fc", 9)])
def test_fc_network(tag, A):
    z = T.load("net")
    net = networks.FullyConnected(_weights(z, tag), A, 10)
    v0, r0, p0, s0 = net.initial_inference(z[tag + "/obs"])
    v1, r1, p1, s1 = net.recurrent_inference(s0, z[tag + "/act"])
    v2, r2, p2, s2 = net.recurrent_inference(s1, z[tag + "/act2"])
    for got, name in ((v0, "v0"), (p0, "p0"), (s0, "s0"), (v1, "v1"), (r1, "r1"), (p1, "p1"), (s1, "s1"),
                      (v2, "v2"), (r2, "r2"), (p2, "p2"), (s2, "s2")):
        np.testing.assert_allclose(got, z[f"{tag}/{name}"], rtol=1e-5, atol=2e-6, err_msg=name)
    np.testing.assert_array_equal(r0, z[tag + "/r0"])
    np.testing.assert_allclose(networks.support_to_scalar(v1, 10), z[tag + "/sv1"], rtol=3e-4, atol=2e-4)


RES = {"tictactoe": (9, 1, (3, 3, 3), False), "connect4": (7, 3, (3, 6, 7), False),
       "gomoku": (121, 6, (3, 11, 11), False), "breakout": (4, 2, (3, 96, 96), "resnet")}


@pytest.mark.parametrize("tag", list(RES))
def test_residual_network(tag):
    z = T.load("net")
    A, blocks, shape, down = RES[tag]
    if tag == "gomoku":
        keys = str(z["gomoku/keys"]).split("\n")
        shapes = [[int(d) for d in s.split("x")] if s else [] for s in z["gomoku/shapes"]]
        sd = seeded_state_dict(keys, shapes)
    else:
        sd = _weights(z, tag)
    net = networks.Residual(sd, shape, A, blocks, 10, down)
    v0, r0, p0, s0 = net.initial_inference(z[tag + "/obs"])
    v1, r1, p1, s1 = net.recurrent_inference(s0, z[tag + "/act"])
    for got, name in ((v0, "v0"), (p0, "p0"), (s0, "s0"), (v1, "v1"), (r1, "r1"), (p1, "p1"), (s1, "s1")):
        np.testing.assert_allclose(got, z[f"{tag}/{name}"], rtol=1e-4, atol=1e-5, err_msg=name)


def test_targets_bit_exact():
    z = T.load("targets")
    for i in range(int(z["n"])):
        pre = f"{i}/"
        K, td, disc, A, players = z[pre + "cfg"]
        disc = int(disc) if float(disc).is_integer() else float(disc)
        slot, step = [int(v) for v in z[pre + "slot_step"]]
        re = z[pre + "reanalysed"]
        tv, tr, tp, ta = targets.make_target(
            z[pre + "root_values"].tolist(), z[pre + "reward_history"].tolist(), z[pre + "to_play_history"].tolist(),
            z[pre + "child_visits"].tolist(), z[pre + "action_history"].tolist(), int(z[pre + "state_index"]),
            int(K), int(td), disc, int(A), reanalysed_root_values=re.tolist() if len(re) else None,
            pad_action=lambda row: rng.pad_action(T.SEED, slot, step, _PAST.pop(0), int(A)))
        assert np.array(tv, dtype=np.float64).tobytes() == z[pre + "target_values"].tobytes()
        assert np.array(tr, dtype=np.float64).tobytes() == z[pre + "target_rewards"].tobytes()
        assert np.array(tp, dtype=np.float64).tobytes() == z[pre + "target_policies"].tobytes()
        np.testing.assert_array_equal(np.array(ta, dtype=np.int32), z[pre + "actions"])
        _PAST[:] = list(range(200))


_PAST = list(range(200))    # the k-th past-the-end row of a make_target call draws pad index k


def test_select_action_and_statistics():
    z = T.load("action")
    for i in range(int(z["n"])):
        pre = f"{i}/"
        actions, visits = z[pre + "actions"].tolist(), z[pre + "visits"].tolist()
        a = mcts.select_action(actions, visits, float(z[pre + "T"]), float(z[pre + "u"]))
        assert a == int(z[pre + "action"]), i
        cv = mcts.search_statistics(actions, visits, int(z[pre + "A"]))
        assert np.array(cv, dtype=np.float64).tobytes() == z[pre + "child_visits"].tobytes()


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    assert rng.philox4x32(0, 0, 0, 0, 0, 0) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert rng.philox4x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == (
        0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert rng.philox4x32(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)
    c = np.arange(5)
    v = rng.philox4x32_np(c, c + 1, c + 2, c + 3, 7, 9)
    for i in range(5):
        assert tuple(int(x[i]) for x in v) == rng.philox4x32(i, i + 1, i + 2, i + 3, 7, 9)
