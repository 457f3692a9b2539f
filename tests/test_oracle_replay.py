"""The oracle ReplayBuffer reproduces the UNMODIFIED reference class (replay_buffer.py:33-220) on a scripted sequence
of save_game / get_batch / update_priorities calls with injected uniforms (tests/golden/replay.npz): initial PER
priorities, FIFO eviction, sampled (game, position) pairs, targets, importance weights, gradient scales and the
priorities after every update - bit-exact."""
import numpy as np
import pytest

import _tables as T
from oracle import replay as oreplay
from oracle import rng

Z = T.load("replay")


def case_config(ci):
    per, alpha, size, batch, K, td, discount, A, players = Z[f"{ci}/cfg"]
    discount = int(discount) if float(discount).is_integer() else float(discount)
    alpha = int(alpha) if float(alpha).is_integer() else float(alpha)
    return bool(per), alpha, int(size), int(batch), int(K), int(td), discount, int(A), int(players)


def load_game(ci, gi):
    g = f"{ci}/game{gi}/"
    vis = Z[g + "visits"]
    cv = [[int(v) / int(row.sum()) if v else 0 for v in row] for row in vis]
    return oreplay.Game(list(Z[g + "observations"]), Z[g + "actions"].tolist(), Z[g + "rewards"].tolist(),
                        Z[g + "to_play"].tolist(), cv, Z[g + "root_values"].tolist())


def test_np_sum_restatement_matches_numpy():
    rs = np.random.RandomState(5)
    for n in [1, 2, 7, 8, 9, 127, 128, 129, 255, 1000, 8193, 20000, 200000]:
        for _ in range(3):
            a = (rs.uniform(0, 10, n) ** 3).astype(np.float32)
            assert np.sum(a).tobytes() == oreplay.np_sum_f32(a).tobytes(), n


@pytest.mark.parametrize("ci", range(int(Z["n"])))
def test_oracle_replay_equals_reference(ci):
    per, alpha, size, batch, K, td, discount, A, players = case_config(ci)
    rb = oreplay.ReplayBuffer(per, alpha, size, batch, K, td, discount, A)
    gi = bi = ui = 0
    for op in str(Z[f"{ci}/script"]).split():
        if op == "save":
            g = load_game(ci, gi)
            rb.save_game(g)
            if per:
                assert g.priorities.tobytes() == Z[f"{ci}/game{gi}/priorities"].tobytes()
                assert np.float32(g.game_priority).tobytes() == Z[f"{ci}/game{gi}/game_priority"].tobytes()
            gi += 1
        elif op == "batch":
            b = f"{ci}/batch{bi}/"
            ug = [rng.replay_uniform(T.SEED, e, bi, rng.STREAM_RGAME) for e in range(batch)]
            up = [rng.replay_uniform(T.SEED, e, bi, rng.STREAM_RPOS) for e in range(batch)]
            index, (obs, act, val, rew, pol, w, gs) = rb.get_batch(
                ug, up, lambda e, row, _bi=bi: rng.pad_action(T.SEED, e, _bi, row, A))
            assert np.array(index, dtype=np.int64).tobytes() == Z[b + "index"].tobytes()
            assert np.array(obs, dtype=np.float32).tobytes() == Z[b + "observations"].tobytes()
            assert np.array(act, dtype=np.int32).tobytes() == Z[b + "actions"].tobytes()
            assert np.array(val, dtype=np.float64).tobytes() == Z[b + "values"].tobytes()
            assert np.array(rew, dtype=np.float64).tobytes() == Z[b + "rewards"].tobytes()
            assert np.array(pol, dtype=np.float64).tobytes() == Z[b + "policies"].tobytes()
            assert np.array(gs, dtype=np.int32).tobytes() == Z[b + "gradient_scale"].tobytes()
            if per:
                assert w.dtype == np.float32 and w.tobytes() == Z[b + "weights"].tobytes()
            assert [rb.total_samples, rb.num_played_games, len(rb.buffer)] == Z[b + "state"].tolist()
            bi += 1
        else:
            u = f"{ci}/update{ui}/"
            if per:
                rb.update_priorities(Z[u + "priorities"], Z[u + "index"].tolist())
                for gid, g in rb.buffer.items():
                    assert g.priorities.tobytes() == Z[u + f"after/{gid}"].tobytes()
                    assert np.float32(g.game_priority).tobytes() == Z[u + f"after_game/{gid}"].tobytes()
            ui += 1
