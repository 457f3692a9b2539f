"""Device replay store (csrc/mzb_replay.cu behind muzero_hypermodel_b200.replay_buffer.ReplayBuffer) vs the golden
outputs of the UNMODIFIED reference ReplayBuffer (tests/golden/replay.npz): the scripted save_game / get_batch /
update_priorities sequence with injected uniforms and with the device's own Philox draws - initial PER priorities,
FIFO eviction, sampled indices, observations, targets, float32 importance weights, gradient scales and updated
priorities are bit-exact."""
import numpy as np
import pytest
import torch

import _tables as T
from oracle import rng
from test_oracle_replay import Z, case_config

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


class _Cfg:
    pass


class _Game:
    pass


def make(ci):
    from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
    per, alpha, size, batch, K, td, discount, A, players = case_config(ci)
    cfg = _Cfg()
    cfg.action_space = list(range(A)); cfg.PER, cfg.PER_alpha, cfg.replay_buffer_size, cfg.batch_size = per, alpha, size, batch
    cfg.num_unroll_steps, cfg.td_steps, cfg.discount, cfg.seed, cfg.stacked_observations = K, td, discount, T.SEED, 0
    cfg.observation_shape = tuple(Z[f"{ci}/game0/observations"].shape[1:])
    cfg.max_moves, cfg.num_simulations = 100, int(Z[f"{ci}/game0/visits"][0].sum())
    return ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg, device=DEV), cfg


def load_game(ci, gi):
    g, pre = _Game(), f"{ci}/game{gi}/"
    g.observation_history = list(Z[pre + "observations"])
    g.action_history, g.reward_history = Z[pre + "actions"].tolist(), Z[pre + "rewards"].tolist()
    g.to_play_history, g.root_values = Z[pre + "to_play"].tolist(), Z[pre + "root_values"].tolist()
    vis = Z[pre + "visits"]
    g.child_visits = [[int(v) / int(row.sum()) if v else 0 for v in row] for row in vis]
    g.priorities = None
    return g


@pytest.mark.parametrize("inject", [True, False])
@pytest.mark.parametrize("ci", range(int(Z["n"])))
def test_device_replay_equals_reference(ci, inject):
    rb, cfg = make(ci)
    per, batch = cfg.PER, cfg.batch_size
    gi = bi = ui = 0
    for op in str(Z[f"{ci}/script"]).split():
        if op == "save":
            rb.save_game(load_game(ci, gi))
            if per:
                pr, gp = rb.game_priorities(gi)
                assert pr.tobytes() == Z[f"{ci}/game{gi}/priorities"].tobytes(), ("initial priorities", gi)
                assert np.float32(gp).tobytes() == Z[f"{ci}/game{gi}/game_priority"].tobytes()
            gi += 1
        elif op == "batch":
            b = f"{ci}/batch{bi}/"
            ug = up = None
            if inject:
                ug = [rng.replay_uniform(T.SEED, e, bi, rng.STREAM_RGAME) for e in range(batch)]
                up = [rng.replay_uniform(T.SEED, e, bi, rng.STREAM_RPOS) for e in range(batch)]
            index, (obs, act, val, rew, pol, w, gs) = rb.get_batch(u_game=ug, u_pos=up)
            assert index.cpu().numpy().tobytes() == Z[b + "index"].tobytes(), ("index", bi)
            assert obs.cpu().numpy().tobytes() == Z[b + "observations"].tobytes()
            assert act.cpu().numpy().tobytes() == Z[b + "actions"].tobytes()
            assert val.cpu().numpy().tobytes() == Z[b + "values"].tobytes()
            assert rew.cpu().numpy().tobytes() == Z[b + "rewards"].tobytes()
            assert pol.cpu().numpy().tobytes() == Z[b + "policies"].tobytes()
            assert gs.cpu().numpy().tobytes() == Z[b + "gradient_scale"].tobytes()
            if per:
                assert w.cpu().numpy().tobytes() == Z[b + "weights"].tobytes(), ("weights", bi)
            else:
                assert w is None
            assert [rb.total_samples, rb.num_played_games, len(rb)] == Z[b + "state"].tolist()
            bi += 1
        else:
            u = f"{ci}/update{ui}/"
            if per:
                rb.update_priorities(Z[u + "priorities"], Z[u + "index"])
                first = rb._info()[4]
                for gid in range(first, first + len(rb)):
                    pr, gp = rb.game_priorities(gid)
                    assert pr.tobytes() == Z[u + f"after/{gid}"].tobytes(), ("updated priorities", gid)
                    assert np.float32(gp).tobytes() == Z[u + f"after_game/{gid}"].tobytes()
            ui += 1


def test_empty_buffer_and_oversize_game_are_refused():
    from muzero_hypermodel_b200._lib import MzbError
    rb, cfg = make(0)
    with pytest.raises(MzbError):
        rb.get_batch()
    g = load_game(0, 0)
    n = 200                                   # longer than max_moves + 1 entries
    g.root_values = [0.0] * n
    g.action_history, g.reward_history, g.to_play_history = [0] * (n + 1), [0.0] * (n + 1), [0] * (n + 1)
    g.child_visits = [[1.0] + [0.0] * (len(cfg.action_space) - 1)] * n
    g.observation_history = [g.observation_history[0]] * (n + 1)
    with pytest.raises(MzbError):
        rb.save_game(g)


@pytest.mark.parametrize("kind", ["tictactoe", "cartpole"])
def test_ingest_from_export_ring_equals_host_path(kind):
    """Device-to-device hop: games exported by the self-play kernels and ingested straight into the store give the same
    batches as the same games decoded to host GameHistory objects and saved one by one."""
    import importlib
    from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
    from muzero_hypermodel_b200.self_play import SelfPlay, decode_export
    cfg = importlib.import_module(f"muzero_hypermodel_b200.games.{kind}").MuZeroConfig()
    if kind == "tictactoe":
        cfg.network = "fullyconnected"
    cfg.num_simulations = 10
    cfg.max_moves = min(cfg.max_moves, 25)
    cfg.replay_buffer_size, cfg.batch_size, cfg.PER = 64, 32, True
    z = T.load("net")
    pre = ("tictactoe_fc" if kind == "tictactoe" else "cartpole") + "/w/"
    sd = {k[len(pre):]: torch.tensor(z[k]) for k in z.files if k.startswith(pre)}
    histories = []
    buffers = []
    for mode in ("device", "host"):
        sp = SelfPlay({"weights": sd}, None, cfg, 5, n_games=16, device=DEV)
        env, _ = sp._setup()
        rb = ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg, device=DEV,
                          record_env=env if mode == "device" else None)
        for _ in range(cfg.max_moves + 1):
            sp.step(1.0, None)
            if mode == "device":
                rb.ingest(env)
            else:
                for gh in decode_export(env):
                    gh.priorities = None
                    rb.save_game(gh)
                    histories.append(gh)
        buffers.append(rb)
    dev_rb, host_rb = buffers
    assert len(dev_rb) == len(host_rb) > 8 and dev_rb.total_samples == host_rb.total_samples
    for _ in range(3):
        ia, ba = dev_rb.get_batch()
        ib, bb = host_rb.get_batch()
        assert torch.equal(ia, ib)
        for x, y in zip(ba, bb):
            assert torch.equal(x, y)


def test_continuous_self_play_feeds_the_device_store():
    """SelfPlay.continuous_self_play (self_play.py:30-108) with a plain-object shared storage: games flow from the
    self-play kernels into the device replay store and a batch can be drawn."""
    from muzero_hypermodel_b200.games.tictactoe import MuZeroConfig
    from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
    from muzero_hypermodel_b200.self_play import SelfPlay
    cfg = MuZeroConfig()
    cfg.network, cfg.num_simulations, cfg.replay_buffer_size, cfg.batch_size = "fullyconnected", 8, 256, 64
    cfg.self_play_delay, cfg.ratio = 0, None

    class Storage:
        def __init__(self):
            self.info = {"training_step": 0, "terminate": False, "weights": None, "num_played_games": 0, "num_played_steps": 0}

        def get_info(self, key):
            return self.info[key]

        def set_info(self, key, value=None):
            self.info[key] = value

    st = Storage()
    sp = SelfPlay({"weights": None}, None, cfg, 3, n_games=64, device=DEV)
    env, _ = sp._setup()
    rb = ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg, device=DEV, record_env=env)
    sp._env = env
    sp.continuous_self_play(st, rb, max_moves=12)
    assert len(rb) >= 64 and st.info["num_played_games"] == rb.num_played_games > 0
    assert st.info["num_played_steps"] == rb.num_played_steps == rb.total_samples
    index, (obs, act, val, rew, pol, w, gs) = rb.get_batch()
    assert obs.shape == (64, 3, 3, 3) and pol.shape == (64, cfg.num_unroll_steps + 1, 9)
    assert float(w.max()) == 1.0 and float(w.min()) > 0
    assert torch.all((obs[:, 2] == 1) | (obs[:, 2] == -1))
    np.testing.assert_allclose(pol.sum(-1).cpu().numpy(), 1.0, rtol=0, atol=1e-12)


def test_get_buffer_round_trip():
    """get_buffer() -> host GameHistory dict (what muzero.py pickles) -> a new store built from it as `initial_buffer`
    draws the same batches: counts, float32 priorities and observations survive the round trip exactly."""
    from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
    rb, cfg = make(1)
    for gi in range(5):
        rb.save_game(load_game(1, gi))
    buf = rb.get_buffer()
    assert sorted(buf) == [0, 1, 2, 3, 4] and all(g.priorities is not None for g in buf.values())
    rb2 = ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, buf, cfg, device=DEV)
    assert len(rb2) == 5 and rb2.total_samples == rb.total_samples
    for gid in range(5):
        np.testing.assert_array_equal(rb.game_priorities(gid)[0], rb2.game_priorities(gid)[0])
    ia, ba = rb.get_batch()
    ib, bb = rb2.get_batch()
    assert torch.equal(ia, ib)
    for x, y in zip(ba, bb):
        assert torch.equal(x, y)


def test_reanalyse_replaces_bootstrap_values():
    """Reanalyse (replay_buffer.py:298-361): a stored game goes through one batched initial_inference on the CUDA
    kernels, its value predictions become reanalysed_predicted_root_values, and later batches bootstrap from them -
    targets equal the oracle's make_target fed the same values (float64 accumulation)."""
    from muzero_hypermodel_b200 import models
    from muzero_hypermodel_b200.games.cartpole import MuZeroConfig
    from muzero_hypermodel_b200.replay_buffer import Reanalyse, ReplayBuffer
    from oracle import targets as otargets
    cfg = MuZeroConfig()
    cfg.PER, cfg.batch_size, cfg.replay_buffer_size, cfg.td_steps, cfg.max_moves = False, 16, 8, 5, 100
    z = T.load("net")
    pre = "cartpole_shipped/w/"
    sd = {k[len(pre):]: torch.tensor(z[k]) for k in z.files if k.startswith(pre)}
    rb = ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg, device=DEV)
    games = []
    for gi in range(3):
        g = load_game(0, gi)
        rb.save_game(g)
        games.append(g)
    ra = Reanalyse({"weights": sd, "num_reanalysed_games": 0}, cfg, device=DEV)
    ra.reanalyse_game(rb, 1)
    assert ra.num_reanalysed_games == 1
    obs = torch.tensor(np.array(games[1].observation_history[:len(games[1].root_values)]), device=DEV)
    assert torch.equal(rb.game_observations(1), obs)
    net = models.MuZeroNetwork(cfg); net.set_weights(sd); net.to(DEV)
    fresh = models.support_to_scalar(net.initial_inference(obs)[0], cfg.support_size).reshape(-1).cpu().numpy()
    assert np.abs(fresh - np.array(games[1].root_values)).max() > 1e-3      # the network disagrees with the stored values
    ug = [(1 + 0.5) / 3] * 16                                                # always game 1 (uniform draw floor(u * 3))
    up = [(i + 0.5) / 16 for i in range(16)]
    index, (_, act, val, rew, pol, w, gs) = rb.get_batch(u_game=ug, u_pos=up)
    g = games[1]
    for b in range(16):
        pos = int(index[b, 1])
        ev, er, ep, ea = otargets.make_target(g.root_values, [float(np.float32(x)) for x in g.reward_history], g.to_play_history,
                                              g.child_visits, g.action_history, pos, cfg.num_unroll_steps, cfg.td_steps,
                                              cfg.discount, 2, reanalysed_root_values=[float(x) for x in fresh],
                                              pad_action=lambda row: int(act[b, row]))
        assert np.array(ev, dtype=np.float64).tobytes() == val[b].cpu().numpy().tobytes(), (b, pos)


def test_one_call_with_more_games_than_slots_equals_sequential_saves():
    """save_games_device with 9 games into 5 slots in ONE call = 9 sequential save_game calls (FIFO eviction inside the
    call, no two games racing for a slot): same counters, same priorities, same batches."""
    rb_seq, cfg = make(1)
    rb_one, _ = make(1)
    games = [load_game(1, gi) for gi in range(9)]
    for g in games:
        rb_seq.save_game(g)
    A = len(cfg.action_space)
    obs, act, rew, tp, rv, vis, start, length = [], [], [], [], [], [], [], []
    for g in games:
        n = len(g.root_values)
        start.append(sum(len(x) for x in act)); length.append(n)
        obs.append(np.stack([np.asarray(o, dtype=np.float32).reshape(-1) for o in g.observation_history]))
        act.append(np.array(g.action_history, dtype=np.int32)); rew.append(np.array(g.reward_history, dtype=np.float32))
        tp.append(np.array(g.to_play_history, dtype=np.int8)); rv.append(np.concatenate([g.root_values, [0.0]]))
        counts = np.rint(np.array(g.child_visits) * cfg.num_simulations).astype(np.uint16)
        vis.append(np.concatenate([counts, np.zeros((1, A), dtype=np.uint16)]))
    dev = lambda x, dt: torch.as_tensor(np.concatenate(x)).to(DEV, dt).contiguous()
    rb_one.save_games_device(start, length, dev(obs, torch.float32), dev(act, torch.int32), dev(rew, torch.float32),
                             dev(tp, torch.int8), dev(rv, torch.float64), torch.from_numpy(np.concatenate(vis)).to(DEV))
    assert rb_one._info() == rb_seq._info() and len(rb_one) == 5
    first = rb_one._info()[4]
    for gid in range(first, first + 5):
        a, b = rb_one.game_priorities(gid), rb_seq.game_priorities(gid)
        assert a[0].tobytes() == b[0].tobytes() and np.float32(a[1]).tobytes() == np.float32(b[1]).tobytes()
    ia, ba = rb_one.get_batch()
    ib, bb = rb_seq.get_batch()
    assert torch.equal(ia, ib) and all(torch.equal(x, y) for x, y in zip(ba, bb))


def test_ingest_into_store_with_other_record_layout_goes_through_host_format():
    """A store built with the reference's 3-argument constructor keeps decoded observations (27 floats for tictactoe)
    while the environment's export ring holds packed records (3 floats): the C entry refuses the copy, `ingest` takes
    the host GameHistory route, and both stores then hold the same games."""
    import ctypes as C
    from muzero_hypermodel_b200 import _lib, models
    from muzero_hypermodel_b200.games.tictactoe import MuZeroConfig
    from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
    from muzero_hypermodel_b200.self_play import SelfPlay
    cfg = MuZeroConfig()
    cfg.network, cfg.num_simulations, cfg.replay_buffer_size = "fullyconnected", 8, 512
    ck = {"num_played_games": 0, "num_played_steps": 0}

    def play():
        torch.manual_seed(0)
        sp = SelfPlay({"weights": None}, None, cfg, 3, n_games=64, device=DEV)
        for _ in range(9):
            sp.step()
        return sp

    sp = play()
    plain = ReplayBuffer(ck, {}, cfg, device=DEV)                         # no record_env: decoded layout
    n = C.c_int32(-1)
    rc = _lib.lib.mzb_env_export_to_replay(sp._env._h, plain._h, C.byref(n), _lib.current_stream())
    assert rc != 0 and b"floats per observation record" in _lib.lib.mzb_last_error()
    assert len(plain) == 0                                                # nothing copied
    got = plain.ingest(sp._env)                                           # host-format fallback
    assert got >= 64 and len(plain) == got
    sp2 = play()
    packed = ReplayBuffer(ck, {}, cfg, device=DEV, record_env=sp2._env)
    assert packed.ingest(sp2._env) == got
    a, b = plain.get_buffer(), packed.get_buffer()
    assert sorted(a) == sorted(b)
    for k in a:
        assert a[k].action_history == b[k].action_history and a[k].root_values == b[k].root_values
        assert all(np.array_equal(x, y) for x, y in zip(a[k].observation_history, b[k].observation_history))
        np.testing.assert_array_equal(a[k].priorities, b[k].priorities)


def test_resume_keeps_checkpoint_counters_and_game_ids():
    """replay_buffer.py:17-24: the counters come from the checkpoint (re-loading the buffer does not play its games
    again) and the buffered games keep their ids; the next saved game continues the numbering."""
    from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
    rb, cfg = make(1)
    for gi in range(5):
        rb.save_game(load_game(1, gi))
    steps5 = rb.num_played_steps
    buf = {7 + k: g for k, g in rb.get_buffer().items()}                  # as pickled by a run that had played 12 games
    ck = {"num_played_games": 12, "num_played_steps": 345}
    rb2 = ReplayBuffer(ck, buf, cfg, device=DEV)
    assert rb2.num_played_games == 12 and rb2.num_played_steps == 345 and len(rb2) == 5
    assert rb2.total_samples == steps5                                    # recomputed from the buffer (:22-24)
    assert sorted(rb2.get_buffer()) == [7, 8, 9, 10, 11]
    idx, batch = rb2.get_batch()
    assert int(idx[:, 0].min()) >= 7 and int(idx[:, 0].max()) <= 11
    rb2.update_priorities(torch.ones((idx.shape[0], cfg.num_unroll_steps + 1)), idx)   # public ids accepted back
    g = load_game(1, 5)
    rb2.save_game(g)
    assert rb2.num_played_games == 13 and rb2.num_played_steps == 345 + len(g.root_values)
    assert sorted(rb2.get_buffer())[-1] == 12
    np.testing.assert_array_equal(rb.game_priorities(2)[0], ReplayBuffer(ck, buf, cfg, device=DEV).game_priorities(9)[0])


def test_save_game_reports_to_plain_and_remote_storages():
    """save_game(game_history, shared_storage) works with plain get_info / set_info objects (Trainer, tests) and
    Ray-style handles (`.remote`)."""
    rb, cfg = make(0)

    class Plain:
        def __init__(self): self.info = {}
        def set_info(self, k, v): self.info[k] = v

    class Remote:
        def __init__(self):
            self.info = {}
            outer = self

            class _M:
                def remote(self, k, v): outer.info[k] = v
            self.set_info = _M()

    for st in (Plain(), Remote()):
        rb.save_game(load_game(0, 0), st)
        assert st.info["num_played_games"] == rb.num_played_games and st.info["num_played_steps"] == rb.num_played_steps


@pytest.mark.parametrize("kind,S", [("tictactoe", 2), ("cartpole", 3)])
def test_get_batch_stacks_observations_like_game_history(kind, S):
    """stacked_observations > 0: get_batch and Reanalyse's game_observations return
    GameHistory.get_stacked_observations(position, S) (self_play.py:514-548, pinned to the reference by
    tests/golden/stacked.npz) of the stored games, from the packed records and from decoded host games alike."""
    import importlib
    from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
    from muzero_hypermodel_b200.self_play import SelfPlay, decode_export
    cfg = importlib.import_module(f"muzero_hypermodel_b200.games.{kind}").MuZeroConfig()
    if kind == "tictactoe":
        cfg.network = "fullyconnected"
    cfg.num_simulations, cfg.stacked_observations = 6, S
    cfg.max_moves = min(cfg.max_moves, 12)
    cfg.replay_buffer_size, cfg.batch_size, cfg.PER = 64, 48, False
    torch.manual_seed(0)
    stores = []
    games = None
    for mode in ("packed", "host"):
        sp = SelfPlay({"weights": None}, None, cfg, 5, n_games=24, device=DEV)
        env, _ = sp._setup()
        rb = ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg, device=DEV,
                          record_env=env if mode == "packed" else None)
        got = []
        for _ in range(cfg.max_moves + 1):
            sp.step(1.0, None)
            if mode == "packed":
                rb.ingest(env)
            else:
                for gh in decode_export(env):
                    rb.save_game(gh)
                    got.append(gh)
        stores.append(rb)
        if mode == "host":
            games = got
    packed, host = stores
    assert len(packed) == len(host) == len(games) > 8
    for rb in stores:
        idx, (obs, *_rest) = rb.get_batch()
        idx, obs = idx.cpu().numpy(), obs.cpu().numpy()
        c, h, w = cfg.observation_shape
        assert obs.shape[1:] == (c * (S + 1) + S, h, w)
        for b in range(idx.shape[0]):
            gh = games[int(idx[b, 0])]
            want = np.asarray(gh.get_stacked_observations(int(idx[b, 1]), S), dtype=np.float32)
            assert np.array_equal(obs[b], want), (b, idx[b])
        g0 = rb.game_observations(3).cpu().numpy()
        want = np.stack([np.asarray(games[3].get_stacked_observations(i, S), dtype=np.float32) for i in range(len(games[3].root_values))])
        assert np.array_equal(g0, want)
