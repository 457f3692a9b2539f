"""CPU self-play baseline: the oracle port of SelfPlay.play_game, one game stream per process.

Test/bench infrastructure (only bench.py's cpu_baseline and `--impl reference` legs run it).  It mirrors
what each of the reference's Ray `SelfPlay` actors executes (self_play.py:31-108 minus the Ray RPC):
one process = one game at a time = one simulation at a time = batch-1 network calls.  The reference is
pure Python and cannot travel to the GPU box (/root/reference is absent there), so this is kind="port".
"""
import os
import time

import numpy as np

from . import games, mcts, networks, rng


def _make(kind, weights, cfg):
    A = len(cfg["action_space"])
    if cfg.get("network", "fullyconnected") == "resnet":
        import torch
        torch.set_num_threads(1)
        net = networks.Residual(weights, cfg["observation_shape"], A, cfg["blocks"], cfg["support_size"],
                                cfg.get("downsample", False))
    else:
        net = networks.FullyConnected(weights, A, cfg["support_size"])
    env = {"cartpole": lambda: games.CartPole(1, seed=cfg["seed"], slot0=cfg["slot"]),
           "tictactoe": lambda: games.TicTacToe(1), "connect4": lambda: games.Connect4(1),
           "gomoku": lambda: games.Gomoku(1),
           "breakout": lambda: games.SyntheticFrames(1, seed=cfg["seed"], slot0=cfg["slot"])}[kind]()
    return net, env, A


def play_for(kind, weights, cfg, max_searches=None, max_seconds=None):
    """Plays games back to back until a budget runs out; returns (simulations, env_steps, seconds)."""
    net, env, A = _make(kind, weights, cfg)
    S = cfg["support_size"]
    rs = np.random.RandomState(cfg["seed"] + cfg["slot"])
    sims = steps = 0
    t0 = time.perf_counter()

    def decode(v, r, p, legal):
        pri = mcts.softmax_f32(p[0][legal])
        return (float(networks.support_to_scalar(v, S)[0, 0]), float(networks.support_to_scalar(r, S)[0, 0]),
                [float(x) for x in pri])

    def recurrent(hidden, action):
        v, r, p, s = net.recurrent_inference(hidden, np.array([action]))
        val, rew, pri = decode(v, r, p, slice(None))
        return val, rew, pri, s

    done_budget = False
    while not done_budget:
        obs = env.reset()
        moves, done = 0, False
        while not done and moves < cfg["max_moves"]:
            legal = env.legal_lists()[0] if hasattr(env, "legal_lists") else list(range(A))
            with np.errstate(divide="ignore", invalid="ignore"):
                v, r, p, s = net.initial_inference(obs)
                val, rew, pri = decode(v, r, p, legal)
            noise = rs.dirichlet([cfg["root_dirichlet_alpha"]] * len(legal))
            res = mcts.search(recurrent, (val, rew, pri, s), legal, int(env.to_play()[0]), n_actions=A,
                              n_players=cfg["n_players"], num_simulations=cfg["num_simulations"],
                              discount=cfg["discount"], pb_c_base=cfg["pb_c_base"], pb_c_init=cfg["pb_c_init"],
                              noise=[float(x) for x in noise], exploration_fraction=cfg["root_exploration_fraction"],
                              tie=lambda n, sim, depth: int(rs.randint(n)))
            action = mcts.select_action(res.root_actions, res.visits, 1.0, float(rs.random_sample()))
            obs, _, d = env.step(np.array([action]))
            done = bool(d[0])
            moves += 1
            steps += 1
            sims += cfg["num_simulations"]
            if (max_searches and steps >= max_searches) or (max_seconds and time.perf_counter() - t0 >= max_seconds):
                done_budget = True
                break
    return sims, steps, time.perf_counter() - t0


def _worker(args):
    kind, weights, cfg, max_searches, max_seconds = args
    os.environ["OMP_NUM_THREADS"] = "1"
    return play_for(kind, weights, cfg, max_searches, max_seconds)


def run_parallel(kind, weights, cfg, n_procs, max_searches=None, max_seconds=None, pool=None):
    """n_procs independent game streams; returns (total simulations, total env steps, wall seconds)."""
    import multiprocessing as mp
    jobs = [(kind, weights, dict(cfg, slot=i), max_searches, max_seconds) for i in range(n_procs)]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(n_procs)
    t0 = time.perf_counter()
    res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    if own:
        pool.close()
        pool.join()
    return sum(r[0] for r in res), sum(r[1] for r in res), wall
