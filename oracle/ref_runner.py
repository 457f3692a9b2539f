"""Time the UNMODIFIED reference `SelfPlay.play_game` (self_play.py:110-184) on the host cores.

Test/bench infrastructure (bench.py's `cpu_baseline` leg and `--impl reference` only).  The reference classes come
from `oracle/_ref/reference_pyc.zip` (oracle/build_ref.py: byte-code compiled from /root/reference, zipimport), under the
two stand-ins of oracle/ref_loader.py: a `ray` stub whose `remote` is the identity decorator and a `gym` shim backed by
oracle.games.CartPoleV1.  It is what each of the reference's Ray `SelfPlay` actors executes minus the Ray RPC at game
boundaries: P independent processes, `torch.set_num_threads(1)` each, one game at a time, batch-1
network calls through the reference's own torch modules.

A "sample" is one `play_game(temperature=1, None, False, "self", 0)` call with `config.max_moves` lowered so that the
call ends after a bounded number of searches (the loop condition is the reference's own, :129-131); simulations are
counted from the returned GameHistory: len(root_values) * num_simulations.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "_ref", "reference_pyc.zip")
GAMES = ("cartpole", "tictactoe", "connect4", "gomoku")          # breakout needs ALE + cv2: not runnable offline


def available(workload=None):
    return os.path.isfile(ARCHIVE) and (workload is None or workload in GAMES)


def _load(game):
    """Reference modules from the archive (never from /root/reference: that path does not exist on the GPU box)."""
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import ref_loader
    ref_loader._install_ray_stub()
    ref_loader._install_gym_shim()
    if ARCHIVE not in sys.path:
        sys.path.insert(0, ARCHIVE)
    import importlib
    sp = importlib.import_module("self_play")
    gm = importlib.import_module(f"games.{game}")
    assert ARCHIVE in sp.__file__ and ARCHIVE in gm.__file__, (sp.__file__, gm.__file__)
    return sp, gm


_state = {}


def _worker(args):
    game, weights, overrides, seed, searches_cap, n_calls = args
    import torch
    torch.set_num_threads(1)
    # the CPU arm: the reference picks "cuda" whenever torch sees a GPU (self_play.py:27); a forked worker of a process
    # that has touched the CUDA driver cannot initialise it anyway - keep the reference on the host cores
    torch.cuda.is_available = lambda: False
    key = (game, tuple(sorted(overrides.items())))
    if _state.get("key") != key:
        sp, gm = _load(game)
        cfg = gm.MuZeroConfig()
        for k, v in overrides.items():
            setattr(cfg, k, v)
        cfg.max_moves = min(cfg.max_moves, searches_cap)
        cfg.selfplay_on_gpu = False
        w = {k: torch.as_tensor(v) for k, v in weights.items()}
        _state.update(key=key, cfg=cfg, worker=sp.SelfPlay({"weights": w}, gm.Game, cfg, seed))
    cfg, worker = _state["cfg"], _state["worker"]
    sims = steps = 0
    t0 = time.perf_counter()
    for _ in range(n_calls):
        gh = worker.play_game(1, None, False, "self", 0)
        steps += len(gh.root_values)
        sims += len(gh.root_values) * cfg.num_simulations
    return sims, steps, time.perf_counter() - t0


def run_parallel(game, weights, overrides, n_procs, searches_cap, n_calls=1, pool=None, seed0=0):
    """n_procs processes x n_calls play_game calls of <= searches_cap searches; returns (simulations, env steps, wall s)."""
    import multiprocessing as mp
    jobs = [(game, weights, overrides, seed0 + i, searches_cap, n_calls) for i in range(n_procs)]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(n_procs)
    t0 = time.perf_counter()
    res = pool.map(_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    if own:
        pool.close()
        pool.join()
    return sum(r[0] for r in res), sum(r[1] for r in res), wall
