#!/usr/bin/env python
"""bench.py - self-play hot path throughput: MCTS simulations/s (and env steps/s) per BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|cartpole|tictactoe|...]
                    [--games G]

Default (`--workload all`): the four other BASELINE configs run first and are nested under `workloads`
(tictactoe-FC at 4,096 games/GPU, connect4 at 16,384 games/GPU, gomoku at 400 simulations, breakout on synthetic
96x96 frames - each with its own value / e2e / roofline / cpu_baseline / parity / clocks), then the HEADLINE workload,
BASELINE.json configs[0] "cartpole FC MuZero (games/cartpole.py defaults, num_simulations=50)" - the configuration the
>= 1e8 simulations/s on 8xB200 target is quoted on - whose keys form the top level of the ONE JSON line printed last.

One self-play MOVE = observe -> MCTS.run (num_simulations simulations, network in the loop) -> select_action ->
Game.step -> GameHistory append -> harvest/auto-reset of every game on every GPU, with the finished games ingested by
the device replay store every 4th move.  One bench STEP = `config.moves_per_step` consecutive moves (cartpole: 25, so
that the driver's 20 steps time > 2 s of self-play and whole episodes, resets and replay ingests fall inside the timed
region; the residual workloads: 1).  `value` = simulations/s over all GPUs with state resident in HBM; `e2e` = the
same searches through the batched MCTS.run entry point with HOST observation / legal-action / to-play buffers (pinned
H2D before, D2H of visit counts + root values after, every step); `roofline` = the dominant kernel: `frac` alone (timed
live with CUDA events around the launch) and `in_step` (the same launches timed inside the timed region); `cpu_baseline`
= the UNMODIFIED reference SelfPlay.play_game (oracle/_ref byte-code archive, kind "reference") on the host cores, with
the oracle port beside it; `parity` = measured agreement of the benchmarked arithmetic with the exact path.
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G_FULL_WAVES = 8 * 148 * 512        # cartpole: 8 full waves of the whole-search kernel's 148 SMs x 2 CTAs x 256 games (ramp-up and
                                    # tail of a launch amortised: 10.86 ns per search against 11.02 at 4 waves, 11.1 at 3)
WORKLOADS = {
    # name: (golden weight tag, config module, games per GPU, (initial FLOP, recurrent FLOP) per BASELINE.md §3,
    #        moves per bench step, default timed steps)
    "cartpole": ("cartpole_shipped", "cartpole", G_FULL_WAVES, (1312, 2752), 25, 20),
    "tictactoe": ("tictactoe_fc", "tictactoe", 4096, (3648, 5952), 9, 20),       # BASELINE configs[1]: 4096 games/GPU
    "connect4": ("connect4", "connect4", 16384, (37372160, 40396160), 1, 3),
    "gomoku": ("gomoku", "gomoku", 4096, (857557760, 892780160), 1, 2),          # 51 GB bf16 hidden-state pool
    "breakout": ("breakout", "breakout", 16384, (34192160, 1532480), 1, 5),
}
NESTED = ("tictactoe", "connect4", "gomoku", "breakout")
# committed `ncu --set full` captures of the dominant kernel at the bench shape (profiles/): DRAM bytes per launch
NCU_CAPTURE = {"cartpole": "r03_ncu_k_search_fc_cartpole.csv", "connect4": "r03_ncu_k_conv_tc_connect4.csv",
               "gomoku": "r03_ncu_k_conv_tc_gomoku.csv", "breakout": "r04_ncu_k_recurrent16.csv"}
# column of the capture that holds the probed layer (a plain tower layer, no residual input): the connect4 capture is of
# four consecutive launches (residual, plain, residual, plain)
NCU_COLUMN = {"connect4": 1}


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed capture, or None."""
    name = NCU_CAPTURE.get(workload)
    if not name:
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, 0
    try:
        for ln in open(os.path.join(ROOT, "profiles", name)):
            c = ln.strip().split(",")
            col = 2 + NCU_COLUMN.get(workload, 0)
            if len(c) > col and c[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and c[1] in unit:
                total += float(c[col]) * unit[c[1]]
                seen += 1
    except (OSError, ValueError):
        return None
    return total if seen == 2 else None


def load_weights(tag):
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "net.npz"))
    if tag == "gomoku":                       # 22 MB of fp32: seeded values, regenerated (tests/_weights.py)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _weights import seeded_state_dict
        keys = str(z["gomoku/keys"]).split("\n")
        shapes = [[int(d) for d in s.split("x")] if s else [] for s in z["gomoku/shapes"]]
        return seeded_state_dict(keys, shapes)
    pre = tag + "/w/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def make_config(workload):
    import importlib
    mod = importlib.import_module(f"muzero_hypermodel_b200.games.{WORKLOADS[workload][1]}")
    cfg = mod.MuZeroConfig()
    if workload == "tictactoe":
        cfg.network = "fullyconnected"
    if cfg.network == "resnet":
        cfg.resnet_precision = "bf16"          # tcgen05 path: bf16 operands, fp32 accumulation
    return cfg


def oracle_cfg(cfg):
    return dict(action_space=list(cfg.action_space), support_size=cfg.support_size, seed=cfg.seed, slot=0,
                max_moves=cfg.max_moves, n_players=len(cfg.players), num_simulations=cfg.num_simulations,
                discount=cfg.discount, pb_c_base=cfg.pb_c_base, pb_c_init=cfg.pb_c_init,
                root_dirichlet_alpha=cfg.root_dirichlet_alpha, root_exploration_fraction=cfg.root_exploration_fraction,
                network=cfg.network, observation_shape=tuple(cfg.observation_shape), blocks=cfg.blocks,
                downsample=cfg.downsample)


# searches per reference play_game call (config.max_moves cap): ~1-2 s of one core per call
REF_SEARCH_CAP = {"cartpole": 24, "tictactoe": 9, "connect4": 2, "gomoku": 1}
REF_OVERRIDES = {"tictactoe": {"network": "fullyconnected"}}


def reference_leg(workload, cfg, calls=1, pool=None, cores=None, warm=True):
    """The UNMODIFIED reference SelfPlay.play_game on every host core (oracle/ref_runner.py), or None if the byte-code
    archive is missing / the game cannot run offline (breakout: ALE).  warm: one untimed call per process first
    (imports torch, builds the reference model) so that start-up does not count against the CPU side."""
    import multiprocessing as mp
    from oracle import ref_runner
    if not ref_runner.available(WORKLOADS[workload][1]):
        return None
    cores = cores or os.cpu_count() or 1
    cap = REF_SEARCH_CAP[workload]
    w = load_weights(WORKLOADS[workload][0])
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores)
    game, over = WORKLOADS[workload][1], REF_OVERRIDES.get(workload, {})
    if warm:
        ref_runner.run_parallel(game, w, over, cores, cap, n_calls=1, pool=pool, seed0=cfg.seed)
    sims, steps, wall = ref_runner.run_parallel(game, w, over, cores, cap, n_calls=calls, pool=pool, seed0=cfg.seed)
    if own:
        pool.close()
        pool.join()
    return {"value": sims / wall, "unit": "simulations/s", "cores": cores, "kind": "reference",
            "env_steps_per_s": steps / wall, "wall_s": wall, "simulations": sims,
            "sample": f"{cores} processes (torch.set_num_threads(1) each) x {calls} call(s) of the unmodified reference "
                      f"SelfPlay.play_game capped at max_moves={cap} ({steps} searches of {cfg.num_simulations} "
                      f"simulations, batch-1 torch modules, same weights/config; Ray RPC omitted; one untimed warm-up call per process)"}


def port_leg(workload, cfg, seconds):
    """Oracle port of play_game on every host core for ~`seconds`."""
    from oracle import cpu_baseline
    cores = os.cpu_count() or 1
    w = load_weights(WORKLOADS[workload][0])
    sims, steps, wall = cpu_baseline.run_parallel(WORKLOADS[workload][1], w, oracle_cfg(cfg), cores, max_seconds=seconds)
    return {"value": sims / wall, "unit": "simulations/s", "cores": cores, "kind": "port",
            "env_steps_per_s": steps / wall,
            "sample": f"{cores} processes x {seconds:.0f} s of oracle play_game ({steps} searches of "
                      f"{cfg.num_simulations} simulations, batch-1 numpy network, same weights/config)"}


def cpu_legs(workload, cfg, seconds):
    """cpu_baseline = the reference when it can run here, with the (faster) numpy port kept beside it."""
    port = port_leg(workload, cfg, min(seconds, 6.0))
    ref = reference_leg(workload, cfg, calls={"cartpole": 4, "tictactoe": 8}.get(workload, 1))
    if ref is None:
        return port, None
    return ref, port


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores - the UNMODIFIED
    SelfPlay.play_game from the byte-code archive (kind "reference"), else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    workload = "cartpole" if args.workload == "all" else args.workload
    cfg = make_config(workload)
    cores = os.cpu_count() or 1
    from oracle import ref_runner
    use_ref = ref_runner.available(WORKLOADS[workload][1])
    pool = mp.get_context("fork").Pool(cores)
    times, sims_total, sample = [], 0, ""
    if use_ref:
        for i in range(args.warmup + args.steps):
            leg = reference_leg(workload, cfg, calls=1, pool=pool, cores=cores, warm=False)
            if i >= args.warmup:
                times.append(leg["wall_s"]); sims_total += leg["simulations"]
            sample = "each step = " + leg["sample"]
    else:
        from oracle import cpu_baseline
        w = load_weights(WORKLOADS[workload][0])
        ocfg = oracle_cfg(cfg)
        searches = 24 if cfg.network == "fullyconnected" else 1
        for i in range(args.warmup + args.steps):
            sims, steps, wall = cpu_baseline.run_parallel(WORKLOADS[workload][1], w, dict(ocfg, seed=ocfg["seed"] + i), cores,
                                                          max_searches=searches, pool=pool)
            if i >= args.warmup:
                times.append(wall); sims_total += sims
        sample = (f"each step = {cores} processes x {searches} searches x {cfg.num_simulations} simulations of the oracle "
                  f"port of SelfPlay.play_game (batch-1 numpy network)")
    pool.close()
    pool.join()
    total = sum(times)
    value = sims_total / total
    line = {"impl": "reference", "metric": "mcts_simulations_per_sec", "value": value, "unit": "simulations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": {"workload": workload_name(workload, cfg), "games_per_gpu": cores,
                       "num_simulations": cfg.num_simulations},
            "cpu_baseline": {"value": value, "unit": "simulations/s", "cores": cores, "kind": "reference" if use_ref else "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "simulations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(workload, cfg):
    return {"cartpole": "cartpole FC MuZero (games/cartpole.py defaults, num_simulations=50)",
            "tictactoe": "tictactoe FC MuZero, two-player, network=fullyconnected, 4096 parallel games/GPU, num_simulations=25",
            "connect4": "connect4 residual-network MuZero (games/connect4.py defaults, 3 blocks x 64 ch, num_simulations=200), 16384 parallel games/GPU",
            "gomoku": "gomoku residual MuZero (games/gomoku.py defaults, 6 blocks x 128 ch, A=121, num_simulations=400)",
            "breakout": "Atari Breakout residual MuZero on synthetic 96x96 frames (games/breakout.py defaults, num_simulations=30)"}[workload]


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def count(self):
        """Samples written so far."""
        try:
            with open(self.f.name) as g:
                return sum(1 for ln in g if ln.count(",") >= 8)
        except OSError:
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def fc_parity(workload, cfg, sp, mcts, n_games=48):
    """Network in the loop: the whole-search kernel against the oracle MCTS driven by the numpy restatement of the
    reference network (oracle as the CHECKER).  Tree arithmetic is bit-exact given equal network outputs; float32
    decode conditioning makes some searches diverge at a near-tie - this reports how many."""
    import numpy as np
    import torch
    from muzero_hypermodel_b200.parity import visit_agreement
    from muzero_hypermodel_b200.search import BatchedMCTS
    from oracle import mcts as omcts, networks as onet, rng
    dev, A, S = sp.device, len(cfg.action_space), cfg.support_size
    rs = np.random.RandomState(9)
    if workload == "cartpole":
        obs = rs.uniform(-0.2, 0.2, (n_games, 1, 1, 4)).astype(np.float32)
        legal = np.ones((n_games, A), dtype=bool)
        to_play = np.zeros(n_games, dtype=np.int8)
    else:
        stones = rs.randint(-1, 2, (n_games, 3, 3))
        tp = rs.choice([-1, 1], size=(n_games, 1, 1))
        obs = np.stack([(stones == 1), (stones == -1), np.broadcast_to(tp, stones.shape)], axis=1).astype(np.float32)
        legal = stones.reshape(n_games, -1) == 0
        legal[np.arange(n_games), rs.randint(A, size=n_games)] |= ~legal.any(1)
        to_play = (tp.reshape(-1) == -1).astype(np.int8)
    noise = np.zeros((n_games, A))
    for g in range(n_games):
        noise[g, legal[g]] = rs.dirichlet([cfg.root_dirichlet_alpha] * int(legal[g].sum()))
    slot = rs.randint(1 << 20, size=n_games).astype(np.int32)
    step = rs.randint(400, size=n_games).astype(np.int32)
    seed = 1234
    eng = BatchedMCTS(cfg, n_games, device=dev, seed=seed)
    out = eng.run(sp.model, torch.tensor(obs, device=dev), torch.tensor(legal, device=dev), torch.tensor(to_play, device=dev),
                  True, noise=torch.tensor(noise, device=dev), slot=torch.tensor(slot, device=dev),
                  step=torch.tensor(step, device=dev))
    onn = onet.FullyConnected({k: v.numpy() for k, v in sp.model.get_weights().items()}, A, S)
    ov, orv = [], []
    for g in range(n_games):
        la = np.nonzero(legal[g])[0].tolist()
        with np.errstate(divide="ignore", invalid="ignore"):
            v, r, p, s = onn.initial_inference(obs[g:g + 1])
        root = (float(onet.support_to_scalar(v, S)[0, 0]), float(onet.support_to_scalar(r, S)[0, 0]),
                [float(x) for x in omcts.softmax_f32(p[0][la])], s)

        def rec(hidden, action):
            v, r, p, s = onn.recurrent_inference(hidden, np.array([action]))
            return (float(onet.support_to_scalar(v, S)[0, 0]), float(onet.support_to_scalar(r, S)[0, 0]),
                    [float(x) for x in omcts.softmax_f32(p[0])], s)

        res = omcts.search(rec, root, la, int(to_play[g]), n_actions=A, n_players=len(cfg.players),
                           num_simulations=cfg.num_simulations, discount=cfg.discount, pb_c_base=cfg.pb_c_base,
                           pb_c_init=cfg.pb_c_init, noise=[float(noise[g, a]) for a in la],
                           exploration_fraction=cfg.root_exploration_fraction,
                           tie=lambda n, sim, depth: rng.tie_index(seed, int(slot[g]), int(step[g]), sim, depth, n))
        vv = np.zeros(A, dtype=np.int64)
        vv[la] = res.visits
        ov.append(vv); orv.append(res.root_value())
    m = visit_agreement(out["visits"], np.array(ov), out["root_value"], np.array(orv))
    m["against"] = "oracle MCTS + numpy restatement of the reference network (network in the loop), same noise and tie-break draws"
    return m


def resnet_parity(workload, cfg, weights, dev, n_games):
    """bf16 tcgen05 search against the fp32 search over the same roots, injected noise and tie-break counters."""
    import numpy as np
    import torch
    from muzero_hypermodel_b200 import models
    from muzero_hypermodel_b200.parity import visit_agreement
    from muzero_hypermodel_b200.search import BatchedMCTS
    from muzero_hypermodel_b200.envs import VectorEnv, game_kind
    A = len(cfg.action_space)
    env = VectorEnv(game_kind(cfg), n_games, cfg.max_moves, seed=99, device=dev)
    # a few random opening moves so the roots are not all the empty board
    rs = np.random.RandomState(4)
    for _ in range(4 if len(cfg.players) == 2 else 1):
        obs, legal, to_play = env.observe()
        lg = legal.cpu().numpy().astype(bool)
        act = np.array([rs.choice(np.nonzero(lg[g])[0]) for g in range(n_games)], dtype=np.int32)
        env.act_step(None, None, forced_action=torch.tensor(act, device=dev))
        env.harvest(False)
    obs, legal, to_play = env.observe()
    lg = legal.cpu().numpy().astype(bool)
    noise = np.zeros((n_games, A))
    for g in range(n_games):
        noise[g, lg[g]] = rs.dirichlet([cfg.root_dirichlet_alpha] * int(lg[g].sum()))
    nz = torch.tensor(noise, device=dev)
    res = {}
    for prec in ("fp32", "bf16"):
        net = models.MuZeroNetwork(cfg)
        net.set_weights(weights)
        net.set_precision(prec)
        net.to(dev).eval()
        eng = BatchedMCTS(cfg, n_games, device=dev, seed=77)
        o = eng.run(net, obs, legal, to_play, True, noise=nz, slot=env.slot, step=env.step_count)
        torch.cuda.synchronize()
        res[prec] = {k: v.clone() for k, v in o.items()}
        del eng, net
    m = visit_agreement(res["bf16"]["visits"], res["fp32"]["visits"], res["bf16"]["root_value"], res["fp32"]["root_value"])
    m["against"] = "the fp32 CUDA-core path of the same library (<= 2e-4 of the reference network), same roots / noise / tie-breaks"
    m["num_simulations"] = int(cfg.num_simulations)
    return m


def run_workload(args, workload, steps, warmup, cpu_seconds, want_cpu, want_collectives):
    """One BASELINE config on this rank's GPU (all ranks in lock-step); returns the result dict on rank 0."""
    import torch
    from muzero_hypermodel_b200 import _lib
    from muzero_hypermodel_b200 import dist as mdist
    from muzero_hypermodel_b200.self_play import SelfPlay

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = make_config(workload)
    dev = torch.device("cuda", local_rank)
    barrier = mdist.barrier

    cpu, cpu_port = (None, None)
    if want_cpu:
        cpu, cpu_port = cpu_legs(workload, cfg, cpu_seconds)

    tag, _, G_default, flops, moves_per_step, _ = WORKLOADS[workload]
    G = args.games or G_default
    S = cfg.num_simulations
    weights = {k: torch.tensor(v) for k, v in load_weights(tag).items()}
    sp = SelfPlay({"weights": weights}, None, cfg, cfg.seed, n_games=G, device=dev, first_slot=mdist.first_slot(rank, G))
    env, mcts = sp._setup()
    is_fc = cfg.network == "fullyconnected"
    fused = is_fc and bool(_lib.lib.mzb_search_fc_is_fused(sp.model.handle())) and not args.modular

    # the games self-play finishes are consumed like in the real pipeline: every 4th move the export ring is ingested,
    # device to device, by the replay store (one 8-byte D2H + the per-game table per ingest)
    from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
    try:
        rb = ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg, device=dev, record_env=env)
    except NotImplementedError:                    # synthetic frames are regenerated, not stored
        rb = None
    moves = [0]
    search_events = []                             # (start, end) CUDA events around the search launch of timed moves

    def move(timed=False):
        if timed and is_fc and len(search_events) < 512:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sp.step(temperature=1.0, temperature_threshold=None, add_exploration_noise=True, export=True,
                    allow_fused=not args.modular, search_events=(a, b))
            search_events.append((a, b))
        else:
            sp.step(temperature=1.0, temperature_threshold=None, add_exploration_noise=True, export=True,
                    allow_fused=not args.modular)
        moves[0] += 1
        if rb is not None and moves[0] % 4 == 0:
            rb.ingest(env)

    def step(timed=False):
        for _ in range(moves_per_step):
            move(timed)

    for _ in range(warmup):
        step()
    if rb is not None:
        rb.ingest(env)
    else:
        sp.drain()
    ingested0 = rb.num_played_games if rb is not None else 0
    mcts.tree.counters(reset=True)
    barrier()

    # ---------------- timed region: K steps (K x moves_per_step whole self-play moves), CUDA events on the launching stream
    sampler = ClockSampler(local_rank) if rank == 0 else None
    c0 = env.counters()
    _lib.lib.mzb_reset_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev[0].record()
    for _ in range(steps):
        step(timed=True)
    ev[1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[1])
    launches = int(_lib.lib.mzb_launch_count())
    c1 = env.counters()
    tc = mcts.tree.counters()
    ingested_end = rb.num_played_games if rb is not None else 0
    # a timed region shorter than a few 50 ms sampling periods would leave the clocks line empty: keep the same load
    # running, untimed, until the sampler has seen it
    extra = 0
    if sampler is not None and sampler.p is not None:
        t_end = time.time() + 3.0
        while sampler.count() < 3 and time.time() < t_end:
            move()
            torch.cuda.synchronize()
            extra += 1
    clocks = sampler.stop() if sampler else None
    if clocks is not None and extra:
        clocks["untimed_load_moves_for_sampling"] = extra
    ms = mdist.max_over_ranks(ms, dev)
    n_moves = steps * moves_per_step
    sims_total = n_moves * G * S * world
    value = sims_total / (ms * 1e-3)
    env_steps = n_moves * G * world / (ms * 1e-3)
    mean_path = tc["path_length_sum"] / max(1, tc["simulations"])
    in_step_ms = None
    if search_events:
        in_step_ms = sum(a.elapsed_time(b) for a, b in search_events) / len(search_events)
    if rb is not None:
        rb.ingest(env)
        ingested = ingested_end - ingested0
    else:
        ingested = len(sp.drain())

    # ---------------- the dominant kernel alone: search launches timed with events (same inputs every launch;
    # the tree store it streams through is G*(S+1)*(24A + 4H) bytes >> L2, no L2 flush needed)
    obs, legal, to_play = env.observe()
    kt = []
    for i in range(3 + 5 if is_fc else 1 + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mcts.run(sp.model, obs, legal, to_play, True, slot=env.slot, step=env.step_count, allow_fused=not args.modular,
                 out=sp._out)
        b.record()
        torch.cuda.synchronize()
        if i >= (3 if is_fc else 1):
            kt.append(a.elapsed_time(b))
    k_ms = sum(kt) / len(kt)
    A = len(cfg.action_space)
    H = cfg.encoding_size if is_fc else int(sp.model.latent_shape[0] * sp.model.latent_shape[1] * sp.model.latent_shape[2])
    hs = 4 if is_fc else 2                           # bytes per hidden-state element (bf16 pool on the resnet path)
    L = mean_path + 1.0                              # nodes on the path incl. root (SURVEY.md §8d counts nodes)
    bytes_per_sim = (L - 1) * A * 20 + L * 24 + 8 * A + 8 + 2 * H * hs
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_per_sim * G * S / (k_ms * 1e-3) / 1e9
    tflops = (flops[1] * S + flops[0]) * G / (k_ms * 1e-3) / 1e12
    parity = None
    if is_fc:
        in_step = None
        if in_step_ms:
            a_in = bytes_per_sim * G * S / (in_step_ms * 1e-3) / 1e9
            in_step = {"achieved": a_in, "frac": a_in / peak, "kernel_ms": in_step_ms, "launches_timed": len(search_events),
                       "note": "the same launch timed with CUDA events inside the timed region (after the env kernels of the move)"}
        roofline = {"bound": "hbm", "kernel": "k_search_fc (whole-search, fused)" if fused else "modular: k_select+k_fc_recurrent+k_expand_backup",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "in_step": in_step,
                    "traffic": ncu_traffic(workload) if fused else None,
                    "traffic_source": ("profiles/" + NCU_CAPTURE[workload]) if workload in NCU_CAPTURE else None,
                    "algorithmic_bytes_per_launch": bytes_per_sim * G * S,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                    "nominal_hbm_frac": achieved / 8000.0,
                    "kernel_ms": k_ms, "bytes_per_sim": bytes_per_sim, "mean_path_nodes": L, "fp32_tflops": tflops,
                    "kernel_share_of_step": k_ms * n_moves / ms if world == 1 else None}
        if rank == 0 and world == 1:
            try:
                parity = fc_parity(workload, cfg, sp, mcts)
            except Exception as e:                    # noqa: BLE001 - a checker failure must not lose the measurement
                parity = {"error": repr(e)}
    else:
        # the dominant kernel alone: the tower's C -> C convolution at the bench batch, timed live with CUDA events
        # (activation buffers 3 x B x rows x C bf16 rotate through HBM; connect4: 3 x 132 MB > L2)
        C_lat, H_lat, W_lat = (int(x) for x in sp.model.latent_shape)
        ws = sp.model._workspace(G, dev)
        iters = 20
        probe = lambda: _lib.check(_lib.lib.mzb_resnet_conv_probe(sp.model.handle(), G, _lib.ptr(ws), ws.numel(), iters,
                                                                   _lib.current_stream()))
        probe()
        torch.cuda.synchronize()
        # replayed from a CUDA graph like the search itself: plain back-to-back launches are host-bound here (three
        # tensor-map encodes + a launch per 80 us kernel) and would time the launch path, not the kernel
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            probe()
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); graph.replay(); b.record()
        torch.cuda.synchronize()
        conv_ms = a.elapsed_time(b) / iters
        del graph
        conv_flop = 2.0 * G * H_lat * W_lat * C_lat * C_lat * 9
        conv_tflops = conv_flop / (conv_ms * 1e-3) / 1e12
        # the same launch against the other roof: it reads every row of the padded NHWC layout ((H+1)(W+1) rows per
        # image) once and writes the H*W pixel rows; at C=64 its intensity (~250 flop/B) sits on the ridge
        conv_bytes = 2.0 * G * C_lat * ((H_lat + 1) * (W_lat + 1) + H_lat * W_lat)
        conv_gbs = conv_bytes / (conv_ms * 1e-3) / 1e9
        tpeak = float(peaks.get("bf16_tflops", 1650.0))
        tpeak_s = float(peaks.get("bf16_tflops_sustained", 1400.0))
        narrow = None
        if C_lat == 16 and getattr(mcts, "_pool", None) is not None:
            # narrow network (breakout): the dominant kernel of a step is the ONE-KERNEL recurrent inference
            # (k_recurrent16, csrc/mzb_tower16.cu) - time it alone on the search's own hidden-state pool (+ its three
            # head-mlp launches), slot 0 -> slot 1 of every game
            pool = mcts._pool
            state = int(pool.shape[2])
            zero_slot = torch.zeros(G, dtype=torch.int32, device=dev)
            act = torch.zeros((G, 1), dtype=torch.int32, device=dev)
            rec = lambda: sp.model.recurrent_inference_fused(pool, act, in_layout=2, in_slot=zero_slot,
                                                            in_row_stride=int(pool.shape[1]) * state, slot_stride=state,
                                                            state_out=pool, out_layout=2, out_row_stride=int(pool.shape[1]) * state,
                                                            out_offset=state)
            for _ in range(3):
                rec()
            torch.cuda.synchronize()
            rgraph = torch.cuda.CUDAGraph()                 # replayed from a graph like the search itself (4 launches
            with torch.cuda.graph(rgraph):                  # per call: eager ctypes launches would time the host)
                for _ in range(10):
                    rec()
            rgraph.replay()
            torch.cuda.synchronize()
            ra, rb_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ra.record(); rgraph.replay(); rb_.record()
            torch.cuda.synchronize()
            rec_ms = ra.elapsed_time(rb_) / 10
            del rgraph
            rec_tf = flops[1] * G / (rec_ms * 1e-3) / 1e12
            narrow = {"kernel": "k_recurrent16 (whole recurrent inference of the 16-channel network, warp per image, mma.sync on "
                                "shared-memory activations) + 3 x k_head_mma", "us": rec_ms * 1e3, "achieved": rec_tf, "peak": tpeak,
                      "unit": "TFLOP/s", "frac": rec_tf / tpeak, "flop_per_launch": flops[1] * G,
                      "limiter": "shared-memory pipe 70 % busy (profiles/r04_ncu_k_recurrent16.csv): the implicit GEMM re-reads each "
                                 "activation row once per tap; HMMA pipe 38 %",
                      "traffic": ncu_traffic(workload), "traffic_source": "profiles/" + NCU_CAPTURE[workload]}
        in_ms = ms / n_moves                              # a whole move inside the timed region
        tflops_in = (flops[1] * S + flops[0]) * G / (in_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": f"k_conv_tc (tcgen05 implicit-GEMM 3x3 conv, {C_lat}->{C_lat} ch, {H_lat}x{W_lat}, batch {G})",
                    "achieved": conv_tflops, "peak": tpeak, "unit": "TFLOP/s", "frac": conv_tflops / tpeak,
                    "in_step": {"achieved": tflops_in, "peak": tpeak_s, "frac": tflops_in / tpeak_s,
                                "note": "every FLOP of the move (BASELINE flop/sim x sims + initial inference) / move time inside the timed "
                                        "region, tree / head / env kernels included, against the SUSTAINED bf16 peak: a lower bound of the "
                                        "convolution's in-step fraction (it carries ~100 % of the FLOPs in < 100 % of the time)"},
                    "traffic": ncu_traffic(workload) if narrow is None else None,
                    "traffic_source": ("profiles/" + NCU_CAPTURE[workload]) if (workload in NCU_CAPTURE and narrow is None) else None,
                    "recurrent_inference": narrow,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)" if peaks else "fallback 1650 TFLOP/s",
                    "kernel_us": conv_ms * 1e3, "flop_per_launch": conv_flop,
                    "hbm_side": {"algorithmic_bytes_per_launch": conv_bytes, "achieved": conv_gbs, "peak": peak, "unit": "GB/s",
                                 "frac": conv_gbs / peak, "flop_per_byte": conv_flop / conv_bytes},
                    "whole_search": {"achieved": tflops, "peak": tpeak_s, "frac": tflops / tpeak_s,
                                     "note": "all FLOPs of the search (BASELINE flop/sim) / search time, helper and tree kernels included; sustained peak"},
                    "search_ms": k_ms, "flop_per_sim": flops[1], "tree_bytes_per_sim": bytes_per_sim, "mean_path_nodes": L,
                    "tree_hbm_gbs": achieved, "search_share_of_step": k_ms * n_moves / ms if world == 1 else None}

    # ---------------- e2e: batched MCTS.run entry point with HOST buffers, copies inside the timed region
    # image observations cross PCIe as the emulator's uint8 frames and are normalised on the device exactly like the
    # reference's wrapper does on the host (float32(frame) / 255, games/breakout.py:141-159): 4x fewer bytes per step
    frames_u8 = workload == "breakout"
    h_obs = torch.empty((G, env.obs_dim), dtype=torch.uint8 if frames_u8 else torch.float32).pin_memory()
    h_legal = torch.empty((G, A), dtype=torch.uint8).pin_memory()
    h_tp = torch.empty(G, dtype=torch.int8).pin_memory()
    h_obs.copy_(((obs * 255.0).round().clamp_(0, 255).to(torch.uint8) if frames_u8 else obs).cpu())
    h_legal.copy_(legal.cpu()); h_tp.copy_(to_play.cpu())
    h_vis = torch.empty((G, A), dtype=torch.int32).pin_memory()
    h_rv = torch.empty(G, dtype=torch.float64).pin_memory()
    d_obs = torch.empty((G, env.obs_dim), dtype=h_obs.dtype, device=dev)
    d_legal, d_tp = torch.empty_like(legal), torch.empty_like(to_play)
    e2e_steps = max(3, min(steps * moves_per_step, 20)) if is_fc else 2

    def e2e_step():
        d_obs.copy_(h_obs, non_blocking=True); d_legal.copy_(h_legal, non_blocking=True); d_tp.copy_(h_tp, non_blocking=True)
        o = mcts.run(sp.model, d_obs, d_legal, d_tp, True, slot=env.slot, step=env.step_count,
                     allow_fused=not args.modular, out=sp._out)
        h_vis.copy_(o["visits"], non_blocking=True); h_rv.copy_(o["root_value"], non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller reads the visit counts before the next move

    for _ in range(2):                          # resnet: first call plain launches, second captures the CUDA graph
        e2e_step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(e2e_steps):
        e2e_step()
    b.record()
    barrier()
    e2e_ms = mdist.max_over_ranks(a.elapsed_time(b), dev)
    assert int(h_vis.sum(1).min()) == S and int(h_vis.sum(1).max()) == S
    e2e = {"value": e2e_steps * G * S * world / (e2e_ms * 1e-3), "unit": "simulations/s",
           "h2d_bytes_per_step": (h_obs.numel() * h_obs.element_size() + h_legal.numel() + h_tp.numel()) * world,
           "observation_dtype": "uint8 frames, /255 on the device" if frames_u8 else "float32",
           "d2h_bytes_per_step": (h_vis.numel() * 4 + h_rv.numel() * 8) * world,
           "searches_timed": e2e_steps,
           "api": ("BatchedMCTS.run == mzb_search_fc" if is_fc else "BatchedMCTS.run == mzb_search_resnet")
                  + " (G x MCTS.run) with pinned host buffers; one search per e2e step"}

    # ---------------- learner-side collectives (off the self-play path): weight refresh + gradient all-reduce of this
    # workload's parameter count, device-timed, max over ranks
    collectives = None
    if world > 1 and want_collectives:
        import torch.distributed as dist  # noqa: F401
        n_param = sum(int(v.numel()) for v in weights.values())
        sd = {k: v.to(dev) for k, v in weights.items()}
        grads = [torch.ones(n_param, device=dev)]
        res = {}
        for name, fn in (("weight_broadcast_ms", lambda: mdist.broadcast_weights(sd, 0, dev)),
                         ("grad_allreduce_ms", lambda: mdist.allreduce_gradients(grads))):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                fn()
            b.record()
            barrier()
            res[name] = mdist.max_over_ranks(a.elapsed_time(b) / 20, dev)
        collectives = dict(res, parameters=n_param, backend="nccl", note="not on the self-play path (games are sharded, no data-path collective)")

    # the learner's step on the games this run just produced (off the self-play path, reported beside it):
    # batch from the device replay store -> forward + backward + optimiser -> priorities written back.  FC family: the
    # one-kernel training step (csrc/mzb_fc_train.cu) against the autograd graph replayed from a CUDA graph.
    trainer_probe = None
    if is_fc and rb is not None and world == 1 and len(rb) >= 16:
        try:
            from muzero_hypermodel_b200.trainer import Trainer
            trainer_probe = {}
            for key, env_val in (("ms_per_step", "1"), ("ms_per_step_autograd_cuda_graph", "0")):
                os.environ["MZB_TRAIN_KERNEL"] = env_val
                tr = Trainer({"weights": weights, "training_step": 0, "optimizer_state": None}, cfg, device=dev)
                for _ in range(4):
                    idx, batch = rb.get_batch(); tr.update_lr(); out = tr.update_weights(batch); rb.update_priorities(out[0], idx)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(20):
                    idx, batch = rb.get_batch(); tr.update_lr(); out = tr.update_weights(batch); rb.update_priorities(out[0], idx)
                b.record(); torch.cuda.synchronize()
                trainer_probe[key] = a.elapsed_time(b) / 20
                del tr
            os.environ.pop("MZB_TRAIN_KERNEL", None)
            trainer_probe.update(batch_size=cfg.batch_size, unroll_steps=cfg.num_unroll_steps, optimizer=cfg.optimizer,
                                 samples_per_s=cfg.batch_size / (trainer_probe["ms_per_step"] * 1e-3),
                                 path="ReplayBuffer.get_batch (device) -> k_fc_train (forward + backward, one launch) -> k_adam_flat -> update_priorities")
        except Exception as e:                        # noqa: BLE001
            trainer_probe = {"error": repr(e)}

    tree_gib = mcts.tree.nbytes / 2**30
    if not is_fc and rank == 0 and world == 1:
        # free the big search state before the fp32 comparison model is built
        del sp, env, mcts, rb
        gc.collect(); torch.cuda.empty_cache()
        try:
            parity = resnet_parity(workload, cfg, weights, dev, {"connect4": 64, "gomoku": 16, "breakout": 64}.get(workload, 64))
        except Exception as e:                        # noqa: BLE001
            parity = {"error": repr(e)}
        sp = env = mcts = rb = None

    line = None
    if rank == 0:
        line = {"metric": "mcts_simulations_per_sec", "value": value, "unit": "simulations/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 network / f64 tree statistics" if is_fc else "bf16 conv operands, f32 accumulate / f64 tree statistics",
                "data": "synthetic",
                "config": {"workload": workload_name(workload, cfg), "games_per_gpu": G,
                           "num_simulations": S, "weights": tag, "moves_per_step": moves_per_step,
                           "l2_policy": f"working set {tree_gib:.1f} GiB tree store per GPU >> 126 MB L2",
                           "path": ("fused whole-search kernel" if fused else "modular kernels") if is_fc else "tree kernels + tcgen05 resnet per simulation",
                           "parallelism": f"games sharded x{world}, no collective"},
                "env_steps_per_sec": env_steps, "e2e": e2e, "roofline": roofline, "gpu_launches": launches,
                "clocks": clocks, "mean_search_path_nodes": L, "timed_region_s": ms * 1e-3,
                "games_finished_in_timed_region": c1["games"] - c0["games"],
                "games_dropped": c1["dropped_games"] - c0["dropped_games"],
                "games_ingested_by_replay_store": ingested}
        if parity is not None:
            line["parity"] = parity
        if trainer_probe is not None:
            line["trainer"] = trainer_probe
        if collectives is not None:
            line["collectives"] = collectives
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if cpu_port is not None:
            line["cpu_baseline_port"] = cpu_port
    del sp, env, mcts, rb
    gc.collect()
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps of the headline workload (default 20)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + list(WORKLOADS))
    ap.add_argument("--games", type=int, default=0, help="games per GPU (default: workload's; single-workload runs only)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nested", action="store_true", help="headline only (skip the other four BASELINE configs)")
    ap.add_argument("--modular", action="store_true", help="force the modular kernels instead of the whole-search kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    head = "cartpole" if args.workload == "all" else args.workload
    if args.steps <= 0:
        args.steps = WORKLOADS[head][5] if args.impl == "ours" else 10

    if args.impl == "reference":
        return reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline

    # CPU legs of the nested workloads first: forked workers must not inherit a CUDA context
    nested = list(NESTED) if (args.workload == "all" and not args.no_nested) else []
    if args.games and nested:
        args.games = 0
    pre_cpu = {}
    if want_cpu:
        for w in nested + [head]:
            try:
                pre_cpu[w] = cpu_legs(w, make_config(w), args.cpu_seconds if w == head else 5.0)
            except Exception as e:                    # noqa: BLE001
                pre_cpu[w] = ({"error": repr(e)}, None)

    import torch
    import torch.distributed as dist
    from muzero_hypermodel_b200 import dist as mdist
    torch.cuda.set_device(local_rank)
    mdist.init(backend="nccl", device=torch.device("cuda", local_rank))

    results = {}
    for w in nested:
        try:
            line = run_workload(args, w, WORKLOADS[w][5], 3, 0.0, False, False)
        except Exception as e:                        # noqa: BLE001 - one config failing must not lose the headline
            import traceback
            line = {"error": repr(e), "traceback": traceback.format_exc()[-1500:]}
            gc.collect(); torch.cuda.empty_cache()
        if rank == 0:
            if isinstance(line, dict) and w in pre_cpu and "error" not in line:
                line["cpu_baseline"], port = pre_cpu[w]
                if port is not None:
                    line["cpu_baseline_port"] = port
            results[w] = line
    line = run_workload(args, head, args.steps, args.warmup, 0.0, False, True)
    if rank == 0:
        if head in pre_cpu:
            line["cpu_baseline"], port = pre_cpu[head]
            if port is not None:
                line["cpu_baseline_port"] = port
        if results:
            line["workloads"] = results
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
