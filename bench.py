#!/usr/bin/env python
"""bench.py - self-play hot path throughput: MCTS simulations/s (and env steps/s) per BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cartpole|tictactoe|connect4|gomoku|breakout]
                    [--games G]

One "step" = one self-play move of every game on every GPU: observe -> MCTS.run (num_simulations
simulations, network in the loop) -> select_action -> Game.step -> GameHistory append -> harvest/auto-reset.
Workload (config.workload): BASELINE.json configs[0], "cartpole FC MuZero (games/cartpole.py defaults,
num_simulations=50)" - the configuration the headline target (>=1e8 simulations/s on 8xB200) is quoted on.
Weights: the reference's shipped cartpole checkpoint (tests/golden/net.npz "cartpole_shipped", 1,532
parameters); observations come from the device CartPole-v1 environments (synthetic games, no dataset).

JSON line keys: see the task contract; `value` = simulations/s over all GPUs with state resident in HBM,
`e2e` = the same searches driven through the batched MCTS.run entry point with HOST observation /
legal-action / to-play buffers (pinned H2D before, D2H of visit counts + root values after, every step),
`roofline` = the dominant kernel alone, timed live with CUDA events: the whole-search kernel against the measured
HBM copy peak (FC workloads; `traffic` = DRAM bytes per launch from the committed ncu capture), or the tower's
C->C tcgen05 convolution replayed from a CUDA graph against the measured bf16 burst peak (resnet workloads;
`roofline.whole_search` keeps the all-in figure), `cpu_baseline` = the oracle port of SelfPlay.play_game on the
host cores (bounded sample), `collectives` (N > 1) = device time of the two learner-side collectives.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (golden weight tag, config module, games per GPU, (initial FLOP, recurrent FLOP) per BASELINE.md §3)
    "cartpole": ("cartpole_shipped", "cartpole", 262144, (1312, 2752)),
    "tictactoe": ("tictactoe_fc", "tictactoe", 4096 * 16, (3648, 5952)),
    "connect4": ("connect4", "connect4", 16384, (37372160, 40396160)),
    "gomoku": ("gomoku", "gomoku", 4096, (857557760, 892780160)),      # 51 GB bf16 hidden-state pool (401 slots x 31 KB x 4096)
    "breakout": ("breakout", "breakout", 16384, (34192160, 1532480)),
}
# committed `ncu --set full` captures of the dominant kernel at the bench shape (profiles/): DRAM bytes per launch
NCU_CAPTURE = {"cartpole": "r01_ncu_k_search_fc_cartpole.csv", "connect4": "r01_ncu_k_conv_tc_connect4.csv"}


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
            if len(c) >= 3 and c[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and c[1] in unit:
                total += float(c[2]) * unit[c[1]]
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


def cpu_leg(workload, cfg, seconds):
    """Oracle port of play_game on every host core for ~`seconds` (the bench's only use of oracle/)."""
    from oracle import cpu_baseline
    cores = os.cpu_count() or 1
    w = load_weights(WORKLOADS[workload][0])
    sims, steps, wall = cpu_baseline.run_parallel(WORKLOADS[workload][1], w, oracle_cfg(cfg), cores, max_seconds=seconds)
    return {"value": sims / wall, "unit": "simulations/s", "cores": cores, "kind": "port",
            "env_steps_per_s": steps / wall,
            "sample": f"{cores} processes x {seconds:.0f} s of oracle play_game ({steps} searches of "
                      f"{cfg.num_simulations} simulations, batch-1 numpy network, same weights/config)"}


def reference_arm(args):
    """--impl reference: the reference's CPU path (oracle port; the Python reference cannot travel)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import cpu_baseline
    cfg = make_config(args.workload)
    cores = os.cpu_count() or 1
    w = load_weights(WORKLOADS[args.workload][0])
    ocfg = oracle_cfg(cfg)
    searches = 24 if cfg.network == "fullyconnected" else 1      # searches per process per step (bounded sample)
    pool = mp.get_context("fork").Pool(cores)
    times, sims_total = [], 0
    for i in range(args.warmup + args.steps):
        sims, steps, wall = cpu_baseline.run_parallel(WORKLOADS[args.workload][1], w, dict(ocfg, seed=ocfg["seed"] + i),
                                                      cores, max_searches=searches, pool=pool)
        if i >= args.warmup:
            times.append(wall)
            sims_total += sims
    pool.close()
    pool.join()
    total = sum(times)
    value = sims_total / total
    sample = (f"each step = {cores} processes x {searches} searches x {cfg.num_simulations} simulations of the oracle "
              f"port of SelfPlay.play_game (batch-1 numpy network)")
    line = {"impl": "reference", "metric": "mcts_simulations_per_sec", "value": value, "unit": "simulations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, cfg), "games_per_gpu": cores,
                       "num_simulations": cfg.num_simulations},
            "cpu_baseline": {"value": value, "unit": "simulations/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "simulations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(workload, cfg):
    return {"cartpole": "cartpole FC MuZero (games/cartpole.py defaults, num_simulations=50)",
            "tictactoe": "tictactoe FC MuZero, two-player, network=fullyconnected, num_simulations=25",
            "connect4": "connect4 residual-network MuZero (games/connect4.py defaults, 3 blocks x 64 ch, num_simulations=200)",
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 50 for FC workloads, 3 for resnets)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cartpole", choices=list(WORKLOADS))
    ap.add_argument("--games", type=int, default=0, help="games per GPU (default: workload's)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--modular", action="store_true", help="force the modular kernels instead of the whole-search kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.steps <= 0:
        args.steps = 50 if args.workload in ("cartpole", "tictactoe") else 3
        if args.impl == "reference":
            args.steps = 10

    if args.impl == "reference":
        return reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = make_config(args.workload)

    # CPU baseline first (rank 0, N=1 only): forked workers must not inherit a CUDA context
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_leg(args.workload, cfg, args.cpu_seconds)

    import numpy as np
    import torch
    import torch.distributed as dist
    from muzero_hypermodel_b200 import _lib
    from muzero_hypermodel_b200 import dist as mdist
    from muzero_hypermodel_b200.self_play import SelfPlay

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    mdist.init(backend="nccl", device=dev)
    barrier = mdist.barrier

    G = args.games or WORKLOADS[args.workload][2]
    S = cfg.num_simulations
    weights = {k: torch.tensor(v) for k, v in load_weights(WORKLOADS[args.workload][0]).items()}
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

    def step():
        sp.step(temperature=1.0, temperature_threshold=None, add_exploration_noise=True, export=True,
                allow_fused=not args.modular)
        moves[0] += 1
        if rb is not None and moves[0] % 4 == 0:
            rb.ingest(env)

    for _ in range(args.warmup):
        step()
    if rb is not None:
        rb.ingest(env)
    else:
        sp.drain()
    ingested0 = rb.num_played_games if rb is not None else 0
    mcts.tree.counters(reset=True)
    barrier()

    # ---------------- timed region: K whole self-play steps, CUDA events on the launching stream
    sampler = ClockSampler(local_rank) if rank == 0 else None
    c0 = env.counters()
    _lib.lib.mzb_reset_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        step()
    ev[1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[1])
    launches = int(_lib.lib.mzb_launch_count())
    c1 = env.counters()
    tc = mcts.tree.counters()
    ingested_end = rb.num_played_games if rb is not None else 0
    # a timed region shorter than a few 50 ms sampling periods (breakout: 3 steps of 19 ms) would leave the clocks
    # line empty: keep the same load running, untimed, until the sampler has seen it
    extra = 0
    if sampler is not None and sampler.p is not None:
        t_end = time.time() + 3.0
        while sampler.count() < 3 and time.time() < t_end:
            step()
            torch.cuda.synchronize()
            extra += 1
    clocks = sampler.stop() if sampler else None
    if clocks is not None and extra:
        clocks["untimed_load_steps_for_sampling"] = extra
    ms = mdist.max_over_ranks(ms, dev)
    sims_total = args.steps * G * S * world
    value = sims_total / (ms * 1e-3)
    env_steps = args.steps * G * world / (ms * 1e-3)
    mean_path = tc["path_length_sum"] / max(1, tc["simulations"])
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
    flops = WORKLOADS[args.workload][3]
    tflops = (flops[1] * S + flops[0]) * G / (k_ms * 1e-3) / 1e12
    if is_fc:
        roofline = {"bound": "hbm", "kernel": "k_search_fc (whole-search, fused)" if fused else "modular: k_select+k_fc_recurrent+k_expand_backup",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(args.workload) if (fused and G == WORKLOADS[args.workload][2]) else None,
                    "traffic_source": ("profiles/" + NCU_CAPTURE[args.workload]) if args.workload in NCU_CAPTURE else None,
                    "algorithmic_bytes_per_launch": bytes_per_sim * G * S,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                    "kernel_ms": k_ms, "bytes_per_sim": bytes_per_sim, "mean_path_nodes": L, "fp32_tflops": tflops,
                    "kernel_share_of_step": k_ms * args.steps / ms if world == 1 else None}
    else:
        # the dominant kernel alone: the tower's C -> C convolution at the bench batch, timed live with CUDA events
        # (activation buffers 3 x B x rows x C bf16 rotate through HBM; connect4: 3 x 132 MB > L2)
        import ctypes as C
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
        conv_flop = 2.0 * G * H_lat * W_lat * C_lat * C_lat * 9
        conv_tflops = conv_flop / (conv_ms * 1e-3) / 1e12
        # the same launch against the other roof: it reads every row of the padded NHWC layout ((H+1)(W+1) rows per
        # image) once and writes the H*W pixel rows; at C=64 its intensity (~250 flop/B) sits on the ridge
        conv_bytes = 2.0 * G * C_lat * ((H_lat + 1) * (W_lat + 1) + H_lat * W_lat)
        conv_gbs = conv_bytes / (conv_ms * 1e-3) / 1e9
        tpeak = float(peaks.get("bf16_tflops", 1650.0))
        tpeak_s = float(peaks.get("bf16_tflops_sustained", 1400.0))
        roofline = {"bound": "tensor", "kernel": f"k_conv_tc (tcgen05 implicit-GEMM 3x3 conv, {C_lat}->{C_lat} ch, {H_lat}x{W_lat}, batch {G})",
                    "achieved": conv_tflops, "peak": tpeak, "unit": "TFLOP/s", "frac": conv_tflops / tpeak,
                    "traffic": ncu_traffic(args.workload),
                    "traffic_source": ("profiles/" + NCU_CAPTURE[args.workload]) if args.workload in NCU_CAPTURE else None,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)" if peaks else "fallback 1650 TFLOP/s",
                    "kernel_us": conv_ms * 1e3, "flop_per_launch": conv_flop,
                    "hbm_side": {"algorithmic_bytes_per_launch": conv_bytes, "achieved": conv_gbs, "peak": peak, "unit": "GB/s",
                                 "frac": conv_gbs / peak, "flop_per_byte": conv_flop / conv_bytes},
                    "whole_search": {"achieved": tflops, "peak": tpeak_s, "frac": tflops / tpeak_s,
                                     "note": "all FLOPs of the search (BASELINE flop/sim) / search time, helper and tree kernels included; sustained peak"},
                    "search_ms": k_ms, "flop_per_sim": flops[1], "tree_bytes_per_sim": bytes_per_sim, "mean_path_nodes": L,
                    "tree_hbm_gbs": achieved, "search_share_of_step": k_ms * args.steps / ms if world == 1 else None}

    # ---------------- e2e: batched MCTS.run entry point with HOST buffers, copies inside the timed region
    h_obs = torch.empty((G, env.obs_dim), dtype=torch.float32).pin_memory()
    h_legal = torch.empty((G, A), dtype=torch.uint8).pin_memory()
    h_tp = torch.empty(G, dtype=torch.int8).pin_memory()
    h_obs.copy_(obs.cpu()); h_legal.copy_(legal.cpu()); h_tp.copy_(to_play.cpu())
    h_vis = torch.empty((G, A), dtype=torch.int32).pin_memory()
    h_rv = torch.empty(G, dtype=torch.float64).pin_memory()
    d_obs, d_legal, d_tp = torch.empty_like(obs), torch.empty_like(legal), torch.empty_like(to_play)
    e2e_steps = max(3, min(args.steps, 10)) if is_fc else 2

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
           "h2d_bytes_per_step": (h_obs.numel() * 4 + h_legal.numel() + h_tp.numel()) * world,
           "d2h_bytes_per_step": (h_vis.numel() * 4 + h_rv.numel() * 8) * world,
           "api": "BatchedMCTS.run == mzb_search_fc (G x MCTS.run) with pinned host buffers"}

    # ---------------- learner-side collectives (off the self-play path): weight refresh + gradient all-reduce of this
    # workload's parameter count, device-timed, max over ranks
    collectives = None
    if world > 1:
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

    if rank == 0:
        line = {"metric": "mcts_simulations_per_sec", "value": value, "unit": "simulations/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 network / f64 tree statistics" if is_fc else "bf16 conv operands, f32 accumulate / f64 tree statistics",
                "data": "synthetic",
                "config": {"workload": workload_name(args.workload, cfg), "games_per_gpu": G,
                           "num_simulations": S, "weights": WORKLOADS[args.workload][0],
                           "l2_policy": f"working set {mcts.tree.nbytes / 2**30:.1f} GiB tree store per GPU >> 126 MB L2",
                           "path": ("fused whole-search kernel" if fused else "modular kernels") if is_fc else "tree kernels + tcgen05 resnet per simulation",
                           "parallelism": f"games sharded x{world}, no collective"},
                "env_steps_per_sec": env_steps, "e2e": e2e, "roofline": roofline, "gpu_launches": launches,
                "clocks": clocks, "mean_search_path_nodes": L,
                "games_finished_in_timed_region": c1["games"] - c0["games"],
                "games_dropped": c1["dropped_games"] - c0["dropped_games"],
                "games_ingested_by_replay_store": ingested}
        if collectives is not None:
            line["collectives"] = collectives
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
