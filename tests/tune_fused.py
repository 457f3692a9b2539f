"""Tuning helper (GPU box): time the whole-search kernel for each MZB_FUSED_EXP flag set given on the command line.

    python tests/tune_fused.py cartpole 262144 0 1 3 7 11 19 35

Every flag set runs in its own process (the flags are read once per process).  Prints the search-launch time
(CUDA events, median of 7 after 3 warm-ups) and the whole self-play step.  Flags >= 4 are timing-only diagnostics."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(wl, G):
    import torch
    import bench
    from muzero_hypermodel_b200.self_play import SelfPlay
    cfg = bench.make_config(wl)
    w = {k: torch.tensor(v) for k, v in bench.load_weights(bench.WORKLOADS[wl][0]).items()}
    sp = SelfPlay({"weights": w}, None, cfg, 0, n_games=G, device="cuda:0")
    env, mcts = sp._setup()
    for _ in range(3):
        sp.step()
    obs, legal, to_play = env.observe()
    ks, ts = [], []
    for i in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mcts.run(sp.model, obs, legal, to_play, True, slot=env.slot, step=env.step_count, out=sp._out)
        b.record(); torch.cuda.synchronize()
        ks.append(a.elapsed_time(b))
    for i in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); sp.step(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    k = sorted(ks)[len(ks) // 2]
    t = sorted(ts)[len(ts) // 2]
    vis = sp._out["visits"].sum(1)
    print(f"exp {os.environ.get('MZB_FUSED_EXP', 'default'):>8} {wl} G={G}: search {k:.3f} ms  step {t:.3f} ms -> "
          f"{G * cfg.num_simulations / t / 1e6:.1f} M sims/s  (visits/game {int(vis.min())}..{int(vis.max())})", flush=True)


if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    exps = sys.argv[3:]
    if os.environ.get("_TUNE_CHILD"):
        one(wl, G)
    else:
        for e in exps or ["default"]:
            env = dict(os.environ, _TUNE_CHILD="1")
            if e != "default":
                env["MZB_FUSED_EXP"] = e
            subprocess.run([sys.executable, os.path.abspath(__file__), wl, str(G)], env=env)
