"""Tuning helper: time the whole-search kernel for the MZB_FUSED_VARIANT in the environment (GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from muzero_hypermodel_b200.self_play import SelfPlay

wl = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
G = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[wl][2]
cfg = bench.make_config(wl)
w = {k: torch.tensor(v) for k, v in bench.load_weights(bench.WORKLOADS[wl][0]).items()}
sp = SelfPlay({"weights": w}, None, cfg, 0, n_games=G, device="cuda:0")
for _ in range(3):
    sp.step()
ts = []
for i in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sp.step(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
print(f"variant {os.environ.get('MZB_FUSED_VARIANT', 'default')} {wl} G={G}: {ms:.3f} ms/step -> {G * cfg.num_simulations / ms / 1e6:.1f} M sims/s")
