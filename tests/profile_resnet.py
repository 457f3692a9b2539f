"""Profiling helper: one short resnet search (few simulations) at bench batch size.  python tests/profile_resnet.py [workload] [G] [sims]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from muzero_hypermodel_b200.self_play import SelfPlay

wl = sys.argv[1] if len(sys.argv) > 1 else "connect4"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
sims = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cfg = bench.make_config(wl)
cfg.num_simulations = sims
w = {k: torch.tensor(v) for k, v in bench.load_weights(bench.WORKLOADS[wl][0]).items()}
sp = SelfPlay({"weights": w}, None, cfg, 0, n_games=G, device="cuda:0")
for i in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sp.step(); b.record(); torch.cuda.synchronize()
    print(f"step {i}: {a.elapsed_time(b):.2f} ms for {sims} sims -> {a.elapsed_time(b)/sims:.3f} ms/sim")
