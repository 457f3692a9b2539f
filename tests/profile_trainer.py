"""Trainer step time (python tests/profile_trainer.py [workload] [batch])."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
from muzero_hypermodel_b200.self_play import SelfPlay
from muzero_hypermodel_b200.trainer import Trainer

wl = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
cfg = bench.make_config(wl)
if len(sys.argv) > 2:
    cfg.batch_size = int(sys.argv[2])
cfg.num_simulations = 10
dev = torch.device("cuda:0")
w = {k: torch.tensor(v) for k, v in bench.load_weights(bench.WORKLOADS[wl][0]).items()}
sp = SelfPlay({"weights": w}, None, cfg, 0, n_games=1024, device=dev)
env, _ = sp._setup()
rb = ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg, device=dev, record_env=env)
while len(rb) < 64:
    sp.step(); rb.ingest(env)
tr = Trainer({"weights": w, "training_step": 0, "optimizer_state": None}, cfg, device=dev)
for _ in range(5):
    idx, batch = rb.get_batch(); tr.update_lr(); out = tr.update_weights(batch); rb.update_priorities(out[0], idx)
torch.cuda.synchronize()
t0 = time.time(); n = 30
for _ in range(n):
    idx, batch = rb.get_batch(); tr.update_lr(); out = tr.update_weights(batch); rb.update_priorities(out[0], idx)
torch.cuda.synchronize()
dt = (time.time() - t0) / n
print(f"{wl}: batch {cfg.batch_size} x {cfg.num_unroll_steps + 1} unrolled steps: {dt * 1e3:.2f} ms per training step "
      f"({cfg.batch_size / dt:.0f} samples/s), loss {out[1]:.3f}")
