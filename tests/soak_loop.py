"""Soak: the whole loop for many moves on one GPU (python tests/soak_loop.py [workload] [games] [moves]) - self-play,
device-to-device ingest, training steps with priority write-back, weight refresh, Reanalyse; checks the counters."""
import math, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from muzero_hypermodel_b200.replay_buffer import Reanalyse, ReplayBuffer
from muzero_hypermodel_b200.self_play import SelfPlay
from muzero_hypermodel_b200.trainer import Trainer

wl = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
moves = int(sys.argv[3]) if len(sys.argv) > 3 else 600
cfg = bench.make_config(wl)
dev = torch.device("cuda:0")
w = {k: torch.tensor(v) for k, v in bench.load_weights(bench.WORKLOADS[wl][0]).items()}
sp = SelfPlay({"weights": w}, None, cfg, 0, n_games=G, device=dev)
env, _ = sp._setup()
rb = ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg, device=dev, record_env=env)
tr = Trainer({"weights": w, "training_step": 0, "optimizer_state": None}, cfg, device=dev)
ra = Reanalyse({"weights": w, "num_reanalysed_games": 0}, cfg, device=dev)
t0 = time.time()
losses = []
for m in range(moves):
    sp.step(1.0, cfg.temperature_threshold)
    if (m + 1) % 4 == 0:
        rb.ingest(env)
    if (m + 1) % 10 == 0 and len(rb) > 0:
        idx, batch = rb.get_batch()
        tr.update_lr()
        out = tr.update_weights(batch)
        if cfg.PER:
            rb.update_priorities(out[0], idx)
        losses.append(out[1])
        assert math.isfinite(out[1]), out
    if (m + 1) % 50 == 0 and len(rb) > 0:
        sp.model.set_weights(tr.model.get_weights())                 # weight refresh
        ra.model.set_weights(tr.model.get_weights())
        ra.reanalyse_game(rb, rb.sample_game(force_uniform=True))
rb.ingest(env)
torch.cuda.synchronize()
c = env.counters()
print(f"{wl}: {moves} moves x {G} games in {time.time() - t0:.1f} s; env steps {c['env_steps']}, games finished {c['games']}, "
      f"dropped {c['dropped_games']}, ingested {rb.num_played_games} (buffer holds {len(rb)}, {rb.total_samples} positions), "
      f"training steps {tr.training_step}, loss first/last {losses[0]:.3f}/{losses[-1]:.3f}, reanalysed {ra.num_reanalysed_games}")
assert c["env_steps"] == moves * G and c["dropped_games"] == 0 and rb.num_played_games == c["games"]
print("soak ok")
