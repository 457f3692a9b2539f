"""Bring-up: the one-kernel FC training step against autograd on REAL replay batches (GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from muzero_hypermodel_b200.replay_buffer import ReplayBuffer
from muzero_hypermodel_b200.self_play import SelfPlay
from muzero_hypermodel_b200.trainer import Trainer
wl = "cartpole"
cfg = bench.make_config(wl); cfg.num_simulations = 10
dev = torch.device("cuda:0")
w = {k: torch.tensor(v) for k, v in bench.load_weights(bench.WORKLOADS[wl][0]).items()}
sp = SelfPlay({"weights": w}, None, cfg, 0, n_games=1024, device=dev)
env, _ = sp._setup()
rb = ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg, device=dev, record_env=env)
while len(rb) < 64:
    sp.step(); rb.ingest(env)
tr = Trainer({"weights": w, "training_step": 0, "optimizer_state": None}, cfg, device=dev)
tr.use_cuda_graph = False
idx, batch = rb.get_batch()
t = lambda x, dt: None if x is None else torch.as_tensor(x).to(dev, dt)
ob, ac, tv, trw, tp, wb, gs = batch
tensors = (t(ob, torch.float32), t(ac, torch.int64), t(tv, torch.float32), t(trw, torch.float32), t(tp, torch.float32), t(wb, torch.float32) if cfg.PER else None, t(gs, torch.float32))
for n, x in zip("obs action tv tr tp w gs".split(), tensors):
    print(n, None if x is None else (tuple(x.shape), x.dtype, x.is_contiguous(), float(x.float().min()), float(x.float().max())))
k = tr._fc_kernel_step(tensors); g1 = tr.flat_grad.clone(); k = [x.clone() for x in k]
r = tr._forward_backward_autograd(tensors); g0 = tr.flat_grad.clone()
print("loss", float(k[0]), float(r[0]), "grad max", float(g0.abs().max()), "err", float((g1 - g0).abs().max()))
print("vl", float((k[1] - r[1]).abs().max()), "rl", float((k[2] - r[2]).abs().max()), "pl", float((k[3] - r[3]).abs().max()))
bad = (g1 - g0).abs() > 1e-4 * g0.abs().max()
print("bad grads", int(bad.sum()), "of", bad.numel(), "first", bad.nonzero()[:10].flatten().tolist())
print("tp row sums", tp.sum(-1).min().item() if torch.is_tensor(tp) else None)
# in-situ time of the one-kernel step (warm caches, back to back) against the autograd path replayed from a CUDA graph
for name, fn in (("k_fc_train + reduce (+ 2 codec launches)", lambda: tr._fc_kernel_step(tensors)),):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) * 10:.1f} us per call")
