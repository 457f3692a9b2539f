"""Per-launch durations of the stem kernels of one Breakout initial inference (torch.profiler / CUPTI).
python tests/profile_stem16.py [B]     (MZB_STEM16_GW / MZB_STEM16_NBUF select the kernel's plan)"""
import os, sys
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_resnet import _model, DEV
net, cfg, z = _model("breakout", precision="bf16")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
obs = torch.rand((B,) + tuple(z["breakout/obs"].shape[1:]), device=DEV)
for _ in range(3):
    net.initial_inference(obs)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    net.initial_inference(obs)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
tot = 0.0
out = []
for e in ev:
    d = e.time_range.end - e.time_range.start
    tot += d
    if d > 15:
        out.append(f"{e.name.replace('void ', '').replace('(anonymous namespace)::', '')[:28]}={d:.0f}")
print(f"GW={os.environ.get('MZB_STEM16_GW', '-')} NBUF={os.environ.get('MZB_STEM16_NBUF', '-')} total {tot / 1e3:.2f} ms: " + " ".join(out))
