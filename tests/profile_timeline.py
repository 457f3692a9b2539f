"""Kernel timeline of one self-play step through torch.profiler (CUPTI): per-kernel durations AS THEY RUN inside the
step (warm caches, graph replay) and the idle gaps between consecutive kernels.
python tests/profile_timeline.py [workload] [G] [sims]"""
import os, sys, collections
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from muzero_hypermodel_b200.self_play import SelfPlay

wl = sys.argv[1] if len(sys.argv) > 1 else "connect4"
G = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[wl][2]
sims = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cfg = bench.make_config(wl)
cfg.num_simulations = sims
w = {k: torch.tensor(v) for k, v in bench.load_weights(bench.WORKLOADS[wl][0]).items()}
sp = SelfPlay({"weights": w}, None, cfg, 0, n_games=G, device="cuda:0")
for _ in range(3):
    sp.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    sp.step()
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
prev_end, busy = None, 0.0
for e in ev:
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    name = e.name.replace("void ", "").replace("(anonymous namespace)::", "")[:48]
    a = agg[name]
    a[0] += 1; a[1] += d
    if prev_end is not None:
        a[2] += max(0.0, s - prev_end)        # idle time in front of this kernel
    prev_end = max(prev_end or 0, e.time_range.end)
    busy += d
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"{wl} G={G} sims={sims}: {len(ev)} kernels, span {span/1e3:.2f} ms, busy {busy/1e3:.2f} ms, idle {(span-busy)/1e3:.2f} ms")
print(f"{'kernel':50s} {'n':>5s} {'avg us':>8s} {'total ms':>9s} {'gap before, avg us':>18s}")
for k, (n, t, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} {n:5d} {t/n:8.1f} {t/1e3:9.2f} {g/n:18.1f}")
if len(sys.argv) > 4 and sys.argv[4] == "seq":
    # launch-by-launch durations of one simulation in the middle of the search (between two k_select launches)
    sel = [i for i, e in enumerate(ev) if "k_select" in e.name]
    if len(sel) > 11:
        print("one simulation, launch by launch (us):")
        for e in ev[sel[10]:sel[11]]:
            print(f"  {e.name.replace('void ', '').replace('(anonymous namespace)::', '')[:60]:60s} {e.time_range.end - e.time_range.start:8.1f}")
