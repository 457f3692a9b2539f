"""Bring-up: per-role clock64 timeline of CTA 0 of the persistent convolution (GPU box)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from muzero_hypermodel_b200 import _lib
from muzero_hypermodel_b200.self_play import SelfPlay
_lib.bind("mzb_conv_tc_debug_buffer", None, [C.c_void_p])
WL = sys.argv[2] if len(sys.argv) > 2 else "connect4"
G = int(sys.argv[3]) if len(sys.argv) > 3 else bench.WORKLOADS[WL][2]
cfg = bench.make_config(WL); cfg.num_simulations = 2
w = {k: torch.tensor(v) for k, v in bench.load_weights(bench.WORKLOADS[WL][0]).items()}
sp = SelfPlay({"weights": w}, None, cfg, 0, n_games=G, device="cuda:0")
sp.step(); torch.cuda.synchronize()
buf = torch.zeros(4 * 32 * 4, dtype=torch.int64, device="cuda:0")
_lib.lib.mzb_conv_tc_debug_buffer(C.c_void_p(buf.data_ptr()))
if len(sys.argv) > 1 and sys.argv[1] == "probe":        # the bench's probe layer (no residual) instead of a search
    ws = sp.model._workspace(G, torch.device("cuda:0"))
    _lib.check(_lib.lib.mzb_resnet_conv_probe(sp.model.handle(), G, _lib.ptr(ws), ws.numel(), 1, _lib.current_stream()))
else:
    sp.step()
torch.cuda.synchronize()
_lib.lib.mzb_conv_tc_debug_buffer(None)
d = buf.cpu().numpy().reshape(4, 32, 4)
t0 = d[0, 0, 0]
print("MMA: it  wait_a_start  a_full  acc_empty  issued(commit)")
for it in range(12): print("  ", it, *(int(x - t0) for x in d[0, it]))
for g in (1, 2):
    print(f"EPI group {g-1}: it  wait_start  acc_full  done")
    for it in range(12): print("  ", it, *(int(x - t0) for x in d[g, it, :3]))
print("PRODUCER: it  start  a_empty_ok")
for it in range(12): print("  ", it, *(int(x - t0) for x in d[3, it, :2]))
