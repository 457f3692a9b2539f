"""Bring-up: run-to-run determinism of the stem towers (shared-memory form vs layer by layer).  python tests/debug_stem16.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_resnet import _model, DEV
from muzero_hypermodel_b200 import _lib
net, cfg, z = _model("breakout", precision="bf16")
rs = np.random.RandomState(5)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 333
obs = torch.tensor(rs.randint(0, 256, size=(B,) + tuple(z["breakout/obs"].shape[1:])).astype(np.float32) / 255.0, device=DEV)
seq = [1, 1, 1, 0, 0, 1, 1, 0, 1]
res = []
for mode in seq:
    _lib.lib.mzb_stem16_enable(mode)
    res.append(net.initial_inference(obs)[3].float().cpu().numpy())
_lib.lib.mzb_stem16_enable(1)
first = {1: res[0], 0: res[3]}
for i, (mode, r) in enumerate(zip(seq, res)):
    d = r != first[mode]
    imgs = np.unique(np.nonzero(d)[0])
    print(f"run {i} mode {mode}: {int(d.sum())} elements differ from the first run of this mode; images {imgs[:20].tolist()} ({len(imgs)})")
