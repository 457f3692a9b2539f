"""Bring-up helper for the tcgen05 convolution: python tests/debug_conv_tc.py  (GPU box)."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_hypermodel_b200 import _lib

fn = _lib.bind("mzb_debug_conv3x3", C.c_int, [C.c_int] * 5 + [C.c_void_p] * 4)


def run(B, H, W, Cin, Cout, x, w):
    yd = np.zeros((B, H, W, Cout), np.float32); yt = np.zeros_like(yd)
    rc = fn(B, H, W, Cin, Cout, x.ctypes.data, w.ctypes.data, yd.ctypes.data, yt.ctypes.data)
    assert rc == 0, _lib.lib.mzb_last_error()
    return yd, yt


rs = np.random.RandomState(0)
for (B, H, W, Cin, Cout) in [(3, 6, 7, 64, 64), (2, 11, 11, 128, 128), (14, 3, 3, 16, 16), (3, 6, 6, 16, 16), (2, 5, 5, 32, 32)]:
    print("=== shape", B, H, W, Cin, Cout)
    # 1) single-tap probes with identity channel mixing: y[., c] = x[shifted ., c]
    for tap in range(9):
        w = np.zeros((Cout, Cin, 3, 3), np.float32)
        for c in range(min(Cin, Cout)):
            w[c, c, tap // 3, tap % 3] = 1.0
        x = rs.randint(-3, 4, (B, H, W, Cin)).astype(np.float32)
        yd, yt = run(B, H, W, Cin, Cout, x, w)
        bad = np.argwhere(yd != yt)
        print(f" tap {tap}: mismatches {len(bad)}/{yd.size}", "" if not len(bad) else f"first {bad[0]} direct {yd[tuple(bad[0])]} tc {yt[tuple(bad[0])]}")
    # 2) random
    w = (rs.randn(Cout, Cin, 3, 3) / np.sqrt(9 * Cin)).astype(np.float32)
    x = rs.randn(B, H, W, Cin).astype(np.float32)
    yd, yt = run(B, H, W, Cin, Cout, x, w)
    print(" random: max abs diff", np.abs(yd - yt).max(), "max |y|", np.abs(yd).max())
