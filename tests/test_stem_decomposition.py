"""The algebra behind the shared-memory stem kernels (csrc/mzb_stem16.cu), checked on the CPU against torch's convolutions:
the pixel-pair form of a C/2 -> C/2 convolution (pack_conv_pair, csrc/mzb_resnet.cu) with its structured zeros, and
DownSample.conv2 (stride 2) as six K = 16 taps on pair rows (pack_s2_mma).  float64, so equality is to rounding."""
import numpy as np
import torch


def _pairs(x):
    """[C, H, W] -> pair rows [H, W/2, 2C]: channels of pixel 2x, then of pixel 2x + 1 (k_conv_s2's in_pair layout)."""
    C, H, W = x.shape
    return x.reshape(C, H, W // 2, 2).transpose(1, 2, 3, 0).reshape(H, W // 2, 2 * C)


def _pad_pairs(p):
    H, Wp, K = p.shape
    out = np.zeros((H + 2, Wp + 2, K))
    out[1:-1, 1:-1] = p
    return out


def test_pixel_pair_convolution_and_its_zero_blocks():
    rs = np.random.RandomState(0)
    c, H, W = 8, 10, 12
    w = rs.randn(c, c, 3, 3)
    x = rs.randn(c, H, W)
    ref = torch.nn.functional.conv2d(torch.tensor(x)[None], torch.tensor(w), padding=1)[0].numpy()
    # pack_conv_pair: output channel po*c + o of pair tap kxp reads input channel pi*c + ci with pixel tap dx = 2 (kxp - 1) + pi - po + 1
    wp = np.zeros((3, 3, 2 * c, 2 * c))                         # [ky][kxp][out][in]
    for po in range(2):
        for pi in range(2):
            for kxp in range(3):
                dx = 2 * (kxp - 1) + pi - po + 1
                if 0 <= dx <= 2:
                    wp[:, kxp, po * c:(po + 1) * c, pi * c:(pi + 1) * c] = w[:, :, :, dx].transpose(2, 0, 1)
    # the structured zeros k_stem_tower16's K = 8 taps rely on: the left tap reads only the SECOND pixel of the pair (inputs c..2c-1)
    # and feeds only the first output pixel; the right tap reads only the first pixel and feeds only the second output pixel
    assert not wp[:, 0, :, :c].any() and not wp[:, 0, c:, :].any()
    assert not wp[:, 2, :, c:].any() and not wp[:, 2, :c, :].any()
    xp = _pad_pairs(_pairs(x))
    out = np.zeros((H, W // 2, 2 * c))
    for ky in range(3):
        for kxp in range(3):
            out += np.einsum("yxk,ok->yxo", xp[ky:ky + H, kxp:kxp + W // 2], wp[ky, kxp])
    np.testing.assert_allclose(out, _pairs(ref), rtol=1e-12, atol=1e-12)


def test_stride_two_convolution_as_six_taps_on_pair_rows():
    rs = np.random.RandomState(1)
    cin, cout, H, W = 8, 16, 12, 16
    w = rs.randn(cout, cin, 3, 3)
    x = rs.randn(cin, H, W)
    ref = torch.nn.functional.conv2d(torch.tensor(x)[None], torch.tensor(w), stride=2, padding=1)[0].numpy()      # [cout, H/2, W/2]
    # pack_s2_mma: tap 2 ky + 0 = the pair row LEFT of the output pixel (its second pixel is column 2 ox - 1: kx = 0),
    #              tap 2 ky + 1 = the pair row AT the output pixel (columns 2 ox, 2 ox + 1: kx = 1, 2)
    taps = np.zeros((6, cout, 2 * cin))
    for ky in range(3):
        taps[2 * ky, :, cin:] = w[:, :, ky, 0]
        taps[2 * ky + 1, :, :cin] = w[:, :, ky, 1]
        taps[2 * ky + 1, :, cin:] = w[:, :, ky, 2]
    xp = _pad_pairs(_pairs(x))                                    # line / pair-column index + 1
    out = np.zeros((cout, H // 2, W // 2))
    for oy in range(H // 2):
        for ox in range(W // 2):
            for ky in range(3):
                line = 2 * oy + ky - 1 + 1                        # input line 2 oy + ky - 1 in the padded array
                out[:, oy, ox] += taps[2 * ky] @ xp[line, ox - 1 + 1] + taps[2 * ky + 1] @ xp[line, ox + 1]
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)


def test_swapped_mma_roles_give_the_same_product():
    """D[cout][row] = W_tap [cout][cin] x X[row + shift][cin]^T summed over taps (k_stem_tower16 / k_recurrent16) is the
    transposed accumulator of the [row][cout] form - the epilogue's stmatrix.trans undoes the transposition."""
    rs = np.random.RandomState(2)
    rows, pitch = 40, 7
    x = np.zeros((rows + 2 * (pitch + 1), 16))
    x[pitch + 1:-(pitch + 1)] = rs.randn(rows, 16)
    w = rs.randn(9, 16, 16)                                       # [tap][cout][cin]
    a = np.zeros((rows, 16))
    d = np.zeros((16, rows))
    for tap in range(9):
        shift = (tap // 3 - 1) * pitch + (tap % 3 - 1)
        xs = x[pitch + 1 + shift:pitch + 1 + shift + rows]
        a += xs @ w[tap].T
        d += w[tap] @ xs.T
    np.testing.assert_allclose(d.T, a, rtol=1e-12, atol=1e-12)
