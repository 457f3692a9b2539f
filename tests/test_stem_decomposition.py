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


def test_shared_memory_layout_properties():
    """The bank arithmetic the kernels' layouts rely on: a shared-memory wavefront serves 128 bytes = eight
    16-byte lanes; an ldmatrix 8 x 8 matrix is conflict-free iff its eight row addresses fall into eight different lanes."""
    lane = lambda byte: (byte // 16) % 8
    # 48-byte rows: any 8 CONSECUTIVE rows are conflict-free (both 16-byte halves)
    for r0 in range(64):
        for half in (0, 16):
            assert len({lane(48 * (r0 + i) + half) for i in range(8)}) == 8
    # ... but 8 consecutive POSITIONS of a 6-wide image (pitch 7: one zero-column row per line) are 8 of 9 consecutive rows and
    # always contain rows r and r + 8, which share a lane: the position-mapped form paid 2 wavefronts per matrix
    W, pitch = 6, 7
    row = lambda p: (p // W + 1) * pitch + p % W
    for p0 in range(0, 32, 8):
        rows = [row(p0 + i) for i in range(8)]
        assert rows[-1] - rows[0] == 8 and len({lane(48 * r) for r in rows}) == 7
    # 32-byte rows without padding conflict two-way; swapping the halves in every second group of four rows repairs it
    assert len({lane(32 * i) for i in range(8)}) == 4
    swz = lambda r, h: 32 * r + 16 * (h ^ ((r >> 2) & 1))
    for r0 in range(64):
        for h in (0, 1):
            assert len({lane(swz(r0 + i, h)) for i in range(8)}) == 8
    # transposed action-plane table of k_recurrent16: row length 72 floats; a half-warp's 8-byte loads (4 channels x 4 position
    # pairs) must hit 16 different 8-byte bank pairs
    kTabP = 72
    for base in range(0, 48, 8):
        assert len({((c * kTabP + base + 2 * q) // 2) % 16 for c in range(4) for q in range(4)}) == 16
