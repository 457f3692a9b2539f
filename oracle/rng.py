"""Counter-based RNG shared by the oracle and the CUDA path (test infrastructure).

The reference draws every random number from numpy's *global* Mersenne-Twister
(`numpy.random.choice` for argmax ties self_play.py:372-378 and action sampling
:245, `numpy.random.dirichlet` :474).  A batched device search cannot consume a
single sequential stream, so the B200 path defines its randomness as a pure
function of (seed, game slot, step, stream, index) through Philox4x32-10
(Salmon et al., SC'11 - the published algorithm, restated here), and parity mode
injects the same draws into the reference (tests/golden/make_golden.py patches
`numpy.random.choice` / `numpy.random.dirichlet`).

Counter layout (4 x u32):  c0 = global game slot, c1 = per-slot step counter
(monotone over episodes), c2 = (stream << 16) | sim, c3 = depth / index.
Key (2 x u32) = (seed & 0xffffffff, seed >> 32).
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF

STREAM_TIE = 0      # argmax tie-break inside select_child        (sim, depth)
STREAM_NOISE = 1    # root Dirichlet noise (device generator)     (0, action / draw)
STREAM_ACTION = 2   # temperature sampling of the played action   (0, 0)
STREAM_RESET = 3    # environment reset (cartpole initial state)  (0, component)
STREAM_PAD = 4      # random padding action in make_target        (0, row)
STREAM_RGAME = 5    # replay: game draw of batch element `slot` in batch `step`
STREAM_RPOS = 6     # replay: position draw of batch element `slot` in batch `step`


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Scalar Philox4x32-10 on Python ints. Returns 4 u32."""
    c0 &= MASK; c1 &= MASK; c2 &= MASK; c3 &= MASK; k0 &= MASK; k1 &= MASK
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox4x32_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 on numpy uint64 arrays (values < 2**32)."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) & np.uint64(MASK) for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint64(k0 & MASK); k1 = np.uint64(k1 & MASK)
    m = np.uint64(MASK); s = np.uint64(32)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        c0, c1, c2, c3 = ((p1 >> s) ^ c1 ^ k0) & m, p1 & m, ((p0 >> s) ^ c3 ^ k1) & m, p0 & m
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return c0, c1, c2, c3


def _key(seed):
    return seed & MASK, (seed >> 32) & MASK


def draw(seed, slot, step, stream, sim=0, idx=0):
    k0, k1 = _key(seed)
    return philox4x32(slot, step, (stream << 16) | (sim & 0xFFFF), idx, k0, k1)


def tie_index(seed, slot, step, sim, depth, n_ties):
    """Index into the tied-argmax set (child order): mulhi(u32, n)."""
    r = draw(seed, slot, step, STREAM_TIE, sim, depth)[0]
    return (r * n_ties) >> 32


def u01_double(a, b):
    """53-bit uniform in [0,1) from two u32 (same construction as numpy's random_sample)."""
    return ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0


def action_uniform(seed, slot, step):
    r = draw(seed, slot, step, STREAM_ACTION)
    return u01_double(r[0], r[1])


def reset_uniforms(seed, slot, step, n=4):
    """n doubles in [0,1) for an environment reset (one Philox call per pair)."""
    out = []
    for i in range((n + 1) // 2):
        r = draw(seed, slot, step, STREAM_RESET, 0, i)
        out += [u01_double(r[0], r[1]), u01_double(r[2], r[3])]
    return out[:n]


def pad_action(seed, slot, step, row, n_actions):
    r = draw(seed, slot, step, STREAM_PAD, 0, row)[0]
    return (r * n_actions) >> 32


def replay_uniform(seed, element, batch, stream):
    """float64 uniform of a replay draw: batch element `element` of the `batch`-th get_batch call."""
    r = draw(seed, element, batch, stream)
    return u01_double(r[0], r[1])
