"""CUDA tree kernels (through the C ABI) vs the reference's golden searches: bit-exact."""
import numpy as np
import pytest
import torch

import _tables as T

pytestmark = pytest.mark.gpu

SHAPES = ["cartpole", "tictactoe", "connect4", "gomoku", "breakout", "flat"]


def child_row_torch(parent_row, action):
    x = (parent_row * 0x9E3779B1 + (action + 1) * 0x85EBCA77) & 0xFFFFFFFF
    x = x ^ (x >> 15)
    x = (x * 0x2C1B3C6D) & 0xFFFFFFFF
    x = x ^ (x >> 12)
    return x % T.N_ROWS


def dfs_from_export(ex, A):
    rows = []

    def walk(node, depth):
        for a in range(A):
            ch = int(ex["child"][node][a])
            if ch == -2:
                continue
            prior = ex["root_prior"][a] if node == 0 else float(ex["prior"][node][a])
            rows.append((depth + 1, a, int(ex["visit"][node][a]), float(ex["value_sum"][node][a]),
                         float(ex["reward"][node][a]), prior, 1 if ch >= 0 else 0))
            if ch >= 0:
                walk(ch, depth + 1)

    walk(0, 0)
    return np.array(rows, dtype=np.float64).reshape(-1, 7)


def run_batch(shape, cases, with_noise, replicate=1):
    from muzero_hypermodel_b200.tree import BatchedTree
    dev = torch.device("cuda:0")
    cases = cases * replicate
    G, A, S = len(cases), shape.A, shape.sims
    tree = BatchedTree(G, A, S, shape.players, shape.discount, T.PB_C_BASE, T.PB_C_INIT, seed=T.SEED, device=dev)
    pri = np.zeros((G, A), dtype=np.float32)
    legal = np.zeros((G, A), dtype=np.uint8)
    noise = np.zeros((G, A), dtype=np.float64)
    for g, c in enumerate(cases):
        pri[g, c.legal] = c.root_priors
        legal[g, c.legal] = 1
        if with_noise:
            noise[g, c.legal] = c.noise
    V = torch.tensor(shape.V, device=dev); Rw = torch.tensor(shape.Rw, device=dev); P = torch.tensor(shape.P, device=dev)
    rows = torch.zeros((G, S + 1), dtype=torch.int64, device=dev)
    rows[:, 0] = torch.tensor([c.root_row for c in cases], device=dev)
    root_reward = Rw[rows[:, 0]].contiguous()
    tree.root_init(root_reward, torch.tensor(pri, device=dev), policy_is_logits=False,
                   legal=torch.tensor(legal, device=dev),
                   to_play=torch.tensor([c.to_play for c in cases], dtype=torch.int8, device=dev),
                   noise=torch.tensor(noise, device=dev) if with_noise else None, alpha=shape.alpha,
                   frac=T.FRAC if with_noise else 0.0,
                   slot=torch.tensor([c.slot for c in cases], dtype=torch.int64, device=dev).to(torch.int32),
                   step=torch.tensor([c.step for c in cases], dtype=torch.int64, device=dev).to(torch.int32))
    parent = torch.empty(G, dtype=torch.int32, device=dev)
    action = torch.empty(G, dtype=torch.int32, device=dev)
    depth = torch.empty(G, dtype=torch.int32, device=dev)
    ar = torch.arange(G, device=dev)
    for sim in range(S):
        tree.select(parent, action, depth)
        crow = child_row_torch(rows[ar, parent.long()], action.long())
        rows[:, sim + 1] = crow
        tree.expand_backup(V[crow].contiguous(), Rw[crow].contiguous(), P[crow].contiguous(), policy_is_logits=False)
    stats = {k: v.cpu().numpy() for k, v in tree.root_stats(full=True).items()}
    return tree, stats, cases


@pytest.mark.parametrize("name", SHAPES)
def test_tree_kernels_match_reference(name):
    z = T.load("tree")
    shape = T.Shape(z, name)
    all_cases = [T.Case(z, name, i) for i in range(shape.n_cases)]
    for with_noise in (True, False):
        cases = [c for c in all_cases if (c.noise is not None) == with_noise]
        if not cases:
            continue
        tree, st, cases = run_batch(shape, cases, with_noise)
        for g, c in enumerate(cases):
            np.testing.assert_array_equal(st["visits"][g][c.legal], c.visits)
            assert st["visits"][g].sum() == shape.sims
            assert st["child_value_sum"][g][c.legal].tobytes() == c.value_sums.tobytes()
            assert st["child_reward"][g][c.legal].astype(np.float64).tobytes() == c.rewards.tobytes()
            assert st["child_prior"][g][c.legal].tobytes() == c.priors.tobytes()
            assert np.float64(st["root_value"][g]).tobytes() == np.float64(c.root[2]).tobytes()
            assert st["max_depth"][g] == int(c.root[3])
            ex = tree.export_game(g)
            assert ex["root_visit"] == shape.sims
            assert np.float64(ex["root_value_sum"]).tobytes() == np.float64(c.root[1]).tobytes()
            dfs = dfs_from_export(ex, shape.A)
            assert dfs.shape == c.dfs.shape
            assert dfs.tobytes() == c.dfs.tobytes()


def test_tree_large_batch_is_replica_invariant():
    """4096 games = the golden cases tiled: every replica must reproduce its case (no cross-game leakage)."""
    z = T.load("tree")
    shape = T.Shape(z, "tictactoe")
    cases = [c for c in (T.Case(z, "tictactoe", i) for i in range(shape.n_cases)) if c.noise is not None]
    rep = 4096 // len(cases) + 1
    tree, st, cases = run_batch(shape, cases, True, replicate=rep)
    for g, c in enumerate(cases):
        np.testing.assert_array_equal(st["visits"][g][c.legal], c.visits)
        assert np.float64(st["root_value"][g]).tobytes() == np.float64(c.root[2]).tobytes()


def test_logits_mode_and_device_noise():
    """policy_is_logits=1: device softmax within 1e-6 of torch; device Dirichlet noise is a distribution."""
    from muzero_hypermodel_b200.tree import BatchedTree
    dev = torch.device("cuda:0")
    G, A, S = 512, 9, 25
    g = torch.Generator(device="cpu").manual_seed(3)
    logits = torch.randn(G, A, generator=g) * 2
    legal = (torch.rand(G, A, generator=g) < 0.7)
    legal[:, 0] = True
    tree = BatchedTree(G, A, S, 2, 1.0, T.PB_C_BASE, T.PB_C_INIT, seed=5, device=dev)
    tree.root_init(torch.zeros(G, device=dev), logits.to(dev), True, legal.to(torch.uint8).to(dev), None, None, 0.3, 0.0)
    pri = tree.root_stats(full=True)["child_prior"].cpu()
    ref = torch.softmax(logits.masked_fill(~legal, -float("inf")), dim=1).double()
    torch.testing.assert_close(pri, ref, rtol=1e-5, atol=1e-7)
    tree.root_init(torch.zeros(G, device=dev), logits.to(dev), True, legal.to(torch.uint8).to(dev), None, None, 0.3, 1.0)
    noise = tree.root_stats(full=True)["child_prior"].cpu()        # frac=1 -> pure noise
    assert torch.all(noise[~legal] == 0)
    torch.testing.assert_close(noise.sum(1), torch.ones(G, dtype=torch.float64), rtol=0, atol=1e-12)
    assert (noise >= 0).all()
    # Dirichlet(0.3) marginals: mean 1/k over the k legal actions, and sparse (most mass on one action)
    k = legal.sum(1).double()
    mean_err = ((noise.sum(0) / G) - (legal.double() / k[:, None]).sum(0) / G).abs().max()
    assert mean_err < 0.05
    assert (noise.max(1).values.mean() > 0.5)


def test_more_than_two_players_rejected():
    from muzero_hypermodel_b200.tree import BatchedTree
    with pytest.raises(NotImplementedError):
        BatchedTree(4, 3, 5, 3, 1.0, 19652, 1.25)


@pytest.mark.parametrize("A,k_legal,alpha", [(2, 2, 0.25), (7, 7, 0.3), (9, 5, 0.1), (121, 121, 0.3), (121, 40, 0.3)])
def test_device_dirichlet_matches_numpy_dirichlet(A, k_legal, alpha):
    """Device-generated exploration noise (Marsaglia-Tsang Gamma from Philox, the mode every bench run uses) against
    `numpy.random.dirichlet([alpha] * k)` (self_play.py:472-477): exact Beta(alpha, (k-1) alpha) marginals by a
    one-sample Kolmogorov-Smirnov test, a two-sample KS test against numpy's own draws, and the first two moments
    incl. the negative covariance between components."""
    from scipy import stats
    from muzero_hypermodel_b200.tree import BatchedTree
    dev = torch.device("cuda:0")
    G = 40000
    legal = np.zeros((G, A), dtype=np.uint8)
    legal[:, np.sort(np.random.RandomState(A).permutation(A)[:k_legal])] = 1
    cols = np.nonzero(legal[0])[0]
    tree = BatchedTree(G, A, 4, 1, 0.997, T.PB_C_BASE, T.PB_C_INIT, seed=77, device=dev)
    pri = torch.full((G, A), 1.0 / A, device=dev)
    tree.root_init(torch.zeros(G, device=dev), pri, False, torch.tensor(legal, device=dev), None, None, alpha, 1.0)
    noise = tree.root_stats(full=True)["child_prior"].cpu().numpy()          # frac = 1: the root priors ARE the noise
    assert (noise[:, legal[0] == 0] == 0).all() and np.allclose(noise.sum(1), 1.0, atol=1e-12)
    x = noise[:, cols]
    k = k_legal
    ref = np.random.RandomState(5).dirichlet([alpha] * k, size=G)
    # moments of Dirichlet(alpha 1_k): mean 1/k, var (k-1)/(k^2 (k alpha + 1)), cov -1/(k^2 (k alpha + 1))
    var = (k - 1) / (k * k * (k * alpha + 1))
    se_mean = np.sqrt(var / G)
    assert np.abs(x.mean(0) - 1.0 / k).max() < 5 * se_mean + 1e-12
    assert abs(x.var(0).mean() - var) < 0.05 * var
    if k > 1:
        cov01 = np.cov(x[:, 0], x[:, 1])[0, 1]
        assert abs(cov01 + 1 / (k * k * (k * alpha + 1))) < 0.15 / (k * k * (k * alpha + 1)) + 4 * var / np.sqrt(G)
    for j in (0, k // 2, k - 1):
        # tiny components underflow towards 0 in both generators; compare on the resolvable range
        p1 = stats.kstest(x[:, j], stats.beta(alpha, (k - 1) * alpha).cdf).pvalue if k > 1 else 1.0
        p2 = stats.ks_2samp(x[:, j], ref[:, j]).pvalue
        assert p1 > 1e-4 and p2 > 1e-4, (j, p1, p2)
    # different (slot, step) counters give different samples; the same counters reproduce them
    tree.root_init(torch.zeros(G, device=dev), pri, False, torch.tensor(legal, device=dev), None, None, alpha, 1.0)
    again = tree.root_stats(full=True)["child_prior"].cpu().numpy()
    assert again.tobytes() == noise.tobytes()
    step = torch.full((G,), 3, dtype=torch.int32, device=dev)
    tree.root_init(torch.zeros(G, device=dev), pri, False, torch.tensor(legal, device=dev), None, None, alpha, 1.0, None, step)
    assert not np.array_equal(tree.root_stats(full=True)["child_prior"].cpu().numpy(), noise)


def test_ddiv_rcp_equals_ddiv_rn():
    """The whole-search kernel divides value_sum by visit counts and Q values by the MinMaxStats range through a
    hoisted reciprocal (csrc/mzb_common.cuh).  It must be the IEEE quotient bit for bit - also at zeros, signed
    values, tiny / huge magnitudes (where it defers to the full division) and for every visit count up to 400."""
    import ctypes as C
    from muzero_hypermodel_b200 import _lib
    from muzero_hypermodel_b200._lib import check, ptr
    _lib.bind("mzb_debug_ddiv_rcp", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p])
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(12)
    n = 1 << 22
    parts_a, parts_b = [], []
    # value sums / visit counts
    parts_a.append(rs.standard_normal(n) * 10.0 ** rs.uniform(-3, 3, n)); parts_b.append(rs.randint(1, 401, n).astype(np.float64))
    # (q - min) / (max - min): both arbitrary positive / signed doubles over a wide exponent range
    parts_a.append(rs.standard_normal(n) * 10.0 ** rs.uniform(-8, 8, n)); parts_b.append(np.abs(rs.standard_normal(n)) * 10.0 ** rs.uniform(-8, 8, n) + 1e-300)
    # mantissa patterns that stress the last rounding step: a = k * b +- 1 ulp
    b3 = rs.uniform(0.5, 2.0, n); k3 = rs.randint(1, 1 << 20, n).astype(np.float64)
    parts_a.append(np.nextafter(k3 * b3, np.where(rs.rand(n) < 0.5, np.inf, -np.inf))); parts_b.append(b3)
    # edge cases: zeros, signed zeros, subnormal / tiny / huge numerators and divisors
    edge_a = np.array([0.0, -0.0, 5e-324, -5e-324, 1e-310, 1e-200, 1e200, 1.7e308, -1.7e308, 1.0, 3.0, 1e-130, 1e130])
    edge_b = np.array([1.0, 3.0, 7.0, 50.0, 1e-310, 1e-200, 1e-130, 1e130, 1e200, 1.7e308, 0.1])
    ea, eb = np.meshgrid(edge_a, edge_b)
    parts_a.append(ea.ravel()); parts_b.append(eb.ravel())
    a = torch.tensor(np.concatenate(parts_a), device=dev)
    b = torch.tensor(np.concatenate(parts_b), device=dev)
    fast, ref = torch.empty_like(a), torch.empty_like(a)
    check(_lib.lib.mzb_debug_ddiv_rcp(ptr(a), ptr(b), a.numel(), ptr(fast), ptr(ref), _lib.current_stream()))
    torch.cuda.synchronize()
    fa, re = fast.cpu().numpy(), ref.cpu().numpy()
    assert fa.tobytes() == re.tobytes(), int((fa.view(np.uint64) != re.view(np.uint64)).sum())
    host = (a.cpu().numpy() / b.cpu().numpy())                        # numpy float64 division = Python's
    assert re.tobytes() == host.tobytes()
