"""Helpers shared by the tree parity tests: the table-driven fake network of tests/golden/tree.npz."""
import os

import numpy as np

from oracle import mcts, rng

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_ROWS = 2048
SEED = 20261018
PB_C_BASE, PB_C_INIT, FRAC = 19652, 1.25, 0.25


def child_row(parent_row, action, n_rows=N_ROWS):
    """Same integer hash as tests/golden/make_golden.py:child_row."""
    x = (parent_row * 0x9E3779B1 + (action + 1) * 0x85EBCA77) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x2C1B3C6D) & 0xFFFFFFFF
    x ^= x >> 12
    return x % n_rows


def child_row_np(parent_row, action, n_rows=N_ROWS):
    x = (parent_row.astype(np.uint64) * np.uint64(0x9E3779B1) + (action.astype(np.uint64) + np.uint64(1)) * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(12)
    return (x % np.uint64(n_rows)).astype(np.int64)


_cache = {}


def load(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
    return _cache[name]


class Shape:
    def __init__(self, z, name):
        A, players, sims, discount, alpha, n_cases = z[f"{name}/meta"]
        self.name = name
        self.A, self.players, self.sims, self.n_cases = int(A), int(players), int(sims), int(n_cases)
        self.discount = int(discount) if float(discount).is_integer() else float(discount)
        self.alpha = float(alpha)
        self.V, self.Rw, self.L, self.P = z[f"{name}/V"], z[f"{name}/Rw"], z[f"{name}/L"], z[f"{name}/P"]


class Case:
    def __init__(self, z, name, i):
        pre = f"{name}/{i}/"
        self.root_row = int(z[pre + "root_row"])
        self.legal = z[pre + "legal"].tolist()
        self.to_play = int(z[pre + "to_play"])
        n = z[pre + "noise"]
        self.noise = n if len(n) else None
        self.slot, self.step = [int(v) for v in z[pre + "slot_step"]]
        self.root_priors = z[pre + "root_priors_f32"]
        self.visits = z[pre + "visits"]
        self.value_sums = z[pre + "value_sums"]
        self.rewards = z[pre + "rewards"]
        self.priors = z[pre + "priors"]
        self.root = z[pre + "root"]          # visit, value_sum, value(), max_depth, root_predicted_value
        self.dfs = z[pre + "dfs"]


def run_oracle(shape, case):
    """oracle.mcts.search on a golden case, fed by the committed tables."""

    def recurrent(hidden_row, action):
        r = child_row(hidden_row, action)
        return float(shape.V[r]), float(shape.Rw[r]), [float(p) for p in shape.P[r]], r

    root = (float(shape.V[case.root_row]), float(shape.Rw[case.root_row]),
            [float(p) for p in case.root_priors], case.root_row)
    return mcts.search(
        recurrent, root, case.legal, case.to_play,
        n_actions=shape.A, n_players=shape.players, num_simulations=shape.sims, discount=shape.discount,
        pb_c_base=PB_C_BASE, pb_c_init=PB_C_INIT,
        noise=None if case.noise is None else [float(x) for x in case.noise], exploration_fraction=FRAC,
        tie=lambda n, sim, depth: rng.tie_index(SEED, case.slot, case.step, sim, depth, n))


def oracle_dfs(res, n_actions):
    """Canonical DFS dump of an oracle tree, same row format as make_golden.dfs_dump."""
    t = res.tree
    rows = []

    def walk(node, depth):
        for a in range(n_actions):
            if not t.exists[node][a]:
                continue
            ch = t.child[node][a]
            rows.append((depth + 1, a, t.visit[node][a], t.value_sum[node][a], t.reward[node][a], t.prior[node][a],
                         1 if ch >= 0 else 0))
            if ch >= 0:
                walk(ch, depth + 1)

    walk(0, 0)
    return np.array(rows, dtype=np.float64).reshape(-1, 7)
