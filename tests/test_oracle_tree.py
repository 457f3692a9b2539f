"""The CPU oracle reproduces the reference MCTS bit-for-bit on the committed golden searches."""
import numpy as np
import pytest

import _tables as T

SHAPES = ["cartpole", "tictactoe", "connect4", "gomoku", "breakout", "flat"]


@pytest.mark.parametrize("name", SHAPES)
def test_oracle_matches_reference_tree(name):
    z = T.load("tree")
    shape = T.Shape(z, name)
    for i in range(shape.n_cases):
        case = T.Case(z, name, i)
        res = T.run_oracle(shape, case)
        assert res.root_actions == case.legal
        np.testing.assert_array_equal(np.array(res.visits, dtype=np.int32), case.visits)
        # float64 scalars: bit-exact, not approximately equal
        assert np.array(res.value_sums, dtype=np.float64).tobytes() == case.value_sums.tobytes()
        assert np.array(res.rewards, dtype=np.float64).tobytes() == case.rewards.tobytes()
        assert np.array(res.priors, dtype=np.float64).tobytes() == case.priors.tobytes()
        assert res.root_visit == int(case.root[0]) == shape.sims
        assert np.float64(res.root_value_sum).tobytes() == np.float64(case.root[1]).tobytes()
        assert np.float64(res.root_value()).tobytes() == np.float64(case.root[2]).tobytes()
        assert res.max_tree_depth == int(case.root[3])
        dfs = T.oracle_dfs(res, shape.A)
        assert dfs.shape == case.dfs.shape
        assert dfs.tobytes() == case.dfs.tobytes()
        assert sum(res.visits) == shape.sims


def test_flat_tables_exercise_ties():
    """The all-equal table must produce ties beyond the first simulation (tie rule is exercised)."""
    z = T.load("tree")
    shape = T.Shape(z, "flat")
    case = T.Case(z, "flat", 3)      # a no-noise case
    assert case.noise is None
    seen = []
    from oracle import mcts, rng

    def tie(n, sim, depth):
        seen.append((n, sim, depth))
        return rng.tie_index(T.SEED, case.slot, case.step, sim, depth, n)

    def rec(h, a):
        r = T.child_row(h, a)
        return float(shape.V[r]), float(shape.Rw[r]), [float(p) for p in shape.P[r]], r

    root = (0.0, 0.0, [float(p) for p in case.root_priors], case.root_row)
    mcts.search(rec, root, case.legal, case.to_play, n_actions=shape.A, n_players=2, num_simulations=shape.sims,
                discount=1, pb_c_base=T.PB_C_BASE, pb_c_init=T.PB_C_INIT, noise=None, tie=tie)
    assert any(sim > 0 and depth > 0 for _, sim, depth in seen)
