"""CPU oracle for the muzero-hypermodel self-play hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`muzero_hypermodel_b200/`) may import this package; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs use it, and only as the checker or the timed CPU baseline.

It restates, on the CPU, the algorithms of the reference hot path
(`/root/reference`, file:line cited per function):

* `oracle.rng`      - the counter-based RNG both sides share (our definition, no
                      reference counterpart: the reference draws from numpy's
                      global Mersenne-Twister, which parity mode replaces by
                      injection, see DESIGN.md "Randomness").
* `oracle.mcts`     - MCTS.run / select_child / ucb_score / backpropagate /
                      Node.expand / add_exploration_noise / MinMaxStats
                      (self_play.py:261-477, 551-568), select_action and
                      store_search_statistics (self_play.py:223-246, 497-512).
* `oracle.games`    - TicTacToe / Connect4 / Gomoku boards (games/*.py) and the
                      gym CartPole-v1 physics the cartpole wrapper calls into.
* `oracle.networks` - FC and residual MuZero networks (models.py:80-619) and the
                      support codec (models.py:641-685).
* `oracle.targets`  - compute_target_value / make_target (replay_buffer.py:222-295).
* `oracle.selfplay` - play_game episode loop (self_play.py:110-184).
* `oracle.replay`   - ReplayBuffer save_game / get_batch / update_priorities
                      (replay_buffer.py:33-220), incl. numpy's float32 pairwise sum.
* `oracle.cpu_baseline` - the port timed as the CPU baseline (one process per host
                      core, batch-1 numpy network).

Parity pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the
build container by `tests/golden/make_golden.py` (imports the unmodified
reference through `oracle.ref_loader`) and committed under `tests/golden/`.
The one unpinned boundary is gym's CartPole physics and ALE (third-party,
unpinned, absent): "parity unpinned" there, see DESIGN.md.
"""
