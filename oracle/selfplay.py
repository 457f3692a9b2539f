"""CPU restatement of SelfPlay.play_game (self_play.py:110-184) around the oracle search (test infrastructure).

One game at a time, like the reference.  The environment is an oracle.games object with G=1, the network
is any callable pair (initial, recurrent) returning (value, reward, priors, hidden) with Python floats, and
every random draw is injected through the counter RNG of oracle/rng.py so the device path can be compared
move for move."""
import numpy as np

from . import mcts, rng


class History:
    """GameHistory fields (self_play.py:485-495)."""

    def __init__(self):
        self.observation_history, self.action_history, self.reward_history, self.to_play_history = [], [], [], []
        self.child_visits, self.root_values = [], []


def play_game(env, initial, recurrent, cfg, temperature, temperature_threshold, seed, slot, noise_fn, step0=0,
              decode_obs=lambda o: o):
    """noise_fn(step, legal_actions) -> f64 Dirichlet sample or None.  Returns History."""
    h = History()
    obs = env.reset()
    h.action_history.append(0)
    h.observation_history.append(decode_obs(obs[0]))
    h.reward_history.append(0)
    h.to_play_history.append(int(env.to_play()[0]))
    done = False
    A = len(cfg.action_space)
    step = step0
    while not done and len(h.action_history) <= cfg.max_moves:
        legal = env.legal_lists()[0] if hasattr(env, "legal_lists") else list(range(A))
        to_play = int(env.to_play()[0])
        root = initial(obs, legal)
        res = mcts.search(recurrent, root, legal, to_play, n_actions=A, n_players=len(cfg.players),
                          num_simulations=cfg.num_simulations, discount=cfg.discount, pb_c_base=cfg.pb_c_base,
                          pb_c_init=cfg.pb_c_init, noise=noise_fn(step, legal),
                          exploration_fraction=cfg.root_exploration_fraction,
                          tie=lambda n, sim, depth, _s=step: rng.tie_index(seed, slot, _s, sim, depth, n))
        T = temperature if not temperature_threshold or len(h.action_history) < temperature_threshold else 0
        action = mcts.select_action(res.root_actions, res.visits, T, rng.action_uniform(seed, slot, step))
        obs, reward, done_arr = env.step(np.array([action]))
        done = bool(done_arr[0])
        h.child_visits.append(mcts.search_statistics(res.root_actions, res.visits, A))
        h.root_values.append(res.root_value())
        h.action_history.append(action)
        h.observation_history.append(decode_obs(obs[0]))
        h.reward_history.append(float(reward[0]))
        h.to_play_history.append(int(env.to_play()[0]))
        step += 1
    h.steps = step
    return h
