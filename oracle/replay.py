"""CPU restatement of the reference ReplayBuffer (test infrastructure; see oracle/__init__.py).

Follows /root/reference/replay_buffer.py: save_game :33-64 (initial PER priorities), get_batch :69-140,
sample_n_games :162-177, sample_position :179-192, update_priorities :202-220; targets come from
oracle/targets.py (:222-295).  Randomness is injected: `numpy.random.choice(ids, n, p=probs)` and
`numpy.random.choice(n, p=probs)` are inverse-CDF draws on one float64 uniform each (numpy's legacy
`RandomState.choice`: cdf = p.astype(float64).cumsum(); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')), and the
uniform (p=None) forms are index = floor(u * n).  The float32 arithmetic of the reference is kept operation by
operation: game probabilities divide by `numpy.sum` of a float32 array (numpy's pairwise summation, restated
in `np_sum_f32` and checked against numpy itself), position probabilities by Python's left-to-right `sum`.
Pinned by tests/golden/replay.npz (outputs of the unmodified reference class).
"""
import numpy as np

from . import targets

F32 = np.float32


def _pairwise_f32(a, lo, n):
    """numpy's FLOAT_pairwise_sum on a[lo:lo+n] (float32 accumulators, PW_BLOCKSIZE 128, unroll 8)."""
    if n < 8:
        res = F32(0.0)
        for i in range(n):
            res = F32(res + a[lo + i])
        return res
    if n <= 128:
        r = [a[lo + j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = F32(r[j] + a[lo + i + j])
            i += 8
        res = F32(F32(F32(r[0] + r[1]) + F32(r[2] + r[3])) + F32(F32(r[4] + r[5]) + F32(r[6] + r[7])))
        while i < n:
            res = F32(res + a[lo + i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return F32(_pairwise_f32(a, lo, n2) + _pairwise_f32(a, lo + n2, n - n2))


def np_sum_f32(a):
    """numpy.sum of a contiguous float32 vector = the pairwise sum of all n elements (checked against numpy for
    n = 1 .. 2e5 in tests/test_oracle_replay.py; the device kernel restates the same order)."""
    a = np.asarray(a, dtype=F32)
    return _pairwise_f32(a, 0, len(a))


def choice_index(probs, u):
    """numpy.random.choice(len(probs), p=probs) for the injected uniform u."""
    cdf = np.zeros(len(probs), dtype=np.float64)
    acc = 0.0
    for i, p in enumerate(probs):
        acc = acc + float(p)               # float64 cumsum, sequential
        cdf[i] = acc
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


class Game:
    """GameHistory fields (self_play.py:485-495) as plain lists."""

    def __init__(self, observation_history, action_history, reward_history, to_play_history, child_visits, root_values):
        self.observation_history, self.action_history = observation_history, action_history
        self.reward_history, self.to_play_history = reward_history, to_play_history
        self.child_visits, self.root_values = child_visits, root_values
        self.priorities = None
        self.game_priority = None


class ReplayBuffer:
    def __init__(self, PER, PER_alpha, replay_buffer_size, batch_size, num_unroll_steps, td_steps, discount, n_actions):
        self.PER, self.alpha, self.size, self.batch_size = PER, PER_alpha, replay_buffer_size, batch_size
        self.K, self.td, self.discount, self.A = num_unroll_steps, td_steps, discount, n_actions
        self.buffer = {}                   # insertion-ordered: game_id -> Game
        self.num_played_games = 0
        self.num_played_steps = 0
        self.total_samples = 0

    def _target_value(self, g, i):
        return targets.compute_target_value(g.root_values, g.reward_history, g.to_play_history, i, self.td, self.discount)

    def save_game(self, g):
        if self.PER:
            pr = [np.abs(rv - self._target_value(g, i)) ** self.alpha for i, rv in enumerate(g.root_values)]
            g.priorities = np.array(pr, dtype="float32")
            g.game_priority = np.max(g.priorities)
        self.buffer[self.num_played_games] = g
        self.num_played_games += 1
        self.num_played_steps += len(g.root_values)
        self.total_samples += len(g.root_values)
        if self.size < len(self.buffer):
            del_id = self.num_played_games - len(self.buffer)
            self.total_samples -= len(self.buffer[del_id].root_values)
            del self.buffer[del_id]

    def sample(self, u_game, u_pos):
        """[(game_id, game_prob | None, position, position_prob | None)] for the injected uniforms."""
        ids = list(self.buffer.keys())
        out = []
        if self.PER:
            gp = np.array([self.buffer[i].game_priority for i in ids], dtype="float32")
            gp = gp / np_sum_f32(gp)
            cdf = np.zeros(len(gp), dtype=np.float64)
            acc = 0.0
            for i, p in enumerate(gp):
                acc = acc + float(p)
                cdf[i] = acc
            cdf /= cdf[-1]
        for b in range(len(u_game)):
            if self.PER:
                k = int(np.searchsorted(cdf, u_game[b], side="right"))
                g = self.buffer[ids[k]]
                tot = F32(0.0)
                for p in g.priorities:
                    tot = F32(tot + p)                       # Python sum() over float32 scalars
                pp = (g.priorities / tot).astype(F32)
                pos = choice_index(pp, u_pos[b])
                out.append((ids[k], gp[k], pos, pp[pos]))
            else:
                k = int(u_game[b] * len(ids))
                g = self.buffer[ids[k]]
                out.append((ids[k], None, int(u_pos[b] * len(g.root_values)), None))
        return out

    def get_batch(self, u_game, u_pos, pad_action=lambda b, row: 0):
        index, obs, act, val, rew, pol, gs, w = [], [], [], [], [], [], [], []
        for b, (gid, gprob, pos, pprob) in enumerate(self.sample(u_game, u_pos)):
            g = self.buffer[gid]
            past = iter(range(self.K + 1))                   # the k-th past-the-end row draws pad index k
            v, r, p, a = targets.make_target(g.root_values, g.reward_history, g.to_play_history, g.child_visits,
                                             g.action_history, pos, self.K, self.td, self.discount, self.A,
                                             pad_action=lambda row, _b=b, _p=past: pad_action(_b, next(_p)))
            index.append([gid, pos])
            obs.append(np.asarray(g.observation_history[pos]))
            act.append(a); val.append(v); rew.append(r); pol.append(p)
            gs.append([min(self.K, len(g.action_history) - pos)] * len(a))
            if self.PER:
                w.append(F32(1) / F32(F32(F32(self.total_samples) * gprob) * pprob))
        weights = None
        if self.PER:
            weights = np.array(w, dtype="float32") / max(w)
        return index, (obs, act, val, rew, pol, weights, gs)

    def update_priorities(self, priorities, index_info):
        for i, (gid, pos) in enumerate(index_info):
            if next(iter(self.buffer)) <= gid:
                g = self.buffer[gid]
                pr = priorities[i, :]
                end = min(pos + len(pr), len(g.priorities))
                g.priorities[pos:end] = pr[:end - pos]
                g.game_priority = np.max(g.priorities)
