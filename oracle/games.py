"""CPU restatement of the reference environments, vectorised over G games (test infrastructure).

Follows /root/reference/games:
  TicTacToe  tictactoe.py:242-305 (+ wrapper reward x20 :143)
  Connect4   connect4.py:219-304  (+ wrapper reward x10 :143)
  Gomoku     gomoku.py:219-291    (+ wrapper reward x1  :149)
  cartpole   cartpole.py:130-173 wraps gym's CartPole-v1; the physics below restates gym's
             classic-control CartPoleEnv (third-party, unpinned in requirements.txt:4 and
             absent here -> "parity unpinned" at that boundary; this file IS the golden
             definition).

The reference classes hold one game each and loop in Python; these hold int32[G, ...] boards
and use whole-array integer arithmetic.  Pinned against the reference classes by
tests/golden/env_*.npz (random action sequences replayed through the reference).
"""
import math

import numpy as np

from . import rng


class _BoardGame:
    H = W = 0
    N_ACTIONS = 0
    REWARD_SCALE = 1

    def __init__(self, n_games):
        self.G = n_games
        self.reset()

    def reset(self, mask=None):
        if mask is None:
            self.board = np.zeros((self.G, self.H, self.W), dtype=np.int32)
            self.player = np.ones(self.G, dtype=np.int32)
        else:
            self.board[mask] = 0
            self.player[mask] = 1
        return self.observation()

    def to_play(self):
        return np.where(self.player == 1, 0, 1).astype(np.int32)

    def observation(self):
        """[G, 3, H, W] float32: own-1 plane, own-(-1) plane, to-play plane (+1 / -1)."""
        p1 = (self.board == 1).astype(np.float32)
        p2 = (self.board == -1).astype(np.float32)
        tp = np.broadcast_to(self.player[:, None, None].astype(np.float32), self.board.shape)
        return np.stack([p1, p2, tp], axis=1)

    def legal_lists(self):
        m = self.legal_mask()
        return [np.nonzero(row)[0].tolist() for row in m]


def _run_of(board, player, length, dx, dy):
    """bool[G]: does `player[g]` own `length` consecutive cells in direction (dx, dy) anywhere."""
    G, H, W = board.shape
    own = board == player[:, None, None]
    xs = range(max(0, -dx * (length - 1)), min(H, H - dx * (length - 1)))
    ys = range(max(0, -dy * (length - 1)), min(W, W - dy * (length - 1)))
    if len(xs) == 0 or len(ys) == 0:
        return np.zeros(G, dtype=bool)
    x0, x1 = xs[0], xs[-1] + 1
    y0, y1 = ys[0], ys[-1] + 1
    acc = np.ones((G, x1 - x0, y1 - y0), dtype=bool)
    for k in range(length):
        acc &= own[:, x0 + k * dx:x1 + k * dx, y0 + k * dy:y1 + k * dy]
    return acc.reshape(G, -1).any(axis=1)


class TicTacToe(_BoardGame):
    H = W = 3
    N_ACTIONS = 9
    REWARD_SCALE = 20

    def legal_mask(self):
        return (self.board.reshape(self.G, 9) == 0)

    def step(self, actions):
        a = np.asarray(actions, dtype=np.int64)
        g = np.arange(self.G)
        self.board[g, a // 3, a % 3] = self.player                      # :256-258 (overwrites)
        win = np.zeros(self.G, dtype=bool)
        for dx, dy in ((0, 1), (1, 0), (1, 1), (1, -1)):                # :276-299, mover only
            win |= _run_of(self.board, self.player, 3, dx, dy)
        done = win | ~self.legal_mask().any(axis=1)                     # :260
        reward = win.astype(np.float64) * self.REWARD_SCALE             # :262, wrapper :143
        self.player = -self.player
        return self.observation(), reward, done


class Connect4(_BoardGame):
    H, W = 6, 7
    N_ACTIONS = 7
    REWARD_SCALE = 10

    def legal_mask(self):
        return self.board[:, 5, :] == 0                                  # :249-254

    def step(self, actions):
        a = np.asarray(actions, dtype=np.int64)
        g = np.arange(self.G)
        col = self.board[g, :, a]                                        # [G, 6], row 0 = bottom
        free = col == 0
        has = free.any(axis=1)                                           # full column: no-op :233-236
        row = np.argmax(free, axis=1)
        gi = g[has]
        self.board[gi, row[has], a[has]] = self.player[has]
        win = np.zeros(self.G, dtype=bool)
        for dx, dy in ((0, 1), (1, 0), (1, 1), (-1, 1)):                 # :259-304, mover only
            win |= _run_of(self.board, self.player, 4, dx, dy)
        done = win | ~self.legal_mask().any(axis=1)
        reward = win.astype(np.float64) * self.REWARD_SCALE
        self.player = -self.player
        return self.observation(), reward, done


class Gomoku(_BoardGame):
    H = W = 11
    N_ACTIONS = 121
    REWARD_SCALE = 1

    def legal_mask(self):
        return self.board.reshape(self.G, 121) == 0

    def step(self, actions):
        a = np.asarray(actions, dtype=np.int64)
        g = np.arange(self.G)
        self.board[g, a // 11, a % 11] = self.player                     # :233-236
        fin = np.zeros(self.G, dtype=bool)
        for colour in (1, -1):                                           # is_finished: EITHER colour :255-283
            c = np.full(self.G, colour, dtype=np.int32)
            for dx, dy in ((1, -1), (1, 0), (1, 1), (0, 1)):
                fin |= _run_of(self.board, c, 5, dx, dy)
        done = fin | ~self.legal_mask().any(axis=1)                      # or board full :284
        reward = done.astype(np.float64) * self.REWARD_SCALE             # 1 if done (also a full-board draw) :241
        self.player = -self.player
        return self.observation(), reward, done


# ----------------------------------------------------------------------------- cartpole

GRAVITY = 9.8
MASSCART = 1.0
MASSPOLE = 0.1
TOTAL_MASS = MASSPOLE + MASSCART
LENGTH = 0.5
POLEMASS_LENGTH = MASSPOLE * LENGTH
FORCE_MAG = 10.0
TAU = 0.02
THETA_THRESHOLD = 12 * 2 * math.pi / 360
X_THRESHOLD = 2.4
MAX_EPISODE_STEPS = 500


def cartpole_physics(x, x_dot, theta, theta_dot, action):
    """One explicit-Euler step of gym's CartPoleEnv in float64 (scalar Python floats)."""
    force = FORCE_MAG if action == 1 else -FORCE_MAG
    costheta = math.cos(theta)
    sintheta = math.sin(theta)
    temp = (force + POLEMASS_LENGTH * theta_dot ** 2 * sintheta) / TOTAL_MASS
    thetaacc = (GRAVITY * sintheta - costheta * temp) / (
        LENGTH * (4.0 / 3.0 - MASSPOLE * costheta ** 2 / TOTAL_MASS))
    xacc = temp - POLEMASS_LENGTH * thetaacc * costheta / TOTAL_MASS
    x = x + TAU * x_dot
    x_dot = x_dot + TAU * xacc
    theta = theta + TAU * theta_dot
    theta_dot = theta_dot + TAU * thetaacc
    return x, x_dot, theta, theta_dot


class CartPoleV1:
    """Single CartPole-v1 episode stream (what `gym.make("CartPole-v1")` gives games/cartpole.py).

    Reset draws U(-0.05, 0.05)^4 from the shared counter RNG keyed by (seed, slot, steps so far)
    instead of gym's private generator, so the device path can start from the same states.
    """

    def __init__(self, seed=None, slot=0):
        self.seed = 0 if seed is None else int(seed)
        self.slot = slot
        self.total_steps = 0
        self.state = None
        self.elapsed = 0

    def reset(self):
        u = rng.reset_uniforms(self.seed, self.slot, self.total_steps, 4)
        self.state = tuple(-0.05 + 0.1 * v for v in u)
        self.elapsed = 0
        return np.array(self.state, dtype=np.float32)

    def step(self, action):
        x, x_dot, theta, theta_dot = cartpole_physics(*self.state, int(action))
        self.state = (x, x_dot, theta, theta_dot)
        self.elapsed += 1
        self.total_steps += 1
        done = bool(x < -X_THRESHOLD or x > X_THRESHOLD
                    or theta < -THETA_THRESHOLD or theta > THETA_THRESHOLD)
        done = done or self.elapsed >= MAX_EPISODE_STEPS       # TimeLimit wrapper of -v1
        return np.array(self.state, dtype=np.float32), 1.0, done


class CartPole:
    """G independent CartPole-v1 games, float64 state, same interface as the board games."""

    N_ACTIONS = 2

    def __init__(self, n_games, seed=0, slot0=0):
        self.G = n_games
        self.envs = [CartPoleV1(seed, slot0 + g) for g in range(n_games)]
        self.reset()

    def reset(self, mask=None):
        for g, e in enumerate(self.envs):
            if mask is None or mask[g]:
                e.reset()
        return self.observation()

    def observation(self):
        return np.array([e.state for e in self.envs], dtype=np.float64).astype(np.float32).reshape(self.G, 1, 1, 4)

    def state64(self):
        return np.array([e.state for e in self.envs], dtype=np.float64)

    def to_play(self):
        return np.zeros(self.G, dtype=np.int32)

    def legal_mask(self):
        return np.ones((self.G, 2), dtype=bool)

    def step(self, actions):
        rew = np.zeros(self.G)
        done = np.zeros(self.G, dtype=bool)
        for g, e in enumerate(self.envs):
            _, rew[g], done[g] = e.step(int(actions[g]))
        return self.observation(), rew, done


class SyntheticFrames:
    """Breakout stand-in of BASELINE.json ("synthetic 96x96 frames"): frames are U[0,1) float32 drawn from the
    shared counter RNG keyed by (seed, slot, step, pixel quad); reward 0; never done before max_moves.
    (The reference wraps the ALE emulator, games/breakout.py:135-185 - third-party, absent: parity N/A.)"""

    N_ACTIONS = 4

    def __init__(self, n_games, seed=0, slot0=0):
        self.G, self.seed, self.slot0 = n_games, seed, slot0
        self.steps = np.zeros(n_games, dtype=np.int64)

    def reset(self, mask=None):
        return self.observation()

    def observation(self):
        n4 = 3 * 96 * 96 // 4
        out = np.empty((self.G, n4, 4), dtype=np.float32)
        for g in range(self.G):
            r = rng.philox4x32_np(np.full(n4, self.slot0 + g), np.full(n4, self.steps[g]),
                                  np.full(n4, (rng.STREAM_RESET << 16) | 1), np.arange(n4), self.seed & rng.MASK,
                                  (self.seed >> 32) & rng.MASK)
            out[g] = (np.stack(r, axis=1) >> np.uint64(8)).astype(np.float32) / np.float32(16777216.0)
        return out.reshape(self.G, 3, 96, 96)

    def to_play(self):
        return np.zeros(self.G, dtype=np.int32)

    def legal_mask(self):
        return np.ones((self.G, 4), dtype=bool)

    def step(self, actions):
        self.steps += 1
        return self.observation(), np.zeros(self.G), np.zeros(self.G, dtype=bool)
