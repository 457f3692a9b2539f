"""`Game` plug-ins with the AbstractGame contract (games/abstract_game.py:14-57) on the device environments.

A `Game(seed)` is one slot of a vectorised device environment (G = 1): step / legal_actions / to_play /
reset return what the reference's wrappers return (same array shapes and dtypes, rewards already scaled).
Self-play at scale does not go through these objects - it drives `envs.VectorEnv` with G games - they keep
single-game callers (muzero.py test mode, diagnose_model.py) working.
"""
import numpy

from .abstract_game import AbstractGame


def make_game_class(kind, max_moves):
    class Game(AbstractGame):
        KIND = kind

        def __init__(self, seed=None):
            import torch
            from ..envs import VectorEnv
            self._torch = torch
            self._seed = 0 if seed is None else int(seed)
            self.env = VectorEnv(kind, 1, max(max_moves, 1) + 1, seed=self._seed, device="cuda")

        def _observation(self):
            obs = self.env.observe()[0][0].cpu().numpy()
            if kind == "cartpole":
                return numpy.array([[obs]])
            c, h, w = self.env.obs_shape
            planes = obs.reshape(c, h, w)
            return planes.astype("int32") if kind == "tictactoe" else planes.astype("float64")

        def step(self, action):
            torch = self._torch
            a = torch.tensor([int(action)], dtype=torch.int32, device=self.env.device)
            _, r, d = self.env.act_step(None, None, forced_action=a, want_outputs=True)
            reward = float(r[0])
            return self._observation(), (int(reward) if kind != "cartpole" else reward), bool(d[0])

        def to_play(self):
            self.env.observe()
            return int(self.env.to_play[0])

        def legal_actions(self):
            self.env.observe()
            return numpy.nonzero(self.env.legal[0].cpu().numpy())[0].tolist()

        def reset(self):
            self.env.harvest(False)      # restart a finished episode
            self.env.reset()
            return self._observation()

        def close(self):
            self.env = None

        def render(self):
            print(self._observation())

        def expert_agent(self):
            """The hard-coded evaluation opponents (games/tictactoe.py:217-225, connect4.py:196-204, gomoku.py) are
            host heuristics over the reference's board objects and never influence training.  SelfPlay honours any
            caller-supplied Game class: pass the reference's own `games.<name>.Game` to play against its expert."""
            raise NotImplementedError(
                f"the device {kind} Game has no expert agent: pass a host Game class with expert_agent() to SelfPlay "
                "(it is played through the AbstractGame contract, searches stay on the device)")

        def human_to_action(self):
            legal = self.legal_actions()
            while True:
                choice = input(f"Enter the action to play for player {self.to_play()} {legal}: ")
                if choice.strip().lstrip("-").isdigit() and int(choice) in legal:
                    return int(choice)
                print("Wrong input, try again")

        def action_to_string(self, action_number):
            if kind == "tictactoe":
                return f"Play row {action_number // 3 + 1}, column {action_number % 3 + 1}"
            if kind == "connect4":
                return f"Play column {action_number + 1}"
            if kind == "gomoku":
                return chr(action_number // 11 + 65) + chr(action_number % 11 + 65)
            return f"{action_number}. " + {0: "Push cart to the left", 1: "Push cart to the right"}[action_number]

    Game.__name__ = "Game"
    return Game
