"""Shared attribute bag behind the per-game MuZeroConfig classes.

The reference defines one flat class per game (e.g. games/cartpole.py:11-127); the hot path only
reads attributes, so the drop-in keeps every attribute name and value and factors the common defaults.
Optional B200-path attributes (read with getattr(cfg, name, default) by the batched self-play):
`num_parallel_games`."""
import datetime
import os

import torch

_DEFAULTS = dict(
    seed=0, max_num_gpus=None, stacked_observations=0, muzero_player=0, opponent=None,
    num_workers=1, selfplay_on_gpu=False, temperature_threshold=None,
    root_exploration_fraction=0.25, pb_c_base=19652, pb_c_init=1.25, support_size=10,
    downsample=False, save_model=True, checkpoint_interval=10, optimizer="Adam", weight_decay=1e-4, momentum=0.9,
    PER=True, PER_alpha=0.5, use_last_model_value=True, reanalyse_on_gpu=False, self_play_delay=0, training_delay=0,
    ratio=None,
)


class ConfigBase:
    GAME = "game"
    VALUES = {}
    TEMPERATURE = None            # None: constant 1; else ((fraction_of_training_steps | absolute, T), ...)

    def __init__(self):
        for k, v in _DEFAULTS.items():
            setattr(self, k, v)
        for k, v in self.VALUES.items():
            setattr(self, k, list(v) if isinstance(v, (list, tuple)) and k.endswith("layers") else v)
        self.action_space = list(range(self.VALUES["n_actions"]))
        self.players = list(range(self.VALUES["n_players"]))
        del self.n_actions, self.n_players
        self.train_on_gpu = torch.cuda.is_available()
        self.results_path = os.path.join(os.path.dirname(os.path.realpath(__file__)), "../results", self.GAME,
                                         datetime.datetime.now().strftime("%Y-%m-%d--%H-%M-%S"))

    def visit_softmax_temperature_fn(self, trained_steps):
        """Temperature of the visit-count distribution used to pick the played action."""
        if self.TEMPERATURE is None:
            return 1
        for bound, temperature in self.TEMPERATURE:
            limit = bound * self.training_steps if bound <= 1 else bound
            if trained_steps < limit:
                return temperature
        return self.TEMPERATURE_FINAL
