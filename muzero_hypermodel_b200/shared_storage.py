"""SharedStorage and the on-disk formats around the hot path (shared_storage.py:8-41, muzero.py:94-112, 315-323, 402-441).

`SharedStorage(checkpoint, config)` is the key/value store the self-play workers, the trainer and Reanalyse talk to
(`get_info` / `set_info` / `get_checkpoint` / `save_checkpoint`), as a plain in-process object: on one B200 the actors
are objects driven from one loop, not Ray processes, and `SelfPlay.continuous_self_play`, `Trainer.
continuous_update_weights` and `Reanalyse.reanalyse` accept it (or a Ray handle) unchanged.

`new_checkpoint()` is the dictionary `model.checkpoint` holds; `save_replay_buffer` / `load_replay_buffer` write and
read `replay_buffer.pkl`.  Files written by the reference load here and vice versa (same keys, same GameHistory fields;
GameHistory objects unpickle against whichever `self_play` module is importable under that name).
"""
import copy
import os
import pickle

import torch

# muzero.py:94-112 - every key the workers read or write, with the reference's initial values
CHECKPOINT_KEYS = {
    "weights": None, "optimizer_state": None, "total_reward": 0, "muzero_reward": 0, "opponent_reward": 0,
    "episode_length": 0, "mean_value": 0, "training_step": 0, "lr": 0, "total_loss": 0, "value_loss": 0,
    "reward_loss": 0, "policy_loss": 0, "num_played_games": 0, "num_played_steps": 0, "num_reanalysed_games": 0,
    "terminate": False,
}
# muzero.py:258-300 - TensorBoard tags -> checkpoint key (the logging loop's mapping, for callers that keep a writer)
TENSORBOARD_TAGS = {
    "1.Total_reward/1.Total_reward": "total_reward", "1.Total_reward/2.Mean_value": "mean_value",
    "1.Total_reward/3.Episode_length": "episode_length", "1.Total_reward/4.MuZero_reward": "muzero_reward",
    "1.Total_reward/5.Opponent_reward": "opponent_reward", "2.Workers/1.Self_played_games": "num_played_games",
    "2.Workers/2.Training_steps": "training_step", "2.Workers/3.Self_played_steps": "num_played_steps",
    "2.Workers/4.Reanalysed_games": "num_reanalysed_games", "2.Workers/6.Learning_rate": "lr",
    "3.Loss/1.Total_weighted_loss": "total_loss", "3.Loss/Value_loss": "value_loss", "3.Loss/Reward_loss": "reward_loss",
    "3.Loss/Policy_loss": "policy_loss",
}


def new_checkpoint(weights=None):
    ck = copy.deepcopy(CHECKPOINT_KEYS)
    ck["weights"] = weights
    return ck


class SharedStorage:
    def __init__(self, checkpoint, config):
        self.config = config
        self.current_checkpoint = copy.deepcopy(checkpoint)

    def save_checkpoint(self, path=None):
        if not path:
            path = os.path.join(self.config.results_path, "model.checkpoint")
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save(self.current_checkpoint, path)

    def get_checkpoint(self):
        return copy.deepcopy(self.current_checkpoint)

    def get_info(self, keys):
        if isinstance(keys, str):
            return self.current_checkpoint[keys]
        if isinstance(keys, list):
            return {key: self.current_checkpoint[key] for key in keys}
        raise TypeError

    def set_info(self, keys, values=None):
        if isinstance(keys, str) and values is not None:
            self.current_checkpoint[keys] = values
        elif isinstance(keys, dict):
            self.current_checkpoint.update(keys)
        else:
            raise TypeError


def load_checkpoint(path):
    """model.checkpoint -> dict (muzero.py:411-416).  Reference checkpoints pickle numpy scalars next to the tensors."""
    return torch.load(path, map_location="cpu", weights_only=False)


def save_replay_buffer(path, buffer, checkpoint):
    """replay_buffer.pkl (muzero.py:315-323): `buffer` = {game_id: GameHistory} (ReplayBuffer.get_buffer())."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump({"buffer": buffer, "num_played_games": checkpoint["num_played_games"],
                     "num_played_steps": checkpoint["num_played_steps"],
                     "num_reanalysed_games": checkpoint["num_reanalysed_games"]}, f)


def load_replay_buffer(path, checkpoint):
    """muzero.py:419-436: returns the buffer dict and restores the three counters into `checkpoint`."""
    with open(path, "rb") as f:
        infos = pickle.load(f)
    checkpoint["num_played_steps"] = infos["num_played_steps"]
    checkpoint["num_played_games"] = infos["num_played_games"]
    checkpoint["num_reanalysed_games"] = infos["num_reanalysed_games"]
    return infos["buffer"]
