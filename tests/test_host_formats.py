"""Host-side pieces around the hot path: GameHistory.get_stacked_observations vs the UNMODIFIED reference
(tests/golden/stacked.npz, self_play.py:514-548), SharedStorage / model.checkpoint / replay_buffer.pkl formats
(shared_storage.py:8-41, muzero.py:94-112, 315-323)."""
import os
import pickle

import numpy as np
import pytest
import torch

import _tables as T


def _history(si):
    from muzero_hypermodel_b200.self_play import GameHistory
    Z = T.load("stacked")
    gh = GameHistory()
    gh.observation_history = list(Z[f"{si}/observations"])
    gh.action_history = Z[f"{si}/actions"].tolist()
    return gh, Z


@pytest.mark.parametrize("si", range(3))
def test_get_stacked_observations_equals_reference(si):
    gh, Z = _history(si)
    n = len(gh.observation_history)
    for S in (0, 1, 2, 4):
        for index in range(-1, n + 2):
            want = Z[f"{si}/S{S}/i{index}"]
            got = np.asarray(gh.get_stacked_observations(index, S))
            assert got.shape == want.shape and got.dtype == want.dtype, (S, index, got.shape, want.shape)
            assert got.tobytes() == want.tobytes(), (S, index)


def test_shared_storage_semantics_and_checkpoint_keys(tmp_path):
    from muzero_hypermodel_b200 import shared_storage as ss
    from muzero_hypermodel_b200.games.cartpole import MuZeroConfig
    cfg = MuZeroConfig()
    cfg.results_path = str(tmp_path / "run")
    ck = ss.new_checkpoint({"w": torch.arange(4.0)})
    assert list(ck) == ["weights", "optimizer_state", "total_reward", "muzero_reward", "opponent_reward", "episode_length",
                        "mean_value", "training_step", "lr", "total_loss", "value_loss", "reward_loss", "policy_loss",
                        "num_played_games", "num_played_steps", "num_reanalysed_games", "terminate"]
    st = ss.SharedStorage(ck, cfg)
    ck["training_step"] = 99                                     # the storage holds its own copy (:16-17)
    assert st.get_info("training_step") == 0
    st.set_info("training_step", 5)
    st.set_info({"lr": 0.01, "total_loss": 1.5})
    assert st.get_info(["training_step", "lr"]) == {"training_step": 5, "lr": 0.01}
    with pytest.raises(TypeError):
        st.get_info(3)
    with pytest.raises(TypeError):
        st.set_info("lr")                                        # string key without a value (:33-41)
    st.save_checkpoint()
    back = ss.load_checkpoint(os.path.join(cfg.results_path, "model.checkpoint"))
    assert back["training_step"] == 5 and torch.equal(back["weights"]["w"], torch.arange(4.0)) and set(back) == set(ck)
    assert set(ss.TENSORBOARD_TAGS.values()) <= set(ck)


def test_replay_buffer_pkl_round_trip(tmp_path):
    from muzero_hypermodel_b200 import shared_storage as ss
    from muzero_hypermodel_b200.self_play import GameHistory
    gh, _ = _history(0)
    gh.reward_history, gh.to_play_history = [0.0] * 7, [0] * 7
    gh.child_visits, gh.root_values = [[0.5, 0.5]] * 6, [0.1] * 6
    gh.priorities, gh.game_priority = np.ones(6, np.float32), np.float32(1)
    ck = ss.new_checkpoint()
    ck.update(num_played_games=4, num_played_steps=24, num_reanalysed_games=2)
    path = str(tmp_path / "r" / "replay_buffer.pkl")
    ss.save_replay_buffer(path, {3: gh}, ck)
    raw = pickle.load(open(path, "rb"))
    assert set(raw) == {"buffer", "num_played_games", "num_played_steps", "num_reanalysed_games"}     # muzero.py:315-323
    ck2 = ss.new_checkpoint()
    buf = ss.load_replay_buffer(path, ck2)
    assert (ck2["num_played_games"], ck2["num_played_steps"], ck2["num_reanalysed_games"]) == (4, 24, 2)
    assert isinstance(buf[3], GameHistory) and buf[3].action_history == gh.action_history
    assert all(np.array_equal(a, b) for a, b in zip(buf[3].observation_history, gh.observation_history))
