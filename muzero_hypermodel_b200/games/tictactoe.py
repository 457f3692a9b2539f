"""tictactoe: MuZeroConfig with the reference's values (games/tictactoe.py) and the Game plug-in."""
from ._config import ConfigBase


class MuZeroConfig(ConfigBase):
    GAME = "tictactoe"
    VALUES = dict(
        observation_shape=(3, 3, 3),
        opponent='expert',
        max_moves=9,
        num_simulations=25,
        discount=1,
        root_dirichlet_alpha=0.1,
        network='resnet',
        blocks=1,
        channels=16,
        reduced_channels_reward=16,
        reduced_channels_value=16,
        reduced_channels_policy=16,
        resnet_fc_reward_layers=[8],
        resnet_fc_value_layers=[8],
        resnet_fc_policy_layers=[8],
        encoding_size=32,
        fc_representation_layers=[],
        fc_dynamics_layers=[16],
        fc_reward_layers=[16],
        fc_value_layers=[],
        fc_policy_layers=[],
        training_steps=1000000,
        batch_size=64,
        value_loss_weight=0.25,
        lr_init=0.003,
        lr_decay_rate=1,
        lr_decay_steps=10000,
        replay_buffer_size=3000,
        num_unroll_steps=20,
        td_steps=20,
        n_actions=9,
        n_players=2,
    )


from ._device_game import make_game_class  # noqa: E402

Game = make_game_class("tictactoe", 9)
