"""connect4: MuZeroConfig with the reference's values (games/connect4.py) and the Game plug-in."""
from ._config import ConfigBase


class MuZeroConfig(ConfigBase):
    GAME = "connect4"
    VALUES = dict(
        observation_shape=(3, 6, 7),
        opponent='expert',
        max_moves=42,
        num_simulations=200,
        discount=1,
        root_dirichlet_alpha=0.3,
        network='resnet',
        blocks=3,
        channels=64,
        reduced_channels_reward=2,
        reduced_channels_value=2,
        reduced_channels_policy=4,
        resnet_fc_reward_layers=[64],
        resnet_fc_value_layers=[64],
        resnet_fc_policy_layers=[64],
        encoding_size=32,
        fc_representation_layers=[],
        fc_dynamics_layers=[64],
        fc_reward_layers=[64],
        fc_value_layers=[],
        fc_policy_layers=[],
        training_steps=100000,
        batch_size=64,
        value_loss_weight=0.25,
        lr_init=0.005,
        lr_decay_rate=1,
        lr_decay_steps=10000,
        replay_buffer_size=10000,
        num_unroll_steps=42,
        td_steps=42,
        n_actions=7,
        n_players=2,
    )


from ._device_game import make_game_class  # noqa: E402

Game = make_game_class("connect4", 42)
