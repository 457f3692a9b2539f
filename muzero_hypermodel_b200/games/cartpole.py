"""cartpole: MuZeroConfig with the reference's values (games/cartpole.py) and the Game plug-in."""
from ._config import ConfigBase


class MuZeroConfig(ConfigBase):
    GAME = "cartpole"
    VALUES = dict(
        observation_shape=(1, 1, 4),
        max_moves=500,
        num_simulations=50,
        discount=0.997,
        root_dirichlet_alpha=0.25,
        network='fullyconnected',
        blocks=1,
        channels=2,
        reduced_channels_reward=2,
        reduced_channels_value=2,
        reduced_channels_policy=2,
        resnet_fc_reward_layers=[],
        resnet_fc_value_layers=[],
        resnet_fc_policy_layers=[],
        encoding_size=8,
        fc_representation_layers=[],
        fc_dynamics_layers=[16],
        fc_reward_layers=[16],
        fc_value_layers=[16],
        fc_policy_layers=[16],
        training_steps=10000,
        batch_size=128,
        value_loss_weight=1,
        lr_init=0.02,
        lr_decay_rate=0.9,
        lr_decay_steps=1000,
        replay_buffer_size=500,
        num_unroll_steps=10,
        td_steps=50,
        ratio=1.5,
        n_actions=2,
        n_players=1,
    )
    TEMPERATURE = ((0.5, 1.0), (0.75, 0.5))
    TEMPERATURE_FINAL = 0.25


from ._device_game import make_game_class  # noqa: E402

Game = make_game_class("cartpole", 500)
