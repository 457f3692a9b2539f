"""gomoku: MuZeroConfig with the reference's values (games/gomoku.py) and the Game plug-in."""
from ._config import ConfigBase


class MuZeroConfig(ConfigBase):
    GAME = "gomoku"
    VALUES = dict(
        observation_shape=(3, 11, 11),
        opponent='random',
        num_workers=2,
        max_moves=121,
        num_simulations=400,
        discount=1,
        root_dirichlet_alpha=0.3,
        network='resnet',
        blocks=6,
        channels=128,
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
        training_steps=10000,
        batch_size=512,
        checkpoint_interval=50,
        value_loss_weight=1,
        lr_init=0.002,
        lr_decay_rate=0.9,
        lr_decay_steps=10000,
        replay_buffer_size=10000,
        num_unroll_steps=121,
        td_steps=121,
        use_last_model_value=False,
        ratio=1,
        n_actions=121,
        n_players=2,
    )
    TEMPERATURE = ((0.5, 1.0), (0.75, 0.5))
    TEMPERATURE_FINAL = 0.25


from ._device_game import make_game_class  # noqa: E402

Game = make_game_class("gomoku", 121)
