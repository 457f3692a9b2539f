"""breakout: MuZeroConfig with the reference's values (games/breakout.py) and the Game plug-in."""
from ._config import ConfigBase


class MuZeroConfig(ConfigBase):
    GAME = "breakout"
    VALUES = dict(
        observation_shape=(3, 96, 96),
        max_moves=2500,
        num_simulations=30,
        discount=0.997,
        root_dirichlet_alpha=0.25,
        network='resnet',
        downsample='resnet',
        blocks=2,
        channels=16,
        reduced_channels_reward=4,
        reduced_channels_value=4,
        reduced_channels_policy=4,
        resnet_fc_reward_layers=[16],
        resnet_fc_value_layers=[16],
        resnet_fc_policy_layers=[16],
        encoding_size=10,
        fc_representation_layers=[],
        fc_dynamics_layers=[16],
        fc_reward_layers=[16],
        fc_value_layers=[],
        fc_policy_layers=[],
        training_steps=1000000,
        batch_size=16,
        checkpoint_interval=500,
        value_loss_weight=0.25,
        lr_init=0.005,
        lr_decay_rate=1,
        lr_decay_steps=350000.0,
        replay_buffer_size=1000000,
        num_unroll_steps=5,
        td_steps=10,
        PER_alpha=1,
        use_last_model_value=False,
        n_actions=4,
        n_players=1,
    )
    TEMPERATURE = ((500e3, 1.0), (750e3, 0.5))
    TEMPERATURE_FINAL = 0.25


class Game:
    """The reference wraps the ALE emulator (games/breakout.py:135-185), which is third-party, absent and
    excluded by BASELINE.json ("synthetic 96x96 frames"): the device path feeds the residual network with
    synthetic frames instead (bench.py --workload breakout)."""

    def __init__(self, seed=None):
        raise NotImplementedError("Breakout needs the ALE emulator; the B200 path benchmarks it on synthetic frames")
