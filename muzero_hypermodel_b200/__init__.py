"""muzero_hypermodel_b200 - B200-native (sm_100a) self-play hot path of muzero-hypermodel.

Drop-in for the reference's hot-path API: `models.MuZeroNetwork`,
`models.support_to_scalar/scalar_to_support`, `self_play.MCTS/Node/GameHistory/SelfPlay`,
`games.<name>.MuZeroConfig/Game`, `replay_buffer.make_target` - backed by hand-written CUDA
kernels behind the C ABI in include/mzb200.h (libmzb200.so).  No CPU fallback.
"""
__version__ = "0.1.0"
