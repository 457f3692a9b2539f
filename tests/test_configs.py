"""MuZeroConfig drop-in: same attributes and values as the reference's per-game classes."""
import pytest

from _configs import product_config
from oracle import ref_loader

GAMES = ["cartpole", "tictactoe", "connect4", "gomoku", "breakout"]
# frozen copy of the hot-path fields of the reference configs, checked everywhere
HOT = {
    "cartpole": dict(observation_shape=(1, 1, 4), A=2, P=1, num_simulations=50, max_moves=500, discount=0.997,
                     root_dirichlet_alpha=0.25, network="fullyconnected", encoding_size=8, td_steps=50, num_unroll_steps=10),
    "tictactoe": dict(observation_shape=(3, 3, 3), A=9, P=2, num_simulations=25, max_moves=9, discount=1,
                      root_dirichlet_alpha=0.1, network="resnet", encoding_size=32, td_steps=20, num_unroll_steps=20),
    "connect4": dict(observation_shape=(3, 6, 7), A=7, P=2, num_simulations=200, max_moves=42, discount=1,
                     root_dirichlet_alpha=0.3, network="resnet", encoding_size=32, td_steps=42, num_unroll_steps=42),
    "gomoku": dict(observation_shape=(3, 11, 11), A=121, P=2, num_simulations=400, max_moves=121, discount=1,
                   root_dirichlet_alpha=0.3, network="resnet", encoding_size=32, td_steps=121, num_unroll_steps=121),
    "breakout": dict(observation_shape=(3, 96, 96), A=4, P=1, num_simulations=30, max_moves=2500, discount=0.997,
                     root_dirichlet_alpha=0.25, network="resnet", encoding_size=10, td_steps=10, num_unroll_steps=5),
}


@pytest.mark.parametrize("name", GAMES)
def test_hot_path_fields(name):
    cfg = product_config(name)
    want = dict(HOT[name])
    assert len(cfg.action_space) == want.pop("A") and cfg.action_space == list(range(len(cfg.action_space)))
    assert len(cfg.players) == want.pop("P")
    for k, v in want.items():
        assert getattr(cfg, k) == v, k
    assert cfg.pb_c_base == 19652 and cfg.pb_c_init == 1.25 and cfg.support_size == 10
    assert cfg.root_exploration_fraction == 0.25 and cfg.stacked_observations == 0


def test_temperature_schedules():
    c = product_config("cartpole")
    assert [c.visit_softmax_temperature_fn(s) for s in (0, 4999, 5000, 7499, 7500, 10 ** 6)] == [1.0, 1.0, 0.5, 0.5, 0.25, 0.25]
    b = product_config("breakout")
    assert [b.visit_softmax_temperature_fn(s) for s in (0, 499999, 500000, 749999, 750000)] == [1.0, 1.0, 0.5, 0.5, 0.25]
    assert product_config("connect4").visit_softmax_temperature_fn(12345) == 1


@pytest.mark.refcheck
@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
@pytest.mark.parametrize("name", GAMES)
def test_every_attribute_equals_reference(name):
    ref = ref_loader.load(f"games.{name}").MuZeroConfig()
    cfg = product_config(name)
    rv = {k: v for k, v in vars(ref).items() if k != "results_path"}
    pv = {k: v for k, v in vars(cfg).items() if k != "results_path"}
    assert set(rv) == set(pv)
    for k in rv:
        assert rv[k] == pv[k] and type(rv[k]) == type(pv[k]), k
    for s in (0, 1, 4999, 5000, 7500, 499999, 500000, 750000, 10 ** 7):
        assert ref.visit_softmax_temperature_fn(trained_steps=s) == cfg.visit_softmax_temperature_fn(trained_steps=s)
