"""Config objects for tests: the product's own MuZeroConfig classes (same attributes as the reference's)."""
import importlib


def product_config(name, **override):
    if name == "tictactoe_fc":
        cfg = importlib.import_module("muzero_hypermodel_b200.games.tictactoe").MuZeroConfig()
        cfg.network = "fullyconnected"
    else:
        cfg = importlib.import_module(f"muzero_hypermodel_b200.games.{name}").MuZeroConfig()
    for k, v in override.items():
        setattr(cfg, k, v)
    return cfg
