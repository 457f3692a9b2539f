"""Game plug-ins: one `MuZeroConfig` + `Game` per module, as `importlib.import_module("games." + name)`
expects (muzero.py:45-47).  Configs carry the reference's attribute names and values; `Game` keeps the
AbstractGame contract on top of the vectorised device environments (csrc/mzb_env.cu)."""
