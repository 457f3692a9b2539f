"""The C-ABI boundary: libmzb200.so loads without a GPU and exports every function include/mzb200.h declares
(no compute calls here); the product refuses to run without CUDA instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mzb200.h")
LIB = os.path.join(ROOT, "muzero_hypermodel_b200", "libmzb200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)            # comments mention function names too
    src = re.sub(r"//[^\n]*", "", src)
    names = re.findall(r"\b(mzb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("mzb_version", "mzb_last_error", "mzb_tree_create", "mzb_tree_select", "mzb_tree_expand_backup",
                 "mzb_search_fc", "mzb_search_resnet", "mzb_fc_initial", "mzb_fc_recurrent", "mzb_resnet_initial",
                 "mzb_resnet_recurrent", "mzb_env_create", "mzb_env_act_step", "mzb_make_target"):
        assert must in names, must
    assert len(names) >= 40


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(LIB), "build first: python __graft_entry__.py"
    lib = ctypes.CDLL(LIB)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    lib.mzb_version.restype = ctypes.c_int
    assert lib.mzb_version() >= 100
    lib.mzb_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.mzb_last_error(), bytes)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present: the refusal path is for CPU-only hosts")
    from muzero_hypermodel_b200 import models
    from muzero_hypermodel_b200.games.cartpole import MuZeroConfig
    net = models.MuZeroNetwork(MuZeroConfig())
    with pytest.raises(Exception):
        net.initial_inference(torch.zeros(1, 1, 1, 4))
