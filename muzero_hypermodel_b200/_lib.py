"""ctypes binding of libmzb200.so (the C ABI declared in include/mzb200.h).

The product path has no CPU fallback: importing this module without the built library raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmzb200.so")


class MzbError(RuntimeError):
    pass


class TreeConfig(C.Structure):
    _fields_ = [
        ("n_games", C.c_int32), ("n_actions", C.c_int32), ("num_simulations", C.c_int32), ("n_players", C.c_int32),
        ("discount", C.c_double), ("pb_c_base", C.c_double), ("pb_c_init", C.c_double),
        ("hidden_floats", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64),
    ]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise MzbError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(muzero_hypermodel_b200/csrc/build.sh). There is no CPU fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64, f64, sz = C.c_void_p, C.c_int32, C.c_uint64, C.c_double, C.c_size_t
    sig = {
        "mzb_version": (C.c_int, []),
        "mzb_last_error": (C.c_char_p, []),
        "mzb_launch_count": (u64, []),
        "mzb_reset_launch_count": (None, []),
        "mzb_philox": (None, [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]),
        "mzb_tree_workspace_bytes": (sz, [C.POINTER(TreeConfig)]),
        "mzb_tree_create": (C.c_int, [C.POINTER(vp), C.POINTER(TreeConfig), vp, sz, C.POINTER(f64)]),
        "mzb_tree_destroy": (C.c_int, [vp]),
        "mzb_tree_hidden_ptr": (vp, [vp]),
        "mzb_tree_root_init": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, vp, f64, f64, vp, vp, vp]),
        "mzb_tree_select": (C.c_int, [vp, vp, vp, vp, vp]),
        "mzb_tree_expand_backup": (C.c_int, [vp, vp, vp, vp, C.c_int, vp]),
        "mzb_tree_root_stats": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "mzb_tree_counters_sync": (C.c_int, [vp, C.POINTER(u64), C.c_int, vp]),
        "mzb_tree_export_game_sync": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib, sig


lib, SIGNATURES = _load()


def bind(name, restype, argtypes):
    """Register one more entry point (modules that add ABI families call this at import)."""
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    SIGNATURES[name] = (restype, argtypes)
    return fn


def check(rc):
    if rc != 0:
        msg = lib.mzb_last_error().decode()
        if rc == -3:
            raise NotImplementedError(msg)
        raise MzbError(f"libmzb200 error {rc}: {msg}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array, None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
