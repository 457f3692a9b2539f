"""Import the UNMODIFIED reference from /root/reference (build container only).

Test infrastructure.  Used by tests/golden/make_golden.py to produce the committed golden
vectors, and by the optional `-m refcheck` tests.  /root/reference does not exist on the GPU
box, so nothing that runs there may call `load()`; `available()` says whether it can.

The reference imports two packages that are not installed here:
  * `ray`  - only the `@ray.remote` decorator and `.remote/.get` glue are touched at import
             (self_play.py:5,11; replay_buffer.py:5,11; trainer.py:5,11; shared_storage.py:4,8).
             A stub whose `remote` is the identity decorator is installed in sys.modules.
  * `gym`  - only games/cartpole.py:136-173 uses it (`make("CartPole-v1")`, old 4-tuple API).
             A shim backed by oracle.games.CartPoleV1 is installed; it is the golden
             definition of the physics ("parity unpinned" at that third-party boundary).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MZB_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "self_play.py"))


def _install_ray_stub():
    if "ray" in sys.modules:
        return
    ray = types.ModuleType("ray")

    def remote(*args, **kwargs):
        if len(args) == 1 and not kwargs and (isinstance(args[0], type) or callable(args[0])):
            return args[0]
        return lambda obj: obj

    ray.remote = remote
    ray.get = lambda x: x
    ray.init = lambda *a, **k: None
    sys.modules["ray"] = ray


def _install_gym_shim():
    if "gym" in sys.modules:
        return
    from oracle.games import CartPoleV1

    gym = types.ModuleType("gym")

    class _Env:
        def __init__(self):
            self._core = CartPoleV1(seed=None)

        def seed(self, seed):
            self._core = CartPoleV1(seed=seed)

        def reset(self):
            return self._core.reset()

        def step(self, action):
            obs, reward, done = self._core.step(action)
            return obs, reward, done, {}

        def close(self):
            pass

    def make(name):
        assert name == "CartPole-v1", name
        return _Env()

    gym.make = make
    sys.modules["gym"] = gym


_loaded = {}


def load(*module_names):
    """Return the named reference modules (e.g. 'self_play', 'models', 'games.tictactoe')."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_ray_stub()
    _install_gym_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    out = []
    for name in module_names:
        if name not in _loaded:
            # The product package also has modules called models / self_play / games; make sure
            # the names resolve to the reference's files, not to anything cached.
            mod = importlib.import_module(name)
            assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), mod.__file__
            _loaded[name] = mod
        out.append(_loaded[name])
    return out[0] if len(out) == 1 else out
