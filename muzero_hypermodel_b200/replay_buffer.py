"""Device-side target construction with the reference's names (replay_buffer.py:222-295).

`make_target_batch` builds the value / reward / policy / action targets of a batch of sampled
positions on the GPU (K11, csrc/mzb_targets.cu) from games stored in the export layout of
`envs.VectorEnv.drain_raw()`.

`ReplayBuffer` is the reference's class (replay_buffer.py:11-220) with the game store, the PER sampling, the batch
assembly and the priority updates on the device (csrc/mzb_replay.cu): same constructor, `save_game`, `get_batch`,
`update_priorities`, counters.  Batches come back as device tensors (what the trainer moves to the GPU anyway,
trainer.py:124-135).  Reanalyse is a later row.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

_vp, _i32 = C.c_void_p, C.c_int32
_i64 = C.c_int64


class ReplayConfig(C.Structure):
    _fields_ = [("n_actions", _i32), ("obs_floats", _i32), ("obs_decode", _i32), ("obs_h", _i32), ("obs_w", _i32),
                ("capacity_games", _i32), ("entry_stride", _i32),
                ("num_unroll_steps", _i32), ("td_steps", _i32), ("per", _i32), ("max_batch", _i32),
                ("stacked_observations", _i32), ("obs_channels", _i32),
                ("per_alpha", C.c_double), ("seed", C.c_uint64)]


_lib.bind("mzb_replay_workspace_bytes", C.c_size_t, [C.POINTER(ReplayConfig)])
_lib.bind("mzb_replay_create", C.c_int, [C.POINTER(_vp), C.POINTER(ReplayConfig), _vp, C.c_size_t, _vp, _vp])
_lib.bind("mzb_replay_destroy", C.c_int, [_vp])
_lib.bind("mzb_replay_save_games", C.c_int, [_vp, _i32, _vp, _vp] + [_vp] * 7 + [_vp])
_lib.bind("mzb_replay_get_batch", C.c_int, [_vp, _i32] + [_vp] * 13 + [_vp])
_lib.bind("mzb_replay_update_priorities", C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp])
_lib.bind("mzb_replay_info", C.c_int, [_vp, C.POINTER(_i64)])
_lib.bind("mzb_env_export_to_replay", C.c_int, [_vp, _vp, C.POINTER(_i32), _vp])
_lib.bind("mzb_replay_set_batch_counter", C.c_int, [_vp, C.c_uint32])
_lib.bind("mzb_replay_game_observations", C.c_int, [_vp, _i64, _vp, C.POINTER(_i32), _vp])
_lib.bind("mzb_replay_set_reanalysed", C.c_int, [_vp, _i64, _vp, _vp])
_lib.bind("mzb_replay_export_game_sync", C.c_int, [_vp, _i64, C.POINTER(_i32)] + [_vp] * 6 + [_vp])
_lib.bind("mzb_replay_game_priorities_sync", C.c_int, [_vp, _i64, _vp, _vp, C.POINTER(_i32), _vp])
_lib.bind("mzb_make_target", C.c_int, [_vp] * 8 + [_i32] + [_vp] * 4 + [_i32, _i32, _i32, _vp, C.c_uint64] + [_vp] * 5)


class DeviceGames:
    """Finished games on the device in entry-array form (what the export ring holds)."""

    def __init__(self, reward, to_play, root_value, visits, action, game_start, game_len, reanalysed=None, device="cuda"):
        dev = torch.device(device)
        t = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x)).to(dev, dt).contiguous()
        self.reward, self.to_play = t(reward, torch.float32), t(to_play, torch.int8)
        self.root_value = t(root_value, torch.float64)
        self.visits = torch.from_numpy(np.ascontiguousarray(visits).astype(np.uint16)).to(dev).contiguous()
        self.action, self.game_start, self.game_len = t(action, torch.int32), t(game_start, torch.int32), t(game_len, torch.int32)
        self.reanalysed = None if reanalysed is None else t(reanalysed, torch.float64)
        self.A = self.visits.shape[1]
        self.device = dev


def make_target_batch(games, batch_game, batch_index, config, seed=0, batch_slot=None, batch_step=None):
    """Returns (target_values [B,K+1] f64, target_rewards [B,K+1] f64, target_policies [B,K+1,A] f64, actions [B,K+1] i32)."""
    dev = games.device
    B = len(batch_game)
    K, td = int(config.num_unroll_steps), int(config.td_steps)
    bg = torch.as_tensor(np.asarray(batch_game, dtype=np.int32)).to(dev)
    bi = torch.as_tensor(np.asarray(batch_index, dtype=np.int32)).to(dev)
    bs = None if batch_slot is None else torch.as_tensor(np.asarray(batch_slot, dtype=np.int64)).to(dev).to(torch.int32)
    bt = None if batch_step is None else torch.as_tensor(np.asarray(batch_step, dtype=np.int64)).to(dev).to(torch.int32)
    # Python's own pow, so discount ** i matches replay_buffer.py:240, 253 bit-for-bit
    dp = torch.tensor([config.discount ** i for i in range(td + 1)], dtype=torch.float64, device=dev)
    tv = torch.empty((B, K + 1), dtype=torch.float64, device=dev)
    tr = torch.empty((B, K + 1), dtype=torch.float64, device=dev)
    tp = torch.empty((B, K + 1, games.A), dtype=torch.float64, device=dev)
    ta = torch.empty((B, K + 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib.mzb_make_target(ptr(games.reward), ptr(games.to_play), ptr(games.root_value), ptr(games.reanalysed),
                                       ptr(games.visits), ptr(games.action), ptr(games.game_start), ptr(games.game_len),
                                       games.A, ptr(bg), ptr(bi), ptr(bs), ptr(bt), B, K, td, ptr(dp),
                                       int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(tv), ptr(tr), ptr(tp), ptr(ta),
                                       _lib.current_stream()))
    return tv, tr, tp, ta


def _set_info(shared_storage, key, value):
    """shared_storage.set_info, Ray actor handle (`.remote`) or plain object (Trainer / Reanalyse / tests)."""
    remote = getattr(shared_storage.set_info, "remote", None)
    return remote(key, value) if remote is not None else shared_storage.set_info(key, value)


class ReplayBuffer:
    """ReplayBuffer(initial_checkpoint, initial_buffer, config) (replay_buffer.py:17-31) with the store on `device`."""

    def __init__(self, initial_checkpoint, initial_buffer, config, device=None, max_batch=None, record_env=None):
        """record_env: an envs.VectorEnv whose export ring will feed this buffer with `ingest` - the store then keeps
        the environment's compact observation records (packed int8 boards) and decodes them in get_batch."""
        self.config = config
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("the replay store lives on a CUDA device (there is no CPU path)")
        self.S = int(getattr(config, "stacked_observations", 0) or 0)
        c, h, w = config.observation_shape
        self.obs_out_shape = (c * (self.S + 1) + self.S, h, w)         # get_stacked_observations (self_play.py:514-548)
        self.A = len(config.action_space)
        self.obs_floats = int(np.prod(config.observation_shape))
        self.K, self.td = int(config.num_unroll_steps), int(config.td_steps)
        self.max_batch = int(max_batch or config.batch_size)
        decode, oh, ow, rec = 0, 0, 0, self.obs_floats
        if record_env is not None:
            rec = int(record_env.rec_floats)
            if record_env.kind in ("tictactoe", "connect4", "gomoku"):
                decode, (oh, ow) = 1, record_env.obs_shape[1:]
            elif rec != self.obs_floats:
                raise NotImplementedError(f"no record decoder for {record_env.kind!r} observations")
        self.decode, self.rec_floats, self.board = decode, rec, (oh, ow)
        self.cfg = ReplayConfig(self.A, rec, decode, oh, ow, int(config.replay_buffer_size), int(config.max_moves) + 2, self.K,
                                self.td, int(bool(config.PER)), self.max_batch, self.S, int(c), float(config.PER_alpha),
                                int(config.seed) & 0xFFFFFFFFFFFFFFFF)
        nbytes = _lib.lib.mzb_replay_workspace_bytes(C.byref(self.cfg))
        if nbytes == 0:
            check(-1)
        # Python's own pow, so discount ** i matches replay_buffer.py:240, 253 bit-for-bit
        dp = (C.c_double * (self.td + 1))(*[config.discount ** i for i in range(self.td + 1)])
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = (self.workspace.data_ptr() + 255) // 256 * 256
            self._h = _vp()
            check(_lib.lib.mzb_replay_create(C.byref(self._h), C.byref(self.cfg), _vp(base), nbytes, dp, _lib.current_stream()))
        self.nbytes = nbytes
        self._record_env_kind = None if record_env is None else record_env.kind
        # Resume (replay_buffer.py:17-24): the counters are the CHECKPOINT's - re-loading the buffered games does not
        # play them again - and the games keep their ids (the keys of `initial_buffer`): public id = device id + _id0.
        # With a consistent checkpoint the newest loaded game is num_played_games - 1 and the next saved game gets id
        # num_played_games, as in the reference.
        self._played0, self._id0 = (0, 0), 0
        initial = sorted((initial_buffer or {}).items())
        for _, game_history in initial:
            self.save_game(game_history)
        n0, s0 = self._info()[1], self._info()[2]
        self._played0 = (int(initial_checkpoint["num_played_games"]) - n0, int(initial_checkpoint["num_played_steps"]) - s0)
        self._id0 = int(initial[0][0]) if initial else int(initial_checkpoint["num_played_games"])

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                _lib.lib.mzb_replay_destroy(self._h)
            except (AttributeError, TypeError):      # interpreter shutdown
                pass
            self._h = None

    # -- counters (replay_buffer.py:20-24)
    def _info(self):
        out = (_i64 * 5)()
        check(_lib.lib.mzb_replay_info(self._h, out))
        return [int(x) for x in out]

    total_samples = property(lambda self: self._info()[0])
    num_played_games = property(lambda self: self._played0[0] + self._info()[1])
    num_played_steps = property(lambda self: self._played0[1] + self._info()[2])

    def __len__(self):
        return self._info()[3]

    # -- save_game (:33-64)
    def save_game(self, game_history, shared_storage=None):
        """One host GameHistory (self_play.py:480-495).  `child_visits` are ratios count / num_simulations; the store
        keeps the integer counts (`game_history.visit_counts` if present, else recovered exactly by rounding)."""
        n = len(game_history.root_values)
        counts = getattr(game_history, "visit_counts", None)
        if counts is None:
            counts = np.rint(np.asarray(game_history.child_visits, dtype=np.float64) * self.config.num_simulations)
        vis = np.zeros((n + 1, self.A), dtype=np.uint16)
        vis[:n] = np.asarray(counts, dtype=np.int64)
        if self.decode:           # [own, other, to-play] planes -> the environment's packed int8 record
            oh, ow = self.board
            raw = np.zeros((n + 1, self.rec_floats * 4), dtype=np.int8)
            for i, o in enumerate(game_history.observation_history):
                o = np.asarray(o)
                raw[i, :oh * ow] = ((o[0] == 1).astype(np.int8) - (o[1] == 1).astype(np.int8)).reshape(-1)
                raw[i, oh * ow] = int(o[2].flat[0])
            obs = raw.view(np.float32)
        else:
            obs = np.stack([np.asarray(o, dtype=np.float32).reshape(-1) for o in game_history.observation_history])
        rv = np.zeros(n + 1, dtype=np.float64)
        rv[:n] = game_history.root_values
        pr = None
        if self.config.PER and getattr(game_history, "priorities", None) is not None:
            pr = np.zeros(n + 1, dtype=np.float32)
            pr[:n] = game_history.priorities
        self.save_games_device([0], [n], self._dev(obs, torch.float32), self._dev(game_history.action_history, torch.int32),
                               self._dev(np.asarray(game_history.reward_history, dtype=np.float32), torch.float32),
                               self._dev(game_history.to_play_history, torch.int8), self._dev(rv, torch.float64),
                               torch.from_numpy(vis).to(self.device), None if pr is None else self._dev(pr, torch.float32))
        if shared_storage:
            _set_info(shared_storage, "num_played_games", self.num_played_games)
            _set_info(shared_storage, "num_played_steps", self.num_played_steps)

    def _dev(self, x, dt):
        return torch.as_tensor(np.ascontiguousarray(x)).to(self.device, dt).contiguous()

    def save_games_device(self, game_start, game_len, obs, action, reward, to_play, root_value, visits, priorities=None):
        """Games already on the device as entry arrays in the export layout (envs.VectorEnv export ring)."""
        n = len(game_len)
        hs = (C.c_int32 * n)(*[int(x) for x in game_start])
        hl = (C.c_int32 * n)(*[int(x) for x in game_len])
        with torch.cuda.device(self.device):
            check(_lib.lib.mzb_replay_save_games(self._h, n, hs, hl, ptr(obs), ptr(action), ptr(reward), ptr(to_play),
                                                 ptr(root_value), ptr(visits), ptr(priorities), _lib.current_stream()))

    def ingest(self, env):
        """Append every finished game in `env`'s export ring (device to device) and empty the ring: the hop that
        replaces `replay_buffer.save_game.remote(game_history)` (self_play.py:52).  Returns the number of games."""
        same_layout = (self.rec_floats == int(env.rec_floats) and self.A == int(env.A)
                       and self.decode == (1 if env.kind in ("tictactoe", "connect4", "gomoku") else 0)
                       and (not self.decode or tuple(self.board) == tuple(env.obs_shape[1:]))
                       and int(self.config.max_moves) >= int(env.max_moves))
        if not same_layout:
            # a store built with the reference's 3-argument constructor keeps decoded observations: take the games
            # through the host format instead of copying records with the wrong stride
            from .self_play import decode_export
            games = decode_export(env)
            for game_history in games:
                self.save_game(game_history)
            return len(games)
        n = _i32()
        with torch.cuda.device(self.device):
            check(_lib.lib.mzb_env_export_to_replay(env._h, self._h, C.byref(n), _lib.current_stream()))
        return n.value

    # -- get_batch (:69-140)
    def get_batch(self, batch_size=None, u_game=None, u_pos=None):
        """Returns (index_batch [B,2] i64, (observation_batch, action_batch, value_batch, reward_batch, policy_batch,
        weight_batch | None, gradient_scale_batch)) as device tensors.  u_game / u_pos [B] f64 inject the draws."""
        B = int(batch_size or self.config.batch_size)
        dev, K, A = self.device, self.K, self.A
        gid = torch.empty(B, dtype=torch.int64, device=dev)
        pos = torch.empty(B, dtype=torch.int32, device=dev)
        obs = torch.empty((B,) + self.obs_out_shape, dtype=torch.float32, device=dev)
        act = torch.empty((B, K + 1), dtype=torch.int32, device=dev)
        val = torch.empty((B, K + 1), dtype=torch.float64, device=dev)
        rew = torch.empty((B, K + 1), dtype=torch.float64, device=dev)
        pol = torch.empty((B, K + 1, A), dtype=torch.float64, device=dev)
        w = torch.empty(B, dtype=torch.float32, device=dev) if self.config.PER else None
        gs = torch.empty((B, K + 1), dtype=torch.int32, device=dev)
        ug = None if u_game is None else torch.as_tensor(u_game, dtype=torch.float64).to(dev).contiguous()
        up = None if u_pos is None else torch.as_tensor(u_pos, dtype=torch.float64).to(dev).contiguous()
        with torch.cuda.device(dev):
            check(_lib.lib.mzb_replay_get_batch(self._h, B, ptr(ug), ptr(up), ptr(gid), ptr(pos), None, None, ptr(obs),
                                                ptr(act), ptr(val), ptr(rew), ptr(pol), ptr(w), ptr(gs),
                                                _lib.current_stream()))
        if self._id0:
            gid = gid + self._id0                   # public game ids (resume keeps the checkpoint's numbering)
        return torch.stack([gid, pos.to(torch.int64)], dim=1), (obs, act, val, rew, pol, w, gs)

    # -- update_priorities (:202-220)
    def update_priorities(self, priorities, index_info):
        pr = torch.as_tensor(priorities).to(self.device, torch.float32).contiguous()
        if torch.is_tensor(index_info):             # stays on the device: no host sync on the learner loop
            idx = index_info.to(self.device, torch.int64)
        else:
            idx = torch.as_tensor(np.asarray(index_info, dtype=np.int64)).to(self.device)
        gid = (idx[:, 0] - self._id0).contiguous()
        pos = idx[:, 1].to(torch.int32).contiguous()
        with torch.cuda.device(self.device):
            check(_lib.lib.mzb_replay_update_priorities(self._h, pr.shape[0], ptr(pr), ptr(gid), ptr(pos), _lib.current_stream()))

    # -- Reanalyse support (:141-160, 194-200, 298-361)
    def sample_game(self, force_uniform=True, u=None):
        """A uniformly drawn buffered game id (the only form Reanalyse uses, :332-334)."""
        if not force_uniform:
            raise NotImplementedError("prioritised single-game draws go through get_batch")
        first, n = self._info()[4], len(self)
        u = float(np.random.random_sample()) if u is None else float(u)
        return self._id0 + first + int(u * n)

    def game_observations(self, game_id):
        """[len, C, H, W] float32 device tensor: the stored observations of a game as network input."""
        n = _i32()
        game_id = int(game_id) - self._id0
        check(_lib.lib.mzb_replay_game_observations(self._h, int(game_id), None, C.byref(n), _lib.current_stream()))
        obs = torch.empty((n.value,) + self.obs_out_shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(_lib.lib.mzb_replay_game_observations(self._h, int(game_id), ptr(obs), C.byref(n), _lib.current_stream()))
        return obs

    def set_reanalysed_values(self, game_id, values):
        """game_history.reanalysed_predicted_root_values = values (float32 [len]); ignored if the game was evicted."""
        v = torch.as_tensor(values).to(self.device, torch.float32).reshape(-1).contiguous()
        with torch.cuda.device(self.device):
            check(_lib.lib.mzb_replay_set_reanalysed(self._h, int(game_id) - self._id0, ptr(v), _lib.current_stream()))

    # -- get_buffer (:66-67): what muzero.py pickles into replay_buffer.pkl and hands back as `initial_buffer`
    def get_buffer(self):
        """{game_id: GameHistory} of every buffered game, host objects in the reference's format (self_play.py:480-495)
        with `priorities` / `game_priority` (PER) and the integer `visit_counts` the device store keeps."""
        from .self_play import GameHistory
        E = int(self.config.max_moves) + 2
        obs = np.empty((E, self.rec_floats), np.float32); act = np.empty(E, np.int32); rew = np.empty(E, np.float32)
        tp = np.empty(E, np.int8); rv = np.empty(E, np.float64); vis = np.empty((E, self.A), np.uint16)
        out = {}
        first, n_games = self._info()[4], len(self)
        for gid in range(first, first + n_games):
            n = _i32()
            check(_lib.lib.mzb_replay_export_game_sync(self._h, gid, C.byref(n), ptr(obs), ptr(act), ptr(rew), ptr(tp), ptr(rv),
                                                       ptr(vis), _lib.current_stream()))
            n = n.value
            gh = GameHistory()
            gh.observation_history = [self._decode_record(obs[i]) for i in range(n + 1)]
            gh.action_history = act[:n + 1].tolist()
            gh.reward_history = [0] + [float(x) for x in rew[1:n + 1]]
            gh.to_play_history = tp[:n + 1].astype(int).tolist()
            v = vis[:n].astype(np.int64)
            tot = v.sum(axis=1)
            gh.child_visits = [[int(v[i, a]) / int(tot[i]) if v[i, a] else 0 for a in range(self.A)] for i in range(n)]
            gh.visit_counts = v.copy()
            gh.root_values = rv[:n].tolist()
            if self.config.PER:
                gh.priorities, gh.game_priority = self.game_priorities(self._id0 + gid)
            out[self._id0 + gid] = gh
        return out

    def _decode_record(self, rec):
        if not self.decode:
            return rec.reshape(tuple(self.config.observation_shape)).copy()
        oh, ow = self.board
        raw = rec.view(np.int8)
        board = raw[:oh * ow].reshape(oh, ow)
        return np.array([board == 1, board == -1, np.full((oh, ow), int(raw[oh * ow]))], dtype=np.float32)

    def game_priorities(self, game_id):
        n = _i32()
        buf = np.zeros(int(self.config.max_moves) + 2, dtype=np.float32)
        gp = np.zeros(1, dtype=np.float32)
        check(_lib.lib.mzb_replay_game_priorities_sync(self._h, int(game_id) - self._id0, ptr(buf), ptr(gp), C.byref(n),
                                                       _lib.current_stream()))
        return buf[:n.value].copy(), gp[0]


class Reanalyse:
    """Reanalyse(initial_checkpoint, config) (replay_buffer.py:298-361) against the device store: whole stored games go
    through one batched `initial_inference` on the CUDA kernels and their fresh value predictions become the bootstrap
    values of later batches, without the games leaving the GPU."""

    def __init__(self, initial_checkpoint, config, device=None):
        from . import models
        self.config = config
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        np.random.seed(config.seed)
        torch.manual_seed(config.seed)
        self.model = models.MuZeroNetwork(config)
        if initial_checkpoint.get("weights") is not None:
            self.model.set_weights(initial_checkpoint["weights"])
        self.model.to(self.device)
        self.model.eval()
        self.num_reanalysed_games = int(initial_checkpoint.get("num_reanalysed_games", 0))

    def reanalyse_game(self, replay_buffer, game_id):
        from . import models
        if self.config.use_last_model_value:
            obs = replay_buffer.game_observations(game_id)
            values = models.support_to_scalar(self.model.initial_inference(obs)[0], self.config.support_size)
            replay_buffer.set_reanalysed_values(game_id, values.reshape(-1))
        self.num_reanalysed_games += 1

    def reanalyse(self, replay_buffer, shared_storage, max_games=None):
        """The reference's loop (:321-361) over plain `get_info` / `set_info` objects."""
        import time
        while shared_storage.get_info("num_played_games") < 1:
            time.sleep(0.1)
        done = 0
        while (shared_storage.get_info("training_step") < self.config.training_steps
               and not shared_storage.get_info("terminate")):
            weights = shared_storage.get_info("weights")
            if weights is not None:
                self.model.set_weights(weights)
            self.reanalyse_game(replay_buffer, replay_buffer.sample_game(force_uniform=True))
            shared_storage.set_info("num_reanalysed_games", self.num_reanalysed_games)
            done += 1
            if max_games is not None and done >= max_games:
                break
