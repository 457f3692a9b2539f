"""Self-play + search with the reference's names (self_play.py), on the batched device path.

  MCTS(config).run(model, observation, legal_actions, to_play, add_exploration_noise, override_root_with=None)
      -> (root Node, {"max_tree_depth", "root_predicted_value"})                       self_play.py:261-362
  Node / GameHistory (MinMaxStats lives per game in the device tree store)             :434-568
  SelfPlay(initial_checkpoint, Game, config, seed)  .play_game(...) -> GameHistory     :11-246
  SelfPlay.play_games(...)                          G games in lock-step per GPU (the B200 path)

`MCTS.run` is the G=1 compatibility wrapper: it runs the same kernels as the batched path and
rebuilds the Python Node graph from the device tree, so callers that walk `root.children`
(select_action, store_search_statistics, diagnose_model.py) keep working unchanged.
"""
import math

import numpy
import torch

from . import models
from .envs import VectorEnv, game_kind
from .search import BatchedMCTS


class Node:
    """Host view of one tree node with the reference's read API (self_play.py:434-450)."""

    def __init__(self, prior):
        self.visit_count = 0
        self.to_play = -1
        self.prior = prior
        self.value_sum = 0
        self.children = {}
        self.hidden_state = None
        self.reward = 0

    def expanded(self):
        return len(self.children) > 0

    def value(self):
        if self.visit_count == 0:
            return 0
        return self.value_sum / self.visit_count

    def expand(self, actions, to_play, reward, policy_logits, hidden_state):
        """Node.expand (self_play.py:452-466) for callers that pre-build a root (diagnose_model.py:57-69)."""
        self.to_play = to_play
        self.reward = reward
        self.hidden_state = hidden_state
        logits = torch.as_tensor(policy_logits).detach().float().cpu().reshape(-1)
        values = torch.softmax(torch.tensor([logits[a] for a in actions]), dim=0).tolist()
        for a, p in zip(actions, values):
            self.children[a] = Node(p)


class GameHistory:
    """Trajectory of one self-played game, field for field the reference's (self_play.py:480-548)."""

    def __init__(self):
        self.observation_history = []
        self.action_history = []
        self.reward_history = []
        self.to_play_history = []
        self.child_visits = []
        self.root_values = []
        self.reanalysed_predicted_root_values = None
        self.priorities = None
        self.game_priority = None

    def store_search_statistics(self, root, action_space):
        if root is not None:
            sum_visits = sum(child.visit_count for child in root.children.values())
            self.child_visits.append([root.children[a].visit_count / sum_visits if a in root.children else 0
                                      for a in action_space])
            self.root_values.append(root.value())
        else:
            self.root_values.append(None)

    def get_stacked_observations(self, index, num_stacked_observations):
        index = index % len(self.observation_history)
        stacked = self.observation_history[index].copy()
        for past in reversed(range(index - num_stacked_observations, index)):
            if 0 <= past:
                previous = numpy.concatenate((self.observation_history[past],
                                              [numpy.ones_like(stacked[0]) * self.action_history[past + 1]]))
            else:
                previous = numpy.concatenate((numpy.zeros_like(self.observation_history[index]),
                                              [numpy.zeros_like(stacked[0])]))
            stacked = numpy.concatenate((stacked, previous))
        return stacked


def _model_device(model):
    return next(model.parameters()).device


class MCTS:
    """`MCTS(config).run(...)`: one search, executed by the batched kernels with G = 1."""

    _cache = {}
    _calls = 0          # searches started without an explicit RNG step: each draws fresh noise / tie-breaks

    def __init__(self, config):
        self.config = config

    def _engine(self, device):
        cfg = self.config
        key = (id(cfg), str(device), cfg.num_simulations, len(cfg.action_space))
        eng = MCTS._cache.get(key)
        if eng is None:
            eng = MCTS._cache[key] = BatchedMCTS(cfg, 1, device=device)
        return eng

    def run(self, model, observation, legal_actions, to_play, add_exploration_noise, override_root_with=None,
            noise=None, slot=None, step=None):
        cfg = self.config
        device = _model_device(model)
        if device.type != "cuda":
            raise RuntimeError("MCTS.run: the B200 hot path has no CPU fallback - move the model to a CUDA device")
        if len(cfg.players) > 2:
            raise NotImplementedError("More than two player mode not implemented.")
        A = len(cfg.action_space)
        eng = self._engine(device)
        if override_root_with:
            return self._run_from_root(model, eng, override_root_with, to_play, add_exploration_noise, noise)
        assert legal_actions, f"Legal actions should not be an empty array. Got {legal_actions}."
        assert set(legal_actions).issubset(set(cfg.action_space)), "Legal actions should be a subset of the action space."
        obs = torch.as_tensor(numpy.array(observation)).float().unsqueeze(0).to(device)
        legal = torch.zeros((1, A), dtype=torch.uint8, device=device)
        legal[0, list(legal_actions)] = 1
        nz = None
        if noise is not None:
            nz = torch.zeros((1, A), dtype=torch.float64, device=device)
            nz[0, list(legal_actions)] = torch.as_tensor(numpy.asarray(noise, dtype=numpy.float64), device=device)
        if step is None:
            # the reference draws from numpy's advancing global stream: a call without counters must not repeat the
            # previous call's Dirichlet sample and tie-breaks, so the call number becomes the RNG step
            step = MCTS._calls & 0x7FFFFFFF
            MCTS._calls += 1
        out = eng.run(model, obs, legal, torch.tensor([to_play], dtype=torch.int8, device=device), add_exploration_noise,
                      noise=nz, slot=None if slot is None else torch.tensor([slot], dtype=torch.int32, device=device),
                      step=torch.tensor([step], dtype=torch.int32, device=device))
        root = self._materialise(eng, 0, list(legal_actions), to_play)
        info = {"max_tree_depth": int(out["max_depth"][0]), "root_predicted_value": float(out["root_predicted_value"][0])}
        return root, info

    def _run_from_root(self, model, eng, root, to_play, add_exploration_noise, noise):
        """override_root_with (self_play.py:276-278): the caller expanded the root itself."""
        cfg = self.config
        device = eng.device
        A = len(cfg.action_space)
        tree = eng.tree
        actions = list(root.children.keys())
        pri = torch.zeros((1, A), dtype=torch.float32, device=device)
        legal = torch.zeros((1, A), dtype=torch.uint8, device=device)
        for a in actions:
            pri[0, a] = float(root.children[a].prior)
            legal[0, a] = 1
        nz = None
        if noise is not None:
            nz = torch.zeros((1, A), dtype=torch.float64, device=device)
            nz[0, actions] = torch.as_tensor(numpy.asarray(noise, dtype=numpy.float64), device=device)
        frac = float(cfg.root_exploration_fraction) if add_exploration_noise else 0.0
        step = torch.tensor([MCTS._calls & 0x7FFFFFFF], dtype=torch.int32, device=device)
        MCTS._calls += 1
        tree.root_init(torch.tensor([float(root.reward)], device=device), pri, False, legal,
                       torch.tensor([to_play], dtype=torch.int8, device=device), nz, cfg.root_dirichlet_alpha, frac,
                       None, step)
        hs = root.hidden_state.to(device).float()
        hidden = torch.empty((1, cfg.num_simulations + 1, hs.numel()), device=device)
        hidden[0, 0] = hs.reshape(-1)
        parent = torch.empty(1, dtype=torch.int32, device=device)
        action = torch.empty(1, dtype=torch.int32, device=device)
        for sim in range(cfg.num_simulations):
            tree.select(parent, action)
            state = hidden[0, int(parent[0])].reshape(hs.shape)
            v, r, pl, ns = model.recurrent_inference(state, action.reshape(1, 1))
            hidden[0, sim + 1] = ns.reshape(-1)
            tree.expand_backup(models.support_to_scalar(v, cfg.support_size).reshape(1).contiguous(),
                               models.support_to_scalar(r, cfg.support_size).reshape(1).contiguous(), pl.contiguous(), True)
        stats = tree.root_stats()
        new_root = self._materialise(eng, 0, actions, to_play, hidden_shape=tuple(hs.shape), hidden=hidden)
        return new_root, {"max_tree_depth": int(stats["max_depth"][0]), "root_predicted_value": None}

    def _materialise(self, eng, game, legal_actions, to_play, hidden_shape=None, hidden=None):
        """Device tree of one game -> the reference's Node graph (children dict in action order)."""
        cfg = self.config
        ex = eng.tree.export_game(game)
        nhwc = None
        if hidden is None:
            hidden = eng.tree.hidden()
            if hidden is None:                         # residual networks: dense NHWC pool of the search driver
                hidden = eng._pool
                c, hh, ww = (cfg.channels, ) + ((-(-cfg.observation_shape[1] // 16), -(-cfg.observation_shape[2] // 16))
                                                if cfg.downsample else tuple(cfg.observation_shape[1:]))
                nhwc = (hh, ww, c)
        players = cfg.players
        A = len(cfg.action_space)

        def build(slot, node, depth, tp):
            node.to_play = tp
            h = hidden[game, slot].float()
            if nhwc is not None:
                node.hidden_state = h.reshape(nhwc).permute(2, 0, 1).unsqueeze(0).contiguous()
            else:
                node.hidden_state = (h.reshape(hidden_shape) if hidden_shape else h.reshape(1, -1)).clone()
            nxt = players[(players.index(tp) + 1) % len(players)]
            for a in (legal_actions if slot == 0 else range(A)):
                prior = float(ex["root_prior"][a]) if slot == 0 else float(ex["prior"][slot][a])
                ch = Node(prior)
                ch.visit_count = int(ex["visit"][slot][a])
                ch.value_sum = float(ex["value_sum"][slot][a]) if ch.visit_count else 0
                ch.reward = float(ex["reward"][slot][a]) if int(ex["child"][slot][a]) >= 0 else 0
                node.children[a] = ch
                cs = int(ex["child"][slot][a])
                if cs >= 0:
                    build(cs, ch, depth + 1, nxt)

        root = Node(0)
        root.visit_count = ex["root_visit"]
        root.value_sum = ex["root_value_sum"]
        root.reward = ex["root_reward"]
        build(0, root, 0, to_play)
        return root


def decode_export(env):
    """Empty the device export ring of `env` into reference-format GameHistory objects (self_play.py:480-495);
    child_visits = count / sum(counts) in float64, bit-identical to store_search_statistics (:497-512)."""
    ne, ng, h = env.drain_raw()
    games = []
    A = env.A
    for k in range(ng):
        s, n = int(h["start"][k]), int(h["len"][k])
        gh = GameHistory()
        gh.observation_history = [env.decode_observation(h["obs"][s + i]) for i in range(n + 1)]
        gh.action_history = h["action"][s:s + n + 1].tolist()
        gh.reward_history = [float(x) for x in h["reward"][s:s + n + 1]]
        gh.reward_history[0] = 0
        gh.to_play_history = h["to_play"][s:s + n + 1].astype(int).tolist()
        v = h["visits"][s:s + n].astype(numpy.int64)
        tot = v.sum(axis=1)
        gh.child_visits = [[int(v[i, a]) / int(tot[i]) if v[i, a] else 0 for a in range(A)] for i in range(n)]
        gh.root_values = h["root_value"][s:s + n].tolist()
        gh.slot = int(h["slot"][k])
        games.append(gh)
    return games


class SelfPlay:
    """Self-play worker.  `play_games` advances G games in lock-step on one GPU; `play_game` keeps the
    reference's one-game signature (self_play.py:110-184) on top of it.

    `Game` is honoured as in the reference: the built-in `games/<name>.Game` classes (they carry a `KIND`) and
    `Game=None` select the vectorised device environment of the config's game; any OTHER `AbstractGame` subclass is
    instantiated as `Game(seed)` and played on the host through the step / legal_actions / to_play contract, one
    move at a time, with every search on the device kernels (G = 1)."""

    def __init__(self, initial_checkpoint, Game, config, seed, n_games=None, device=None, first_slot=0):
        self.config = config
        self.Game = Game
        self.seed = seed
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("SelfPlay: the B200 hot path has no CPU fallback")
        numpy.random.seed(seed)
        torch.manual_seed(seed)
        self.model = models.MuZeroNetwork(config)
        if initial_checkpoint is not None and initial_checkpoint.get("weights") is not None:
            self.model.set_weights(initial_checkpoint["weights"])
        self.model.to(self.device)
        self.model.eval()
        self.G = int(n_games if n_games is not None else getattr(config, "num_parallel_games", 1))
        self.first_slot = first_slot
        self._env = None
        self._mcts = None
        self._solo = None
        self._game = None
        self.num_played_games = 0
        self.num_played_steps = 0
        self.host_game = Game is not None and getattr(Game, "KIND", None) is None
        if self.host_game and n_games not in (None, 1):
            raise NotImplementedError("a caller-supplied Game class is played one game at a time on the host "
                                      "(n_games = 1); the vectorised path runs the built-in device environments")

    @property
    def game(self):
        """The host `Game(seed)` object (reference: self.game), created on first use."""
        if self._game is None:
            if self.Game is None:
                import importlib
                self.Game = importlib.import_module(f"{__package__}.games.{game_kind(self.config)}").Game
            self._game = self.Game(self.seed)
        return self._game

    # -- batched device loop
    def _setup(self):
        if self._env is None:
            cfg = self.config
            self._env = VectorEnv(game_kind(cfg), self.G, cfg.max_moves, seed=self.seed, first_slot=self.first_slot,
                                  device=self.device)
            self._mcts = BatchedMCTS(cfg, self.G, device=self.device, seed=self.seed)
        return self._env, self._mcts

    def step(self, temperature=1.0, temperature_threshold=None, add_exploration_noise=True, export=True,
             allow_fused=True, search_events=None):
        """One move of every game: observe -> search -> select_action + Game.step + record -> harvest.
        search_events: optional (start, end) CUDA events recorded around the search launch (bench: in-step timing)."""
        env, mcts = self._setup()
        obs, legal, to_play = env.observe_stacked(int(getattr(self.config, "stacked_observations", 0) or 0))
        if search_events is not None:
            search_events[0].record()
        out = mcts.run(self.model, obs, legal, to_play, add_exploration_noise, slot=env.slot, step=env.step_count,
                       allow_fused=allow_fused, out=getattr(self, "_out", None))
        if search_events is not None:
            search_events[1].record()
        self._out = out
        env.act_step(out["visits"], out["root_value"], legal, temperature, temperature_threshold)
        env.harvest(export)

    def play_games(self, n_moves, temperature=1.0, temperature_threshold=None, drain_every=4):
        """Advance every game by n_moves moves; returns the GameHistory objects of games that finished.
        The export ring is drained to the host every `drain_every` moves (0: never - the caller drains)."""
        done = []
        for i in range(n_moves):
            self.step(temperature, temperature_threshold)
            if drain_every and (i + 1) % drain_every == 0:
                done += self.drain()
        if drain_every:
            done += self.drain()
        return done

    def drain(self):
        env, _ = self._setup()
        games = decode_export(env)
        self.num_played_games += len(games)
        self.num_played_steps += sum(len(g.root_values) for g in games)
        return games

    # -- the actor loop (self_play.py:30-108)
    def continuous_self_play(self, shared_storage, replay_buffer, test_mode=False, max_moves=None):
        """The reference's actor loop over G lock-step games: refresh the weights, play, hand finished games to the
        replay buffer, respect self_play_delay / ratio.  `shared_storage` is anything with get_info / set_info (plain
        methods or Ray-style `.remote`); a replay buffer with `ingest` (the device store, replay_buffer.ReplayBuffer)
        receives the games device to device, any other through `save_game(game_history, shared_storage)`.
        test_mode (:54-88): greedy games, one at a time, against config.opponent; their statistics go to the shared
        storage.  `max_moves` bounds the loop for callers without a trainer (None: until training_steps / terminate):
        moves of the lock-step games, or whole games in test mode."""
        import time

        def call(method, *args):
            remote = getattr(method, "remote", None)
            if remote is None:
                return method(*args)
            import ray
            out = remote(*args)
            return ray.get(out) if method.__name__.startswith("get") else out

        cfg = self.config
        moves = 0
        while (call(shared_storage.get_info, "training_step") < cfg.training_steps
               and not call(shared_storage.get_info, "terminate")):
            weights = call(shared_storage.get_info, "weights")
            if weights is not None:
                self.model.set_weights(weights)
            if test_mode:
                gh = self.play_game(0, cfg.temperature_threshold, False,
                                    "self" if len(cfg.players) == 1 else cfg.opponent, cfg.muzero_player)
                call(shared_storage.set_info, {"episode_length": len(gh.action_history) - 1,
                                               "total_reward": sum(gh.reward_history),
                                               "mean_value": numpy.mean([v for v in gh.root_values if v])})
                if 1 < len(cfg.players):
                    mine = [gh.to_play_history[i - 1] == cfg.muzero_player for i in range(len(gh.reward_history))]
                    call(shared_storage.set_info,
                         {"muzero_reward": sum(r for r, m in zip(gh.reward_history, mine) if m),
                          "opponent_reward": sum(r for r, m in zip(gh.reward_history, mine) if not m)})
            elif self.host_game:
                gh = self.play_game(cfg.visit_softmax_temperature_fn(trained_steps=call(shared_storage.get_info, "training_step")),
                                    cfg.temperature_threshold, False, "self", 0)
                call(replay_buffer.save_game, gh, shared_storage)
            else:
                env, _ = self._setup()
                T = cfg.visit_softmax_temperature_fn(trained_steps=call(shared_storage.get_info, "training_step"))
                self.step(T, cfg.temperature_threshold)
                if hasattr(replay_buffer, "ingest"):
                    n = replay_buffer.ingest(env)
                    if n:
                        call(shared_storage.set_info, "num_played_games", replay_buffer.num_played_games)
                        call(shared_storage.set_info, "num_played_steps", replay_buffer.num_played_steps)
                else:
                    for game_history in self.drain():
                        call(replay_buffer.save_game, game_history, shared_storage)
            moves += 1
            if max_moves is not None and moves >= max_moves:
                break
            if not test_mode and cfg.self_play_delay:
                time.sleep(cfg.self_play_delay)
            if not test_mode and cfg.ratio:
                while (call(shared_storage.get_info, "training_step") / max(1, call(shared_storage.get_info, "num_played_steps")) < cfg.ratio
                       and call(shared_storage.get_info, "training_step") < cfg.training_steps
                       and not call(shared_storage.get_info, "terminate")):
                    time.sleep(0.5)
        self.close_game()

    # -- reference-compatible single game
    def play_game(self, temperature, temperature_threshold, render, opponent, muzero_player):
        """One complete game (self_play.py:110-184).  Self-play of a built-in game runs entirely on the device
        (G = 1: environment, search, action selection and history stay on the GPU); a caller-supplied Game class,
        an opponent ("human" / "expert" / "random") or rendering take the host loop, whose searches are `MCTS.run` on
        the same kernels.  stacked_observations > 0 is assembled on the device from the running histories."""
        cfg = self.config
        if self.host_game or render or (opponent != "self" and len(cfg.players) > 1):
            return self._play_game_host(temperature, temperature_threshold, render, opponent, muzero_player)
        if self._solo is None:
            # kept across calls: its environment step counter (the RNG stream position) keeps advancing, so repeated
            # games with unchanged weights differ like the reference's do under numpy's advancing global stream
            self._solo = SelfPlay(None, self.Game, cfg, self.seed, n_games=1, device=self.device, first_slot=self.first_slot)
            self._solo.model = self.model
        solo = self._solo
        for _ in range(cfg.max_moves + 1):
            solo.step(float(temperature), temperature_threshold)
            games = solo.drain()
            if games:
                return games[0]
        raise RuntimeError("game did not finish within max_moves")

    def _play_game_host(self, temperature, temperature_threshold, render, opponent, muzero_player):
        """play_game over the AbstractGame contract on the host (any Game class, any opponent)."""
        cfg, game = self.config, self.game
        gh = GameHistory()
        observation = game.reset()
        gh.action_history.append(0)
        gh.observation_history.append(observation)
        gh.reward_history.append(0)
        gh.to_play_history.append(game.to_play())
        done = False
        if render:
            game.render()
        while not done and len(gh.action_history) <= cfg.max_moves:
            shape = numpy.array(observation).shape
            assert len(shape) == 3, f"Observation should be 3 dimensionnal instead of {len(shape)} dimensionnal. Got observation of shape: {shape}"
            assert shape == tuple(cfg.observation_shape), f"Observation should match the observation_shape defined in MuZeroConfig. Expected {cfg.observation_shape} but got {shape}."
            stacked = gh.get_stacked_observations(-1, cfg.stacked_observations)
            if opponent == "self" or muzero_player == game.to_play():
                root, info = MCTS(cfg).run(self.model, stacked, game.legal_actions(), game.to_play(), True)
                greedy = temperature_threshold and len(gh.action_history) >= temperature_threshold
                action = self.select_action(root, 0 if greedy else temperature)
                if render:
                    print(f'Tree depth: {info["max_tree_depth"]}')
                    print(f"Root value for player {game.to_play()}: {root.value():.2f}")
            else:
                action, root = self.select_opponent_action(opponent, stacked)
            observation, reward, done = game.step(action)
            if render:
                print(f"Played action: {game.action_to_string(action)}")
                game.render()
            gh.store_search_statistics(root, cfg.action_space)
            gh.action_history.append(action)
            gh.observation_history.append(observation)
            gh.reward_history.append(reward)
            gh.to_play_history.append(game.to_play())
        return gh

    def select_opponent_action(self, opponent, stacked_observations):
        """The evaluation opponents (self_play.py:189-221)."""
        game = self.game
        if opponent == "human":
            root, info = MCTS(self.config).run(self.model, stacked_observations, game.legal_actions(), game.to_play(), True)
            print(f'Tree depth: {info["max_tree_depth"]}')
            print(f"Root value for player {game.to_play()}: {root.value():.2f}")
            print(f"Player {game.to_play()} turn. MuZero suggests {game.action_to_string(self.select_action(root, 0))}")
            return game.human_to_action(), root
        if opponent == "expert":
            return game.expert_agent(), None
        if opponent == "random":
            legal = game.legal_actions()
            assert legal, f"Legal actions should not be an empty array. Got {legal}."
            assert set(legal).issubset(set(self.config.action_space)), "Legal actions should be a subset of the action space."
            return numpy.random.choice(legal), None
        raise NotImplementedError('Wrong argument: "opponent" argument should be "self", "human", "expert" or "random"')

    def close_game(self):
        if self._game is not None and hasattr(self._game, "close"):
            self._game.close()
        self._env = None
        self._mcts = None
        self._solo = None
        self._game = None

    @staticmethod
    def select_action(node, temperature):
        """Host-side select_action on a Node (self_play.py:223-246); the batched path does this on the device."""
        visit_counts = numpy.array([child.visit_count for child in node.children.values()], dtype="int32")
        actions = [action for action in node.children.keys()]
        if temperature == 0:
            return actions[numpy.argmax(visit_counts)]
        if temperature == float("inf"):
            return numpy.random.choice(actions)
        dist = visit_counts ** (1 / temperature)
        dist = dist / sum(dist)
        return numpy.random.choice(actions, p=dist)
