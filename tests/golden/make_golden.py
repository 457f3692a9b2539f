"""Generate the committed golden vectors by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py [family ...]      families: tree env codec net targets action episode replay trainer

The reference ships no tests or golden vectors, so every fixture here is an
output of the reference's own code under /root/reference, imported through
oracle.ref_loader (ray stub + gym shim).  Randomness is injected: `numpy.random.choice` and
`numpy.random.dirichlet` are patched with the counter-based rules of oracle/rng.py, and for
the tree family `models.support_to_scalar` is patched to pass the table value through, so a
search is a pure function of the committed tables.

The .npz files written next to this script are what `-m "not gpu"` tests check the oracle
against and what `-m gpu` tests check the CUDA path against.  /root/reference is never read
at test time.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader, rng  # noqa: E402

torch.set_num_threads(1)

SHAPES = {
    # name: (A, players, sims, discount, dirichlet alpha, n_cases)
    "cartpole": (2, 1, 50, 0.997, 0.25, 24),
    "tictactoe": (9, 2, 25, 1, 0.1, 24),
    "connect4": (7, 2, 200, 1, 0.3, 8),
    "gomoku": (121, 2, 400, 1, 0.3, 2),
    "breakout": (4, 1, 30, 0.997, 0.25, 8),
    "flat": (5, 2, 40, 1, 0.3, 6),          # all-equal logits and values: ties at every level
}
N_ROWS = 2048
SEED = 20261018


def child_row(parent_row, action, n_rows=N_ROWS):
    """Integer hash (parent row, action) -> row; mirrored in tests/_tables.py for the CUDA path."""
    x = (parent_row * 0x9E3779B1 + (action + 1) * 0x85EBCA77) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x2C1B3C6D) & 0xFFFFFFFF
    x ^= x >> 12
    return x % n_rows


class _Cfg:
    pass


def make_cfg(A, players, sims, discount, alpha):
    c = _Cfg()
    c.action_space = list(range(A))
    c.players = list(range(players))
    c.num_simulations = sims
    c.discount = discount
    c.root_dirichlet_alpha = alpha
    c.root_exploration_fraction = 0.25
    c.pb_c_base = 19652
    c.pb_c_init = 1.25
    c.support_size = 10
    return c


class TableModel:
    """Fake network: outputs are table rows; hidden state is the row index."""

    def __init__(self, V, Rw, L):
        self.V, self.Rw, self.L = V, Rw, L
        self._p = torch.zeros(1)
        self.sim = 0
        self.depth = 0

    def parameters(self):
        yield self._p

    def _out(self, row):
        return (torch.tensor([[self.V[row]]]), torch.tensor([[self.Rw[row]]]),
                torch.tensor(self.L[row:row + 1]), torch.tensor([[row]], dtype=torch.int64))

    def initial_inference(self, observation):
        return self._out(int(observation.flatten()[0].item()))

    def recurrent_inference(self, hidden, action):
        out = self._out(child_row(int(hidden.item()), int(action.item())))
        self.sim += 1           # one recurrent call closes one simulation (self_play.py:340)
        self.depth = 0
        return out


def dfs_dump(root, n_actions):
    """Canonical dump of every created child Node: DFS in action order."""
    rows = []

    def walk(node, depth):
        for a, ch in node.children.items():
            rows.append((depth + 1, a, ch.visit_count, float(ch.value_sum), float(ch.reward), float(ch.prior),
                         1 if ch.expanded() else 0))
            if ch.expanded():
                walk(ch, depth + 1)

    walk(root, 0)
    return np.array(rows, dtype=np.float64).reshape(-1, 7)


def gen_tree():
    self_play, models = ref_loader.load("self_play", "models")
    out = {}
    rs = np.random.RandomState(SEED)
    orig_choice, orig_dir, orig_s2s = np.random.choice, np.random.dirichlet, models.support_to_scalar
    try:
        for name, (A, players, sims, discount, alpha, n_cases) in SHAPES.items():
            if name == "flat":
                V = np.zeros(N_ROWS, dtype=np.float32)
                Rw = np.zeros(N_ROWS, dtype=np.float32)
                L = np.zeros((N_ROWS, A), dtype=np.float32)
            else:
                V = rs.uniform(-3, 3, N_ROWS).astype(np.float32)
                Rw = (rs.uniform(-1, 1, N_ROWS) * (rs.uniform(size=N_ROWS) < 0.5)).astype(np.float32)
                L = rs.normal(0, 1.5, (N_ROWS, A)).astype(np.float32)
            # interior priors exactly as Node.expand computes them (1-D f32 torch softmax)
            P = np.stack([torch.softmax(torch.tensor([torch.tensor(L)[r][a] for a in range(A)]), dim=0).numpy()
                          for r in range(N_ROWS)]) if A <= 16 else \
                np.stack([torch.softmax(torch.tensor(L[r]), dim=0).numpy() for r in range(N_ROWS)])
            out[f"{name}/V"], out[f"{name}/Rw"], out[f"{name}/L"], out[f"{name}/P"] = V, Rw, L, P
            out[f"{name}/meta"] = np.array([A, players, sims, discount, alpha, n_cases], dtype=np.float64)
            cfg = make_cfg(A, players, sims, discount, alpha)
            for case in range(n_cases):
                root_row = int(rs.randint(N_ROWS))
                if case % 3 == 0 or A == 2:
                    legal = list(range(A))
                else:
                    k = int(rs.randint(1, A + 1))
                    legal = sorted(rs.choice(A, size=k, replace=False).tolist())
                to_play = int(rs.randint(players))
                use_noise = case % 4 != 3
                noise = rs.dirichlet([alpha] * len(legal)) if use_noise else None
                slot, step = 1000 + case, 7 * case
                model = TableModel(V, Rw, L)

                # depth must advance on every select_child call, tie or not
                def choice_counting(a, size=None, replace=True, p=None, _m=model):
                    a = list(a)
                    i = rng.tie_index(SEED, slot, step, _m.sim, _m.depth, len(a)) if len(a) > 1 else 0
                    _m.depth += 1
                    return a[i]

                np.random.choice = choice_counting
                np.random.dirichlet = lambda alphas, _n=noise: _n
                models.support_to_scalar = lambda logits, support_size: logits
                root, info = self_play.MCTS(cfg).run(
                    model, np.full((1, 1, 1), root_row, dtype=np.float32), legal, to_play, use_noise)
                root_priors = torch.softmax(torch.tensor([torch.tensor(L[root_row:root_row + 1])[0][a] for a in legal]),
                                            dim=0).numpy()
                pre = f"{name}/{case}/"
                out[pre + "root_row"] = np.int64(root_row)
                out[pre + "legal"] = np.array(legal, dtype=np.int32)
                out[pre + "to_play"] = np.int64(to_play)
                out[pre + "noise"] = noise if use_noise else np.zeros(0)
                out[pre + "slot_step"] = np.array([slot, step], dtype=np.int64)
                out[pre + "root_priors_f32"] = root_priors
                out[pre + "visits"] = np.array([c.visit_count for c in root.children.values()], dtype=np.int32)
                out[pre + "value_sums"] = np.array([c.value_sum for c in root.children.values()], dtype=np.float64)
                out[pre + "rewards"] = np.array([c.reward for c in root.children.values()], dtype=np.float64)
                out[pre + "priors"] = np.array([c.prior for c in root.children.values()], dtype=np.float64)
                out[pre + "root"] = np.array([root.visit_count, root.value_sum, root.value(),
                                              info["max_tree_depth"], info["root_predicted_value"]], dtype=np.float64)
                out[pre + "dfs"] = dfs_dump(root, A)
    finally:
        np.random.choice, np.random.dirichlet, models.support_to_scalar = orig_choice, orig_dir, orig_s2s
    np.savez_compressed(os.path.join(HERE, "tree.npz"), **out)
    print("tree.npz:", len(out), "arrays")


def gen_action():
    """select_action (self_play.py:223-246) and store_search_statistics (:497-512)."""
    self_play = ref_loader.load("self_play")
    rs = np.random.RandomState(SEED + 1)
    out = {}
    orig_choice = np.random.choice
    cases = []
    try:
        for i in range(400):
            A = int(rs.choice([2, 4, 7, 9, 121]))
            k = int(rs.randint(1, A + 1))
            actions = sorted(rs.choice(A, size=k, replace=False).tolist())
            sims = int(rs.choice([25, 30, 50, 200, 400]))
            visits = rs.multinomial(sims, rs.dirichlet([0.5] * k)).astype(np.int32)
            T = [0, 0.25, 0.5, 1.0, float("inf")][i % 5]
            u = float(rs.random_sample()) if i % 7 else [0.0, 0.999999999999][i % 2]
            root = self_play.Node(0)
            for a, v in zip(actions, visits):
                root.children[a] = self_play.Node(0.1)
                root.children[a].visit_count = int(v)

            def choice(a, size=None, replace=True, p=None, _u=u):
                a = list(a)
                if p is None:
                    return a[int(_u * len(a))]
                cdf = np.cumsum(p)
                cdf /= cdf[-1]
                return a[int(np.searchsorted(cdf, _u, side="right"))]

            np.random.choice = choice
            act = self_play.SelfPlay.select_action(root, T)
            gh = self_play.GameHistory()
            root.visit_count = sims
            root.value_sum = 1.5
            gh.store_search_statistics(root, list(range(A)))
            cases.append((A, actions, visits, T, u, int(act), np.array(gh.child_visits[0], dtype=np.float64)))
    finally:
        np.random.choice = orig_choice
    # cross-check the injected inverse-CDF against numpy's own choice(p=...) on a seeded generator
    ok = 0
    for (A, actions, visits, T, u, act, cv) in cases[:200]:
        if T in (0, float("inf")):
            continue
        p = visits.astype(np.float64) ** (1 / T)
        p = p / sum(p)
        st = np.random.RandomState(5)
        u0 = np.random.RandomState(5).random_sample()
        a_np = st.choice(actions, p=p)
        cdf = np.cumsum(p); cdf /= cdf[-1]
        assert a_np == actions[int(np.searchsorted(cdf, u0, side="right"))]
        ok += 1
    out["n"] = np.int64(len(cases))
    for i, (A, actions, visits, T, u, act, cv) in enumerate(cases):
        out[f"{i}/A"] = np.int64(A)
        out[f"{i}/actions"] = np.array(actions, dtype=np.int32)
        out[f"{i}/visits"] = visits
        out[f"{i}/T"] = np.float64(T)
        out[f"{i}/u"] = np.float64(u)
        out[f"{i}/action"] = np.int64(act)
        out[f"{i}/child_visits"] = cv
    np.savez_compressed(os.path.join(HERE, "action.npz"), **out)
    print("action.npz:", len(cases), "cases; numpy.choice equivalence verified on", ok)


def gen_env():
    """Random action sequences through the reference TicTacToe / Connect4 / Gomoku classes."""
    out = {}
    rs = np.random.RandomState(SEED + 2)
    for name, n_games in (("tictactoe", 64), ("connect4", 64), ("gomoku", 24)):
        mod = ref_loader.load(f"games.{name}")
        A = {"tictactoe": 9, "connect4": 7, "gomoku": 121}[name]
        for g in range(n_games):
            game = mod.Game(seed=0)
            obs = game.reset()
            acts, rews, dones, boards, obss, legals, tps = [], [], [], [np.array(game.env.board)], [np.array(obs, dtype=np.float32)], [], []
            m = np.zeros(A, dtype=np.uint8); m[game.legal_actions()] = 1
            legals.append(m); tps.append(game.to_play())
            done = False
            illegal_ok = (g % 8 == 7)            # some games also play illegal moves (overwrite / full column)
            while not done and len(acts) < 130:
                la = game.legal_actions()
                a = int(rs.randint(A)) if (illegal_ok and rs.uniform() < 0.3) or not la else int(rs.choice(la))
                obs, r, done = game.step(a)
                acts.append(a); rews.append(r); dones.append(done)
                boards.append(np.array(game.env.board)); obss.append(np.array(obs, dtype=np.float32))
                m = np.zeros(A, dtype=np.uint8); m[game.legal_actions()] = 1
                legals.append(m); tps.append(game.to_play())
            pre = f"{name}/{g}/"
            out[pre + "actions"] = np.array(acts, dtype=np.int32)
            out[pre + "rewards"] = np.array(rews, dtype=np.float64)
            out[pre + "dones"] = np.array(dones, dtype=np.uint8)
            out[pre + "boards"] = np.array(boards, dtype=np.int32)
            out[pre + "obs"] = np.array(obss, dtype=np.float32)
            out[pre + "legal"] = np.array(legals, dtype=np.uint8)
            out[pre + "to_play"] = np.array(tps, dtype=np.int32)
        out[f"{name}/n"] = np.int64(n_games)
    # cartpole through the reference wrapper + gym shim (physics = oracle.games; parity unpinned there)
    mod = ref_loader.load("games.cartpole")
    for g in range(8):
        game = mod.Game(seed=g)
        obs = game.reset()
        obss, acts, dones = [np.array(obs, dtype=np.float32)], [], []
        done = False
        while not done:
            a = int(rs.randint(2)) if g < 6 else int(obss[-1].flatten()[2] > 0)   # last two: a balancing policy
            obs, r, done = game.step(a)
            assert r == 1.0
            obss.append(np.array(obs, dtype=np.float32)); acts.append(a); dones.append(done)
        out[f"cartpole/{g}/obs"] = np.array(obss, dtype=np.float32)
        out[f"cartpole/{g}/actions"] = np.array(acts, dtype=np.int32)
    out["cartpole/n"] = np.int64(8)
    np.savez_compressed(os.path.join(HERE, "env.npz"), **out)
    print("env.npz:", len(out), "arrays")


def gen_codec():
    models = ref_loader.load("models")
    rs = np.random.RandomState(SEED + 3)
    logits = np.concatenate([
        rs.normal(0, 3, (256, 21)),
        np.eye(21) * 30 - 15,                                   # near one-hot rows -> exact integers
        np.log(np.eye(21)[10:11] + 0.0),                        # log one-hot(centre): -inf / 0  -> -0.0
        np.zeros((1, 21)),
    ]).astype(np.float32)
    with np.errstate(divide="ignore"):
        s = models.support_to_scalar(torch.tensor(logits), 10).numpy()
    x = np.concatenate([
        rs.normal(0, 30, 200), rs.normal(0, 1, 100), np.arange(-12, 13, dtype=np.float64),
        [0.0, -0.0, 1e-8, -1e-8, 120.9, 121.0, 143.0, -143.0, 1e4, -1e4],
    ]).astype(np.float32).reshape(5, -1)
    sup = models.scalar_to_support(torch.tensor(x), 10).numpy()
    np.savez_compressed(os.path.join(HERE, "codec.npz"), logits=logits, scalars=s, x=x, support=sup)
    print("codec.npz")


CONFIG_NAMES = ("cartpole", "tictactoe", "connect4", "gomoku", "breakout")


def _ref_config(name, **override):
    if name == "tictactoe_fc":
        cfg = ref_loader.load("games.tictactoe").MuZeroConfig()
        cfg.network = "fullyconnected"
    else:
        cfg = ref_loader.load(f"games.{name}").MuZeroConfig()
    for k, v in override.items():
        setattr(cfg, k, v)
    return cfg


def load_shipped_cartpole_weights():
    import numpy.core.multiarray  # noqa: F401
    path = os.path.join(ref_loader.REFERENCE_ROOT, "results/cartpole/model.checkpoint")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    return ck["weights"]


def gen_net():
    """initial / recurrent inference of the reference models: shipped cartpole weights + seeded inits."""
    models = ref_loader.load("models")
    out = {}
    rs = np.random.RandomState(SEED + 4)
    specs = [("cartpole_shipped", "cartpole", 64), ("cartpole", "cartpole", 64), ("tictactoe_fc", "tictactoe_fc", 64),
             ("tictactoe", "tictactoe", 16), ("connect4", "connect4", 8), ("gomoku", "gomoku", 2), ("breakout", "breakout", 2)]
    for tag, cname, B in specs:
        cfg = _ref_config(cname)
        torch.manual_seed(0)
        net = models.MuZeroNetwork(cfg)
        if tag == "cartpole_shipped":
            net.set_weights(load_shipped_cartpole_weights())
        elif cfg.network == "resnet":
            # give batch-norm non-trivial running statistics and affine terms (eval mode uses them)
            g = torch.Generator().manual_seed(1)
            for k, v in net.state_dict().items():
                if k.endswith("running_mean"):
                    v.copy_(torch.randn(v.shape, generator=g) * 0.1)
                elif k.endswith("running_var"):
                    v.copy_(torch.rand(v.shape, generator=g) + 0.5)
                elif ".bn" in k and k.endswith(".weight"):
                    v.copy_(torch.rand(v.shape, generator=g) + 0.5)
                elif ".bn" in k and k.endswith(".bias"):
                    v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        if tag == "gomoku":
            # 22 MB of fp32 weights: seeded numpy values regenerated by the tests (tests/_weights.py)
            from _weights import seeded_tensor
            sd = net.state_dict()
            net.load_state_dict({k: torch.tensor(seeded_tensor(k, v.shape)) for k, v in sd.items()})
            out["gomoku/keys"] = np.array("\n".join(sd.keys()))
            out["gomoku/shapes"] = np.array(["x".join(str(d) for d in v.shape) for v in sd.values()])
        net.eval()
        C, H, W = cfg.observation_shape
        if cname.startswith("cartpole"):
            obs = rs.uniform(-0.05, 0.05, (B, C, H, W)).astype(np.float32) * np.array([1, 10, 1, 10], dtype=np.float32)
        elif cname == "breakout":
            obs = rs.uniform(0, 1, (B, C, H, W)).astype(np.float32)
        else:
            stones = rs.randint(-1, 2, (B, H, W))
            tp = rs.choice([-1, 1], size=(B, 1, 1))
            obs = np.stack([(stones == 1), (stones == -1), np.broadcast_to(tp, stones.shape)], axis=1).astype(np.float32)
        A = len(cfg.action_space)
        with torch.no_grad():
            v0, r0, p0, s0 = net.initial_inference(torch.tensor(obs))
            act = rs.randint(A, size=(B, 1))
            v1, r1, p1, s1 = net.recurrent_inference(s0, torch.tensor(act))
            act2 = rs.randint(A, size=(B, 1))
            v2, r2, p2, s2 = net.recurrent_inference(s1, torch.tensor(act2))
            sv0 = models.support_to_scalar(v0, cfg.support_size); sr0 = models.support_to_scalar(r0, cfg.support_size)
            sv1 = models.support_to_scalar(v1, cfg.support_size); sr1 = models.support_to_scalar(r1, cfg.support_size)
        pre = tag + "/"
        if tag != "gomoku":
            for k, v in net.get_weights().items():
                out[pre + "w/" + k] = v.numpy()
        out[pre + "obs"] = obs
        out[pre + "act"], out[pre + "act2"] = act.astype(np.int32), act2.astype(np.int32)
        for nm, t in (("v0", v0), ("r0", r0), ("p0", p0), ("s0", s0), ("v1", v1), ("r1", r1), ("p1", p1), ("s1", s1),
                      ("v2", v2), ("r2", r2), ("p2", p2), ("s2", s2), ("sv0", sv0), ("sr0", sr0), ("sv1", sv1), ("sr1", sr1)):
            out[pre + nm] = t.numpy()
        print(tag, "params", sum(p.numel() for p in net.parameters()))
    np.savez_compressed(os.path.join(HERE, "net.npz"), **out)
    print("net.npz", os.path.getsize(os.path.join(HERE, "net.npz")) // 1024, "KiB")


def gen_targets():
    self_play, replay_buffer = ref_loader.load("self_play", "replay_buffer")
    rs = np.random.RandomState(SEED + 5)
    out = {}
    orig_choice = np.random.choice
    n = 0
    try:
        for cname, players in (("cartpole", 1), ("tictactoe", 2), ("connect4", 2), ("breakout", 1)):
            cfg = _ref_config(cname)
            rb = replay_buffer.ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg)
            A = len(cfg.action_space)
            for case in range(12):
                T = int(rs.randint(1, min(cfg.max_moves, 90) + 1))
                gh = self_play.GameHistory()
                gh.action_history = [0] + rs.randint(A, size=T).tolist()
                gh.reward_history = [0] + np.round(rs.uniform(-1, 2, T), 3).tolist()
                gh.to_play_history = [int(i % players) for i in range(T + 1)]
                if players == 2 and case % 3 == 0:      # irregular turn order is legal for the formula
                    gh.to_play_history = rs.randint(2, size=T + 1).tolist()
                gh.root_values = rs.normal(0, 3, T).tolist()
                cv = rs.dirichlet([0.7] * A, size=T)
                gh.child_visits = [row.tolist() for row in cv]
                if case % 4 == 1:
                    gh.reanalysed_predicted_root_values = rs.normal(0, 3, T)
                for state_index in sorted(set([0, T - 1, T // 2, max(0, T - cfg.td_steps), int(rs.randint(T))])):
                    slot, step = 50 + n, state_index

                    rows_past = []

                    def choice2(a, size=None, replace=True, p=None):
                        a = list(a)
                        row = len(rows_past)
                        rows_past.append(row)
                        return a[rng.pad_action(SEED, slot, step, row, len(a))]

                    np.random.choice = choice2
                    tv, tr, tp, ta = rb.make_target(gh, state_index)
                    pre = f"{n}/"
                    out[pre + "cfg"] = np.array([cfg.num_unroll_steps, cfg.td_steps, cfg.discount, A, players], dtype=np.float64)
                    out[pre + "state_index"] = np.int64(state_index)
                    out[pre + "slot_step"] = np.array([slot, step], dtype=np.int64)
                    out[pre + "action_history"] = np.array(gh.action_history, dtype=np.int32)
                    out[pre + "reward_history"] = np.array(gh.reward_history, dtype=np.float64)
                    out[pre + "to_play_history"] = np.array(gh.to_play_history, dtype=np.int32)
                    out[pre + "root_values"] = np.array(gh.root_values, dtype=np.float64)
                    out[pre + "child_visits"] = cv
                    out[pre + "reanalysed"] = (np.array(gh.reanalysed_predicted_root_values, dtype=np.float64)
                                               if gh.reanalysed_predicted_root_values is not None else np.zeros(0))
                    out[pre + "target_values"] = np.array(tv, dtype=np.float64)
                    out[pre + "target_rewards"] = np.array(tr, dtype=np.float64)
                    out[pre + "target_policies"] = np.array(tp, dtype=np.float64)
                    out[pre + "actions"] = np.array(ta, dtype=np.int32)
                    n += 1
    finally:
        np.random.choice = orig_choice
    out["n"] = np.int64(n)
    np.savez_compressed(os.path.join(HERE, "targets.npz"), **out)
    print("targets.npz:", n, "cases")


def obs_row(obs, n_rows=N_ROWS):
    """Observation -> table row: crc32 of the float32 bytes (mirrored in tests/_tables.py)."""
    import zlib
    return zlib.crc32(np.ascontiguousarray(np.asarray(obs, dtype=np.float32)).tobytes()) % n_rows


class EpisodeModel(TableModel):
    """Table model whose root row is a hash of the observation: a whole game is a pure function of the tables."""

    def __init__(self, V, Rw, L):
        super().__init__(V, Rw, L)
        self.step = -1

    def initial_inference(self, observation):
        self.step += 1          # one initial inference per move (self_play.py:288-292)
        self.sim = 0
        self.depth = 0
        return self._out(obs_row(observation.cpu().numpy()[0]))


def gen_episode():
    """G7: whole games through the UNMODIFIED SelfPlay.play_game (self_play.py:110-184) - reset, observation
    stacking, search, select_action, Game.step, store_search_statistics, history appends and the loop bounds -
    with the reference's own Game classes, a table-driven network and injected randomness."""
    self_play, models = ref_loader.load("self_play", "models")
    rs = np.random.RandomState(SEED + 7)
    out = {}
    orig = (np.random.choice, np.random.dirichlet, models.support_to_scalar, self_play.MCTS.run, self_play.SelfPlay.__init__)
    # game, simulations override, max_moves override, (temperature, temperature_threshold) per episode
    plans = [("tictactoe", 25, None, [(1.0, None), (1.0, 4), (0.5, None), (0, None)]),
             ("connect4", 40, None, [(1.0, None), (1.0, 10), (0.25, None)]),
             ("gomoku", 20, 30, [(1.0, None), (0.5, 6)]),
             ("cartpole", 50, None, [(1.0, None), (0.5, 12), (0.25, None)]),
             ("cartpole", 20, 9, [(1.0, None)])]            # episode cut by max_moves (self_play.py:129-131)
    n = 0
    try:
        for gname, sims, max_moves, episodes in plans:
            gmod = ref_loader.load(f"games.{gname}")
            cfg = gmod.MuZeroConfig()
            cfg.num_simulations = sims
            if max_moves is not None:
                cfg.max_moves = max_moves
            A = len(cfg.action_space)
            V = rs.uniform(-3, 3, N_ROWS).astype(np.float32)
            Rw = (rs.uniform(-1, 1, N_ROWS) * (rs.uniform(size=N_ROWS) < 0.5)).astype(np.float32)
            L = rs.normal(0, 1.5, (N_ROWS, A)).astype(np.float32)
            P = np.stack([torch.softmax(torch.tensor([torch.tensor(L)[r][a] for a in range(A)]), dim=0).numpy()
                          for r in range(N_ROWS)]) if A <= 16 else \
                np.stack([torch.softmax(torch.tensor(L[r]), dim=0).numpy() for r in range(N_ROWS)])
            tab = f"tab{len([k for k in out if k.endswith('/V')])}"
            out[f"{tab}/V"], out[f"{tab}/Rw"], out[f"{tab}/L"], out[f"{tab}/P"] = V, Rw, L, P
            for ei, (T, T_thr) in enumerate(episodes):
                slot, seed_env = 300 + n, 11 * n + 5
                model = EpisodeModel(V, Rw, L)
                roots = []          # per move: (row, legal, torch softmax over the legal logits = Node.expand's priors)

                def choice(a, size=None, replace=True, p=None, _m=model):
                    a = list(a)
                    if p is None and not isinstance(a[0], (int, np.integer)):
                        raise AssertionError("unexpected numpy.random.choice use")
                    if p is not None or _m.sim >= cfg.num_simulations:
                        # select_action (self_play.py:240-244): inverse CDF on the injected uniform / uniform pick
                        u = rng.action_uniform(SEED, slot, _m.step)
                        if p is None:
                            return a[int(u * len(a))]
                        cdf = np.cumsum(p)
                        cdf /= cdf[-1]
                        return a[int(np.searchsorted(cdf, u, side="right"))]
                    i = rng.tie_index(SEED, slot, _m.step, _m.sim, _m.depth, len(a)) if len(a) > 1 else 0
                    _m.depth += 1
                    return a[i]

                noises = {}

                def dirichlet(alphas, _m=model, _slot=slot):
                    nz = np.random.RandomState([SEED & 0x7FFFFFFF, _slot, _m.step]).dirichlet(alphas)
                    noises[_m.step] = nz
                    return nz

                def run(self, mdl, observation, legal_actions, to_play, add_noise, override_root_with=None, _orig=orig[3]):
                    row = obs_row(np.asarray(observation, dtype=np.float32))
                    pri = torch.softmax(torch.tensor([torch.tensor(L[row:row + 1])[0][a] for a in legal_actions]), dim=0).numpy()
                    roots.append((row, list(legal_actions), pri))
                    return _orig(self, mdl, observation, legal_actions, to_play, add_noise, override_root_with)

                def init(self, checkpoint, Game, config, seed, _m=model):
                    self.config, self.game, self.model = config, Game(seed), _m

                np.random.choice, np.random.dirichlet = choice, dirichlet
                models.support_to_scalar = lambda logits, support_size: logits
                self_play.MCTS.run = run
                self_play.SelfPlay.__init__ = init
                sp = self_play.SelfPlay(None, gmod.Game, cfg, seed_env)
                gh = sp.play_game(T, T_thr, False, "self", 0)
                pre = f"{n}/"
                out[pre + "meta"] = np.array([A, len(cfg.players), sims, cfg.max_moves, cfg.discount, cfg.root_dirichlet_alpha,
                                              float(T), float(T_thr or 0), slot, seed_env], dtype=np.float64)
                out[pre + "game"] = np.array(gname)
                out[pre + "table"] = np.array(tab)
                out[pre + "actions"] = np.array(gh.action_history, dtype=np.int32)
                out[pre + "rewards"] = np.array(gh.reward_history, dtype=np.float64)
                out[pre + "to_play"] = np.array(gh.to_play_history, dtype=np.int32)
                out[pre + "observations"] = np.array([np.asarray(o, dtype=np.float32) for o in gh.observation_history])
                out[pre + "child_visits"] = np.array(gh.child_visits, dtype=np.float64)
                out[pre + "root_values"] = np.array(gh.root_values, dtype=np.float64)
                out[pre + "root_rows"] = np.array([r[0] for r in roots], dtype=np.int64)
                for mi, (_, lg, pri) in enumerate(roots):
                    out[pre + f"root_priors/{mi}"] = pri
                    out[pre + f"noise/{mi}"] = noises[mi]
                print(gname, "episode", ei, "moves", len(gh.action_history) - 1, "T", T, T_thr)
                n += 1
    finally:
        (np.random.choice, np.random.dirichlet, models.support_to_scalar, self_play.MCTS.run,
         self_play.SelfPlay.__init__) = orig
    out["n"] = np.int64(n)
    np.savez_compressed(os.path.join(HERE, "episode.npz"), **out)
    print("episode.npz:", n, "games")


def gen_replay():
    """The unmodified reference ReplayBuffer (replay_buffer.py:33-220): save_game with initial PER priorities and FIFO
    eviction, get_batch (sample_n_games / sample_position / make_target / weights / gradient scale) and
    update_priorities, as a scripted sequence of operations with injected uniforms."""
    self_play, replay_buffer = ref_loader.load("self_play", "replay_buffer")
    rs = np.random.RandomState(SEED + 9)
    out = {}
    orig_choice = np.random.choice
    cases = [("cartpole", True, 1), ("tictactoe", True, 2), ("connect4", False, 2), ("cartpole", True, 1)]
    try:
        for ci, (cname, per, players) in enumerate(cases):
            cfg = _ref_config(cname)
            cfg.PER = per
            cfg.batch_size = 24
            cfg.replay_buffer_size = 5 if ci < 3 else 400
            if ci == 3:
                cfg.PER_alpha = 1
            A = len(cfg.action_space)
            rb = replay_buffer.ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg)
            ctx = {"batch": -1, "b": -1, "row": 0}

            def choice(a, size=None, replace=True, p=None):
                if size is not None:                                     # sample_n_games (:162-177)
                    ids = list(a)
                    ctx["b"] = -1
                    u = [rng.replay_uniform(SEED, b, ctx["batch"], rng.STREAM_RGAME) for b in range(size)]
                    if p is None:
                        return np.array([ids[int(x * len(ids))] for x in u])
                    cdf = np.cumsum(np.asarray(p, dtype=np.float64)); cdf /= cdf[-1]
                    return np.array([ids[int(np.searchsorted(cdf, x, side="right"))] for x in u])
                if isinstance(a, (int, np.integer)):                     # sample_position (:179-192)
                    ctx["b"] += 1
                    ctx["row"] = 0
                    u = rng.replay_uniform(SEED, ctx["b"], ctx["batch"], rng.STREAM_RPOS)
                    if p is None:
                        return int(u * a)
                    cdf = np.cumsum(np.asarray(p, dtype=np.float64)); cdf /= cdf[-1]
                    return int(np.searchsorted(cdf, u, side="right"))
                a = list(a)                                              # make_target padding action (:291)
                k = ctx["row"]
                ctx["row"] += 1
                return a[rng.pad_action(SEED, ctx["b"], ctx["batch"], k, len(a))]

            np.random.choice = choice
            pre = f"{ci}/"
            out[pre + "cfg"] = np.array([int(per), cfg.PER_alpha, cfg.replay_buffer_size, cfg.batch_size, cfg.num_unroll_steps,
                                         cfg.td_steps, cfg.discount, A, players], dtype=np.float64)
            out[pre + "game"] = np.array(cname)
            n_games = 9 if ci < 3 else 300
            script = (["save"] * 3 + ["batch", "batch", "update", "batch"] + ["save"] * 4 + ["batch", "update", "save", "save", "batch"]
                      if ci < 3 else ["save"] * n_games + ["batch", "update", "batch"])
            gi = bi = ui = 0
            last_index = None
            ops = []
            for op in script:
                if op == "save":
                    T = int(rs.randint(1, min(cfg.max_moves, 90) + 1))
                    gh = self_play.GameHistory()
                    gh.action_history = [0] + rs.randint(A, size=T).tolist()
                    # float32-representable rewards (game rewards are small integers; the device store keeps float32)
                    gh.reward_history = [0] + np.round(rs.uniform(-1, 2, T), 3).astype(np.float32).astype(np.float64).tolist()
                    gh.to_play_history = [int(i % players) for i in range(T + 1)]
                    gh.root_values = rs.normal(0, 3, T).tolist()
                    vis = rs.multinomial(cfg.num_simulations, rs.dirichlet([0.7] * A), size=T)
                    gh.child_visits = [[int(v) / int(row.sum()) if v else 0 for v in row] for row in vis]
                    gh.observation_history = [rs.uniform(-1, 1, cfg.observation_shape).astype(np.float32) for _ in range(T + 1)]
                    g = f"{pre}game{gi}/"
                    out[g + "actions"] = np.array(gh.action_history, dtype=np.int32)
                    out[g + "rewards"] = np.array(gh.reward_history, dtype=np.float64)
                    out[g + "to_play"] = np.array(gh.to_play_history, dtype=np.int32)
                    out[g + "root_values"] = np.array(gh.root_values, dtype=np.float64)
                    out[g + "visits"] = vis.astype(np.int32)
                    out[g + "observations"] = np.array(gh.observation_history)
                    rb.save_game(gh)
                    if per:
                        out[g + "priorities"] = np.array(gh.priorities, dtype=np.float32)
                        out[g + "game_priority"] = np.float32(gh.game_priority)
                    gi += 1
                elif op == "batch":
                    ctx["batch"] = bi
                    index, (obs, act, val, rew, pol, w, gs) = rb.get_batch()
                    b = f"{pre}batch{bi}/"
                    out[b + "index"] = np.array(index, dtype=np.int64)
                    out[b + "observations"] = np.array(obs, dtype=np.float32)
                    out[b + "actions"] = np.array(act, dtype=np.int32)
                    out[b + "values"] = np.array(val, dtype=np.float64)
                    out[b + "rewards"] = np.array(rew, dtype=np.float64)
                    out[b + "policies"] = np.array(pol, dtype=np.float64)
                    out[b + "gradient_scale"] = np.array(gs, dtype=np.int32)
                    if per:
                        assert w.dtype == np.float32
                        out[b + "weights"] = w
                    out[b + "state"] = np.array([rb.total_samples, rb.num_played_games, len(rb.buffer)], dtype=np.int64)
                    last_index = index
                    bi += 1
                else:                                                    # update_priorities with trainer-like values
                    pr = np.abs(rs.normal(0, 2, (len(last_index), cfg.num_unroll_steps + 1))).astype(np.float32) ** cfg.PER_alpha
                    pr = pr.astype(np.float32)
                    if per:
                        rb.update_priorities(pr, last_index)
                    u = f"{pre}update{ui}/"
                    out[u + "priorities"] = pr
                    out[u + "index"] = np.array(last_index, dtype=np.int64)
                    if per:
                        for gid, gh in rb.buffer.items():
                            out[u + f"after/{gid}"] = np.array(gh.priorities, dtype=np.float32)
                            out[u + f"after_game/{gid}"] = np.float32(gh.game_priority)
                    ui += 1
                ops.append(op)
            out[pre + "script"] = np.array(" ".join(ops))
            print(cname, "PER" if per else "uniform", "games", gi, "batches", bi, "updates", ui)
    finally:
        np.random.choice = orig_choice
    out["n"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "replay.npz"), **out)
    print("replay.npz", os.path.getsize(os.path.join(HERE, "replay.npz")) // 1024, "KiB")


def gen_trainer():
    """Two consecutive steps of the UNMODIFIED reference Trainer.update_weights (trainer.py:124-255) on CPU: losses,
    priorities and the updated weights, for a batch drawn by the reference ReplayBuffer."""
    self_play, replay_buffer, trainer, models = ref_loader.load("self_play", "replay_buffer", "trainer", "models")
    rs = np.random.RandomState(SEED + 11)
    out = {}
    orig_choice = np.random.choice
    cases = [("cartpole", {}), ("tictactoe", {}), ("connect4", {"optimizer": "SGD", "blocks": 1, "channels": 16, "lr_init": 0.01}),
             ("breakout", {"batch_size": 4, "num_unroll_steps": 2})]
    try:
        for ci, (cname, over) in enumerate(cases):
            cfg = _ref_config(cname, **over)
            cfg.batch_size = over.get("batch_size", 12)
            cfg.train_on_gpu = False
            A, players = len(cfg.action_space), len(cfg.players)
            rb = replay_buffer.ReplayBuffer({"num_played_games": 0, "num_played_steps": 0}, {}, cfg)
            np.random.choice = orig_choice
            for _ in range(4):
                T = int(rs.randint(3, min(cfg.max_moves, 30) + 1))
                gh = self_play.GameHistory()
                gh.action_history = [0] + rs.randint(A, size=T).tolist()
                gh.reward_history = [0] + np.round(rs.uniform(-1, 2, T), 3).astype(np.float32).astype(np.float64).tolist()
                gh.to_play_history = [int(i % players) for i in range(T + 1)]
                gh.root_values = rs.normal(0, 3, T).tolist()
                vis = rs.multinomial(cfg.num_simulations, rs.dirichlet([0.7] * A), size=T)
                gh.child_visits = [[int(v) / int(row.sum()) if v else 0 for v in row] for row in vis]
                gh.observation_history = [rs.uniform(0, 1, cfg.observation_shape).astype(np.float32) for _ in range(T + 1)]
                rb.save_game(gh)
            np.random.seed(ci)
            index, batch = rb.get_batch()
            torch.manual_seed(100 + ci)
            model = models.MuZeroNetwork(cfg)
            w0 = {k: v.clone() for k, v in model.state_dict().items()}
            tr = trainer.Trainer({"weights": w0, "training_step": 0, "optimizer_state": None}, cfg)
            pre = f"{ci}/"
            out[pre + "game"] = np.array(cname)
            out[pre + "over"] = np.array(repr(over))
            for k, v in w0.items():
                out[pre + "w0/" + k] = v.numpy()
            names = ["observation", "action", "value", "reward", "policy", "weight", "gradient_scale"]
            for nm, arr in zip(names, batch):
                if arr is not None:
                    out[pre + "batch/" + nm] = np.array(arr, dtype=np.float32 if nm != "action" else np.int64)
            for step in range(2):
                tr.update_lr()
                pr, total, vl, rl, pl = tr.update_weights(batch)
                out[pre + f"step{step}/losses"] = np.array([total, vl, rl, pl, tr.optimizer.param_groups[0]["lr"]], dtype=np.float64)
                out[pre + f"step{step}/priorities"] = np.array(pr, dtype=np.float32)
            for k, v in tr.model.state_dict().items():
                out[pre + "w2/" + k] = v.detach().numpy()
            print(cname, over, "losses", out[pre + "step0/losses"][:4], out[pre + "step1/losses"][:4])
    finally:
        np.random.choice = orig_choice
    out["n"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "trainer.npz"), **out)
    print("trainer.npz", os.path.getsize(os.path.join(HERE, "trainer.npz")) // 1024, "KiB")


def gen_stacked():
    """GameHistory.get_stacked_observations (self_play.py:514-548) of the UNMODIFIED reference: every index (incl. the
    -1 play_game uses and indices past the end, which wrap) for S in {0, 1, 2, 4} on three observation shapes."""
    self_play = ref_loader.load("self_play")
    rs = np.random.RandomState(SEED + 11)
    out = {}
    shapes = [((1, 1, 4), np.float32, 2), ((3, 3, 3), np.int32, 9), ((3, 6, 7), np.float64, 7)]
    for si, (shape, dt, A) in enumerate(shapes):
        n = 7
        gh = self_play.GameHistory()
        gh.observation_history = [(rs.randint(-1, 2, size=shape) if dt != np.float32 else rs.randn(*shape)).astype(dt)
                                  for _ in range(n)]
        gh.action_history = [0] + rs.randint(A, size=n - 1).tolist()
        out[f"{si}/observations"] = np.stack(gh.observation_history)
        out[f"{si}/actions"] = np.array(gh.action_history, dtype=np.int64)
        for S in (0, 1, 2, 4):
            for index in list(range(-1, n + 2)):
                out[f"{si}/S{S}/i{index}"] = np.asarray(gh.get_stacked_observations(index, S))
    out["n"] = np.array(len(shapes))
    np.savez_compressed(os.path.join(HERE, "stacked.npz"), **out)
    print("stacked.npz:", len(out), "arrays")


FAMILIES = {"stacked": gen_stacked, "trainer": gen_trainer, "tree": gen_tree, "action": gen_action, "env": gen_env, "codec": gen_codec, "net": gen_net,
            "targets": gen_targets, "episode": gen_episode, "replay": gen_replay}

if __name__ == "__main__":
    todo = sys.argv[1:] or list(FAMILIES)
    for fam in todo:
        FAMILIES[fam]()
