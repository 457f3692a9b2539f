"""CPU restatement of the reference MCTS (test infrastructure; see oracle/__init__.py).

Follows /root/reference/self_play.py:
  MCTS.run :261-362, select_child :364-379, ucb_score :381-405, backpropagate :407-431,
  Node.expand :452-466, Node.add_exploration_noise :468-477, MinMaxStats :551-568,
  SelfPlay.select_action :223-246, GameHistory.store_search_statistics :497-512.

The reference builds a graph of Python `Node` objects; this restatement keeps one flat
edge table per search (the layout the CUDA tree store uses) but performs the SAME
float64 operations in the SAME order, so every scalar it produces is bit-identical to
the reference's when both consume the same network outputs, Dirichlet noise and
tie-break draws (pinned by tests/golden/tree_*.npz, generated from the reference).

Node k>0 is the node created by simulation k (slot = simulation index); node 0 is the root.
Edge (n, a) holds what the reference stores on the child Node reached from n by a:
prior, visit_count, value_sum, reward, and the child's slot (-1 = not expanded).
"""
import math


class SearchResult:
    """Everything a caller of MCTS.run can observe, as flat lists."""

    def __init__(self):
        self.root_actions = []      # root children in insertion order (= legal_actions order)
        self.visits = []            # per root child
        self.value_sums = []        # per root child (f64)
        self.rewards = []           # per root child
        self.priors = []            # per root child, after noise (f64)
        self.root_visit = 0
        self.root_value_sum = 0.0
        self.max_tree_depth = 0
        self.minimum = float("inf")
        self.maximum = -float("inf")
        self.tree = None            # the full Tree for deep comparisons

    def root_value(self):
        # Node.value self_play.py:446-449
        return self.root_value_sum / self.root_visit if self.root_visit else 0


class Tree:
    def __init__(self, n_actions, n_sims):
        n = n_sims + 1
        self.A = n_actions
        self.prior = [[0.0] * n_actions for _ in range(n)]
        self.visit = [[0] * n_actions for _ in range(n)]
        self.value_sum = [[0.0] * n_actions for _ in range(n)]
        self.reward = [[0.0] * n_actions for _ in range(n)]
        self.child = [[-1] * n_actions for _ in range(n)]
        self.exists = [[False] * n_actions for _ in range(n)]   # root: legal only; interior: all
        self.hidden = [None] * n
        self.depth = [0] * n
        self.root_visit = 0
        self.root_value_sum = 0.0
        self.root_reward = 0.0


def default_tie(n_ties, sim, depth):
    """Deterministic stand-in when no tie rule is injected: first maximum."""
    return 0


def search(
    recurrent,
    root_value_reward_priors_hidden,
    legal_actions,
    to_play,
    *,
    n_actions,
    n_players,
    num_simulations,
    discount,
    pb_c_base,
    pb_c_init,
    noise=None,
    exploration_fraction=0.25,
    tie=default_tie,
):
    """One MCTS.run (self_play.py:261-362).

    recurrent(hidden, action) -> (value, reward, priors[n_actions], hidden) with value/reward
    Python floats (the f32 `.item()` the reference takes, :344-345) and priors the f32 softmax
    values widened to float (:461-463).
    root_value_reward_priors_hidden = (root_predicted_value, reward, priors over `legal_actions`,
    hidden) from the initial inference, priors being softmax over the LEGAL logits only (:303-309).
    noise: f64 Dirichlet sample over legal_actions or None (:310-314, :468-477).
    tie(n_ties, sim, depth) -> index into the tied set, in child order (:372-378 injected).
    """
    assert legal_actions, f"Legal actions should not be an empty array. Got {legal_actions}."
    assert set(legal_actions).issubset(set(range(n_actions))), "Legal actions should be a subset of the action space."
    if n_players not in (1, 2):
        raise NotImplementedError("More than two player mode not implemented.")

    t = Tree(n_actions, num_simulations)
    _, root_reward, root_priors, root_hidden = root_value_reward_priors_hidden
    t.root_reward = root_reward
    t.hidden[0] = root_hidden
    for a, p in zip(legal_actions, root_priors):
        t.prior[0][a] = p
        t.exists[0][a] = True
    if noise is not None:
        frac = exploration_fraction
        for a, n in zip(legal_actions, noise):
            t.prior[0][a] = t.prior[0][a] * (1 - frac) + n * frac

    minimum, maximum = float("inf"), -float("inf")
    max_tree_depth = 0
    root_order = list(legal_actions)
    all_actions = list(range(n_actions))

    for sim in range(num_simulations):
        node = 0
        parent_visits = t.root_visit
        path = []                                  # (node, action) edges walked
        depth = 0
        while True:
            order = root_order if node == 0 else all_actions
            # ucb_score for each child, self_play.py:381-405
            scores = []
            pb_c0 = math.log((parent_visits + pb_c_base + 1) / pb_c_base) + pb_c_init
            sq = math.sqrt(parent_visits)
            for a in order:
                n = t.visit[node][a]
                pb_c = pb_c0 * (sq / (n + 1))
                s = pb_c * t.prior[node][a]
                if n > 0:
                    v = t.value_sum[node][a] / n
                    q = t.reward[node][a] + discount * (v if n_players == 1 else -v)
                    if maximum > minimum:
                        q = (q - minimum) / (maximum - minimum)
                    s = s + q
                else:
                    s = s + 0
                scores.append(s)
            best = max(scores)
            tied = [a for a, s in zip(order, scores) if s == best]
            a = tied[tie(len(tied), sim, depth)] if len(tied) > 1 else tied[0]
            path.append((node, a))
            depth += 1
            nxt = t.child[node][a]
            if nxt < 0:
                break
            parent_visits = t.visit[node][a]
            node = nxt

        # leaf expansion, self_play.py:339-352
        parent, action = path[-1]
        value, reward, priors, hidden = recurrent(t.hidden[parent], action)
        new = sim + 1
        t.child[parent][action] = new
        t.reward[parent][action] = reward
        t.hidden[new] = hidden
        t.depth[new] = depth
        for b in all_actions:
            t.prior[new][b] = priors[b]
            t.exists[new][b] = True

        # backpropagate, self_play.py:407-431 (leaf first, root last)
        for k in range(len(path) - 1, -1, -1):
            pn, pa = path[k]
            # the node reached by edge (pn, pa) is at tree depth k+1; the leaf at depth len(path)
            same = ((len(path) - (k + 1)) % 2 == 0)
            if n_players == 1:
                t.value_sum[pn][pa] += value
                t.visit[pn][pa] += 1
                q = t.reward[pn][pa] + discount * (t.value_sum[pn][pa] / t.visit[pn][pa])
                maximum = max(maximum, q); minimum = min(minimum, q)
                value = t.reward[pn][pa] + discount * value
            else:
                t.value_sum[pn][pa] += value if same else -value
                t.visit[pn][pa] += 1
                q = t.reward[pn][pa] + discount * -(t.value_sum[pn][pa] / t.visit[pn][pa])
                maximum = max(maximum, q); minimum = min(minimum, q)
                value = (-t.reward[pn][pa] if same else t.reward[pn][pa]) + discount * value
        # ... and the root itself
        same = (len(path) % 2 == 0)
        if n_players == 1:
            t.root_value_sum += value
            t.root_visit += 1
            q = t.root_reward + discount * (t.root_value_sum / t.root_visit)
        else:
            t.root_value_sum += value if same else -value
            t.root_visit += 1
            q = t.root_reward + discount * -(t.root_value_sum / t.root_visit)
        maximum = max(maximum, q); minimum = min(minimum, q)
        max_tree_depth = max(max_tree_depth, depth)

    r = SearchResult()
    r.root_actions = root_order
    r.visits = [t.visit[0][a] for a in root_order]
    r.value_sums = [t.value_sum[0][a] for a in root_order]
    r.rewards = [t.reward[0][a] for a in root_order]
    r.priors = [t.prior[0][a] for a in root_order]
    r.root_visit = t.root_visit
    r.root_value_sum = t.root_value_sum
    r.max_tree_depth = max_tree_depth
    r.minimum, r.maximum = minimum, maximum
    r.tree = t
    return r


def select_action(actions, visit_counts, temperature, u):
    """SelfPlay.select_action (self_play.py:223-246) with the random draw injected as u in [0,1).

    numpy.random.choice(actions, p=p) is inverse-CDF sampling on one uniform
    (cdf = p.cumsum(); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')), SURVEY.md §8 a9.
    """
    n = len(actions)
    if temperature == 0:
        best = 0
        for i in range(1, n):
            if visit_counts[i] > visit_counts[best]:
                best = i
        return actions[best]
    if temperature == float("inf"):
        return actions[int(u * n)]
    d = [float(v) ** (1 / temperature) for v in visit_counts]
    tot = 0
    for x in d:            # Python's left-to-right sum()
        tot = tot + x
    p = [x / tot for x in d]
    c = 0.0
    cdf = []
    for x in p:
        c = c + x
        cdf.append(c)
    last = cdf[-1]
    cdf = [x / last for x in cdf]
    for i, x in enumerate(cdf):
        if x > u:
            return actions[i]
    return actions[-1]


def search_statistics(root_actions, visits, n_actions):
    """GameHistory.store_search_statistics (self_play.py:497-512): visit-count policy."""
    total = sum(visits)
    by_action = dict(zip(root_actions, visits))
    return [by_action[a] / total if a in by_action else 0 for a in range(n_actions)]


def softmax_f32(logits):
    """float32 softmax the way Node.expand applies it (self_play.py:459-461) - numpy f32.

    torch's CPU softmax is not bit-reproducible by numpy; tree parity therefore injects the
    priors (golden fixtures carry torch's), this helper serves end-to-end tolerance checks."""
    import numpy as np
    x = np.asarray(logits, dtype=np.float32)
    e = np.exp(x - x.max())
    return (e / e.sum(dtype=np.float32)).astype(np.float32)
