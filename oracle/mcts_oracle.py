"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's sequential PUCT search
(woker/solo_play.py::HivePlayer with none_queue=False) on top of the C environment oracle.

Follows, line by line:
  solo_play.py:110-151  action            -> MctsOracle.action
  solo_play.py:153-165  search_moves      -> MctsOracle.action (loop)
  solo_play.py:167-247  search_my_move    -> MctsOracle._search
  solo_play.py:294-335  select_action_q_and_u -> MctsOracle._select
  solo_play.py:337-374  apply_temperature / calc_policy

The arithmetic is done with the same NumPy scalar types as the reference (float32 priors, Python
float W/Q, np.float64 sqrt) so that NumPy's promotion rules reproduce the reference's precision
chain bit for bit.  Pinned against the real HivePlayer by
tests/golden/mcts_cases.npz (oracle/gen_golden_mcts.py) and, in the build container, by
tests/test_oracle_vs_reference.py.
"""
import zlib

import numpy as np

ACTION_SPACE = 1584
MAX_GAME_LENGTH = 55
c_puct = 0.7
dirichlet_alpha = 0.3
noise_eps = 0.25
virtual_loss = 1
tau_decay_rate = 0.01


def hash_net(planes_hwc):
    """Deterministic stand-in for the network: (12,12,56) planes -> (p float32[1584], v float).
    Peaky priors so that searches go several plies deep."""
    key = zlib.crc32(np.ascontiguousarray(planes_hwc, dtype=np.float64).tobytes())
    rng = np.random.RandomState(key)
    p = (rng.rand(ACTION_SPACE) ** 6).astype(np.float32)
    p /= p.sum()
    v = float(rng.rand() * 2.0 - 1.0)
    return p, v


def _splitmix64_np(x):
    """splitmix64 over uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def device_hash_net(planes_hwc, salt=0):
    """NumPy twin of the library's on-device stand-in network (mcts_hash_eval_dev, csrc/hive_mcts.cu): a pure
    function of the bf16 CHW planes the device holds and a salt.  (12,12,56) planes -> (p float32[1584], v float)."""
    chw = np.ascontiguousarray(np.asarray(planes_hwc, dtype=np.float32).transpose(2, 0, 1)).reshape(-1)
    bf16 = (chw.view(np.uint32) >> 16).astype(np.uint64)                 # exact: the planes hold small integers
    j = np.arange(bf16.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _splitmix64_np((j << np.uint64(16)) | bf16).sum(dtype=np.uint64)
        key = _splitmix64_np(np.uint64(salt) ^ h)
        a = np.arange(1, ACTION_SPACE + 1, dtype=np.uint64)
        u = _splitmix64_np(key ^ (a * np.uint64(0x9E3779B97F4A7C15))) >> np.uint64(40)
    x = u.astype(np.float32) * np.float32(1.0 / 16777216.0)
    x2 = x * x
    x4 = x2 * x2
    p = (x4 * x2).astype(np.float32)
    with np.errstate(over="ignore"):
        k = _splitmix64_np(key ^ np.uint64(0x5bf03635)) >> np.uint64(11)
    v = float(np.float64(k) * (1.0 / 4503599627370496.0) - 1.0)
    return p, v


class _Edge:
    __slots__ = ("n", "w", "q", "p")

    def __init__(self):
        self.n, self.w, self.q, self.p = 0, 0, 0, 0


class _Node:
    def __init__(self):
        self.a = {}          # insertion ordered like defaultdict(ActionStats)
        self.sum_n = 0
        self.p = None

    def edge(self, action):
        if action not in self.a:
            self.a[action] = _Edge()
        return self.a[action]


class MctsOracle:
    def __init__(self, net, sims):
        self.net = net                       # f(planes_hwc float64) -> (p float32[1584], v float)
        self.sims = sims
        self.tree = {}
        self.noise_log = []                  # the Dirichlet rows drawn at the root, in order

    def action(self, env):
        """env: oracle.hive_oracle.OracleEnv.  Returns (action, policy float64[1584], sum_n)."""
        self.tree = {}
        self.noise_log = []
        for _ in range(self.sims):
            self._search(env.clone(), True)
        policy, sum_all = self._calc_policy(env)
        p = self._apply_temperature(policy, int(env.turn + 1) / 2)
        my_action = int(np.random.choice(range(ACTION_SPACE), p=p))
        return my_action, policy, sum_all

    # solo_play.py:167-247
    def _search(self, env, is_root):
        if env.game_is_over():
            player = 0 if env.turn % 2 == 1 else 1
            w = env.winner
            if player == 0:
                if w == 1:
                    return 1
                elif w == 2:
                    return -1
            else:
                if w == 1:
                    return -1
                elif w == 2:
                    return 1
            return 5
        elif env.turn >= MAX_GAME_LENGTH:
            return 5
        state = env.state_key
        if state not in self.tree:
            leaf_p, leaf_v = self.net(env.encode_board())
            node = _Node()
            node.p = leaf_p
            self.tree[state] = node
            return leaf_v
        action_t = self._select(env, is_root)
        node = self.tree[state]
        st = node.edge(action_t)
        node.sum_n += virtual_loss
        st.n += virtual_loss
        st.w += -virtual_loss
        st.q = st.w / st.n
        env.move(action_t)
        leaf_v = self._search(env, False)
        reach_max = False
        if leaf_v == 5:
            leaf_v = 1
            reach_max = True
        leaf_v = -leaf_v
        node.sum_n += -virtual_loss + 1
        st.n += -virtual_loss + 1
        st.w += virtual_loss + leaf_v
        st.q = st.w / st.n
        if reach_max:
            leaf_v = 5
        return leaf_v

    # solo_play.py:294-335
    def _select(self, env, is_root):
        actions = env.actions().tolist()
        if len(actions) == 0:
            return -1
        node = self.tree[env.state_key]
        if node.p is not None:
            tot_p = 1e-8
            for mov in actions:
                mov_p = node.p[mov]
                node.edge(mov).p = mov_p
                tot_p += mov_p
            for a_s in node.a.values():
                a_s.p /= tot_p
            node.p = None
        xx_ = np.sqrt(node.sum_n + 1)
        e = noise_eps
        best_s, best_a = -999, None
        if is_root:
            noise = np.random.dirichlet([dirichlet_alpha] * len(node.a))
            self.noise_log.append(noise)
        i = 0
        for action, a_s in node.a.items():
            p_ = a_s.p
            if is_root:
                p_ = (1 - e) * p_ + e * noise[i]
                i += 1
            b = a_s.q + c_puct * p_ * xx_ / (1 + a_s.n)
            if b > best_s:
                best_s, best_a = b, action
        return best_a

    # solo_play.py:351-374
    def _calc_policy(self, env):
        node = self.tree[env.state_key]
        policy = np.zeros(ACTION_SPACE)
        policy_t = np.zeros(ACTION_SPACE)
        w = []
        for action, a_s in node.a.items():
            policy[action] = a_s.n
            policy_t[action] = a_s.p
            w.append(a_s.w)
        sum_all = np.sum(policy)
        policy /= np.sum(policy)
        if np.max(w) < 0:
            policy = policy_t
        return policy, sum_all

    # solo_play.py:337-349
    @staticmethod
    def _apply_temperature(policy, turn):
        tau = np.power(tau_decay_rate, turn)
        if tau < 0.1:
            tau = 0
        if tau == 0:
            action = np.argmax(policy)
            ret = np.zeros(ACTION_SPACE)
            ret[action] = 1.0
            return ret
        ret = np.power(policy, 1 / tau)
        ret /= np.sum(ret)
        return ret

    def root_stats(self, env):
        """(actions int32[E], n int32[E], w float64[E], q float64[E], p float32[E], sum_n, n_nodes)"""
        node = self.tree[env.state_key]
        acts = np.array(list(node.a.keys()), dtype=np.int32)
        n = np.array([e.n for e in node.a.values()], dtype=np.int32)
        w = np.array([float(e.w) for e in node.a.values()], dtype=np.float64)
        q = np.array([float(e.q) for e in node.a.values()], dtype=np.float64)
        p = np.array([np.float32(e.p) for e in node.a.values()], dtype=np.float32)
        return acts, n, w, q, p, int(node.sum_n), len(self.tree)
