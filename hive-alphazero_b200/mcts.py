"""Host-side mirror of the reference's search interface (woker/solo_play.py::HivePlayer) on top of
the batched PUCT kernels (include/hive_b200.h, mcts_*).

* ``MctsBatch``  -- one tree per game of a ``HiveBatch``; searches advance in waves
                    (descend -> leaf evaluation -> expand/backup).
* ``HivePlayer`` -- the reference's class: ``action(env) -> (move, [pi, sum_n])`` with the same
                    hyper-parameters, the same use of ``np.random`` (Dirichlet rows for the root,
                    the final ``np.random.choice``), driven by a ``GamePlay``.
"""
import ctypes

import numpy as np

from . import config as C
from ._capi import check, lib
from .env import GamePlay, HiveBatch, _bf16_to_f32


class _LeafEnv:
    """What expand_and_evaluate (solo_play.py:260-278) needs from an env: encode_board()."""

    def __init__(self, planes_hwc):
        self._p = planes_hwc

    def encode_board(self, player="N"):
        return self._p


class MctsBatch:
    def __init__(self, batch, sims, edges_per_sim=0):
        if not isinstance(batch, HiveBatch):
            raise TypeError("MctsBatch needs a HiveBatch")
        self.batch, self.n, self.sims = batch, batch.n, int(sims)
        h = ctypes.c_void_p()
        check(lib().mcts_create(batch._h, self.sims, int(edges_per_sim), ctypes.byref(h)), "mcts_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().mcts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, sims=None, max_turn=C.MAX_GAME_LENGTH, noise_seed=0x5EED):
        sims = self.sims if sims is None else int(sims)
        check(lib().mcts_set_params(self._h, sims, int(max_turn), int(noise_seed)), "mcts_set_params")
        self.sims = sims

    def set_root_noise(self, noise):
        """noise[n][rows][cols] float64 (one Dirichlet row per root visit) or None = device sampling."""
        if noise is None:
            check(lib().mcts_set_root_noise_host(self._h, None, 0, 0), "mcts_set_root_noise_host")
            return
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        if noise.ndim != 3 or noise.shape[0] != self.n:
            raise ValueError("noise must have shape (n_trees, rows, cols)")
        check(lib().mcts_set_root_noise_host(self._h, noise.ctypes.data, noise.shape[1], noise.shape[2]),
              "mcts_set_root_noise_host")

    # ---- wave protocol
    def begin(self, tree_mask=None):
        m = None if tree_mask is None else np.ascontiguousarray(tree_mask, dtype=np.uint8)
        check(lib().mcts_begin(self._h, None if m is None else m.ctypes.data), "mcts_begin")

    def descend(self):
        pending = ctypes.c_int(0)
        check(lib().mcts_descend(self._h, ctypes.byref(pending)), "mcts_descend")
        return pending.value

    def expand(self):
        check(lib().mcts_expand(self._h), "mcts_expand")

    def leaf_planes_host(self):
        planes = np.empty((self.n, C.STATE_FEATURES * 144), dtype=np.uint16)
        mask = np.empty(self.n, dtype=np.uint8)
        check(lib().mcts_leaf_planes_host(self._h, planes.ctypes.data, mask.ctypes.data), "mcts_leaf_planes_host")
        return planes, mask

    def set_leaf_eval_host(self, policy, value):
        p = np.ascontiguousarray(policy, dtype=np.float32)
        v = np.ascontiguousarray(value, dtype=np.float64)
        check(lib().mcts_set_leaf_eval_host(self._h, p.ctypes.data, v.ctypes.data), "mcts_set_leaf_eval_host")

    @property
    def dev_leaf_planes(self): return lib().mcts_dev_leaf_planes(self._h)
    @property
    def dev_leaf_policy(self): return lib().mcts_dev_leaf_policy(self._h)
    @property
    def dev_leaf_value(self): return lib().mcts_dev_leaf_value(self._h)
    @property
    def dev_pending_mask(self): return lib().mcts_dev_pending_mask(self._h)
    @property
    def launches(self): return lib().mcts_launch_count(self._h)
    @property
    def stream_ptr(self): return lib().mcts_stream(self._h)

    def errors(self):
        """OR of (1 << code) over the trees of the running search: 2 node arena, 4 edge arena, 8 depth."""
        flags = ctypes.c_uint32(0)
        check(lib().mcts_error_host(self._h, ctypes.byref(flags)), "mcts_error_host")
        return flags.value

    def tree_stats(self, max_trees=256):
        """dict(edges_per_node, select_depth, nodes_per_tree, sims_per_tree) of the last search (first max_trees trees)."""
        out = (ctypes.c_double * 4)()
        check(lib().mcts_tree_stats_host(self._h, int(max_trees), out), "mcts_tree_stats_host")
        return dict(edges_per_node=out[0], select_depth=out[1], nodes_per_tree=out[2], sims_per_tree=out[3])

    def search_host(self, evaluate, tree_mask=None):
        """Full search with a host evaluator ``evaluate(leaf_env) -> (p float32[1584], v float)`` called
        once per new position (expand_and_evaluate, solo_play.py:260-278). Returns the number of waves."""
        self.begin(tree_mask)
        p = np.zeros((self.n, C.ACTION_SPACE), dtype=np.float32)
        v = np.zeros(self.n, dtype=np.float64)
        waves = 0
        while self.descend() > 0:
            planes, mask = self.leaf_planes_host()
            for t in np.nonzero(mask)[0]:
                hwc = _bf16_to_f32(planes[t]).reshape(C.STATE_FEATURES, 12, 12).transpose(1, 2, 0).astype(np.float64)
                pt, vt = evaluate(_LeafEnv(hwc))
                p[t] = np.asarray(pt, dtype=np.float32).reshape(-1)
                v[t] = float(np.asarray(vt).reshape(-1)[0])
            self.set_leaf_eval_host(p, v)
            self.expand()
            waves += 1
        return waves

    def search_device(self, evaluate_device, tree_mask=None, graph=None):
        """Full search with a device evaluator ``evaluate_device(planes_ptr, policy_ptr, value_ptr,
        mask_ptr, n)`` (all device pointers) called once per wave.  Every wave finishes at least one
        simulation of every unfinished tree, so `sims` waves always suffice: the loop runs without
        reading anything back and checks once at the end.

        graph: a ``WaveGraph`` (see below) replays one captured wave instead of re-issuing its ~60
        launches from Python every time -- worthwhile when a wave is short (small batches)."""
        self.begin(tree_mask)
        waves = 0
        if graph is not None:
            graph.ensure(self, evaluate_device, tree_mask is not None)
            for _ in range(self.sims):
                graph.replay()
            waves += self.sims
        else:
            for _ in range(self.sims):
                check(lib().mcts_descend(self._h, None), "mcts_descend")
                evaluate_device(self.dev_leaf_planes, self.dev_leaf_policy, self.dev_leaf_value, self.dev_pending_mask, self.n)
                self.expand()
                waves += 1
        while self.descend() > 0:                            # normally returns 0 at once
            evaluate_device(self.dev_leaf_planes, self.dev_leaf_policy, self.dev_leaf_value, self.dev_pending_mask, self.n)
            self.expand()
            waves += 1
        return waves

    def actions(self):
        """Only the chosen moves (int32[n]) -- skips the 12.7 KB/tree policy read-back.
        Raises SearchError if any tree of the search was cut short (arena / depth)."""
        action = np.empty(self.n, dtype=np.int32)
        check(lib().mcts_policy_host(self._h, None, action.ctypes.data, None), "mcts_policy_host")
        return action

    def policy(self, out=None):
        """(pi float64[n,1584], action int32[n], sum_n int32[n]) -- calc_policy + apply_temperature.
        Raises SearchError if any tree of the search was cut short (arena / depth).
        `out`: (pi, action, sum_n) arrays to fill instead of fresh ones (page-locked ones make the download several times faster)."""
        if out is not None:
            pi, action, sum_n = out
            assert pi.shape == (self.n, C.ACTION_SPACE) and pi.dtype == np.float64 and pi.flags.c_contiguous
            assert action.dtype == np.int32 and sum_n.dtype == np.int32 and len(action) == self.n and len(sum_n) == self.n
        else:
            pi = np.empty((self.n, C.ACTION_SPACE), dtype=np.float64)
            action = np.empty(self.n, dtype=np.int32)
            sum_n = np.empty(self.n, dtype=np.int32)
        check(lib().mcts_policy_host(self._h, pi.ctypes.data, action.ctypes.data, sum_n.ctypes.data), "mcts_policy_host")
        return pi, action, sum_n

    def root_stats(self, t, max_edges=256):
        a = np.zeros(max_edges, dtype=np.int32); n = np.zeros(max_edges, dtype=np.int32)
        w = np.zeros(max_edges, dtype=np.float64); q = np.zeros(max_edges, dtype=np.float64)
        p = np.zeros(max_edges, dtype=np.float32); info = np.zeros(6, dtype=np.int32)
        check(lib().mcts_root_stats_host(self._h, int(t), max_edges, a.ctypes.data, n.ctypes.data, w.ctypes.data,
                                         q.ctypes.data, p.ctypes.data, info.ctypes.data), "mcts_root_stats_host")
        k = int(info[0])
        return dict(action=a[:k], n=n[:k], w=w[:k], q=q[:k], p=p[:k], sum_n=int(info[1]), n_nodes=int(info[2]),
                    sims_done=int(info[3]), error=int(info[4]), root_selects=int(info[5]))


class HashEvaluator:
    """evaluate_device callback backed by the library's deterministic stand-in network (mcts_hash_eval_dev):
    policy / value are a pure hash of the leaf's planes and `salt`.  For parity tests of the device leaf path."""

    def __init__(self, salt, stream_ptr):
        self.salt, self.stream_ptr, self.calls = int(salt), stream_ptr, 0

    def __call__(self, planes_ptr, policy_ptr, value_ptr, mask_ptr, n):
        check(lib().mcts_hash_eval_dev(planes_ptr, policy_ptr, value_ptr, mask_ptr, int(n), self.salt, self.stream_ptr),
              "mcts_hash_eval_dev")
        self.calls += 1


class WaveGraph:
    """One search wave (descend -> leaf evaluation by the environment kernels -> network -> expand) captured
    once as a CUDA graph on the stream shared by the library handles and torch, then replayed.  All buffers of
    a wave are fixed arenas, so the captured launches stay valid for every wave of every search of this tree
    batch with this evaluator."""

    def __init__(self, torch_stream):
        import torch
        self._torch, self.stream = torch, torch_stream
        self.g, self.key = None, None

    def ensure(self, mcts, evaluate_device, masked):
        key = (id(mcts), id(evaluate_device), bool(masked), mcts.sims)
        if self.g is not None and self.key == key:
            return
        torch = self._torch
        # one eager wave first: lets torch pick its algorithms / allocate outside the capture
        check(lib().mcts_descend(mcts._h, None), "mcts_descend")
        evaluate_device(mcts.dev_leaf_planes, mcts.dev_leaf_policy, mcts.dev_leaf_value, mcts.dev_pending_mask, mcts.n)
        mcts.expand()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream, capture_error_mode="relaxed"):
            check(lib().mcts_descend(mcts._h, None), "mcts_descend")
            evaluate_device(mcts.dev_leaf_planes, mcts.dev_leaf_policy, mcts.dev_leaf_value, mcts.dev_pending_mask, mcts.n)
            mcts.expand()
        self.g, self.key = g, key

    def replay(self):
        self.g.replay()


class HivePlayer:
    """Drop-in for woker/solo_play.py::HivePlayer (sequential search semantics).

    Set ``net`` to a callable ``planes(B,56,12,12 float32 ndarray) -> (p (B,1584), v (B,1))`` or
    override ``expand_and_evaluate_with_net(env) -> (p, v)`` like the reference's tests do.
    """

    def __init__(self, pipes=None, reward=False):
        self.moves = []
        self.pipe_pool = pipes
        self.none_queue = True
        self.net = None
        self.simulation_num_per_move = C.simulation_num_per_move
        self.reward = reward
        self.main_key_state = None
        self.max_depth = None
        self._mcts = None
        self._mcts_key = None

    def reset(self):
        pass                                   # the device tree is rebuilt by every action()

    def expand_and_evaluate_with_net(self, env):           # solo_play.py:249-258
        if self.net is None:
            raise RuntimeError("HivePlayer.net is not set")
        planes = np.asarray(env.encode_board()).transpose(2, 0, 1)[None].astype(np.float32)
        p, v = self.net(planes)
        return np.asarray(p, dtype=np.float32).reshape(-1), float(np.asarray(v).reshape(-1)[0])

    def expand_and_evaluate(self, env):                    # solo_play.py:260-278
        return self.expand_and_evaluate_with_net(env)

    def _tree_for(self, env):
        key = (id(env._batch), self.simulation_num_per_move)
        if self._mcts is None or self._mcts_key != key:
            if self._mcts is not None:
                self._mcts.close()
            self._mcts = MctsBatch(env._batch, self.simulation_num_per_move)
            self._mcts_key = key
        return self._mcts

    def search_moves(self, env):                           # solo_play.py:153-165
        """simulation_num_per_move simulations from env's position (sequential semantics, none_queue=False).  Returns
        (max leaf value, first leaf value) like the reference -- here (root Q max, root Q of the first edge), the values
        its callers only print; the tree stays on the device for calc_policy."""
        if not isinstance(env, GamePlay):
            raise TypeError("HivePlayer needs a hive_b200.GamePlay")
        m = self._tree_for(env)
        sims = self.simulation_num_per_move
        n_edges = max(len(env.actions()), 1)
        # the reference draws one Dirichlet row per root visit from the global NumPy stream
        noise = np.zeros((1, max(sims - 1, 1), max(n_edges, 1)), dtype=np.float64)
        if len(env.actions()) > 0:
            for r in range(sims - 1):
                noise[0, r] = np.random.dirichlet([C.dirichlet_alpha] * n_edges)
        m.set_root_noise(noise)
        m.search_host(self.expand_and_evaluate)
        st = m.root_stats(0)
        q = st["q"]
        return (float(q.max()) if len(q) else 0.0), (float(q[0]) if len(q) else 0.0)

    def calc_policy(self, env):                            # solo_play.py:351-374
        """(policy float64[1584], sum of visit counts) of the last search_moves(env)."""
        pi, _, sum_n = self._tree_for(env).policy()
        return pi[0], float(sum_n[0])

    def action(self, env, non_queue=True):                 # solo_play.py:110-151
        if not isinstance(env, GamePlay):
            raise TypeError("HivePlayer.action needs a hive_b200.GamePlay")
        self.max_depth = env.state.turn
        self.main_key_state = env.state_key
        self.search_moves(env)
        policy, sum_all = self.calc_policy(env)
        p = self.apply_temperature(policy, int(env.state.turn + 1) / 2)
        my_action = int(np.random.choice(range(C.ACTION_SPACE), p=p))
        return my_action, [list(policy), sum_all]

    def apply_temperature(self, policy, turn):             # solo_play.py:337-349
        tau = np.power(C.tau_decay_rate, turn)
        if tau < 0.1:
            tau = 0
        if tau == 0:
            action = np.argmax(policy)
            ret = np.zeros(C.ACTION_SPACE)
            ret[action] = 1.0
            return ret
        ret = np.power(policy, 1 / tau)
        ret /= np.sum(ret)
        return ret

    def finish_game(self, z):                              # solo_play.py:376-384
        for move in self.moves:
            move += [z]
