"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/mcts_cases.npz by running the UNMODIFIED
reference player (woker/solo_play.py::HivePlayer, none_queue=False) with the deterministic
hash-net of oracle/mcts_oracle.py.  Build container only.

Each case: a position reached by replaying `prefix` (actions from reset, so that plane history
matches), np.random.seed(seed), `sims` simulations.  Stored: the root's edges (action, N, W, Q, P),
sum_n, number of tree nodes, the returned action and policy.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from oracle.mcts_oracle import hash_net  # noqa: E402

MAX_E = 192
MAX_PREFIX = 64


def run_case(prefix_seed, plies, seed, sims):
    solo = rh.load_player()
    rng = np.random.RandomState(prefix_seed)
    env = rh.new_env()
    prefix = []
    for _ in range(plies):
        if env.game_is_over():
            break
        acts = env.actions()
        a = int(acts[rng.randint(len(acts))]) if acts else -1
        prefix.append(a)
        env.move(a)
    if env.game_is_over() or env.state.turn >= 55:
        return None          # the reference itself cannot search from a finished game (calc_policy divides by zero)
    pl = solo.HivePlayer()
    pl.none_queue = False
    pl.simulation_num_per_move = sims
    pl.expand_and_evaluate_with_net = lambda e: hash_net(e.encode_board())
    np.random.seed(seed)
    with rh.quiet():
        action, (policy, sum_all) = pl.action(env)
    node = pl.tree[env.state_key]
    acts = list(node.a.keys())
    return dict(prefix=prefix, seed=seed, sims=sims, turn=int(env.state.turn), action=int(action),
                policy=np.array(policy, dtype=np.float64), sum_all=float(sum_all),
                e_action=np.array(acts, dtype=np.int32), e_n=np.array([node.a[a].n for a in acts], dtype=np.int32),
                e_w=np.array([float(node.a[a].w) for a in acts], dtype=np.float64),
                e_q=np.array([float(node.a[a].q) for a in acts], dtype=np.float64),
                e_p=np.array([np.float32(node.a[a].p) for a in acts], dtype=np.float32),
                sum_n=int(node.sum_n), n_nodes=len(pl.tree))


def main():
    cases = [(11, 1, 1, 40), (12, 4, 2, 50), (13, 10, 3, 60), (14, 17, 4, 60), (15, 24, 5, 50), (16, 33, 6, 60),
             (17, 46, 7, 60), (0, 50, 8, 60), (1, 52, 9, 50), (20, 8, 10, 80), (21, 29, 11, 40), (2, 48, 12, 70),
             (3, 53, 13, 30), (4, 40, 14, 64)]
    out = []
    for c in cases:
        r = run_case(*c)
        if r is None:
            print(c, 'skipped: finished game'); continue
        out.append(r)
        print(c, "turn", r["turn"], "edges", len(r["e_action"]), "sum_n", r["sum_n"], "nodes", r["n_nodes"],
              "maxN", int(r["e_n"].max()), flush=True)
    n = len(out)
    prefix = np.full((n, MAX_PREFIX), -2, dtype=np.int32)
    e_action = np.full((n, MAX_E), -2, dtype=np.int32)
    e_n = np.zeros((n, MAX_E), dtype=np.int32)
    e_w = np.zeros((n, MAX_E), dtype=np.float64)
    e_q = np.zeros((n, MAX_E), dtype=np.float64)
    e_p = np.zeros((n, MAX_E), dtype=np.float32)
    for i, r in enumerate(out):
        prefix[i, :len(r["prefix"])] = r["prefix"]
        k = len(r["e_action"])
        e_action[i, :k] = r["e_action"]; e_n[i, :k] = r["e_n"]; e_w[i, :k] = r["e_w"]
        e_q[i, :k] = r["e_q"]; e_p[i, :k] = r["e_p"]
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "mcts_cases.npz"),
        prefix=prefix, n_prefix=np.array([len(r["prefix"]) for r in out], dtype=np.int32),
        seed=np.array([r["seed"] for r in out], dtype=np.int32), sims=np.array([r["sims"] for r in out], dtype=np.int32),
        turn=np.array([r["turn"] for r in out], dtype=np.int32), action=np.array([r["action"] for r in out], dtype=np.int32),
        policy=np.array([r["policy"] for r in out]), sum_all=np.array([r["sum_all"] for r in out]),
        n_edges=np.array([len(r["e_action"]) for r in out], dtype=np.int32),
        e_action=e_action, e_n=e_n, e_w=e_w, e_q=e_q, e_p=e_p,
        sum_n=np.array([r["sum_n"] for r in out], dtype=np.int32),
        n_nodes=np.array([r["n_nodes"] for r in out], dtype=np.int32))


if __name__ == "__main__":
    main()
