"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/mcts_deep.npz: searches of 250 and 500 simulations run by the
UNMODIFIED reference player (woker/solo_play.py::HivePlayer, none_queue=False, hash-net), the depths of BASELINE
configs[3] / [4].  They pin what the 24-80 simulation cases of mcts_cases.npz cannot: transposition merges deep in
the tree, hundreds of nodes, edge-arena pressure.  Build container only (about a minute per case)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.gen_golden_mcts import MAX_E, MAX_PREFIX, run_case  # noqa: E402


def main():
    # (prefix seed, plies, np.random seed, sims)
    cases = [(31, 12, 21, 250), (32, 30, 22, 250), (33, 22, 23, 500), (34, 44, 24, 500)]
    out = []
    for c in cases:
        r = run_case(*c)
        if r is None:
            print(c, "skipped: finished game"); continue
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
        os.path.join(ROOT, "tests", "golden", "mcts_deep.npz"),
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
