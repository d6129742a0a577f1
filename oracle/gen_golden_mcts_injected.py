"""TEST INFRASTRUCTURE ONLY -- tests/golden/mcts_injected.npz: searches of the UNMODIFIED reference player
from INJECTED positions (oracle/ref_harness.inject) that random play does not reach: roots without any legal
action (the pass edge -1 and the policy[-1] quirk of solo_play.py:298-300,360-362), a root one ply before a
drawn game, a tall-stack position.  Build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from oracle.mcts_oracle import hash_net  # noqa: E402

MAX_E = 192


def main():
    E = np.load(os.path.join(ROOT, "tests", "golden", "edge_positions.npz"))
    solo = rh.load_player()
    idx = [int(i) for i in np.nonzero((E["n_legal"] == 0) & (E["done"] == 0) & (E["turn"] < 54))[0]]
    idx += [0, 1]                                              # the hand-made 5-high stack positions
    idx += [int(i) for i in np.nonzero((E["n_legal"] > 40) & (E["levels"].max(axis=1) >= 2) & (E["done"] == 0) & (E["turn"] < 50))[0][:3]]
    out = []
    for i in idx:
        env = rh.inject(int(E["turn"][i]), E["cells"][i], E["levels"][i])
        if env.game_is_over():
            continue
        pl = solo.HivePlayer()
        pl.none_queue = False
        pl.simulation_num_per_move = 24
        pl.expand_and_evaluate_with_net = lambda e: hash_net(e.encode_board())
        np.random.seed(1000 + i)
        with rh.quiet():
            action, (policy, sum_all) = pl.action(env)
        node = pl.tree[env.state_key]
        acts = list(node.a.keys())
        out.append(dict(i=i, action=int(action), policy=np.array(policy, dtype=np.float64), sum_all=float(sum_all),
                        e_action=acts, e_n=[node.a[a].n for a in acts], e_w=[float(node.a[a].w) for a in acts],
                        e_p=[float(np.float32(node.a[a].p)) for a in acts], sum_n=int(node.sum_n), n_nodes=len(pl.tree)))
        print(i, "turn", int(E["turn"][i]), "legal", int(E["n_legal"][i]), "edges", acts[:4], "sum_n", node.sum_n, "nodes", len(pl.tree),
              "action", action, flush=True)
    n = len(out)
    ea = np.full((n, MAX_E), -2, dtype=np.int32); en = np.zeros((n, MAX_E), dtype=np.int32)
    ew = np.zeros((n, MAX_E)); ep = np.zeros((n, MAX_E), dtype=np.float32)
    for j, r in enumerate(out):
        k = len(r["e_action"])
        ea[j, :k] = r["e_action"]; en[j, :k] = r["e_n"]; ew[j, :k] = r["e_w"]; ep[j, :k] = r["e_p"]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mcts_injected.npz"),
                        edge_index=np.array([r["i"] for r in out], dtype=np.int32), sims=np.full(n, 24, dtype=np.int32),
                        seed=np.array([1000 + r["i"] for r in out], dtype=np.int32),
                        action=np.array([r["action"] for r in out], dtype=np.int32), policy=np.array([r["policy"] for r in out]),
                        sum_all=np.array([r["sum_all"] for r in out]), n_edges=np.array([len(r["e_action"]) for r in out], dtype=np.int32),
                        e_action=ea, e_n=en, e_w=ew, e_p=ep, sum_n=np.array([r["sum_n"] for r in out], dtype=np.int32),
                        n_nodes=np.array([r["n_nodes"] for r in out], dtype=np.int32))


if __name__ == "__main__":
    main()
