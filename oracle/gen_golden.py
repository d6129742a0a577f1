"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the
UNMODIFIED Python reference (/root/reference, build container only).

    python oracle/gen_golden.py            # rewrites tests/golden/

Fixtures (all produced by GamePlay / HivePlayer themselves, nothing restated):

  env_rollouts.npz   per-ply records of seeded games:
      policy "uniform": SURVEY 8d Config-1 rule, rng=RandomState(seed)
      policy "beetle" : same rng, but prefers Beetle moves onto occupied cells
                        (builds 2..5-high stacks, planes 24-29)
      arrays: game_start[g], game_seed[g], game_policy[g] and per ply
      turn, cells[22], levels[22], n_legal, legal[<=160] (padded -1),
      planes[56][18] (bit-packed little-endian, plane 31 zeroed), plane31,
      done, winner, action (taken from this ply, -2 at the last record), key (str)
  env_pins.json      transcript hashes / legal sums
  mcts_cases.npz     HivePlayer root statistics with a deterministic hash-net
                     (see oracle/mcts_oracle.py for the net), written by
                     oracle/gen_golden_mcts.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

MAX_LEGAL = 160


def beetle_bias_pick(env, acts, rng):
    """Prefer (p=0.75) a legal Beetle move whose destination is occupied."""
    if not acts:
        return -1
    onto = [a for a in acts if (a % 11) in (1, 2) and env.board_matrix[(a // 11) // 12, (a // 11) % 12].has_pieces()]
    if onto and rng.rand() < 0.75:
        return int(onto[rng.randint(len(onto))])
    beetle = [a for a in acts if (a % 11) in (1, 2)]
    if beetle and rng.rand() < 0.5:
        return int(beetle[rng.randint(len(beetle))])
    return int(acts[rng.randint(len(acts))])


def play(seed, policy):
    rng = np.random.RandomState(seed)
    env = rh.new_env()
    recs = []
    while True:
        done, winner = rh.status(env)
        turn, cells, levels = rh.position(env)
        bits, tval = rh.planes_bits(env)
        acts = list(env.actions())
        rec = dict(turn=turn, cells=cells, levels=levels, legal=acts, planes=bits,
                   plane31=tval, done=done, winner=winner, key=env.state_key, action=-2)
        recs.append(rec)
        if done or turn >= 55:
            break
        if policy == "uniform":
            a = int(acts[rng.randint(len(acts))]) if acts else -1
        else:
            a = beetle_bias_pick(env, acts, rng)
        rec["action"] = a
        env.move(a)
    return recs


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    games = [(s, "uniform") for s in range(8)] + [(100 + s, "beetle") for s in range(16)]
    cols = {k: [] for k in ("turn", "cells", "levels", "n_legal", "legal", "planes", "plane31",
                            "done", "winner", "action", "key")}
    game_start, game_seed, game_policy = [], [], []
    pins = {}
    max_h = 0
    for seed, policy in games:
        recs = play(seed, policy)
        game_start.append(len(cols["turn"]))
        game_seed.append(seed)
        game_policy.append(policy)
        for r in recs:
            assert len(r["legal"]) <= MAX_LEGAL
            leg = np.full(MAX_LEGAL, -1, dtype=np.int32)
            leg[:len(r["legal"])] = r["legal"]
            cols["turn"].append(r["turn"]); cols["cells"].append(r["cells"]); cols["levels"].append(r["levels"])
            cols["n_legal"].append(len(r["legal"])); cols["legal"].append(leg)
            cols["planes"].append(r["planes"]); cols["plane31"].append(r["plane31"])
            cols["done"].append(r["done"]); cols["winner"].append(r["winner"])
            cols["action"].append(r["action"]); cols["key"].append(r["key"])
            max_h = max(max_h, int(r["levels"].max()) + 1)
        if policy == "uniform":
            tr = [r["action"] for r in recs if r["action"] != -2]
            pins[str(seed)] = dict(first8=tr[:8], legal_sum=int(sum(len(r["legal"]) for r in recs[:len(tr)])),
                                   sha16=rh.transcript_hash(tr), plies=len(tr))
        print(seed, policy, len(recs), "plies; tallest stack so far", max_h, flush=True)
    game_start.append(len(cols["turn"]))
    np.savez_compressed(
        os.path.join(out_dir, "env_rollouts.npz"),
        game_start=np.array(game_start, dtype=np.int32), game_seed=np.array(game_seed, dtype=np.int32),
        game_policy=np.array(game_policy), turn=np.array(cols["turn"], dtype=np.int32),
        cells=np.array(cols["cells"], dtype=np.uint8), levels=np.array(cols["levels"], dtype=np.uint8),
        n_legal=np.array(cols["n_legal"], dtype=np.int32), legal=np.array(cols["legal"], dtype=np.int32),
        planes=np.array(cols["planes"], dtype=np.uint8), plane31=np.array(cols["plane31"], dtype=np.int32),
        done=np.array(cols["done"], dtype=np.uint8), winner=np.array(cols["winner"], dtype=np.uint8),
        action=np.array(cols["action"], dtype=np.int32), key=np.array(cols["key"]))
    with open(os.path.join(out_dir, "env_pins.json"), "w") as f:
        json.dump(dict(source="GamePlay rollouts, rng=np.random.RandomState(seed), SURVEY.md Appendix D",
                       reset_actions=[858, 859, 861, 863, 866], pins=pins), f, indent=1)
    print("tallest stack in fixtures:", max_h)


if __name__ == "__main__":
    main()
