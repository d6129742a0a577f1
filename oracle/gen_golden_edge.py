"""TEST INFRASTRUCTURE ONLY -- tests/golden/edge_positions.npz: hand-built and random INJECTED
positions evaluated by the unmodified Python reference (oracle/ref_harness.inject): 5-high stacks
(planes 24-29), doubly surrounded queens (drawn game), positions without legal actions, Grasshoppers
on every one of the 144 origins with runs crossing the board edge (the raw-delta is_straight_line of
move_checker.py:249-265), random connected hives with random stacks.  Build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

MAX_LEGAL = 200


def nbrs(c):
    q, r = divmod(c, 12)
    return [((q - 1) % 12) * 12 + r, ((q + 1) % 12) * 12 + r, q * 12 + (r + 1) % 12, q * 12 + (r - 1) % 12,
            ((q - 1) % 12) * 12 + (r - 1) % 12, ((q + 1) % 12) * 12 + (r + 1) % 12]


def random_hive(rng, n_cells, origin=None):
    cells = [int(rng.randint(144)) if origin is None else origin]
    while len(cells) < n_cells:
        c = cells[rng.randint(len(cells))]
        x = nbrs(c)[rng.randint(6)]
        if x not in cells:
            cells.append(x)
    return cells


def random_position(rng, hopper_origin=None):
    """Random connected hive; beetles may sit on top of other pieces (up to 5 high)."""
    cells = np.full(22, 255, dtype=np.uint8)
    levels = np.zeros(22, dtype=np.uint8)
    ground = [p for p in range(22) if (p % 11) not in (1, 2)]
    rng.shuffle(ground)
    k = int(rng.randint(3, len(ground) + 1))
    ground = ground[:k]
    if hopper_origin is not None and 5 not in ground:
        ground[0] = 5
    hive = random_hive(rng, len(ground), hopper_origin)
    if hopper_origin is not None:
        ground.remove(5)
        ground.insert(0, 5)                       # white G0 sits on the requested origin
    for p, c in zip(ground, hive):
        cells[p] = c
    height = {c: 1 for c in hive}
    for b in (1, 2, 12, 13):
        u = rng.rand()
        if u < 0.55:                              # climb on something
            c = hive[rng.randint(len(hive))]
            if hopper_origin is not None and c == hopper_origin:
                continue
            if height[c] < 5:
                cells[b] = c; levels[b] = height[c]; height[c] += 1
        elif u < 0.8:                             # on the ground next to the hive
            c = hive[rng.randint(len(hive))]
            free = [x for x in nbrs(c) if x not in height]
            if free:
                x = free[rng.randint(len(free))]
                cells[b] = x; levels[b] = 0; height[x] = 1; hive.append(x)
    turn = int(rng.randint(3, 55))
    return turn, cells, levels


def handmade():
    out = []
    # 5-high stack: black ant at (6,6) under four beetles; a few neighbours
    c = np.full(22, 255, dtype=np.uint8); l = np.zeros(22, dtype=np.uint8)
    c[19] = 78; c[1] = 78; l[1] = 1; c[12] = 78; l[12] = 2; c[2] = 78; l[2] = 3; c[13] = 78; l[13] = 4
    c[0] = 79; c[11] = 77; c[8] = 66; c[3] = 90
    out.append((20, c, l)); out.append((21, c.copy(), l.copy()))
    # both queens surrounded (drawn): queens adjacent, every other neighbour occupied
    c = np.full(22, 255, dtype=np.uint8); l = np.zeros(22, dtype=np.uint8)
    wq, bq = 78, 79
    ring = sorted((set(nbrs(wq)) | set(nbrs(bq))) - {wq, bq})
    pieces = [1, 2, 3, 4, 5, 6, 7, 8, 12, 13, 14, 15]
    c[0] = wq; c[11] = bq
    for p, x in zip(pieces, ring):
        c[p] = x
    out.append((30, c, l)); out.append((31, c.copy(), l.copy()))
    # only white queen surrounded
    c2 = c.copy(); c2[11] = 255
    free = [x for x in nbrs(wq) if x not in c2.tolist()]
    for p, x in zip([16, 17, 18], free):
        c2[p] = x
    out.append((33, c2, l.copy()))
    # side to move has every piece on board and pinned in a line: few or no actions
    c = np.full(22, 255, dtype=np.uint8); l = np.zeros(22, dtype=np.uint8)
    line = [6 * 12 + r for r in range(1, 12)]
    for p, x in zip(range(11), line):
        c[p] = x
    c[11] = 5 * 12 + 0
    out.append((41, c, l))
    return out


def main():
    rng = np.random.RandomState(20261018)
    pos = handmade()
    for o in range(144):                              # a Grasshopper on every origin
        pos.append(random_position(rng, hopper_origin=o))
    for _ in range(160):
        pos.append(random_position(rng))
    cols = {k: [] for k in ("turn", "cells", "levels", "n_legal", "legal", "planes", "plane31", "done", "winner", "key")}
    tallest, draws, passes = 0, 0, 0
    for turn, cells, levels in pos:
        env = rh.inject(turn, cells, levels)
        done, winner = rh.status(env)
        bits, tval = rh.planes_bits(env)
        acts = list(env.actions())
        assert len(acts) <= MAX_LEGAL
        leg = np.full(MAX_LEGAL, -1, dtype=np.int32); leg[:len(acts)] = acts
        cols["turn"].append(turn); cols["cells"].append(cells); cols["levels"].append(levels)
        cols["n_legal"].append(len(acts)); cols["legal"].append(leg); cols["planes"].append(bits); cols["plane31"].append(tval)
        cols["done"].append(done); cols["winner"].append(winner); cols["key"].append(env.state_key)
        tallest = max(tallest, int(levels.max()) + 1)
        draws += int(done and winner == 0); passes += int(len(acts) == 0)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "edge_positions.npz"),
                        turn=np.array(cols["turn"], dtype=np.int32), cells=np.array(cols["cells"], dtype=np.uint8),
                        levels=np.array(cols["levels"], dtype=np.uint8), n_legal=np.array(cols["n_legal"], dtype=np.int32),
                        legal=np.array(cols["legal"], dtype=np.int32), planes=np.array(cols["planes"], dtype=np.uint8),
                        plane31=np.array(cols["plane31"], dtype=np.int32), done=np.array(cols["done"], dtype=np.uint8),
                        winner=np.array(cols["winner"], dtype=np.uint8), key=np.array(cols["key"]))
    print(len(pos), "positions; tallest stack", tallest, "; drawn", draws, "; without legal actions", passes,
          "; max legal", max(cols["n_legal"]))


if __name__ == "__main__":
    main()
