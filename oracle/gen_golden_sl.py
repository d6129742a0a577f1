"""TEST INFRASTRUCTURE ONLY -- golden vectors for the recorded-game ingest.

Runs the UNMODIFIED reference ``woker/sl.py::get_buffer`` (:146-231) on synthetic game records and
stores what it returned in tests/golden/sl_ingest.npz:

    python oracle/gen_golden_sl.py

A record is the reference's own format, one row per recorded move: [piece, x, y, player, bot] with piece in
{"Q","B1","B2","S1","S2","G1".."G3","A1".."A3"}, x in index_char (H..S), y in index_number (7..18), player
"W"/"B", bot 0/1.  The games are random legal games played by the reference's GamePlay (so the records are
legal), then edited to cover get_buffer's branches:

  * plain games to the end (decisive or not), with and without bot-weighted moves (BOT_WEIGHT 0.24),
  * a game truncated before the end (value 0 for every row),
  * games in which a side was skipped (the next recorded mover is not the side to move -> skip_turn
    before the move, sl.py:157-161), and one whose missing move makes a later row illegal,
  * a game holding an illegal row (-> "CCC": the whole game is discarded, data == []).

Stored per game: the record rows, and per returned sample the 56 planes bit-packed (plane 31 = turn kept
apart), the policy's single non-zero (index, weight), the value and [game_len_for_side, counter].
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

SHORT = ["Q", "B1", "B2", "S1", "S2", "G1", "G2", "G3", "A1", "A2", "A3"]


def random_record(seed, bot_white=False, bot_black=False, decisive_bias=False, skip_at=()):
    """Plays a random legal game with the reference env and returns its record rows.  At the turns in
    `skip_at` the side to move is skipped without a row (the record then shows the same colour twice)."""
    rng = np.random.RandomState(seed)
    env = rh.new_env()
    rows = []
    while not env.game_is_over() and env.state.turn < 55:
        if env.state.turn in skip_at:
            skip_at = tuple(t for t in skip_at if t != env.state.turn)
            env.skip_turn()
            continue
        acts = env.actions()
        if not acts:
            env.move(-1)
            continue
        a = int(acts[rng.randint(len(acts))])
        if decisive_bias:
            # prefer moves next to the opponent queen so that some games end by surround
            opp = env.black_pieces_set if env.state.player() == 0 else env.white_pieces_set
            qtile = list(opp.values())[0][0]
            if qtile.axial_coords != (99, 99):
                near = [x for x in acts if env.board_matrix[(x // 11) // 12, (x // 11) % 12] in qtile.adjacent_tiles]
                if near and rng.rand() < 0.85:
                    a = int(near[rng.randint(len(near))])
        cell, k = divmod(a, 11)
        q, r = divmod(cell, 12)
        tile = env.board_matrix[q, r]
        player = "W" if env.state.player() == 0 else "B"
        bot = int((player == "W" and bot_white) or (player == "B" and bot_black))
        rows.append([SHORT[k], tile.core_index[0], tile.core_index[1], player, bot])
        env.move(a)
    return rows


def run_reference(rows):
    rh.load()
    os.makedirs("/tmp/hive_sl_scratch", exist_ok=True)
    os.chdir("/tmp/hive_sl_scratch")
    from woker import sl  # noqa
    with rh.quiet():
        data, _ = sl.get_buffer([list(r) for r in rows])
    return data


def pack(data):
    n = len(data)
    planes = np.zeros((n, 56, 18), dtype=np.uint8)
    plane31 = np.zeros(n, dtype=np.int32)
    pol_idx = np.zeros(n, dtype=np.int32)
    pol_w = np.zeros(n, dtype=np.float64)
    value = np.zeros(n, dtype=np.int32)
    lens = np.zeros((n, 2), dtype=np.int32)
    for i, (state, policy, v, gl) in enumerate(data):
        p = np.asarray(state, dtype=np.float64)
        assert p.shape == (12, 12, 56)
        chw = p.transpose(2, 0, 1).reshape(56, 144)
        assert np.all(chw[31] == chw[31][0])
        plane31[i] = int(chw[31][0])
        rest = np.delete(chw, 31, axis=0)
        assert np.all((rest == 0) | (rest == 1))
        u = chw.astype(np.uint8)
        u[31] = 0
        planes[i] = np.packbits(u, axis=1, bitorder="little")
        pol = np.asarray(policy, dtype=np.float64)
        nz = np.nonzero(pol)[0]
        assert len(nz) == 1 and pol.shape == (1584,)
        pol_idx[i], pol_w[i] = nz[0], pol[nz[0]]
        value[i] = v
        lens[i] = gl
    return planes, plane31, pol_idx, pol_w, value, lens


def main():
    games = []
    games.append(("plain_0", random_record(0)))
    games.append(("plain_bot_white", random_record(1, bot_white=True)))
    games.append(("plain_bot_black", random_record(2, bot_black=True)))
    for s in range(3, 40):                                   # find decisive games
        rows = random_record(s, decisive_bias=True, bot_white=(s % 2 == 0))
        d = run_reference(rows)
        if d and d[0][2] != 0:
            games.append(("decisive_%d" % s, rows))
        if sum(1 for n, _ in games if n.startswith("decisive")) >= 3:
            break
    base = random_record(50)
    games.append(("truncated", base[:17]))
    games.append(("skipped_once", random_record(53, skip_at=(12,))[:30]))       # same colour twice -> skip_turn
    games.append(("skipped_thrice", random_record(54, skip_at=(9, 20, 21), bot_black=True)))
    dropped = [r for i, r in enumerate(base) if i != 9]      # a move missing: skip_turn, then (here) an illegal row -> discarded
    games.append(("dropped_move", dropped[:30]))
    bad = [list(r) for r in random_record(52)[:20]]
    bad[12] = ["A3", "H", "7", bad[12][3], 0]                # far corner: not a legal target
    games.append(("illegal_row", bad))
    games.append(("empty", []))

    out = {"names": np.array([n for n, _ in games])}
    starts, rec_starts = [0], [0]
    P, T, I, W, V, L, R = [], [], [], [], [], [], []
    for name, rows in games:
        data = run_reference(rows)
        planes, plane31, pi, pw, v, ln = pack(data)
        print("%-18s rows %3d  samples %3d  values %s" % (name, len(rows), len(data), sorted(set(v.tolist()))))
        P.append(planes); T.append(plane31); I.append(pi); W.append(pw); V.append(v); L.append(ln)
        R += [[str(x) for x in r] for r in rows]
        starts.append(starts[-1] + len(data)); rec_starts.append(rec_starts[-1] + len(rows))
    out.update(sample_start=np.array(starts), record_start=np.array(rec_starts),
               records=np.array(R, dtype="U4").reshape(-1, 5),
               planes=np.concatenate(P), plane31=np.concatenate(T), policy_index=np.concatenate(I),
               policy_weight=np.concatenate(W), value=np.concatenate(V), lens=np.concatenate(L))
    path = os.path.join(ROOT, "tests", "golden", "sl_ingest.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
