"""Recorded-game ingest (woker/sl.py:146-231).

tests/golden/sl_ingest.npz holds what the UNMODIFIED reference get_buffer returned for a set of synthetic records
(oracle/gen_golden_sl.py): plain, bot-weighted, decisive, truncated, skipped-turn, illegal-row and empty games.
CPU: the record -> action mapping and a restatement of get_buffer on the C oracle environment reproduce the
golden samples.  GPU: the product's get_buffer (facade, one game) and get_buffers (all games in lock step on one
HiveBatch) reproduce them bit for bit.
"""
import os

import numpy as np
import pytest

import hive_b200
from hive_b200 import config as C

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "sl_ingest.npz"))
NAMES = [str(x) for x in G["names"]]


def record(i):
    a, b = int(G["record_start"][i]), int(G["record_start"][i + 1])
    return [[str(r[0]), str(r[1]), str(r[2]), str(r[3]), int(r[4])] for r in G["records"][a:b]]


def golden(i):
    a, b = int(G["sample_start"][i]), int(G["sample_start"][i + 1])
    bits = np.unpackbits(G["planes"][a:b], axis=2, bitorder="little")[:, :, :144].astype(np.float64)
    bits[:, 31, :] = G["plane31"][a:b, None]
    return dict(planes=bits, idx=G["policy_index"][a:b], w=G["policy_weight"][a:b], value=G["value"][a:b], lens=G["lens"][a:b])


def rows_to_arrays(data):
    m = len(data)
    planes = np.zeros((m, 56, 144))
    idx, w, value, lens = np.zeros(m, np.int64), np.zeros(m), np.zeros(m, np.int64), np.zeros((m, 2), np.int64)
    for i, (state, policy, v, gl) in enumerate(data):
        planes[i] = np.asarray(state, dtype=np.float64).transpose(2, 0, 1).reshape(56, 144)
        p = np.asarray(policy)
        assert p.shape == (1584,)
        nz = np.nonzero(p)[0]
        assert len(nz) == 1
        idx[i], w[i], value[i], lens[i] = nz[0], p[nz[0]], v, gl
    return dict(planes=planes, idx=idx, w=w, value=value, lens=lens)


def assert_same(got, want, name):
    assert len(got["idx"]) == len(want["idx"]), name
    for k in ("planes", "idx", "w", "value", "lens"):
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), (name, k)


def test_fixture_covers_the_branches():
    assert {"truncated", "skipped_once", "skipped_thrice", "dropped_move", "illegal_row", "empty"} <= set(NAMES)
    assert any(n.startswith("decisive") for n in NAMES) and any("bot" in n for n in NAMES)
    vals = set(int(v) for v in G["value"])
    assert vals == {-1, 0, 1}
    assert C.BOT_WEIGHT in set(float(x) for x in G["policy_weight"])


def test_record_action_mapping_and_piece_ids():
    assert hive_b200.decode_piece("Q") == "<class 'pieces.Queen'>0"
    assert hive_b200.decode_piece("G3") == "<class 'pieces.Grasshopper'>2"
    assert hive_b200.decode_piece("B1") == "<class 'pieces.Beetle'>0"
    assert hive_b200.record_action(["Q", "H", "7", "W", 0]) == 0
    assert hive_b200.record_action(["A3", "S", "18", "B", 0]) == 1583
    with pytest.raises(ValueError):
        hive_b200.record_action(["Q", "Z", "7", "W", 0])
    with pytest.raises(IndexError):
        hive_b200.record_action(["G5", "H", "7", "W", 0])
    for i, name in enumerate(NAMES):                       # games the reference kept: row j <-> sample j
        want = golden(i)
        rec = record(i)
        if len(want["idx"]) == len(rec):
            assert [hive_b200.record_action(r) for r in rec] == want["idx"].tolist(), name


def oracle_get_buffer(game):
    """sl.py:146-231 restated on the C oracle environment (test infrastructure)."""
    from oracle.hive_oracle import OracleEnv
    env = OracleEnv()
    spp, wc, bc = [], 0, 0
    for row in game:
        side = (env.turn + 1) % 2
        if (side == 1 and row[3] == "W") or (side == 0 and row[3] == "B"):
            env.move(-1)
        if row[3] == "W":
            wc += 1
            counter = wc
        else:
            bc += 1
            counter = bc
        a = hive_b200.record_action(row)
        if a not in env.actions().tolist():
            spp = []
            break
        policy = np.zeros(1584)
        policy[a] = C.BOT_WEIGHT if row[4] == 1 else 1
        spp.append([env.encode_board().tolist(), policy, row[3], counter])
        env.move(a)
    vw = 0
    if env.game_is_over():
        vw = 1 if env.winner == 1 else -1 if env.winner == 2 else 0
    return [[s, p.tolist(), 0 if vw == 0 else (vw if pl == "W" else -vw), [wc if pl == "W" else bc, c]] for s, p, pl, c in spp]


@pytest.mark.parametrize("i", range(len(NAMES)))
def test_oracle_restatement_reproduces_reference_samples(i):
    assert_same(rows_to_arrays(oracle_get_buffer(record(i))), golden(i), NAMES[i])


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(NAMES)))
def test_get_buffer_facade_reproduces_reference_samples(i):
    rec = record(i)
    data, game = hive_b200.get_buffer(rec)
    assert game is rec
    assert_same(rows_to_arrays(data), golden(i), NAMES[i])


@pytest.mark.gpu
def test_get_buffers_lockstep_batch_reproduces_reference_samples():
    recs = [record(i) for i in range(len(NAMES))] * 3         # 3 copies of every game in one batch
    res = hive_b200.get_buffers(recs)
    assert res.ticks <= max(len(r) for r in recs) + 3          # one env step per recorded move (+ skipped turns)
    for j, rec in enumerate(recs):
        i = j % len(NAMES)
        assert_same(rows_to_arrays(res.rows(j)), golden(i), NAMES[i])
        assert res.discarded[j] == (NAMES[i] in ("dropped_move", "illegal_row"))
    assert res.n_samples() == 3 * len(G["value"])
