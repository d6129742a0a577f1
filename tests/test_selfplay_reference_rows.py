"""Content pin of the self-play game loop: the drop-in self_play_buffer (GamePlay + HivePlayer on the GPU)
must return, row for row, what the UNMODIFIED reference loop returned under the same np.random seed and the same
hash-net (tests/golden/selfplay_rows.npz, written by oracle/gen_golden_selfplay.py from
woker/self_play_with_train.py:146-219): planes, pi, value (draw / cut game => -1 for both sides), [game_len, idx] --
and write_play_file must put exactly those rows on disk.  Plus the evaluator's report and gate (f3)."""
import json
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def golden():
    p = os.path.join(os.path.dirname(__file__), "golden", "selfplay_rows.npz")
    return np.load(p)


def _player_factory(hb, sims):
    from oracle.mcts_oracle import hash_net

    def make(pipes):
        pl = hb.HivePlayer(pipes=pipes)
        pl.simulation_num_per_move = sims
        pl.expand_and_evaluate_with_net = lambda e: hash_net(e.encode_board())
        return pl
    return make


@pytest.mark.gpu
def test_self_play_buffer_rows_equal_the_reference(golden, tmp_path):
    import hive_b200 as hb
    G = golden
    sims = int(G["sims"])
    starts = G["game_start"]
    decisive = 0
    for gi, seed in enumerate(G["seeds"]):
        np.random.seed(int(seed))
        cur = [None]
        data, (value_white,) = hb.self_play_buffer(cur, make_player=_player_factory(hb, sims))
        assert cur == [None]                                   # the pipe bundle goes back to the pool
        lo, hi = int(starts[gi]), int(starts[gi + 1])
        assert value_white == int(G["value_white"][gi]) and len(data) == hi - lo
        decisive += value_white != 0
        for i, (state, policy, value, lens) in enumerate(data):
            chw = np.asarray(state, dtype=np.float64).transpose(2, 0, 1).reshape(56, 144)
            assert (chw[31] == G["plane31"][lo + i]).all()
            chw[31] = 0
            assert (np.packbits(chw.astype(np.uint8), axis=1, bitorder="little") == G["planes"][lo + i]).all(), (gi, i)
            assert (np.asarray(policy, dtype=np.float64) == G["pi"][lo + i]).all(), (gi, i)
            assert value == int(G["value"][lo + i]) and list(lens) == G["lens"][lo + i].tolist(), (gi, i)
        if value_white == 0:
            assert all(row[2] == -1 for row in data)           # self_play.py:188-189
    # the on-disk form (self_play.py:100-112, sl.py:49-60): the rows as they are
    last = [(np.asarray(r[0]), r[1], r[2], r[3]) for r in data[:3]]
    bf16 = [((np.asarray(r[0], dtype=np.float32).transpose(2, 0, 1).reshape(-1).view(np.uint32) >> 16).astype(np.uint16), r[1], r[2], r[3])
            for r in last]
    path = hb.write_play_file(bf16, directory=str(tmp_path))
    rows = json.load(open(path))
    assert len(rows) == 3 and rows[0][0] == data[0][0] and rows[0][1] == [float(x) for x in data[0][1]]
    assert rows[2][2] == data[2][2] and rows[2][3] == list(data[2][3])


def test_evaluation_report_and_gate():
    """woker/evaluation.py:66-88 on a four-game fixture: white win rate = share of +1, mean length, distinct final keys."""
    import hive_b200 as hb
    win_lose = [1, -1, 0, 1]
    lens = [[31], [55], [55], [40]]
    keys = ["k1", "k2", "k2", "k3"]
    r = hb.evaluation_report(win_lose, lens, keys)
    # the reference's own expressions
    assert r["white_win_rate"] == len(np.where(np.array(win_lose) == 1)[0]) / len(win_lose)
    assert r["mean_game_len"] == np.round(np.mean(lens), 2)
    assert r["distinct_final_positions"] == len(np.unique(keys)) / len(win_lose)
    assert r["counter"] == {-1: 1, 0: 1, 1: 2} and r["total_games"] == 4
    assert hb.accept_new_network(56, 44) == (True, 0.56)
    assert hb.accept_new_network(54, 46)[0] is False
    assert hb.accept_new_network(0, 0) == (False, 0.0)
