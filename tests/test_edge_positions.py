"""Injected edge-case positions evaluated by the REAL reference (tests/golden/edge_positions.npz,
oracle/gen_golden_edge.py): 5-high stacks (planes 24-29), doubly surrounded queens, positions
without legal actions, a Grasshopper on each of the 144 origins (runs crossing the board edge),
random connected hives with random stacks.  Checked against the oracle, the emulated kernels (CPU)
and the real kernels through the C ABI (GPU)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E = np.load(os.path.join(ROOT, "tests", "golden", "edge_positions.npz"))
N = len(E["turn"])


def _expect_planes(i):
    return E["planes"][i], int(E["plane31"][i])


def _check(i, legal, planes_u8, done, winner, key=None):
    n = E["n_legal"][i]
    assert list(legal) == E["legal"][i][:n].tolist(), i
    bits, t31 = _expect_planes(i)
    assert (planes_u8[31] == t31).all(), i
    p = planes_u8.copy()
    p[31] = 0
    assert (np.packbits(p, axis=1, bitorder="little") == bits).all(), i
    assert bool(done) == bool(E["done"][i]) and int(winner) == int(E["winner"][i]), i
    if key is not None:
        assert key == str(E["key"][i]), i


def test_fixture_covers_the_edge_cases():
    assert int(E["levels"].max()) == 4                               # a 5-high stack
    assert int(((E["done"] == 1) & (E["winner"] == 0)).sum()) >= 1     # both queens surrounded
    assert int((E["n_legal"] == 0).sum()) >= 1                         # side to move must pass
    planes = np.unpackbits(E["planes"], axis=2, bitorder="little")[:, :, :144]
    assert planes[:, 26].any() or planes[:, 29].any()                  # "top of a 5-stack" beetle planes are live


def test_oracle_on_edge_positions():
    from oracle.hive_oracle import OracleEnv
    o = OracleEnv()
    for i in range(N):
        o.load(int(E["turn"][i]), E["cells"][i], E["levels"][i])
        done = o.game_is_over()
        _check(i, o.actions(), o.planes(), done, o.winner, o.state_key)


def test_emulated_kernels_on_edge_positions():
    from tests.emu.emu import EmuBatch
    n = 31
    for start in range(0, N, n):
        idx = list(range(start, min(N, start + n)))
        b = EmuBatch(len(idx), sched_seed=start)
        for g, i in enumerate(idx):
            b.load(g, int(E["turn"][i]), E["cells"][i], E["levels"][i])
        for g, i in enumerate(idx):
            _check(i, b.actions(g), b.planes_u8(g), b.done(g), b.winner(g))


@pytest.mark.gpu
def test_device_kernels_on_edge_positions():
    import hive_b200
    b = hive_b200.HiveBatch(N)
    for i in range(N):
        b.load_state(i, int(E["turn"][i]), E["cells"][i], E["levels"][i])
    acts = b.actions()
    planes = (b.planes_bf16().astype(np.uint32) << 16).view(np.float32).astype(np.uint8)
    turn, winner, done = b.status()
    for i in range(N):
        _check(i, acts[i], planes[i], done[i], winner[i], b.state_key(i))
