"""Live differential run of the C oracle against the unmodified Python reference.  Only possible
in the build container (/root/reference); skipped elsewhere -- the committed golden vectors carry
the same evidence to the GPU box."""
import numpy as np
import pytest

from oracle import ref_harness as rh
from oracle.hive_oracle import OracleEnv

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


@pytest.mark.parametrize("seed", [7001, 7002])
def test_live_game(seed):
    rng = np.random.RandomState(seed)
    env = rh.new_env()
    o = OracleEnv()
    while True:
        done, winner = rh.status(env)
        turn, cells, levels = rh.position(env)
        ot, oc, ol = o.position()
        assert ot == turn and (oc == cells).all() and (ol == levels).all()
        la = np.array(env.actions(), dtype=np.int32)
        assert la.tolist() == o.actions().tolist()
        bits, tval = rh.planes_bits(env)
        pl = o.planes()
        assert (pl[31] == tval).all()
        pl[31] = 0
        assert (np.packbits(pl, axis=1, bitorder="little") == bits).all()
        assert o.state_key == env.state_key
        assert o.game_is_over() == done and o.winner == winner
        if done or turn >= 55:
            break
        a = int(la[rng.randint(len(la))]) if len(la) else -1
        env.move(a)
        o.move(a)
