"""The host-driven game loop inside the library (hive_host_loop_*): native driver threads, one CUDA-graph launch per part
and step, host policy between the steps -- the loop woker/self_play.py:54-56,116-193 runs around GamePlay.  Checked
against the oracle: the games of every part must sit exactly where the oracle's replay of the same policy puts them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hive_b200
    return hive_b200


def _check_part(b, oracles):
    acts = b.actions()
    planes = (b.planes_bf16().astype(np.uint32) << 16).view(np.float32).astype(np.uint8)
    turn, winner, done = b.status()
    for g, o in enumerate(oracles):
        assert turn[g] == o.turn and bool(done[g]) == o.game_is_over() and winner[g] == o.winner, g
        assert acts[g].tolist() == o.actions().tolist(), g
        assert (planes[g] == o.planes()).all(), g


def test_random_policy_twin_matches_oracle_replay(hb):
    from oracle.hive_oracle import OracleEnv
    n, parts, threads, steps, seed = 200, 3, 2, 75, 4242          # 75 steps: every game passes turn 55 and is reset once
    loop = hb.HostLoop(n, parts=parts, threads=threads)
    assert loop.parts == parts and loop.threads == threads
    s0 = loop.env_steps()
    r = loop.run(steps, seed=seed, max_turn=55)
    assert r["seconds"] > 0
    expect_steps = 0
    for i in range(parts):
        b, first = loop.part(i)
        oracles = [OracleEnv() for _ in range(b.n)]
        episodes = [0] * b.n
        pseed = seed + 77 * (i + 1)
        for _ in range(steps):
            for g, o in enumerate(oracles):
                if o.game_is_over() or o.turn >= 55:
                    o.reset(); episodes[g] += 1
                else:
                    o.move(o.pick_action(pseed, g + b.n * episodes[g]))
                    expect_steps += 1
        _check_part(b, oracles)
    assert loop.env_steps() - s0 == expect_steps
    # a second run continues from where the first one stopped (the loop keeps its per-game episode counters)
    loop.run(3, seed=seed, max_turn=55)
    loop.close()


def test_python_policy_callback(hb):
    """A caller-supplied policy (lowest legal action; reset when over) driven through the same native loop."""
    from oracle.hive_oracle import OracleEnv
    n, steps = 70, 30
    calls = []

    def policy(part, first, mask, count, status, actions):
        calls.append(part)
        for g in range(len(count)):
            done, turn = (status[g] >> 16) & 0xFF, status[g] & 0xFF
            if done or turn >= 55:
                actions[g] = -3
            elif count[g] == 0:
                actions[g] = -1
            else:
                w = int(np.flatnonzero(mask[g])[0])
                m = int(mask[g][w])
                actions[g] = w * 64 + (m & -m).bit_length() - 1

    loop = hb.HostLoop(n, parts=2, threads=2)
    loop.run(steps, policy=policy)
    assert len(calls) == 2 * steps
    for i in range(2):
        b, first = loop.part(i)
        oracles = [OracleEnv() for _ in range(b.n)]
        for _ in range(steps):
            for o in oracles:
                if o.game_is_over() or o.turn >= 55:
                    o.reset()
                else:
                    la = o.actions()
                    o.move(int(la[0]) if len(la) else -1)
        _check_part(b, oracles)
    loop.close()
