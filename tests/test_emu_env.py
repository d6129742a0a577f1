"""The product's CUDA device source, run on the CPU lock-step warp emulator (tests/emu), against
the oracle and the reference's golden vectors.  This is the no-GPU safety net for kernel logic;
the `-m gpu` tests repeat the comparison on the real device through the C ABI."""
import numpy as np
import pytest

from oracle.hive_oracle import OracleEnv
from tests.emu.emu import EmuBatch


def _compare(e, g, o):
    ot, oc, ol = o.position()
    assert e.turn(g) == ot
    assert (e.cells(g) == oc).all() and (e.levels(g) == ol).all()
    assert e.actions(g).tolist() == o.actions().tolist()
    assert e.count[g] == len(o.actions())
    assert (e.planes_u8(g) == o.planes()).all()
    assert e.done(g) == o.game_is_over() and e.winner(g) == o.winner


@pytest.mark.parametrize("seed", range(0, 40))
def test_uniform_games(seed):
    rng = np.random.RandomState(9000 + seed)
    e, o = EmuBatch(1, sched_seed=seed), OracleEnv()
    while True:
        _compare(e, 0, o)
        if o.game_is_over() or o.turn >= 55:
            break
        acts = o.actions()
        a = int(acts[rng.randint(len(acts))]) if len(acts) else -1
        o.move(a)
        e.step(np.array([a], dtype=np.int32))


def test_golden_replay_including_tall_stacks(golden):
    G = golden
    starts = G["game_start"]
    for g in range(len(starts) - 1):
        e = EmuBatch(1, sched_seed=100 + g)
        for i in range(starts[g], starts[g + 1]):
            n = G["n_legal"][i]
            assert e.turn(0) == G["turn"][i]
            assert e.actions(0).tolist() == G["legal"][i][:n].tolist()
            pl = e.planes_u8(0)
            assert (pl[31] == G["plane31"][i]).all()
            pl[31] = 0
            assert (np.packbits(pl, axis=1, bitorder="little") == G["planes"][i]).all()
            assert e.done(0) == bool(G["done"][i]) and e.winner(0) == G["winner"][i]
            if G["action"][i] != -2:
                e.step(np.array([G["action"][i]], dtype=np.int32))


def test_random_policy_matches_oracle_rule():
    n, seed = 3, 0xC0FFEE
    e = EmuBatch(n, sched_seed=5)
    oracles = [OracleEnv() for _ in range(n)]
    episodes = [0] * n
    for _ in range(70):
        expect = []
        for g in range(n):
            o = oracles[g]
            if o.game_is_over() or o.turn >= 55:
                o.reset(); episodes[g] += 1
                expect.append(-2)
            else:
                a = o.pick_action(seed, g + n * episodes[g])
                o.move(a)
                expect.append(a)
        e.step_random(seed, max_turn=55, auto_reset=1)
        assert e.chosen.tolist() == expect
        for g in range(n):
            _compare(e, g, oracles[g])


def test_noop_and_masked_reset():
    e = EmuBatch(2, sched_seed=3)
    o = OracleEnv()
    e.step(np.array([858, -2], dtype=np.int32))
    o.move(858)
    _compare(e, 0, o)
    assert e.turn(1) == 1
    e.reset(mask=np.array([1, 0], dtype=np.uint8))
    o.reset()
    _compare(e, 0, o)
    e.step(np.array([-1, -1], dtype=np.int32))      # pass: turn advances, no history push
    o.move(-1)
    _compare(e, 0, o)
    _compare(e, 1, o)


def test_batch_of_forty_games_two_ctas():
    """40 games = one full 32-game CTA + a partial one, the games of different ages (lane <-> game in the analyse and
    encode phases: every lane of a warp sits at another position; CTA-wide flood / move queues shared by 32 games)."""
    n = 40
    e = EmuBatch(n, sched_seed=11)
    oracles = [OracleEnv() for _ in range(n)]
    rngs = [np.random.RandomState(500 + g) for g in range(n)]
    for ply in range(62):
        actions = np.full(n, -2, dtype=np.int32)
        for g, o in enumerate(oracles):
            if ply % 3 == 0 or g < 3:
                _compare(e, g, o)
            if ply < g % 7 or o.game_is_over() or o.turn >= 55:     # late starters; finished games idle (NOOP)
                continue
            acts = o.actions()
            a = int(acts[rngs[g].randint(len(acts))]) if len(acts) else -1
            o.move(a)
            actions[g] = a
        e.step(actions)
    for g, o in enumerate(oracles):
        _compare(e, g, o)


def test_rollout_kernel_equals_single_steps():
    """hive_rollout_kernel (n steps in one launch, every CTA looping on its own and storing its planes itself) leaves the
    batch exactly where n single random steps leave it -- records, legal masks, counts, status and planes."""
    n, seed = 37, 0xBEEF
    a, b = EmuBatch(n, sched_seed=21), EmuBatch(n, sched_seed=22)
    for _ in range(3):
        a.step_random(seed); b.step_random(seed)
    for _ in range(9):
        a.step_random(seed)
    b.step_random_multi(seed, 9)
    assert (a.recs == b.recs).all() and (a.legal == b.legal).all() and (a.count == b.count).all()
    assert (a.status == b.status).all() and (a.planes == b.planes).all()
    b.step_random_multi(seed, 60)                       # across the turn-55 cut: resets inside the loop
    for _ in range(60):
        a.step_random(seed)
    assert (a.recs == b.recs).all() and (a.planes == b.planes).all() and (a.legal == b.legal).all()


def test_delta_plane_store_equals_full_store():
    """hive_planes_delta_kernel (only the 32-byte sectors that differ from what the planes arena holds are rewritten;
    HIVE_B200_DELTA_STORE=1) leaves the same planes as hive_planes_kernel (all 16 KB of every evaluated game rewritten) after every step:
    random play across resets at the turn-55 cut, masked resets, and masked evaluations of loaded positions (games that
    are not evaluated in a launch keep their planes and their shadow)."""
    n, seed = 37, 0xD17A
    a, b = EmuBatch(n, sched_seed=31, full_store=False), EmuBatch(n, sched_seed=32)
    assert a.shadow is not None and b.shadow is None
    for step in range(64):
        a.step_random(seed); b.step_random(seed)
        assert (a.planes == b.planes).all(), step
        if step % 9 == 4:                                  # a few games start over
            mask = (np.arange(n) % 5 == step % 5).astype(np.uint8)
            a.reset(mask); b.reset(mask)
            assert (a.planes == b.planes).all(), step
        if step % 11 == 7:                                 # one game jumps to another game's position (OP_EVAL, masked)
            src, dst = step % n, (3 * step + 1) % n
            turn, cells, levels = a.turn(src), a.cells(src), a.levels(src)
            a.load(dst, turn, cells, levels); b.load(dst, turn, cells, levels)
            assert (a.planes == b.planes).all(), step
    assert (a.recs == b.recs).all() and (a.legal == b.legal).all()
    # the shadow is the bit image of the planes: expanding it reproduces them (plane 31 = turn in every cell)
    sh = a.shadow
    bits = np.unpackbits(sh.view(np.uint8).reshape(n, 56, 20)[:, :, :18], axis=2, bitorder="little")
    vals = a.planes.astype(np.uint32) << 16
    vals = vals.view(np.float32).reshape(n, 56, 144)
    keep = np.arange(56) != 31
    assert (vals[:, keep] == bits[:, keep]).all()
    assert (vals[:, 31] == sh[:, 155].astype(np.float32)[:, None]).all()


def test_queue_rollout_equals_single_steps():
    """hive_rollout_q_kernel + hive_planes_q_kernel (CTAs take (group, step) tickets; the emulator runs them one after the
    other, so at most two steps per call) leave the batch where single steps leave it -- records, legal masks, planes."""
    n, seed = 70, 0xFACE                                  # three groups, the last one ragged
    a, b = EmuBatch(n, sched_seed=41), EmuBatch(n, sched_seed=42)
    for rnd in range(30):                                 # 60 plies: across the turn-55 cut
        a.step_random(seed); a.step_random(seed)
        b.step_random_queue(seed, 2)
        assert (a.recs == b.recs).all() and (a.legal == b.legal).all() and (a.count == b.count).all(), rnd
        assert (a.status == b.status).all() and (a.planes == b.planes).all(), rnd
    a.step_random(seed); b.step_random_queue(seed, 1)
    assert (a.recs == b.recs).all() and (a.planes == b.planes).all()


def test_compact_legal_lists_equal_the_masks():
    """EnvArgs::lists (what the host-driven loop downloads instead of the masks): every game's decoded list -- live, idle or
    just reset -- is its legal mask's set bits in ascending order."""
    n = 45                                                # two groups, the second ragged
    e = EmuBatch(n, sched_seed=61)
    for step in range(40):
        if step % 7 == 3:                                 # some games idle (NOOP): listed again from their unchanged masks
            acts = np.full(n, -2, dtype=np.int32)
            for g in range(0, n, 3):
                la = e.actions(g)
                acts[g] = la[0] if len(la) else -1
            e.step(acts)
        elif step % 11 == 5:
            e.reset((np.arange(n) % 4 == 1).astype(np.uint8))
        else:
            e.step_random(0xC0FFEE, auto_reset=1)
        lists = e.legal_lists()
        for g in range(n):
            assert lists[g] is not None and lists[g] == e.actions(g).tolist(), (step, g)
