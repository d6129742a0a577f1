"""Parity tests proper: the sm_100a kernels, called through the C ABI (HiveBatch / GamePlay),
against the CPU oracle on identical seeded inputs and against the reference's golden vectors.
Bit-exact bar: legal sets, planes, turn/winner/done, state_key, transcripts."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hive_b200
    return hive_b200


def _u8_planes(bf16):
    return (bf16.astype(np.uint32) << 16).view(np.float32).astype(np.uint8)


def test_lockstep_256_games_vs_oracle(hb):
    from oracle.hive_oracle import OracleEnv
    n = 256
    b = hb.HiveBatch(n)
    oracles = [OracleEnv() for _ in range(n)]
    rngs = [np.random.RandomState(40000 + g) for g in range(n)]
    live = np.ones(n, dtype=bool)
    plies = 0
    while live.any():
        mask, count = b.legal_mask()
        bits = np.unpackbits(mask.view(np.uint8), axis=1, bitorder="little")[:, :1584]
        planes = _u8_planes(b.planes_bf16())
        turn, winner, done = b.status()
        actions = np.full(n, -2, dtype=np.int32)
        for g in range(n):
            if not live[g]:
                continue
            o = oracles[g]
            la = o.actions()
            assert np.nonzero(bits[g])[0].tolist() == la.tolist(), (g, o.turn)
            assert count[g] == len(la)
            assert (planes[g] == o.planes()).all(), (g, o.turn)
            od = o.game_is_over()
            assert turn[g] == o.turn and bool(done[g]) == od and winner[g] == o.winner
            if od or o.turn >= 55:
                live[g] = False
                assert b.state_key(g) == o.state_key
                continue
            a = int(la[rngs[g].randint(len(la))]) if len(la) else -1
            o.move(a)
            actions[g] = a
            plies += 1
        b.step(actions)
    assert plies > n * 40


def test_golden_replay_through_gameplay_facade(hb, golden):
    G = golden
    starts = G["game_start"]
    env = hb.GamePlay()
    for g in range(len(starts) - 1):
        env.new_game()
        for i in range(starts[g], starts[g + 1]):
            n = G["n_legal"][i]
            assert env.state.turn == G["turn"][i]
            assert env.actions() == G["legal"][i][:n].tolist()
            p = env.encode_board()
            assert p.shape == (12, 12, 56) and p.dtype == np.float64
            chw = p.transpose(2, 0, 1).reshape(56, 144)
            assert (chw[31] == G["plane31"][i]).all()
            chw = chw.astype(np.uint8)
            chw[31] = 0
            assert (np.packbits(chw, axis=1, bitorder="little") == G["planes"][i]).all()
            assert env.state_key == str(G["key"][i])
            assert env.game_is_over() == bool(G["done"][i])
            w = env.state.winner
            assert (0 if w is None else 1 if w == hb.config.PIECE_WHITE else 2) == G["winner"][i]
            if G["action"][i] != -2:
                env.move(int(G["action"][i]))


def test_known_answers_and_facade_contract(hb):
    env = hb.GamePlay(debug=True)
    assert env.state.turn == 1 and env.state.player() == 0
    assert env.actions() == [858, 859, 861, 863, 866]
    assert env.state_key == "." * 144 + "0"
    assert env.decode_action(858) == ("<class 'pieces.Queen'>0", ("N", "13"))
    with pytest.raises(ValueError):
        env.move(0)
    env.move(858)
    assert env.turn() == 2 and env.player() == 1
    assert env.actions() == [726, 727, 729, 731, 734]
    with pytest.raises(KeyError):
        env.encode_board("W")               # only the side to move is encoded (env_hive.py:313-318)
    twin = copy.deepcopy(env)
    env.move(726)
    assert twin.state.turn == 2 and twin.actions() == [726, 727, 729, 731, 734]
    assert env.state.turn == 3 and len(env.actions()) == 14
    assert env.state_key.replace(".", "") == "Q0q00"
    key_before = env.state_key
    env.skip_turn()                          # env_hive.py:493-496 leaves state_key untouched
    assert env.state.turn == 4 and env.state_key == key_before
    assert not env.game_is_over() and env.state.winner is None


def test_load_state_positions_from_golden(hb, golden):
    from oracle.hive_oracle import OracleEnv
    G = golden
    idx = np.argsort(-G["levels"].max(axis=1).astype(np.int32), kind="stable")[:64]            # the tallest stacks first
    b = hb.HiveBatch(len(idx))
    o = OracleEnv()
    assert G["levels"][idx[0]].max() >= 3
    for g, i in enumerate(idx):
        b.load_state(g, int(G["turn"][i]), G["cells"][i], G["levels"][i])
    acts = b.actions()
    planes = _u8_planes(b.planes_bf16())
    for g, i in enumerate(idx):
        n = G["n_legal"][i]
        assert acts[g].tolist() == G["legal"][i][:n].tolist()
        o.load(int(G["turn"][i]), G["cells"][i], G["levels"][i])     # history cleared in both
        assert (planes[g] == o.planes()).all()
        t, c, l = b.dump_state(g)
        assert t == G["turn"][i] and (c == G["cells"][i]).all() and (l == G["levels"][i]).all()
    with pytest.raises(hb.HiveError):
        bad = G["levels"][idx[0]].copy()
        bad[:] = 3
        b.load_state(0, 9, G["cells"][idx[0]], bad)


def test_step_random_full_size_16384(hb):
    """BASELINE config 2 size: on-device policy, auto-reset; a 512-game subsample is replayed by
    the oracle move for move, the whole batch is checked through size-independent properties."""
    import torch
    from oracle.hive_oracle import OracleEnv
    n, seed, steps, sub = 16384, 20261018, 130, 512
    b = hb.HiveBatch(n)
    chosen = torch.empty(n, dtype=torch.int32, device="cuda")
    oracles = [OracleEnv() for _ in range(sub)]
    episodes = [0] * sub
    total_steps = 0
    for it in range(steps):
        b.step_random(seed, max_turn=55, auto_reset=True, chosen_dev_ptr=chosen.data_ptr())
        b.sync()
        ch = chosen.cpu().numpy()
        total_steps += int((ch != -2).sum())
        for g in range(sub):
            o = oracles[g]
            if o.game_is_over() or o.turn >= 55:
                o.reset(); episodes[g] += 1
                assert ch[g] == -2
            else:
                a = o.pick_action(seed, g + n * episodes[g])
                assert ch[g] == a, (it, g)
                o.move(a)
    mask, count = b.legal_mask()
    planes = _u8_planes(b.planes_bf16())
    turn, winner, done = b.status()
    for g in range(sub):
        o = oracles[g]
        bits = np.unpackbits(mask[g].view(np.uint8), bitorder="little")[:1584]
        assert np.nonzero(bits)[0].tolist() == o.actions().tolist()
        assert (planes[g] == o.planes()).all()
        assert turn[g] == o.turn and bool(done[g]) == o.game_is_over()
    # properties over all 16,384 games
    pop = np.unpackbits(mask.view(np.uint8), axis=1, bitorder="little").sum(axis=1)
    assert (pop == count).all()
    assert (turn >= 1).all() and (turn <= 55).all()
    assert (planes[:, 30] == (planes[:, 11] | planes[:, 23])).all()
    assert (planes[:, 31] == turn[:, None]).all()
    assert (planes[:, 11] & planes[:, 23]).sum(axis=1).max() <= 4         # only stacks overlap
    steps_ctr, episodes_ctr = b.counters()
    assert int(steps_ctr.sum()) == total_steps
    assert (episodes_ctr[:sub] == np.array(episodes)).all()


def test_host_policy_twin_matches_device_policy(hb):
    """hive_host_pick_actions (e2e path with host buffers) == hive_step_random (resident path)."""
    n, seed = 1024, 77
    dev, host = hb.HiveBatch(n), hb.HiveBatch(n)
    episodes = np.zeros(n, dtype=np.uint32)
    actions = np.empty(n, dtype=np.int32)
    packed = np.empty(n, dtype=np.uint32)
    for _ in range(120):
        mask, count = host.legal_mask()
        host.status_packed_into(packed.ctypes.data)
        hb.host_pick_actions(mask, count, packed, episodes, seed, 55, actions)
        host.step(actions)
        dev.step_random(seed, max_turn=55, auto_reset=True)
    m1, c1 = dev.legal_mask()
    m2, c2 = host.legal_mask()
    assert (m1 == m2).all() and (c1 == c2).all()
    assert (dev.planes_bf16() == host.planes_bf16()).all()
    assert [a.tolist() for a in dev.status()] == [a.tolist() for a in host.status()]


def test_multi_step_graph_equals_single_steps(hb):
    """hive_step_random_multi (one CUDA graph, slices free-running across steps) == the same number of
    hive_step_random calls."""
    n, seed = 4096, 4242
    a, b = hb.HiveBatch(n), hb.HiveBatch(n)
    for _ in range(3):
        a.step_random_multi(seed, 37)
        for _ in range(37):
            b.step_random(seed, 55, True)
    m1, c1 = a.legal_mask()
    m2, c2 = b.legal_mask()
    assert (m1 == m2).all() and (c1 == c2).all()
    assert (a.planes_bf16() == b.planes_bf16()).all()
    assert [x.tolist() for x in a.status()] == [x.tolist() for x in b.status()]
    assert [x.tolist() for x in a.counters()] == [x.tolist() for x in b.counters()]


def test_planes_as_bits_expand_to_the_planes(hb):
    """hive_bits_host: the bit rows of the games the last launch evaluated expand to exactly their bf16 planes -- after single
    steps, multi-step graphs of odd and even length (the two bit-plane buffers), masked resets and with idle games."""
    from importlib import import_module
    expand = import_module("hive-alphazero_b200.env").bits_to_planes_bf16
    n, seed = 1024, 31337
    b = hb.HiveBatch(n)

    def check(evaluated=None):
        bits, planes = b.planes_bits(), b.planes_bf16()
        ok = bits[:, 31, 1] == 1
        if evaluated is not None:
            assert (ok == evaluated).all()
        assert ok.any() and (expand(bits[ok]) == planes[ok]).all()

    check(np.ones(n, dtype=bool))
    for steps in (1, 1, 6, 7, 2):
        if steps == 1:
            b.step_random(seed, 55, True)
        else:
            b.step_random_multi(seed, steps)
        check(np.ones(n, dtype=bool))
    mask = (np.arange(n) % 3 == 0).astype(np.uint8)
    b.reset(mask)
    check(mask.astype(bool))
    acts = np.full(n, -2, dtype=np.int32)                   # everybody idle but a few
    m, c = b.legal_mask()
    for g in range(0, n, 5):
        la = np.nonzero(np.unpackbits(m[g].view(np.uint8), bitorder="little")[:1584])[0]
        acts[g] = la[0] if len(la) else -1
    b.step(acts)
    check(acts != -2)


def test_delta_plane_store_equals_full_store(hb, monkeypatch):
    """HIVE_B200_DELTA_STORE=1 (hive_planes_delta_kernel: only the 32-byte sectors that differ from what the planes arena
    holds are rewritten) leaves the same planes as the default full store after single steps, multi-step graphs across
    resets, masked resets and hive_copy_state."""
    n, seed = 2048, 777
    a = hb.HiveBatch(n)
    monkeypatch.setenv("HIVE_B200_DELTA_STORE", "1")
    b = hb.HiveBatch(n)
    c = hb.HiveBatch(n)
    monkeypatch.delenv("HIVE_B200_DELTA_STORE")
    for rnd in range(3):
        for _ in range(5):
            a.step_random(seed, 55, True); b.step_random(seed, 55, True)
        assert (a.planes_bf16() == b.planes_bf16()).all()
        a.step_random_multi(seed, 41); b.step_random_multi(seed, 41)
        assert (a.planes_bf16() == b.planes_bf16()).all()
        mask = (np.arange(n) % 7 == rnd).astype(np.uint8)
        a.reset(mask); b.reset(mask)
        assert (a.planes_bf16() == b.planes_bf16()).all()
    # a copied game brings its planes and their bit image along: the next delta store of the copy starts from them
    for g in (0, 5, n - 1):
        c.copy_state_from(g, b, (g + 3) % n)
    for _ in range(4):
        c.step_random(seed, 55, False)
    ref = hb.HiveBatch(n)
    for g in (0, 5, n - 1):
        ref.copy_state_from(g, a, (g + 3) % n)
    for _ in range(4):
        ref.step_random(seed, 55, False)
    assert (c.planes_bf16() == ref.planes_bf16()).all()
    assert (a.legal_mask()[0] == b.legal_mask()[0]).all()


@pytest.mark.parametrize("mode", ["1", "2"])
def test_queue_rollout_equals_step_graphs(hb, monkeypatch, mode):
    """HIVE_B200_ROLLOUT_QUEUE=1 (two persistent kernels whose CTAs take (group, step) tickets) and =2 (one launch per
    step, chained by programmatic dependent launch + per-group flags) leave the batch where the default per-step graphs
    leave it; a rollout whose waits time out is reported by hive_sync instead of hanging the GPU."""
    n, seed = 4096 + 17, 991                              # a ragged last group
    a = hb.HiveBatch(n)
    monkeypatch.setenv("HIVE_B200_ROLLOUT_QUEUE", mode)
    b = hb.HiveBatch(n)
    monkeypatch.delenv("HIVE_B200_ROLLOUT_QUEUE")
    for steps in (2, 37, 60, 5):
        a.step_random_multi(seed, steps); b.step_random_multi(seed, steps)
        a.sync(); b.sync()
        m1, c1 = a.legal_mask()
        m2, c2 = b.legal_mask()
        assert (m1 == m2).all() and (c1 == c2).all(), steps
        assert (a.planes_bf16() == b.planes_bf16()).all(), steps
        assert [x.tolist() for x in a.status()] == [x.tolist() for x in b.status()], steps
        assert [x.tolist() for x in a.counters()] == [x.tolist() for x in b.counters()], steps


def test_async_host_step_graph_replay_matches_device_policy(hb):
    """hive_step_host_async from one fixed set of page-locked buffers (replayed as one CUDA graph from the second
    call on) == hive_step_random; the same loop from pageable buffers (plain path) gives the same games."""
    import torch
    n, seed = 4096, 99
    dev, pinned, pageable = hb.HiveBatch(n), hb.HiveBatch(n), hb.HiveBatch(n)

    def buffers(pin):
        t = [torch.empty((n, 25), dtype=torch.int64), torch.empty(n, dtype=torch.int32), torch.empty(n, dtype=torch.int32),
             torch.empty(n, dtype=torch.int32)]
        return [x.pin_memory() for x in t] if pin else t

    runs = []
    for b, pin in ((pinned, True), (pageable, False)):
        mask_h, count_h, status_h, actions_h = buffers(pin)
        mask, count = mask_h.numpy().view(np.uint64), count_h.numpy()
        status, actions = status_h.numpy().view(np.uint32), actions_h.numpy()
        episodes = np.zeros(n, dtype=np.uint32)
        b.legal_into(mask_h.data_ptr(), count_h.data_ptr())
        b.status_packed_into(status_h.data_ptr())
        for _ in range(70):
            b.wait_results()                                   # downloads landed; the plane store may still run
            hb.host_pick_actions(mask, count, status, episodes, seed, 55, actions)
            b.step_async_ptr(actions_h.data_ptr(), mask_h.data_ptr(), count_h.data_ptr(), status_h.data_ptr())
        b.sync()
        runs.append((mask.copy(), count.copy(), status.copy()))
    for _ in range(70):
        dev.step_random(seed, max_turn=55, auto_reset=True)
    m1, c1 = dev.legal_mask()
    for b, (mask, count, status) in zip((pinned, pageable), runs):
        m2, c2 = b.legal_mask()
        assert (m1 == m2).all() and (c1 == c2).all()
        assert (mask == m2).all() and (count == c2).all()          # what the step downloaded == the device arrays
        assert (dev.planes_bf16() == b.planes_bf16()).all()
        assert [a.tolist() for a in dev.status()] == [a.tolist() for a in b.status()]
        assert [x.tolist() for x in dev.counters()] == [x.tolist() for x in b.counters()]


def test_facade_attributes_callers_read():
    """human_play, white_pieces_set / black_pieces_set, board_matrix, history_white / history_black and the cached
    turn / state_key (env_hive.py:33-39,185-194; solo_play.py:127-134 reads the piece keys and board_matrix[..].core_index)."""
    import hive_b200 as hb
    from oracle.hive_oracle import OracleEnv
    env, o = hb.GamePlay(), OracleEnv()
    rng = np.random.RandomState(5)
    pushed = {0: [], 1: []}
    for ply in range(14):
        side = env.state.player()
        pushed[side].insert(0, env.encode_board()[:, :, [11, 23]].copy())       # what this evaluation pushed (env_hive.py:436-445)
        assert env.state.turn == o.turn and env.turn() == o.turn
        hist = env.history_white if side == 0 else env.history_black
        assert len(hist) == min(len(pushed[side]), 4)
        for a, b in zip(hist, pushed[side]):
            assert (a == b).all()
        la = env.actions()
        a = int(la[rng.randint(len(la))])
        env.move(a); o.move(a)
    keys = list(env.white_pieces_set.keys())
    assert keys[0] == "<class 'pieces.Queen'>0" and keys[10] == "<class 'pieces.Ant'>2" and len(keys) == 11
    _, cells, levels = env.position()
    for color, pset in enumerate((env.white_pieces_set, env.black_pieces_set)):
        for k, (key, (tile, level, name)) in enumerate(pset.items()):
            c = int(cells[color * 11 + k])
            assert level == int(levels[color * 11 + k])
            if c == 255:
                assert tile.axial_coords == (99, 99)
            else:
                assert tile is env.board_matrix[c // 12, c % 12] and tile.core_index == (hb.config.index_char[c // 12], hb.config.index_number[c % 12])
    key = env.state_key
    env.human_play()                                            # nothing moved by hand: same position, same outputs
    assert env.state_key == key and env.actions() == o.actions().tolist()
    # public search entry points of the player (solo_play.py:153-165,351-374)
    from oracle.mcts_oracle import hash_net
    pl = hb.HivePlayer()
    pl.simulation_num_per_move = 12
    pl.expand_and_evaluate_with_net = lambda e: hash_net(e.encode_board())
    np.random.seed(3)
    pl.search_moves(env)
    policy, sum_all = pl.calc_policy(env)
    np.random.seed(3)
    _, (policy2, sum_all2) = pl.action(env)
    assert (policy == np.asarray(policy2)).all() and sum_all == sum_all2 == 11.0
