"""The DEVICE leaf-evaluation path of the search -- MctsBatch.search_device, plain and replayed through WaveGraph,
LeafEvaluator-style callbacks writing into mcts_dev_leaf_policy / mcts_dev_leaf_value, SplitEvaluator's pointer
arithmetic -- against the sequential oracle of the reference's HivePlayer (woker/solo_play.py:167-291).

The evaluator is the library's deterministic stand-in network on the device (mcts_hash_eval_dev; NumPy twin
oracle.mcts_oracle.device_hash_net), root noise is recorded from the oracle's own np.random.dirichlet draws.
Bar: per-edge N, W, Q, P equal as floats, node counts and sum N equal, for every tree."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hive_b200
    return hive_b200


def _positions(hb, n, seed, stream=None, max_ply=40):
    """n games at different depths (2..40 plies of random play), on the device and in the oracle.  A search from a
    finished game has no root node (in the reference too: HivePlayer.action raises), so a draw of games in which one
    ended is thrown away and the next seed is tried -- deterministic, the oracle decides."""
    from oracle.hive_oracle import OracleEnv
    for attempt in range(50):
        envs = [OracleEnv() for _ in range(n)]
        rng = np.random.RandomState(seed + 1000 * attempt)
        plies = []
        for ply in range(max_ply):
            acts = np.full(n, -2, dtype=np.int32)
            for t, e in enumerate(envs):
                la = e.actions()
                if ply < 2 + (t * 37) % (max_ply - 1) and not e.game_is_over():
                    acts[t] = la[rng.randint(len(la))] if len(la) else -1
                    e.move(int(acts[t]))
            plies.append(acts)
        if not any(e.game_is_over() for e in envs):
            break
    else:
        raise AssertionError("no draw without a finished game")
    b = hb.HiveBatch(n, stream=stream)
    for acts in plies:
        b.step(acts)
    return b, envs


def _oracle_expectation(envs, sims, nets, seed0):
    from oracle.mcts_oracle import MctsOracle
    noise = np.zeros((len(envs), sims, 256))
    expect = []
    for t, e in enumerate(envs):
        np.random.seed(seed0 + t)
        o = MctsOracle(nets[t], sims)
        o.action(e)
        expect.append(o.root_stats(e))
        for r, row in enumerate(o.noise_log):
            noise[t, r, :len(row)] = row
    return noise, expect


def _compare(m, expect, sims):
    pi, action, sum_n = m.policy()
    for t, (acts, nn, w, q, p, s_n, n_nodes) in enumerate(expect):
        st = m.root_stats(t)
        assert st["error"] == 0 and st["sims_done"] == sims, t
        assert st["action"].tolist() == acts.tolist() and st["n"].tolist() == nn.tolist(), t
        assert (st["w"] == w).all() and (st["q"] == q).all() and (st["p"] == p).all(), t
        assert st["sum_n"] == s_n and st["n_nodes"] == n_nodes and sum_n[t] == s_n, t


def test_search_device_64_trees_50_sims_plain_and_graph(hb):
    import torch
    from oracle.mcts_oracle import device_hash_net
    n, sims, salt = 64, 50, 0xA5
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        b, envs = _positions(hb, n, 21, stream=stream.cuda_stream)
        noise, expect = _oracle_expectation(envs, sims, [lambda pl: device_hash_net(pl, salt)] * n, 700)
        m = hb.MctsBatch(b, sims)
        m.set_root_noise(noise)
        ev = hb.HashEvaluator(salt, m.stream_ptr)
        waves = m.search_device(ev)
        assert waves >= sims and m.errors() == 0
        _compare(m, expect, sims)
        # the same search replayed from one captured wave (what self-play and bench.py run)
        g = hb.WaveGraph(stream)
        m.search_device(ev, graph=g)
        _compare(m, expect, sims)
        m.search_device(ev, graph=g)                       # the captured graph stays valid for the next search
        _compare(m, expect, sims)
        # masked search: the trees outside the mask keep their statistics, the others are searched again
        mask = (np.arange(n) % 2 == 0).astype(np.uint8)
        m.search_device(ev, tree_mask=mask)
        _compare(m, expect, sims)


def test_search_device_with_delta_plane_store(hb, monkeypatch):
    """The same comparison with HIVE_B200_DELTA_STORE=1: the search copies a root's planes and their shadow row into its
    working batch, and every leaf evaluation rewrites only the sectors that differ from the slot's previous leaf."""
    from oracle.mcts_oracle import device_hash_net
    monkeypatch.setenv("HIVE_B200_DELTA_STORE", "1")
    n, sims, salt = 16, 50, 0x3C
    b, envs = _positions(hb, n, 77)
    noise, expect = _oracle_expectation(envs, sims, [lambda pl: device_hash_net(pl, salt)] * n, 900)
    m = hb.MctsBatch(b, sims)
    m.set_root_noise(noise)
    m.search_device(hb.HashEvaluator(salt, m.stream_ptr))
    assert m.errors() == 0
    _compare(m, expect, sims)
    m.search_device(hb.HashEvaluator(salt, m.stream_ptr))          # again from the same roots: same trees
    _compare(m, expect, sims)


@pytest.mark.parametrize("sims", [250, 500])
def test_search_device_8_trees_deep(hb, sims):
    from oracle.mcts_oracle import device_hash_net
    n, salt = 8, 0x51 + sims
    b, envs = _positions(hb, n, 5 + sims, max_ply=36)
    noise, expect = _oracle_expectation(envs, sims, [lambda pl: device_hash_net(pl, salt)] * n, 1300 + sims)
    m = hb.MctsBatch(b, sims)
    m.set_root_noise(noise)
    m.search_device(hb.HashEvaluator(salt, m.stream_ptr))
    assert m.errors() == 0
    _compare(m, expect, sims)
    assert max(e[6] for e in expect) > sims // 3           # the trees really grow to hundreds of nodes


def test_split_evaluator_two_networks(hb):
    """Rows [0, k) by one network, rows [k, n) by another (the evaluator match): a wrong row offset in
    SplitEvaluator would hand a tree the other network's numbers."""
    from oracle.mcts_oracle import device_hash_net
    n, k, sims = 24, 10, 40
    b, envs = _positions(hb, n, 77)
    nets = [(lambda pl: device_hash_net(pl, 1111)) if t < k else (lambda pl: device_hash_net(pl, 2222)) for t in range(n)]
    noise, expect = _oracle_expectation(envs, sims, nets, 4200)
    m = hb.MctsBatch(b, sims)
    m.set_root_noise(noise)
    ev = hb.SplitEvaluator(hb.HashEvaluator(1111, m.stream_ptr), hb.HashEvaluator(2222, m.stream_ptr), k)
    m.search_device(ev)
    _compare(m, expect, sims)
    # and the two networks do differ on these positions
    p1, _ = device_hash_net(envs[0].encode_board(), 1111)
    p2, _ = device_hash_net(envs[0].encode_board(), 2222)
    assert (p1 != p2).any()


def test_hash_evaluator_matches_numpy_twin_and_respects_mask(hb):
    import torch
    from oracle.mcts_oracle import device_hash_net
    n = 16
    b, envs = _positions(hb, n, 3)
    b.sync()
    policy = torch.full((n, 1584), -1.0, dtype=torch.float32, device="cuda")
    value = torch.full((n,), -7.0, dtype=torch.float64, device="cuda")
    mask = torch.tensor([t % 3 != 0 for t in range(n)], dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    hb.HashEvaluator(99, s)(b.dev_planes, policy.data_ptr(), value.data_ptr(), mask.data_ptr(), n)
    torch.cuda.synchronize()
    for t, e in enumerate(envs):
        if t % 3 == 0:
            assert float(value[t]) == -7.0 and float(policy[t].max()) == -1.0
        else:
            p, v = device_hash_net(e.encode_board(), 99)
            assert (policy[t].cpu().numpy() == p).all() and float(value[t]) == v


def test_arena_overflow_raises(hb):
    """A tree that outgrows its edge arena stops and the read-back of the search fails loudly."""
    n, sims = 8, 32
    b, envs = _positions(hb, n, 9, max_ply=24)
    m = hb.MctsBatch(b, sims, edges_per_sim=1)
    m.set_root_noise(np.full((n, sims, 256), 1.0 / 64))
    m.search_device(hb.HashEvaluator(5, m.stream_ptr))
    assert m.errors() & 4
    with pytest.raises(hb.SearchError):
        m.actions()
    with pytest.raises(hb.SearchError):
        m.policy()
    for t in range(n):
        st = m.root_stats(t)
        if st["error"]:
            assert st["n"].sum() == st["sum_n"] == st["sims_done"] - 1 and (st["n"] >= 0).all()
    # a roomy arena on the same positions: no flag
    m2 = hb.MctsBatch(b, sims)
    m2.set_root_noise(np.full((n, sims, 256), 1.0 / 64))
    m2.search_device(hb.HashEvaluator(5, m2.stream_ptr))
    assert m2.errors() == 0
    m2.policy()


def test_device_noise_fresh_per_search_also_under_graph_replay(hb):
    """On-device Dirichlet rows: two consecutive searches from the same slots must not replay the same noise, also when
    the wave is a replayed CUDA graph (the stream position is read from device memory, not baked into the graph)."""
    import torch
    n, sims = 32, 40
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        b, envs = _positions(hb, n, 31, stream=stream.cuda_stream, max_ply=20)
        m = hb.MctsBatch(b, sims)
        m.set_root_noise(None)
        ev = hb.HashEvaluator(8, m.stream_ptr)
        g = hb.WaveGraph(stream)
        runs = []
        for _ in range(3):
            m.search_device(ev, graph=g)
            runs.append([m.root_stats(t)["n"].copy() for t in range(n)])
        for t in range(n):
            assert runs[0][t].sum() == runs[1][t].sum() == sims - 1
        differ01 = sum(int((runs[0][t] != runs[1][t]).any()) for t in range(n))
        differ12 = sum(int((runs[1][t] != runs[2][t]).any()) for t in range(n))
        assert differ01 >= n // 2 and differ12 >= n // 2
