"""PUCT search on the real device through the C ABI (MctsBatch / HivePlayer) against the
reference's own results (golden) and the oracle.  Visit counts bit-exact given identical network
outputs (the deterministic hash-net) and recorded root noise."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "mcts_cases.npz"))


@pytest.fixture(scope="module")
def hb():
    import hive_b200
    return hive_b200


@pytest.mark.parametrize("i", range(len(G["seed"])))
def test_hiveplayer_facade_reproduces_reference(hb, i):
    """Same call sequence as the reference: np.random.seed(s); HivePlayer.action(env)."""
    from oracle.mcts_oracle import hash_net
    env = hb.GamePlay()
    for a in G["prefix"][i][:G["n_prefix"][i]]:
        env.move(int(a))
    pl = hb.HivePlayer()
    pl.none_queue = False
    pl.simulation_num_per_move = int(G["sims"][i])
    pl.expand_and_evaluate_with_net = lambda e: hash_net(e.encode_board())
    np.random.seed(int(G["seed"][i]))
    action, (policy, sum_all) = pl.action(env)
    st = pl._mcts.root_stats(0)
    k = G["n_edges"][i]
    assert st["error"] == 0
    assert st["action"].tolist() == G["e_action"][i][:k].tolist() and st["n"].tolist() == G["e_n"][i][:k].tolist()
    assert (st["w"] == G["e_w"][i][:k]).all() and (st["q"] == G["e_q"][i][:k]).all() and (st["p"] == G["e_p"][i][:k]).all()
    assert st["sum_n"] == G["sum_n"][i] and st["n_nodes"] == G["n_nodes"][i]
    assert action == G["action"][i] and (np.array(policy) == G["policy"][i]).all() and sum_all == G["sum_all"][i]


def test_batch_of_trees_vs_oracle(hb):
    from oracle.hive_oracle import OracleEnv
    from oracle.mcts_oracle import MctsOracle, hash_net
    n, sims = 24, 48
    b = hb.HiveBatch(n)
    envs = [OracleEnv() for _ in range(n)]
    rng = np.random.RandomState(11)
    for ply in range(40):
        acts = np.full(n, -2, dtype=np.int32)
        for t, e in enumerate(envs):
            la = e.actions()
            if ply < 2 + (t * 37) % 39 and not e.game_is_over():
                acts[t] = la[rng.randint(len(la))] if len(la) else -1
                e.move(int(acts[t]))
        b.step(acts)
    noise = np.zeros((n, sims, 256))
    expect = []
    for t, e in enumerate(envs):
        np.random.seed(500 + t)
        o = MctsOracle(hash_net, sims)
        o.action(e)
        expect.append(o.root_stats(e))
        for r, row in enumerate(o.noise_log):
            noise[t, r, :len(row)] = row
    m = hb.MctsBatch(b, sims)
    m.set_root_noise(noise)
    m.search_host(lambda leaf: hash_net(leaf.encode_board()))
    pi, action, sum_n = m.policy()
    for t in range(n):
        st = m.root_stats(t)
        acts, nn, w, q, p, s_n, n_nodes = expect[t]
        assert st["error"] == 0 and st["sims_done"] == sims
        assert st["action"].tolist() == acts.tolist() and st["n"].tolist() == nn.tolist(), t
        assert (st["w"] == w).all() and (st["q"] == q).all() and (st["p"] == p).all()
        assert st["sum_n"] == s_n and st["n_nodes"] == n_nodes and sum_n[t] == s_n


def test_device_noise_search_runs_and_is_consistent(hb):
    """On-device Dirichlet sampling: no reference stream to match, so check invariants."""
    from oracle.mcts_oracle import hash_net
    n, sims = 64, 32
    b = hb.HiveBatch(n)
    for _ in range(12):
        b.step_random(99, 55, True)
    m = hb.MctsBatch(b, sims)
    m.set_root_noise(None)
    m.search_host(lambda leaf: hash_net(leaf.encode_board()))
    pi, action, sum_n = m.policy()
    legal = b.actions()
    for t in range(n):
        st = m.root_stats(t)
        assert st["error"] == 0 and st["sims_done"] == sims and st["sum_n"] == sims - 1
        assert st["action"].tolist() == legal[t].tolist()            # root edges == legal actions
        assert st["n"].sum() == sims - 1 and abs(pi[t].sum() - 1.0) < 1e-9
        assert action[t] in legal[t]


GI = np.load(os.path.join(ROOT, "tests", "golden", "mcts_injected.npz"))
EP = np.load(os.path.join(ROOT, "tests", "golden", "edge_positions.npz"))


@pytest.mark.parametrize("j", range(len(GI["seed"])))
def test_facade_on_injected_roots_incl_pass_edge(hb, j):
    """Roots without legal actions (pass edge -1, policy[-1] quirk) and tall-stack roots, searched by the real
    reference from injected positions."""
    from oracle.mcts_oracle import hash_net
    i = int(GI["edge_index"][j])
    env = hb.GamePlay()
    env.load_position(int(EP["turn"][i]), EP["cells"][i], EP["levels"][i])
    pl = hb.HivePlayer()
    pl.none_queue = False
    pl.simulation_num_per_move = int(GI["sims"][j])
    pl.expand_and_evaluate_with_net = lambda e: hash_net(e.encode_board())
    np.random.seed(int(GI["seed"][j]))
    action, (policy, sum_all) = pl.action(env)
    st = pl._mcts.root_stats(0)
    k = int(GI["n_edges"][j])
    assert st["error"] == 0 and st["action"].tolist() == GI["e_action"][j][:k].tolist()
    assert st["n"].tolist() == GI["e_n"][j][:k].tolist() and (st["w"] == GI["e_w"][j][:k]).all() and (st["p"] == GI["e_p"][j][:k]).all()
    assert st["n_nodes"] == GI["n_nodes"][j] and action == GI["action"][j] and (np.array(policy) == GI["policy"][j]).all()


GD = np.load(os.path.join(ROOT, "tests", "golden", "mcts_deep.npz"))


@pytest.mark.parametrize("i", range(len(GD["seed"])))
def test_hiveplayer_facade_reproduces_reference_at_250_and_500_sims(hb, i):
    """BASELINE configs[3] / [4] depths: searches of 250 and 500 simulations by the real reference player."""
    from oracle.mcts_oracle import hash_net
    env = hb.GamePlay()
    for a in GD["prefix"][i][:GD["n_prefix"][i]]:
        env.move(int(a))
    pl = hb.HivePlayer()
    pl.none_queue = False
    pl.simulation_num_per_move = int(GD["sims"][i])
    pl.expand_and_evaluate_with_net = lambda e: hash_net(e.encode_board())
    np.random.seed(int(GD["seed"][i]))
    action, (policy, sum_all) = pl.action(env)
    st = pl._mcts.root_stats(0)
    k = GD["n_edges"][i]
    assert st["error"] == 0 and st["sims_done"] == GD["sims"][i]
    assert st["action"].tolist() == GD["e_action"][i][:k].tolist() and st["n"].tolist() == GD["e_n"][i][:k].tolist()
    assert (st["w"] == GD["e_w"][i][:k]).all() and (st["q"] == GD["e_q"][i][:k]).all() and (st["p"] == GD["e_p"][i][:k]).all()
    assert st["sum_n"] == GD["sum_n"][i] and st["n_nodes"] == GD["n_nodes"][i]
    assert action == GD["action"][i] and (np.array(policy) == GD["policy"][i]).all() and sum_all == GD["sum_all"][i]
