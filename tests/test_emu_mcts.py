"""The PUCT search kernels (verbatim device source on the CPU SIMT emulator) against the
reference's own HivePlayer results (tests/golden/mcts_cases.npz) and the Python MCTS oracle.
Bar: per-edge N, W, Q, P equal as floats, node count, returned move and policy identical."""
import os

import numpy as np
import pytest

from oracle.hive_oracle import OracleEnv
from oracle.mcts_oracle import MctsOracle, hash_net
from tests.emu.emu import EmuBatch, EmuMcts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "mcts_cases.npz"))


def _recorded_noise(seed, sims, k, rows=None):
    np.random.seed(seed)
    noise = np.zeros((1, sims, 256))
    for r in range(sims - 1):
        noise[0, r, :k] = np.random.dirichlet([0.3] * k)
    return noise


@pytest.mark.parametrize("i", range(len(G["seed"])))
def test_oracle_matches_reference_player(i):
    env = OracleEnv()
    for a in G["prefix"][i][:G["n_prefix"][i]]:
        env.move(int(a))
    m = MctsOracle(hash_net, int(G["sims"][i]))
    np.random.seed(int(G["seed"][i]))
    action, policy, sum_all = m.action(env)
    acts, n, w, q, p, sum_n, n_nodes = m.root_stats(env)
    k = G["n_edges"][i]
    assert acts.tolist() == G["e_action"][i][:k].tolist() and n.tolist() == G["e_n"][i][:k].tolist()
    assert (w == G["e_w"][i][:k]).all() and (q == G["e_q"][i][:k]).all() and (p == G["e_p"][i][:k]).all()
    assert sum_n == G["sum_n"][i] and n_nodes == G["n_nodes"][i] and action == G["action"][i]
    assert (policy == G["policy"][i]).all() and sum_all == G["sum_all"][i]


@pytest.mark.parametrize("i", [0, 1, 4, 7, 8, 12, 13])
def test_kernels_match_reference_player(i):
    b = EmuBatch(1, sched_seed=i)
    for a in G["prefix"][i][:G["n_prefix"][i]]:
        b.step(np.array([a], dtype=np.int32))
    sims, k = int(G["sims"][i]), int(G["n_edges"][i])
    m = EmuMcts(b, sims)
    m.set_noise(_recorded_noise(int(G["seed"][i]), sims, k))
    m.search(hash_net)
    st = m.root_stats(0)
    pi, action, sum_n = m.policy()
    assert st["error"] == 0 and st["sims_done"] == sims
    assert st["action"].tolist() == G["e_action"][i][:k].tolist() and st["n"].tolist() == G["e_n"][i][:k].tolist()
    assert (st["w"] == G["e_w"][i][:k]).all() and (st["q"] == G["e_q"][i][:k]).all() and (st["p"] == G["e_p"][i][:k]).all()
    assert st["sum_n"] == G["sum_n"][i] and st["n_nodes"] == G["n_nodes"][i]
    assert action[0] == G["action"][i] and (pi[0] == G["policy"][i]).all() and sum_n[0] == G["sum_n"][i]


@pytest.mark.parametrize("full_store", [True, False])
def test_three_trees_in_one_batch(full_store):
    """Independent trees in one launch (different positions, separate noise) vs the oracle.  full_store=False: the
    delta plane store (the search copies a root's planes AND their shadow row into its working batch)."""
    n, sims = 3, 24
    b = EmuBatch(n, sched_seed=77, full_store=full_store)
    envs = [OracleEnv() for _ in range(n)]
    rng = np.random.RandomState(3)
    for ply in range(9):
        acts = np.zeros(n, dtype=np.int32)
        for t, e in enumerate(envs):
            la = e.actions()
            acts[t] = la[rng.randint(len(la))] if (len(la) and ply < 3 + 3 * t) else -2
            if acts[t] != -2:
                e.move(int(acts[t]))
        b.step(acts)
    noise = np.zeros((n, sims, 256))
    expect = []
    for t, e in enumerate(envs):
        k = len(e.actions())
        np.random.seed(100 + t)
        o = MctsOracle(hash_net, sims)
        o.action(e)
        expect.append(o.root_stats(e))
        for r, row in enumerate(o.noise_log):
            noise[t, r, :k] = row
    m = EmuMcts(b, sims)
    m.set_noise(noise)
    m.search(hash_net)
    for t in range(n):
        st = m.root_stats(t)
        acts, nn, w, q, p, sum_n, n_nodes = expect[t]
        assert st["action"].tolist() == acts.tolist() and st["n"].tolist() == nn.tolist()
        assert (st["w"] == w).all() and (st["p"] == p).all() and st["n_nodes"] == n_nodes


# ---- searches from injected positions (pass-edge roots, tall stacks): reference results in mcts_injected.npz
GI = np.load(os.path.join(ROOT, "tests", "golden", "mcts_injected.npz"))
EP = np.load(os.path.join(ROOT, "tests", "golden", "edge_positions.npz"))


def _check_injected(j, st, pi, action, sum_n):
    k = int(GI["n_edges"][j])
    assert st["action"].tolist() == GI["e_action"][j][:k].tolist() and st["n"].tolist() == GI["e_n"][j][:k].tolist()
    assert (st["w"] == GI["e_w"][j][:k]).all() and (st["p"] == GI["e_p"][j][:k]).all()
    assert st["sum_n"] == GI["sum_n"][j] and st["n_nodes"] == GI["n_nodes"][j]
    assert (pi == GI["policy"][j]).all() and int(np.argmax(pi)) == GI["action"][j] and sum_n == GI["sum_n"][j]


def test_fixture_has_pass_edge_roots():
    assert int((GI["e_action"][:, 0] == -1).sum()) >= 1


@pytest.mark.parametrize("j", range(len(GI["seed"])))
def test_oracle_on_injected_roots(j):
    i = int(GI["edge_index"][j])
    env = OracleEnv()
    env.load(int(EP["turn"][i]), EP["cells"][i], EP["levels"][i])
    m = MctsOracle(hash_net, int(GI["sims"][j]))
    np.random.seed(int(GI["seed"][j]))
    action, policy, sum_all = m.action(env)
    acts, n, w, q, p, sum_n, n_nodes = m.root_stats(env)
    _check_injected(j, dict(action=acts, n=n, w=w, p=p, sum_n=sum_n, n_nodes=n_nodes), policy, action, sum_n)
    assert action == GI["action"][j]


@pytest.mark.parametrize("j", range(len(GI["seed"])))
def test_kernels_on_injected_roots(j):
    i = int(GI["edge_index"][j])
    b = EmuBatch(1, sched_seed=40 + j)
    b.load(0, int(EP["turn"][i]), EP["cells"][i], EP["levels"][i])
    sims, k = int(GI["sims"][j]), max(int(EP["n_legal"][i]), 1)
    m = EmuMcts(b, sims)
    noise = np.zeros((1, sims, 256))
    if EP["n_legal"][i] > 0:
        np.random.seed(int(GI["seed"][j]))
        for r in range(sims - 1):
            noise[0, r, :k] = np.random.dirichlet([0.3] * k)
    m.set_noise(noise)
    m.search(hash_net)
    st = m.root_stats(0)
    pi, action, sum_n = m.policy()
    assert st["error"] == 0
    _check_injected(j, st, pi[0], action[0], sum_n[0])


# ---- robustness: a tree that outgrows its arena must stop cleanly and be flagged (never a silent short search)
def test_edge_arena_overflow_is_flagged_and_leaves_clean_statistics():
    b = EmuBatch(2, sched_seed=5)
    rng = np.random.RandomState(4)
    env = OracleEnv()
    for _ in range(14):
        la = env.actions()
        a = int(la[rng.randint(len(la))])
        env.move(a)
        b.step(np.array([a, a], dtype=np.int32))
    sims = 24
    m = EmuMcts(b, sims, edges_per_sim=1)              # 24 + 256 edges: the arena fills after a few expansions
    m.set_noise(np.full((2, sims, 256), 1.0 / 64))
    m.search(hash_net)
    assert m.errors() & 4                               # edge arena flagged for the batch
    for t in range(2):
        st = m.root_stats(t)
        assert st["error"] == 2 and st["sims_done"] < sims
        # the aborted simulation took its virtual losses back: counts are those of the finished simulations only
        assert st["n"].sum() == st["sum_n"] == st["sims_done"] - 1 and (st["n"] >= 0).all()
        assert np.isfinite(st["w"]).all() and (np.abs(st["w"]) <= st["n"] + 1e-9).all()
        live = st["n"] > 0
        assert (st["q"][live] == st["w"][live] / st["n"][live]).all() and (st["q"][~live] == 0).all()


def test_device_noise_differs_between_searches_of_one_slot():
    """Root Dirichlet noise generated in the kernel: a fresh stream per search of the same slot (the reference draws
    fresh np.random.dirichlet rows every time), identical for identical (seed, slot, search number)."""
    b = EmuBatch(1, sched_seed=9)
    rng = np.random.RandomState(8)
    env = OracleEnv()
    for _ in range(9):
        la = env.actions()
        a = int(la[rng.randint(len(la))])
        env.move(a)
        b.step(np.array([a], dtype=np.int32))
    flat = lambda planes: (np.full(1584, 1.0 / 1584, dtype=np.float32), 0.0)      # flat net: the noise decides
    m = EmuMcts(b, 40)
    m.search(flat)
    n1 = m.root_stats(0)["n"].copy()
    m.search(flat)
    n2 = m.root_stats(0)["n"].copy()
    assert n1.sum() == n2.sum() == 39 and (n1 != n2).any()
    m2 = EmuMcts(b, 40)                                 # a new handle restarts the slot's stream
    m2.search(flat)
    assert (m2.root_stats(0)["n"] == n1).all()


# ---- 250 / 500 simulations (BASELINE configs[3] / [4] depths) run by the real reference player: mcts_deep.npz
GD = np.load(os.path.join(ROOT, "tests", "golden", "mcts_deep.npz"))


def _check_deep(i, st, pi, action, sum_n):
    k = int(GD["n_edges"][i])
    assert st["action"].tolist() == GD["e_action"][i][:k].tolist() and st["n"].tolist() == GD["e_n"][i][:k].tolist()
    assert (st["w"] == GD["e_w"][i][:k]).all() and (st["q"] == GD["e_q"][i][:k]).all() and (st["p"] == GD["e_p"][i][:k]).all()
    assert st["sum_n"] == GD["sum_n"][i] and st["n_nodes"] == GD["n_nodes"][i]
    assert action == GD["action"][i] and (pi == GD["policy"][i]).all() and sum_n == GD["sum_n"][i]


@pytest.mark.parametrize("i", range(len(GD["seed"])))
def test_oracle_matches_reference_player_deep(i):
    assert int(GD["sims"][i]) in (250, 500)
    env = OracleEnv()
    for a in GD["prefix"][i][:GD["n_prefix"][i]]:
        env.move(int(a))
    m = MctsOracle(hash_net, int(GD["sims"][i]))
    np.random.seed(int(GD["seed"][i]))
    action, policy, sum_all = m.action(env)
    acts, n, w, q, p, sum_n, n_nodes = m.root_stats(env)
    _check_deep(i, dict(action=acts, n=n, w=w, q=q, p=p, sum_n=sum_n, n_nodes=n_nodes), policy, action, sum_n)


@pytest.mark.parametrize("i", [1, 3])
def test_kernels_match_reference_player_deep(i):
    b = EmuBatch(1, sched_seed=200 + i)
    for a in GD["prefix"][i][:GD["n_prefix"][i]]:
        b.step(np.array([a], dtype=np.int32))
    sims, k = int(GD["sims"][i]), int(GD["n_edges"][i])
    m = EmuMcts(b, sims, edges_per_sim=160)
    m.set_noise(_recorded_noise(int(GD["seed"][i]), sims, k))
    m.search(hash_net)
    pi, action, sum_n = m.policy()
    assert m.errors() == 0
    _check_deep(i, m.root_stats(0), pi[0], action[0], sum_n[0])
