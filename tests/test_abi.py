"""The C-ABI library loads and exports every symbol include/hive_b200.h declares."""
import ctypes
import os
import re

import pytest

import hive_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "hive_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b((?:hive|mcts|net)_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_builds_and_exports_header_symbols():
    path = hive_b200.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), "missing export: " + name
    for name in hive_b200.ENV_SYMBOLS:
        assert name in declared


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hive_b200.HiveError) as e:
        hive_b200.HiveBatch(4)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "hive-alphazero_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "hive_oracle" not in text.replace(
                    "oracle/hive_oracle.c", ""), f


def test_host_policy_twin_pure_host_code():
    """hive_host_pick_actions needs no GPU: check it (and its worker pool) against the rule
    a = A[splitmix64(seed ^ gid<<32 ^ turn) % len(A)] computed with the oracle's splitmix64."""
    import numpy as np
    from oracle.hive_oracle import splitmix64
    n, seed, max_turn = 5000, 99, 55
    rng = np.random.RandomState(0)
    bits = (rng.rand(n, 1584) < 0.03)
    bits[::7] = False                                        # some games without legal actions
    mask = np.packbits(np.pad(bits, ((0, 0), (0, 16))), axis=1, bitorder="little").view(np.uint64).copy()
    count = bits.sum(axis=1).astype(np.int32)
    turn = rng.randint(1, 60, size=n).astype(np.uint32)
    done = (rng.rand(n) < 0.05).astype(np.uint32)
    packed = (turn | (done << 16)).astype(np.uint32)
    episodes = rng.randint(0, 5, size=n).astype(np.uint32)
    ep0 = episodes.copy()
    actions = np.empty(n, dtype=np.int32)
    hive_b200.host_pick_actions(mask, count, packed, episodes, seed, max_turn, actions)
    for g in range(n):
        if done[g] or turn[g] >= max_turn:
            assert actions[g] == -3 and episodes[g] == ep0[g] + 1
        elif count[g] == 0:
            assert actions[g] == -1 and episodes[g] == ep0[g]
        else:
            legal = np.nonzero(bits[g])[0]
            h = splitmix64(seed ^ ((g + n * int(ep0[g])) << 32) ^ int(turn[g]))
            assert actions[g] == legal[h % len(legal)]
