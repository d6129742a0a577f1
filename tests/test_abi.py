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


def test_host_policy_twin_from_compact_lists():
    """hive_host_pick_actions_lists (the k-th legal action read from the 96-byte-per-game compact lists that the host-driven
    loop downloads) returns exactly the actions of the mask scan; games of a group flagged as overflow are left untouched."""
    import ctypes
    import numpy as np
    from importlib import import_module
    L = import_module("hive-alphazero_b200._capi").lib()
    n, seed, max_turn = 1000, 11, 55                       # 32 groups, the last one ragged
    rng = np.random.RandomState(5)
    bits = rng.rand(n, 1584) < 0.035
    bits[::9] = False
    mask = np.packbits(np.pad(bits, ((0, 0), (0, 16))), axis=1, bitorder="little").view(np.uint64).copy()
    count = bits.sum(axis=1).astype(np.int32)
    packed = (rng.randint(1, 60, size=n).astype(np.uint32) | ((rng.rand(n) < 0.05).astype(np.uint32) << 16)).astype(np.uint32)
    # the device's encoding, restated: per group 32 headers of 12 bytes, then one byte (id & 255) per action
    nb = (n + 31) // 32
    lists = np.zeros((nb, 3072), dtype=np.uint8)
    flagged = np.zeros(n, dtype=bool)
    for b in range(nb):
        off = 0
        over = b == 3                                       # pretend group 3 did not fit
        for lane in range(32):
            g = b * 32 + lane
            ids = np.nonzero(bits[g])[0] if g < n else np.zeros(0, dtype=np.int64)
            hdr = lists[b, lane * 12:lane * 12 + 12]
            hdr[0], hdr[1] = off & 255, off >> 8
            for p in range(7):
                hdr[2 + p] = min(int((ids < 256 * (p + 1)).sum()), 255)
            hdr[9] = 1 if over else 0
            if not over:
                lists[b, 384 + off:384 + off + len(ids)] = ids & 255
            if g < n:
                flagged[g] = over
            off += len(ids)
        assert off <= 3072 - 384
    ep_a = rng.randint(0, 5, size=n).astype(np.uint32)
    ep_b = ep_a.copy()
    a = np.empty(n, dtype=np.int32)
    b_ = np.full(n, -77, dtype=np.int32)
    hive_b200.host_pick_actions(mask, count, packed, ep_a, seed, max_turn, a)
    nov = ctypes.c_int(0)
    assert L.hive_host_pick_actions_lists(n, lists.ctypes.data, packed.ctypes.data, ep_b.ctypes.data, ctypes.c_uint64(seed), max_turn,
                                          b_.ctypes.data, ctypes.byref(nov)) == 0
    resets = a == -3
    keep = ~flagged | resets                                # (a finished game is reset whatever its group's flag says)
    assert (a[keep] == b_[keep]).all() and (ep_a == ep_b).all()
    assert (b_[flagged & ~resets] == -77).all() and nov.value == int((flagged & ~resets).sum())
