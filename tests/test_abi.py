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
