"""Per-phase SM clocks of hive_step_kernel (library built with -DHIVE_PHASE_CLOCKS, see build_variants.py):
    python profiles/build_variants.py clk:-DHIVE_PHASE_CLOCKS
    HIVE_B200_LIB=$PWD/hive-alphazero_b200/lib/variants/lib_clk.so python profiles/phase_probe.py
Prints the mean clocks a CTA spends in each phase (thread 0's view: barrier to barrier), alone on the GPU
(profile_step: un-sliced launches) and inside the running rollout graph."""
import ctypes
import sys
sys.path.insert(0, '.')
import numpy as np
import hive_b200
from importlib import import_module
L = import_module("hive-alphazero_b200._capi").lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
b = hive_b200.HiveBatch(n)
for _ in range(140):
    b.step_random(20261018, 55, True)
b.sync()
buf = (ctypes.c_ulonglong * 8)()
names = ["prologue", "pieces", "flood", "moves", "encode", "finalize"]


def report(tag):
    L.hive_phase_clocks(buf, 1)
    c = np.array(list(buf), dtype=np.float64)
    ctas = max(c[7], 1)
    print(tag, {k: int(c[i] / ctas) for i, k in enumerate(names)}, "total", int(c[:6].sum() / ctas), "clocks/CTA; CTAs", int(ctas))


L.hive_phase_clocks(buf, 1)
ms = [b.profile_step(20261018, 55) for _ in range(4)]
report("alone  ")
print("   step kernel us:", [round(m["step"] * 1e3, 1) for m in ms])
b.step_random_multi(20261018, 20, 55, True)
b.sync()
L.hive_phase_clocks(buf, 1)
b.step_random_multi(20261018, 20, 55, True)
b.sync()
report("graph  ")
