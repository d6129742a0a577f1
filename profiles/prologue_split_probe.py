import ctypes, sys
sys.path.insert(0, '.')
import numpy as np
import hive_b200
from importlib import import_module
L = import_module("hive-alphazero_b200._capi").lib()
b = hive_b200.HiveBatch(16384)
for _ in range(140): b.step_random(20261018, 55, True)
b.sync()
buf = (ctypes.c_ulonglong * 8)()
b.step_random_multi(20261018, 20, 55, True); b.sync()
L.hive_phase_clocks(buf, 1)
b.step_random_multi(20261018, 20, 55, True); b.sync()
L.hive_phase_clocks(buf, 1)
c = list(buf); n = c[7]
print('CTAs', n, 'prologue total', c[0] / n, 'loads', (c[6] & 0xFFFFFFFF) / n, 'pick', (c[6] >> 32) / n)
