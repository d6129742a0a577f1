"""Per-CTA timeline of the env step (experiment build with -DHIVE_TRACE, see hive_env_kernel.cuh).
usage: HIVE_B200_LIB=<variant .so> python profiles/trace_probe.py out.npz [steps]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hive_b200
from hive_b200 import _capi

out, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 10
L = _capi.lib()
b = hive_b200.HiveBatch(16384)
for _ in range(140):                      # positions of every age
    b.step_random(7, 55, True)
b.step_random_multi(7, steps, 55, True)
b.step_random_multi(7, steps, 55, True)
b.sync()
cap = 400000
assert L.hive_trace_start(cap) == 0
b.step_random_multi(7, steps, 55, True)
b.sync()
buf = np.zeros(cap, dtype=[("t0", "<u8"), ("t1", "<u8"), ("sm", "<u4"), ("kernel", "<u4"), ("g_offset", "<u4"), ("block", "<u4")])
n = L.hive_trace_read(buf.ctypes.data_as(ctypes.c_void_p), cap)
buf = buf[:n]
np.savez_compressed(out, trace=buf)
print("records", n, "span us", (buf["t1"].max() - buf["t0"].min()) / 1e3, "per step", (buf["t1"].max() - buf["t0"].min()) / 1e3 / steps)
