"""Resident rollout rate of the library HIVE_B200_LIB names (A/B builds): us per 16,384-game step of the graph pipeline."""
import sys, time
sys.path.insert(0, '.')
import torch
import hive_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
b = hive_b200.HiveBatch(n)
for _ in range(140):
    b.step_random(20261018, 55, True)
b.sync()
for _ in range(3):
    b.step_random_multi(20261018, 20, 55, True)
b.sync()
t0 = time.perf_counter()
reps = 40
for _ in range(reps):
    b.step_random_multi(20261018, 20, 55, True)
b.sync()
dt = time.perf_counter() - t0
print("us per step %.2f  (%.1f M env-steps/s)" % (dt / (reps * 20) * 1e6, n * reps * 20 / dt / 1e6))
