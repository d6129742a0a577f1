"""Queue-driven rollout (HIVE_B200_ROLLOUT_QUEUE=1) against per-step graphs: equality of the batch after n steps, then timing.
usage: python profiles/queue_probe.py [n_games] [n_steps] [reps]"""
import os
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
import hive_b200
seed = 4242
os.environ["HIVE_B200_ROLLOUT_QUEUE"] = os.environ.get("QMODE", "1")
a = hive_b200.HiveBatch(n)
os.environ["HIVE_B200_ROLLOUT_QUEUE"] = "0"
b = hive_b200.HiveBatch(n)
for _ in range(3):
    a.step_random_multi(seed, steps); a.sync()
    b.step_random_multi(seed, steps); b.sync()
print("synced", flush=True)
m1, c1 = a.legal_mask(); m2, c2 = b.legal_mask()
print("equal:", bool((m1 == m2).all() and (c1 == c2).all()), bool((a.planes_bf16() == b.planes_bf16()).all()),
      [x.tolist() for x in a.status()] == [x.tolist() for x in b.status()], flush=True)
for tag, h in (("queue", a), ("graph", b)):
    for _ in range(3):
        h.step_random_multi(seed, steps)
    h.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        h.step_random_multi(seed, steps)
    h.sync()
    dt = time.perf_counter() - t0
    print(tag, round(dt / (reps * steps) * 1e6, 2), "us per step", flush=True)
