"""Un-sliced environment step for the ncu --set full capture (whole 16,384-game batch per kernel, so the 264 MB of
planes do not fit the 126 MB L2 and dram__bytes_* are the step's real HBM traffic).  The games are advanced
`warm` random steps first (positions of every game age), then the CUDA profiler range covers `reps` steps issued
by hive_profile_step (step kernel -> plane store, one launch each, on one stream):

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:hive_ \
        -o gpurun_out/step python profiles/step_probe.py
"""
import sys
sys.path.insert(0, '.')
import torch
import hive_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 140
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
b = hive_b200.HiveBatch(n)
for _ in range(warm):
    b.step_random(20261018, 55, True)
b.sync()
for _ in range(3):
    b.profile_step(20261018, 55)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ms = [b.profile_step(20261018, 55) for _ in range(reps)]
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print({k: round(sum(m[k] for m in ms) / len(ms) * 1e3, 1) for k in ms[0]}, 'us per kernel; n =', n)
