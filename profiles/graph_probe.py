"""The resident rollout as ncu sees a whole CUDA graph (--graph-profiling graph): one multi-step graph launch =
`steps` env steps of 16,384 games with all slices overlapping, caches left alone (--cache-control none), so
dram__bytes_* are the pipeline's real HBM traffic per launch (divide by `steps`).  Run with HIVE_B200_SPLIT_GRAPHS=0
(one graph for all slices):

    HIVE_B200_SPLIT_GRAPHS=0 ncu --graph-profiling graph --cache-control none --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,\
smsp__issue_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum --profile-from-start off \
        --csv --log-file gpurun_out/graph.csv python profiles/graph_probe.py
"""
import sys
sys.path.insert(0, '.')
import torch
import hive_b200
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
b = hive_b200.HiveBatch(16384)
for _ in range(140):
    b.step_random(20261018, 55, True)
for _ in range(3):
    b.step_random_multi(20261018, steps, 55, True)
b.sync()
torch.cuda.synchronize()
torch.cuda.profiler.start()
b.step_random_multi(20261018, steps, 55, True)
b.step_random_multi(20261018, steps, 55, True)
b.sync()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", steps)
