# One GPU call: parity tests, the bench line, the ncu launch list, one --set full capture of the env kernels and of the
# network kernels, and the graph-level HBM traffic of the running rollout.
# usage (from the repo root on the GPU box): bash profiles/capture.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-selfplay --chunk 1 --min-window-ms 1"
timeout 300 $CMD > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list rc=$?"
# the un-sliced step (whole 16,384-game batch per kernel: real HBM traffic), mid-game positions
timeout 300 python profiles/step_probe.py > gpurun_out/${tag}_step_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:hive_ -c 4 -o gpurun_out/${tag}_step_prof python profiles/step_probe.py > gpurun_out/${tag}_step_ncu.log 2>&1
echo "full capture (env) rc=$?"; tail -1 gpurun_out/${tag}_step_plain.log
# graph-level traffic of the running pipeline
HIVE_B200_SPLIT_GRAPHS=0 timeout 600 ncu --graph-profiling graph --cache-control none --clock-control none \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum \
  --profile-from-start off --csv --log-file gpurun_out/${tag}_graph_level_ncu.csv python profiles/graph_probe.py > gpurun_out/${tag}_graph_ncu.log 2>&1
echo "graph-level rc=$?"
# the network kernels (one 2,048-board forward: trunk convolutions + the three head kernels)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hive_conv3x3|hive_head" -s 80 -c 6 -o gpurun_out/${tag}_net_prof python profiles/selfplay_probe.py 2048 50 1 tc > gpurun_out/${tag}_net_ncu.log 2>&1
echo "full capture (net) rc=$?"
ls -la gpurun_out | tail -14
