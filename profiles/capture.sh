# One GPU call: parity tests, the bench line, the ncu launch list and one --set full capture of the env kernels.
# usage (from the repo root on the GPU box): bash profiles/capture.sh <tag>
tag=${1:-v7}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-selfplay --chunk 1"
timeout 300 $CMD > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list rc=$?"
# the un-sliced step (whole 16,384-game batch per kernel: real HBM traffic), mid-game positions
timeout 300 python profiles/step_probe.py > gpurun_out/${tag}_step_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:hive_ -o gpurun_out/${tag}_step_prof python profiles/step_probe.py > gpurun_out/${tag}_step_ncu.log 2>&1
echo "full capture rc=$?"; tail -1 gpurun_out/${tag}_step_plain.log
ls -la gpurun_out | tail -12
