for s in 1 2 3 4; do echo "SLICES=$s"; HIVE_B200_SLICES=$s python profiles/pipe_probe.py | tail -1; done
for c in 1 2 3; do echo "STORE_CTAS=$c"; HIVE_B200_STORE_CTAS=$c python profiles/pipe_probe.py | tail -1; done
echo "SKIP_PLANES"; HIVE_B200_EXPERIMENT_SKIP_PLANES=1 python profiles/pipe_probe.py | tail -1
echo "QUEUE=1"; HIVE_B200_ROLLOUT_QUEUE=1 python profiles/pipe_probe.py | tail -1
echo "QUEUE=2"; HIVE_B200_ROLLOUT_QUEUE=2 python profiles/pipe_probe.py | tail -1
