# stream priorities, carve-out and slices x store CTAs with the final step kernel
echo default; python profiles/pipe_probe.py | tail -1
for p in 1 2; do echo "PRIO=$p"; HIVE_B200_PRIO=$p python profiles/pipe_probe.py | tail -1; done
echo NO_CARVEOUT; HIVE_B200_NO_CARVEOUT=1 python profiles/pipe_probe.py | tail -1
for sc in "3 3" "4 3" "4 4" "2 4"; do set -- $sc; echo "SLICES=$1 STORE_CTAS=$2"; HIVE_B200_SLICES=$1 HIVE_B200_STORE_CTAS=$2 python profiles/pipe_probe.py | tail -1; done
echo default; python profiles/pipe_probe.py | tail -1
