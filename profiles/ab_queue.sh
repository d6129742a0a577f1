for v in f72 f64; do export HIVE_B200_LIB=$PWD/hive-alphazero_b200/lib/variants/lib_$v.so
for cfg in "1" "2"; do echo "$v FLOW STORE=$cfg"; QMODE=2 HIVE_B200_ROLL_STORE_CTAS=$cfg timeout 60 python profiles/queue_probe.py 16384 20 40 2>&1 | grep "queue\|equal\|rror"; done
echo "$v FLOW NOSTORE 1"; QMODE=2 HIVE_B200_ROLL_STORE_CTAS=1 HIVE_B200_EXPERIMENT_Q_NOSTORE=1 timeout 60 python profiles/queue_probe.py 16384 20 40 2>&1 | grep "queue\|rror"
done
