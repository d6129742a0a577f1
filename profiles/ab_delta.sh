HIVE_B200_ROLLOUT_QUEUE=1 timeout 120 python -m pytest tests/test_gpu_env.py -x -q -m gpu -k "multi_step or full_size" 2>&1 | tail -3
bash profiles/ab_generic.sh "- 4 HIVE_B200_ROLLOUT_QUEUE=1" "- 4 HIVE_B200_ROLLOUT_QUEUE=1 HIVE_B200_STORE_CTAS=1" "- 4 HIVE_B200_ROLLOUT_QUEUE=1 HIVE_B200_STORE_CTAS=3" "- 4 HIVE_B200_ROLLOUT_QUEUE=1 HIVE_B200_ROLL_CTAS=1" "- 4"
