for v in p92c p92 p53c p72c p63c; do export HIVE_B200_LIB=$PWD/hive-alphazero_b200/lib/variants/lib_$v.so
echo "--- $v"; HIVE_B200_CONV_PAIR=1 timeout 150 python -m pytest tests/test_net.py -x -q -m gpu -k "tensor_core_trunk or bf16_device" 2>&1 | tail -1
HIVE_B200_CONV_PAIR=1 timeout 120 python profiles/selfplay_probe.py 2048 50 1 tc 2>&1 | grep "net fwd"
HIVE_B200_CONV_PAIR=1 timeout 120 python profiles/selfplay_probe.py 512 50 1 tc 2>&1 | grep "net fwd"
done
unset HIVE_B200_LIB; echo "--- single-CTA"; timeout 120 python profiles/selfplay_probe.py 512 50 1 tc 2>&1 | grep "net fwd"
