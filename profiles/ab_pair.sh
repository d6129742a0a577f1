timeout 200 python -m pytest tests/test_net.py -x -q -m gpu 2>&1 | tail -1
timeout 120 python profiles/selfplay_probe.py 2048 50 1 tc 2>&1 | grep "net fwd"
timeout 120 python profiles/selfplay_probe.py 512 50 1 tc 2>&1 | grep "net fwd"
