for s in 2 4 5 6 8; do echo "SLICES=$s"; HIVE_B200_SLICES=$s python profiles/pipe_probe.py | tail -1; done
for s in 2 4; do echo "bench SLICES=$s"; HIVE_B200_SLICES=$s python bench.py --steps 20 --warmup 5 --no-selfplay --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"; done
