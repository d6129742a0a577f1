for c in 2 3 4; do echo "STORE_CTAS=$c"; HIVE_B200_STORE_CTAS=$c python bench.py --steps 20 --warmup 5 --no-selfplay --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"; done
