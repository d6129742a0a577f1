# host-driven path: parts x driver threads with the r02k step kernel
for pd in "4 4" "6 3" "6 6" "8 4" "8 8" "3 3" "12 4"; do set -- $pd; echo "parts=$1 drivers=$2"; python bench.py --steps 20 --warmup 5 --no-selfplay --no-cpu-baseline --e2e-parts $1 --e2e-drivers $2 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); e=d['e2e']; print(round(e['value']/1e6,1), 'M env-steps/s; wait share', round(e.get('wait_share_of_thread_time',0),2), 'policy share', round(e.get('policy_share_of_thread_time',0),2))"; done
