# usage: ab_e2e.sh "<parts> <ENV=.. ENV=..>" ...   -- one bench run (no self-play, no CPU arm) per configuration of the host-driven path
mkdir -p gpurun_out
i=0
for cfg in "$@"; do i=$((i+1)); set -- $cfg; parts=$1; shift
  env "$@" timeout 200 python bench.py --no-selfplay --no-cpu-baseline --e2e-parts $parts > gpurun_out/e2e_$i.log 2> gpurun_out/e2e_$i.err
  echo "parts=$parts $*"; python -c "
import json
d=json.loads(open('gpurun_out/e2e_$i.log').read().strip().splitlines()[-1])
print('   resident', round(d['value']/1e6,1), 'M/s   e2e', round(d['e2e']['value']/1e6,1), 'M/s')
" 2>&1 | tail -1; done
