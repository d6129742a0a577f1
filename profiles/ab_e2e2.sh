# usage: ab_e2e2.sh "<threads> <parts>" ...   -- host-driven path only (resident part shortened)
for cfg in "$@"; do set -- $cfg
  HIVE_B200_E2E_THREADS=$1 timeout 300 python bench.py --no-selfplay --no-cpu-baseline --steps 20 --warmup 3 --min-window-ms 10 --e2e-parts $2 > gpurun_out/e2e.json 2> gpurun_out/e2e.err
  python -c "
import json
d=json.loads(open('gpurun_out/e2e.json').read().strip().splitlines()[-1]); e=d['e2e']
print('threads $1 parts $2: e2e %.1f M/s  policy %.2f wait %.2f pcie %.1f GB/s  (%s)' % (e['value']/1e6, e['policy_share_of_thread_time'], e['wait_share_of_thread_time'], e['pcie_gbs_this_rank'], e['bound']))
" 2>&1 | tail -1; done
