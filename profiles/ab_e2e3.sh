# usage: ab_e2e3.sh "<parts> <drivers> [ENV=..]..." ...  -- e2e only matters here
i=0
for cfg in "$@"; do i=$((i+1)); set -- $cfg; parts=$1; drv=$2; shift; shift
  env "$@" timeout 200 python bench.py --no-selfplay --no-cpu-baseline --e2e-parts $parts --e2e-drivers $drv > gpurun_out/e3_$i.log 2> gpurun_out/e3_$i.err
  echo "parts=$parts drivers=$drv $*"; python -c "
import json
d=json.loads(open('gpurun_out/e3_$i.log').read().strip().splitlines()[-1])
e=d['e2e']
print('   e2e', round(e['value']/1e6,1), 'M/s  policy', round(e['policy_share_of_thread_time'],2), 'wait', round(e['wait_share_of_thread_time'],2), 'pcie', round(e['pcie_gbs_this_rank'],1), ' resident', round(d['value']/1e6,1))
" 2>&1 | tail -1; done
