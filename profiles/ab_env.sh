# usage: ab_env.sh "<ENV=.. ENV=..>" ...   -- one bench run per environment string, default library
i=0
for e in "$@"; do i=$((i+1)); env $e timeout 200 python bench.py --no-selfplay --no-cpu-baseline > gpurun_out/bench_env_$i.log 2> gpurun_out/bench_env_$i.err; echo "$e"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_env_$i.log').read().strip().splitlines()[-1])
r=d['roofline']; print(round(d['value']/1e6,2), round(d['ms_per_step']*1e3,2), {k:round(x,1) for k,x in r['dominant_kernel']['per_kernel_us'].items()}, round(d['e2e']['value']/1e6,1))
"; done
