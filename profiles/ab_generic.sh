# usage: ab_generic.sh "<variant|-> <parts> [ENV=..]..." ...  -- one bench run (no self-play, no CPU arm) per configuration;
# variant "-" = the default library, else hive-alphazero_b200/lib/variants/lib_<variant>.so
mkdir -p gpurun_out
i=0
for cfg in "$@"; do i=$((i+1)); set -- $cfg; v=$1; parts=$2; shift; shift
  libenv=""; [ "$v" != "-" ] && libenv="HIVE_B200_LIB=$PWD/hive-alphazero_b200/lib/variants/lib_$v.so"
  env $libenv "$@" timeout 200 python bench.py --no-selfplay --no-cpu-baseline --e2e-parts $parts > gpurun_out/ab_$i.log 2> gpurun_out/ab_$i.err
  echo "variant=$v parts=$parts $*"; python -c "
import json
d=json.loads(open('gpurun_out/ab_$i.log').read().strip().splitlines()[-1])
r=d['roofline']
print('   resident', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step']*1e3,1), 'us', {k:round(x,1) for k,x in r['dominant_kernel']['per_kernel_us'].items()}, ' e2e', round(d['e2e']['value']/1e6,1), 'M/s')
" 2>&1 | tail -1; done
