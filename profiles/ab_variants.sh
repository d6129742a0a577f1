for v in "$@"; do HIVE_B200_LIB=$PWD/hive-alphazero_b200/lib/variants/lib_$v.so timeout 200 python bench.py --no-selfplay --no-cpu-baseline > gpurun_out/bench_enc_$v.log 2> gpurun_out/bench_enc_$v.err; echo $v; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_enc_$v.log').read().strip().splitlines()[-1])
r=d['roofline']; print(round(d['value']/1e6,2), round(d['ms_per_step']*1e3,2), round(r['write_only_stream_gbs']), {k:round(x,1) for k,x in r['dominant_kernel']['per_kernel_us'].items()}, round(d['e2e']['value']/1e6,1))
"; done
