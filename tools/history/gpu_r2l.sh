#!/bin/bash
mkdir -p gpurun_out
V=bayesian_torch_b200/build/variants
for rep in 1 2; do
for v in default pp0pin1 pp1pin0 pp0pin0; do
  if [ $v = default ]; then unset BT_LIB_VARIANT; else export BT_LIB_VARIANT=$PWD/$V/libbtb200_$v.so; fi
  timeout 300 python bench.py --dtype bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  timeout 300 python bench.py --dtype fp32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench32_$v.json 2>> gpurun_out/bench_$v.err
  python - <<P
import json
d=json.load(open('gpurun_out/bench_$v.json')); e=json.load(open('gpurun_out/bench32_$v.json'))
fam=lambda x: {k:round(f['ms'],3) for k,f in x['roofline']['families'].items()}
print('$v rep$rep bf16', round(d['ms_per_step'],4), fam(d), '| fp32', round(e['ms_per_step'],4), fam(e))
P
done
done
unset BT_LIB_VARIANT
for v in default pp0pin1 pp1pin0 pp0pin0; do
  if [ $v = default ]; then unset BT_LIB_VARIANT; else export BT_LIB_VARIANT=$PWD/$V/libbtb200_$v.so; fi
  timeout 300 python tools/bench_layers.py --quick --out gpurun_out/layers_$v.json > gpurun_out/layers_$v.log 2>&1
  python - <<P
import json
print('$v', [(r['config'][:24], round(r.get('fwd_us',0),1)) for r in json.load(open('gpurun_out/layers_$v.json'))])
P
done
