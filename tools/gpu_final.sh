#!/bin/bash
# final verification: gpu tests, smoke, bench (sigma cache on / off, direct off), per-layer rooflines, ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --maxfail=12 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
BT_DISABLE_SIGMA_CACHE=1 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nosigma.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 400 python tools/bench_layers.py --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
tail -4 gpurun_out/t_all.log; tail -2 gpurun_out/smoke.log
for f in bench bench_nosigma; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3), round(d['e2e']['value']), d['roofline']['frac'], d['roofline'].get('traffic'))"; done; tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_reference.json
python - <<'P'
import json
for r in json.load(open('gpurun_out/layers.json')):
    print(r['config'], '| fwd_us', round(r.get('fwd_us',0),1), 'tf', round(r.get('tflops',0),1), 'frac_t', round(r.get('frac_tensor_burst',0),3), 'gbs', round(r.get('gbs',0)), 'kl_us', round(r.get('kl_us',0),1), 'kl_frac', round(r.get('kl_frac_hbm',0),3))
P
