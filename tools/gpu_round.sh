#!/bin/bash
# standard evidence run: gpu tests, smoke, bench (+ A/B), per-layer rooflines, ncu launch list, ncu --set full of one step's
# Bayesian-layer launches, direct-kernel phase probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --maxfail=12 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
BT_DISABLE_DIRECT=1 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nodirect.json 2>> gpurun_out/bench.err
timeout 400 python tools/bench_layers.py --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:bt_(fused|ws|direct)' -s 21 -c 21 -f -o gpurun_out/prof_fused \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
timeout 300 python tools/direct_probe.py gpurun_out/direct_probe.json > gpurun_out/direct_probe.log 2>&1; echo "rc=$?" >> gpurun_out/direct_probe.log
tail -4 gpurun_out/t_all.log; tail -2 gpurun_out/smoke.log
for f in bench bench_nodirect; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3), round(d['e2e']['value']))"; done; tail -3 gpurun_out/bench.err
python - <<'P'
import json
for r in json.load(open('gpurun_out/layers.json')):
    print(r['config'], '| fwd_us', round(r.get('fwd_us',0),1), 'tf', round(r.get('tflops',0),1), 'frac_t', round(r.get('frac_tensor_burst',0),3), 'gbs', round(r.get('gbs',0)), 'kl_us', round(r.get('kl_us',0),1))
P
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_full.log
