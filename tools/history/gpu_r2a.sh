#!/bin/bash
# round-2 call A: TMA probe + TMA kernels + full gpu suite + quick benches + memcheck of the new kernels
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_tma.py -q -k probe -p no:cacheprovider > gpurun_out/t_probe.log 2>&1; echo "rc=$?" >> gpurun_out/t_probe.log
timeout 600 python -m pytest tests/test_gpu_tma.py -q -k "not probe" -p no:cacheprovider > gpurun_out/t_tma.log 2>&1; echo "rc=$?" >> gpurun_out/t_tma.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_tma.py --maxfail=30 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
BT_DISABLE_TMA=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_notma.json 2>> gpurun_out/bench.err
timeout 400 python tools/bench_layers.py --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
for fam in tma generic; do
  timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py $fam > gpurun_out/sanitize_memcheck_${fam}_r02a.log 2>&1
  echo "memcheck $fam rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_${fam}_r02a.log | tail -1)"
done
echo "== probe"; tail -5 gpurun_out/t_probe.log
echo "== tma"; tail -15 gpurun_out/t_tma.log
echo "== all"; tail -15 gpurun_out/t_all.log
tail -2 gpurun_out/smoke.log
for f in bench bench_notma; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3), round(d['e2e']['value']))"; done; tail -3 gpurun_out/bench.err
python - <<'P'
import json
for r in json.load(open('gpurun_out/layers.json')):
    print(r['config'], '| fwd_us', round(r.get('fwd_us',0),1), 'tf', round(r.get('tflops',0),1), 'frac_t', round(r.get('frac_tensor_burst',0),3), 'gbs', round(r.get('gbs',0)), 'kl_us', round(r.get('kl_us',0),1), 'par', r.get('parity_rel_rms'))
P
