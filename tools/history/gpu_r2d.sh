#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_configs.py -p no:cacheprovider > gpurun_out/t_d1.log 2>&1; echo "rc=$?" >> gpurun_out/t_d1.log
timeout 900 python -m pytest tests/test_gpu_configs.py -q -p no:cacheprovider > gpurun_out/t_d2.log 2>&1; echo "rc=$?" >> gpurun_out/t_d2.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 400 python tools/bench_layers.py --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
echo "== d1"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/t_d1.log | tail -30
echo "== d2"; grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/t_d2.log | tail -30
grep -E "C3_bench|C4_|tma_mc" gpurun_out/parity.jsonl | cut -c1-260
grep tma_vs_other gpurun_out/parity.jsonl | grep float32 | cut -c1-400 | head -6
python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print('HEAD', d['dtype'], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])
r=d['roofline']; print(' roofline frac', round(r['frac'],3), 'kernel_ms', round(r['kernel_ms_per_step'],3), 'eager', round(r['eager_step_ms'],3), 'share', round(r['kernel_share_of_step'],3), 'tensor', round(r['tensor']['frac'],3))
for k,f in r['families'].items(): print('   ', k, f['launches'], round(f['ms'],3), 'hbm', round(f['hbm_frac'],3), 'tens', round(f['tensor_frac'],3))
b=d['bf16']; print('BF16', round(b['value']), round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value']))
r=b['roofline']; print(' roofline frac', round(r['frac'],3), 'kernel_ms', round(r['kernel_ms_per_step'],3), 'share', round(r['kernel_share_of_step'],3), 'tensor', round(r['tensor']['frac'],3))
for k,f in r['families'].items(): print('   ', k, f['launches'], round(f['ms'],3), 'hbm', round(f['hbm_frac'],3), 'tens', round(f['tensor_frac'],3))
print(d.get('cpu_baseline'))
P
cut -c1-600 gpurun_out/bench_reference.json; tail -3 gpurun_out/bench.err
python - <<'P'
import json
for r in json.load(open('gpurun_out/layers.json')):
    print(r['config'], '| fwd_us', round(r.get('fwd_us',0),1), 'tf', round(r.get('tflops',0),1), 'frac_t', round(r.get('frac_tensor_burst',0),3), 'gbs', round(r.get('gbs',0)), 'kl_us', round(r.get('kl_us',0),1), 'par', r.get('parity_rel_rms'))
P
