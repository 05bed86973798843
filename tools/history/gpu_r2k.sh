#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_tma.py tests/test_gpu_direct.py tests/test_gpu_model.py tests/test_gpu_layers.py -q -p no:cacheprovider > gpurun_out/t_k1.log 2>&1; echo "rc=$?" >> gpurun_out/t_k1.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 300 python tools/bench_layers.py --quick --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16.csv \
    python bench.py --dtype bf16 --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp32.csv \
    python bench.py --dtype fp32 --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch_fp32.log 2>&1
echo "== k1"; grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/t_k1.log | tail -30
python - <<P
import json
d=json.load(open('gpurun_out/bench.json'))
print('bench HEAD', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
r=d['roofline']; print(' frac', round(r['frac'],3), 'kernel_ms', round(r['kernel_ms_per_step'],3), 'share', round(r['kernel_share_of_step'],3), 'tensor', round(r['tensor']['frac'],3))
for k,f in r['families'].items(): print('   ', k, f['launches'], round(f['ms'],3), 'hbm', round(f['hbm_frac'],3), 'tens', round(f['tensor_frac'],3))
b=d['bf16']; print(' BF16', round(b['value']), round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value']))
r=b['roofline']; print(' frac', round(r['frac'],3), 'kernel_ms', round(r['kernel_ms_per_step'],3), 'share', round(r['kernel_share_of_step'],3), 'tensor', round(r['tensor']['frac'],3))
for k,f in r['families'].items(): print('   ', k, f['launches'], round(f['ms'],3), 'hbm', round(f['hbm_frac'],3), 'tens', round(f['tensor_frac'],3))
for fn in ('layers',):
    print(fn)
    try:
        for r in json.load(open('gpurun_out/%s.json'%fn)):
            print('  ', r['config'], '| fwd_us', round(r.get('fwd_us',0),1), 'tf', round(r.get('tflops',0),1), 'frac_t', round(r.get('frac_tensor_burst',0),3), 'par', r.get('parity_rel_rms'))
    except Exception as e: print(e)
P
tail -3 gpurun_out/bench.err
python - <<P
import csv
for fn in ('launches_bf16','launches_fp32'):
    rows=[r for r in csv.reader(open('gpurun_out/%s.csv'%fn)) if len(r)>10 and r[0].isdigit()]
    print(fn, len(rows),'launches')
    for r in rows[-28:]:
        print(r[4][:60].ljust(60), r[-1])
P
