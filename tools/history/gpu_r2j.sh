#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_tma.py -q -p no:cacheprovider -k "pool" > gpurun_out/t_j1.log 2>&1; echo "rc=$?" >> gpurun_out/t_j1.log
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_configs.py -q -p no:cacheprovider -x > gpurun_out/t_j2.log 2>&1; echo "rc=$?" >> gpurun_out/t_j2.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16.csv \
    python bench.py --dtype bf16 --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp32.csv \
    python bench.py --dtype fp32 --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch_fp32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:bt_(tma|tms|dtma)_kernel' -s 21 -c 9 -f \
    -o gpurun_out/prof_j python bench.py --dtype bf16 --profile --steps 1 --warmup 1 > gpurun_out/ncu_full_j.log 2>&1
echo "== j1"; grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/t_j1.log | tail -30
echo "== j2"; grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/t_j2.log | tail -30
python - <<P
import json
d=json.load(open('gpurun_out/bench.json'))
print('bench HEAD', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
r=d['roofline']; print(' frac', round(r['frac'],3), 'kernel_ms', round(r['kernel_ms_per_step'],3), 'share', round(r['kernel_share_of_step'],3), 'tensor', round(r['tensor']['frac'],3))
for k,f in r['families'].items(): print('   ', k, f['launches'], round(f['ms'],3), 'hbm', round(f['hbm_frac'],3), 'tens', round(f['tensor_frac'],3))
b=d['bf16']; print(' BF16', round(b['value']), round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value']))
r=b['roofline']; print(' frac', round(r['frac'],3), 'kernel_ms', round(r['kernel_ms_per_step'],3), 'share', round(r['kernel_share_of_step'],3), 'tensor', round(r['tensor']['frac'],3))
for k,f in r['families'].items(): print('   ', k, f['launches'], round(f['ms'],3), 'hbm', round(f['hbm_frac'],3), 'tens', round(f['tensor_frac'],3))
P
tail -3 gpurun_out/bench.err
python - <<P
import csv
for fn in ('launches_bf16','launches_fp32'):
    rows=[r for r in csv.reader(open('gpurun_out/%s.csv'%fn)) if len(r)>10 and r[0].isdigit()]
    print(fn, len(rows),'launches')
    for r in rows[-30:]:
        print(r[4][:60].ljust(60), r[-1])
P
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_full_j.log
