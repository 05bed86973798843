#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
timeout 600 python -m pytest tests/test_gpu_tma.py -q -p no:cacheprovider > gpurun_out/t_g1.log 2>&1; echo "rc=$?" >> gpurun_out/t_g1.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_tma.py --deselect tests/test_gpu_configs.py -p no:cacheprovider > gpurun_out/t_g2.log 2>&1; echo "rc=$?" >> gpurun_out/t_g2.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
BT_DISABLE_DTMA=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nodtma.json 2>> gpurun_out/bench.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16.csv \
    python bench.py --dtype bf16 --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp32.csv \
    python bench.py --dtype fp32 --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch_fp32.log 2>&1
echo "== g1"; grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/t_g1.log | tail -30
echo "== g2"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/t_g2.log | tail -30
grep -E "dtma" gpurun_out/parity.jsonl | cut -c1-330
for f in bench bench_nodtma; do python - <<P
import json
d=json.load(open('gpurun_out/$f.json'))
print('$f HEAD', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
r=d['roofline']; print(' frac', round(r['frac'],3), 'kernel_ms', round(r['kernel_ms_per_step'],3), 'share', round(r['kernel_share_of_step'],3), 'tensor', round(r['tensor']['frac'],3))
for k,f in r['families'].items(): print('   ', k, f['launches'], round(f['ms'],3), 'hbm', round(f['hbm_frac'],3), 'tens', round(f['tensor_frac'],3))
b=d['bf16']; print(' BF16', round(b['value']), round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value']))
r=b['roofline']; print(' frac', round(r['frac'],3), 'kernel_ms', round(r['kernel_ms_per_step'],3), 'share', round(r['kernel_share_of_step'],3), 'tensor', round(r['tensor']['frac'],3))
for k,f in r['families'].items(): print('   ', k, f['launches'], round(f['ms'],3), 'hbm', round(f['hbm_frac'],3), 'tens', round(f['tensor_frac'],3))
P
done; tail -3 gpurun_out/bench.err
