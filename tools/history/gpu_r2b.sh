#!/bin/bash
# round-2 call B: new parity tests (grad, configs, fresh-eps graphs), tf32 truncation model, fp32 bench, TMA-preferred bench
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_grad.py tests/test_gpu_layers.py tests/test_gpu_model.py tests/test_gpu_direct.py tests/test_gpu_kernels.py -q -p no:cacheprovider > gpurun_out/t_b1.log 2>&1; echo "rc=$?" >> gpurun_out/t_b1.log
timeout 900 python -m pytest tests/test_gpu_configs.py -q -p no:cacheprovider > gpurun_out/t_b2.log 2>&1; echo "rc=$?" >> gpurun_out/t_b2.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
BT_TMA_PREFER=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmaprefer.json 2>> gpurun_out/bench.err
BT_TMA_PREFER=1 BT_TMA_MODE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmaprefer_res.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --dtype fp32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2>> gpurun_out/bench.err
BT_TMA_MODE=1 timeout 300 python bench.py --dtype fp32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32_res.json 2>> gpurun_out/bench.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp32.csv \
    python bench.py --dtype fp32 --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch_fp32.log 2>&1
echo "== b1"; tail -25 gpurun_out/t_b1.log
echo "== b2"; tail -25 gpurun_out/t_b2.log
for f in bench bench_tmaprefer bench_tmaprefer_res bench_fp32 bench_fp32_res; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3), round(d['e2e']['value']))"; done; tail -3 gpurun_out/bench.err
cat gpurun_out/parity.jsonl | grep -v tma_vs_other | head -80
