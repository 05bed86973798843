#!/bin/bash
# quick A/B: direct tests + phase probe + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_direct.py tests/test_gpu_model.py -q -m gpu -p no:cacheprovider > gpurun_out/t_direct.log 2>&1; echo "rc=$?" >> gpurun_out/t_direct.log
timeout 400 python tools/direct_probe.py gpurun_out/direct_probe.json > gpurun_out/direct_probe.log 2>&1; echo "rc=$?" >> gpurun_out/direct_probe.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/t_direct.log; cut -c1-400 gpurun_out/direct_probe.log | grep -v "fast_ws\|BN': 128\|X': 3\|SLOTS"; python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print('bench', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3))"; tail -2 gpurun_out/bench.err
