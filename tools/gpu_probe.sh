#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_direct.py -q -m gpu -p no:cacheprovider > gpurun_out/t_direct.log 2>&1; echo "rc=$?" >> gpurun_out/t_direct.log
timeout 400 python tools/direct_probe.py gpurun_out/direct_probe.json > gpurun_out/direct_probe.log 2>&1; echo "rc=$?" >> gpurun_out/direct_probe.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/t_direct.log; cat gpurun_out/direct_probe.log | cut -c1-400; python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print('bench', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3))"; tail -2 gpurun_out/bench.err
