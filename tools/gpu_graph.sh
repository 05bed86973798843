#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 200 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2>> gpurun_out/bench.err
timeout 300 python tools/small_s_check.py gpurun_out/small_s.json > gpurun_out/small_s.log 2>&1; echo "rc=$?" >> gpurun_out/small_s.log
tail -3 gpurun_out/t_all.log
for f in bench bench_nograph; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'])"; done; tail -3 gpurun_out/bench.err; cut -c1-250 gpurun_out/small_s.log
