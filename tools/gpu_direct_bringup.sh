#!/bin/bash
# direct-mode bring-up: parity tests with both descriptor base-offset encodings, then the full suite and A/B bench
mkdir -p gpurun_out
for bo in 0 1; do
  BT_DIRECT_BO=$bo timeout 300 python -m pytest tests/test_gpu_direct.py -q -m gpu -p no:cacheprovider > gpurun_out/t_direct_bo$bo.log 2>&1
  echo "rc=$?" >> gpurun_out/t_direct_bo$bo.log
done
BO=0
if ! grep -q " passed" gpurun_out/t_direct_bo0.log || grep -q "failed" gpurun_out/t_direct_bo0.log; then
  if grep -q " passed" gpurun_out/t_direct_bo1.log && ! grep -q "failed" gpurun_out/t_direct_bo1.log; then BO=1; fi
fi
echo "selected BO=$BO" > gpurun_out/direct_bo.txt
export BT_DIRECT_BO=$BO
timeout 500 python -m pytest tests -q -m gpu --maxfail=12 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
BT_DISABLE_DIRECT=1 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nodirect.json 2>> gpurun_out/bench.err
timeout 400 python tools/bench_layers.py --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
for b in 0 1; do echo "== BO=$b"; tail -4 gpurun_out/t_direct_bo$b.log; done
cat gpurun_out/direct_bo.txt; tail -3 gpurun_out/t_all.log
for f in bench bench_nodirect; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3))"; done; tail -3 gpurun_out/bench.err
