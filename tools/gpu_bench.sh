#!/bin/bash
# tests + bench (A/B switches) + per-layer rooflines + ncu launch list + one full capture of the fused launches
mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -m gpu --maxfail=8 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
if ! grep -q " passed" gpurun_out/t_all.log || grep -q "failed" gpurun_out/t_all.log; then tail -30 gpurun_out/t_all.log; fi
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
BT_DISABLE_WS=1 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nows.json 2>> gpurun_out/bench.err
timeout 400 python tools/bench_layers.py --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:bt_(fused|ws)' -s 21 -c 21 -o gpurun_out/prof_fused \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/t_all.log; for f in bench bench_nows; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3))"; done; tail -3 gpurun_out/bench.err
