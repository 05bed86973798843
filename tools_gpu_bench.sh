#!/bin/bash
# tests + bench + per-layer rooflines + ncu launch list + one full capture of the fused kernel launches
mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -m gpu --maxfail=8 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
if ! grep -q " passed" gpurun_out/t_all.log || grep -q "failed" gpurun_out/t_all.log; then tail -30 gpurun_out/t_all.log; fi
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 400 python tools/bench_layers.py --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bt_fused -s 21 -c 21 -o gpurun_out/prof_fused \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/t_all.log; cat gpurun_out/bench.json; tail -12 gpurun_out/layers.log; tail -3 gpurun_out/bench.err
