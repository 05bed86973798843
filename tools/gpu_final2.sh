#!/bin/bash
# last call of round 2: cluster test, final bench (both arms already recorded earlier), launch lists, ncu full (traffic), C4 bench
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_model.py tests/test_gpu_configs.py -q -p no:cacheprovider > gpurun_out/t_v.log 2>&1; echo "rc=$?" >> gpurun_out/t_v.log
grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/t_v.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
for dt in fp32 bf16; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${dt}.csv \
      python bench.py --dtype $dt --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch_${dt}.log 2>&1
  timeout 900 ncu --set full --clock-control none -k 'regex:bt_(fused|ws|direct|tma|tms|dtma)' -s 21 -c 21 -f \
      -o gpurun_out/prof_${dt} python bench.py --dtype $dt --profile --steps 1 --warmup 1 > gpurun_out/ncu_full_${dt}.log 2>&1
  ncu -i gpurun_out/prof_${dt}.ncu-rep --page raw --csv > gpurun_out/prof_${dt}_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_${dt}.ncu-rep
done
timeout 600 python tools/bench_layers.py --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 300 python tools/small_s_check.py > gpurun_out/small_s.log 2>&1
cut -c1-300 gpurun_out/bench.json; echo; tail -5 gpurun_out/small_s.log | cut -c1-200
