#!/bin/bash
# final evidence run of a round: gpu tests, smoke, bench (both arms), per-layer rooflines, ncu launch lists, ncu --set full of
# one step's Bayesian-layer launches (fp32 headline model + bf16 model), sanitizer.  usage: tools/gpu_evidence.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl gpurun_out/*.ncu-rep
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 600 python tools/bench_layers.py --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
for dt in fp32 bf16; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${dt}.csv \
      python bench.py --dtype $dt --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch_${dt}.log 2>&1
  # (no --import-source: the report of 21 launches of these kernels with source exceeds what gpurun brings back; the
  #  source-level captures of single kernels are taken separately, profiles/r02_log.md)
  timeout 900 ncu --set full --clock-control none -k 'regex:bt_(fused|ws|direct|tma|tms|dtma)' -s 21 -c 21 -f \
      -o gpurun_out/prof_${dt} python bench.py --dtype $dt --profile --steps 1 --warmup 1 > gpurun_out/ncu_full_${dt}.log 2>&1
  ncu -i gpurun_out/prof_${dt}.ncu-rep --page raw --csv > gpurun_out/prof_${dt}_raw.csv 2>/dev/null
  [ $(stat -c %s gpurun_out/prof_${dt}.ncu-rep) -gt 20000000 ] && rm -f gpurun_out/prof_${dt}.ncu-rep
done
bash tools/gpu_sanitize.sh $tag > gpurun_out/sanitize_summary_${tag}.txt 2>&1
tail -4 gpurun_out/t_all.log; tail -2 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench.json; cat gpurun_out/sanitize_summary_${tag}.txt; ls -la gpurun_out/*.ncu-rep gpurun_out/*_raw.csv
