#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -p no:cacheprovider 2>&1 | tail -3
for fam in dtma pool; do
  timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py $fam > gpurun_out/sanitize_memcheck_${fam}_try.log 2>&1
  echo "memcheck $fam rc=$? $(grep -E "ERROR SUMMARY" gpurun_out/sanitize_memcheck_${fam}_try.log | tail -1)"; grep -E "path=" gpurun_out/sanitize_memcheck_${fam}_try.log
done
timeout 300 python bench.py --dtype bf16 --steps 20 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
timeout 300 python tools/small_s_check.py > gpurun_out/small_s.log 2>&1; tail -6 gpurun_out/small_s.log | cut -c1-260
