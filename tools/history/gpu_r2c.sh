#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_configs.py -p no:cacheprovider > gpurun_out/t_c1.log 2>&1; echo "rc=$?" >> gpurun_out/t_c1.log
timeout 900 python -m pytest tests/test_gpu_configs.py -q -p no:cacheprovider > gpurun_out/t_c2.log 2>&1; echo "rc=$?" >> gpurun_out/t_c2.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
echo "== c1"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/t_c1.log | tail -40
echo "== c2"; grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/t_c2.log | tail -30
tail -2 gpurun_out/smoke.log
grep -E "C3_bench|C4_|convt_golden|lstm|C2_full|C5_full" gpurun_out/parity.jsonl | cut -c1-260
