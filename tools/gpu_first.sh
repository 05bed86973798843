#!/bin/bash
# first GPU bring-up: diagnostics, then the gpu test files one process each (a trap poisons the context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python tests/gpu_diag.py > gpurun_out/diag.log 2>&1; echo "diag rc=$?" >> gpurun_out/diag.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --maxfail=40 -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/t_kernels.log
timeout 900 python -m pytest tests/test_gpu_layers.py -q -m gpu --maxfail=60 -p no:cacheprovider > gpurun_out/t_layers.log 2>&1; echo "rc=$?" >> gpurun_out/t_layers.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/t_model.log 2>&1; echo "rc=$?" >> gpurun_out/t_model.log
tail -5 gpurun_out/diag.log; tail -3 gpurun_out/t_kernels.log; tail -3 gpurun_out/t_layers.log; tail -3 gpurun_out/t_model.log
