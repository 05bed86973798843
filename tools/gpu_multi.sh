#!/bin/bash
# 2-GPU check: sharded MC inference == single rank, then the bench at N=2 (torchrun, NCCL)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "rc=$?" >> gpurun_out/multi_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?" >> gpurun_out/bench_2gpu.err
grep -E "rank|rc=" gpurun_out/multi_check.log | tail -5; cat gpurun_out/bench_2gpu.json | cut -c1-400; tail -3 gpurun_out/bench_2gpu.err
