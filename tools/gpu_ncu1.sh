#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bt_direct -s 2 -c 1 -o gpurun_out/prof_direct_l1 -f \
    python tools/direct_one.py layer1 > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bt_fused -s 2 -c 1 -o gpurun_out/prof_fused_l3 -f \
    python tools/direct_one.py layer3 > gpurun_out/ncu_l3.log 2>&1
tail -3 gpurun_out/ncu_l1.log gpurun_out/ncu_l3.log; ls -la gpurun_out/*.ncu-rep
