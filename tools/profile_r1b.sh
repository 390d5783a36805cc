#!/bin/bash
# Round-1 (second pass) ncu captures of the memory-bound 1x1 conv, a trunk 3x3 and the head wgrad (run under gpurun).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 6 -c 1 -o gpurun_out/prof_r1b_conv1x1_64_256_res \
    python tools/bench_one.py 8 200 336 64 256 1 1 res fwd > gpurun_out/ncu_b1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 6 -c 1 -o gpurun_out/prof_r1b_conv1x1_128_512_res \
    python tools/bench_one.py 16 100 168 128 512 1 1 res fwd > gpurun_out/ncu_b2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 6 -c 1 -o gpurun_out/prof_r1b_conv3x3_64_64 \
    python tools/bench_one.py 8 200 336 64 64 3 1 none fwd > gpurun_out/ncu_b3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 6 -c 1 -o gpurun_out/prof_r1b_wgrad3x3_256 \
    python tools/bench_one.py 16 100 168 256 256 3 1 none wgrad > gpurun_out/ncu_b4.log 2>&1
tail -n 1 gpurun_out/ncu_b1.log gpurun_out/ncu_b2.log gpurun_out/ncu_b3.log gpurun_out/ncu_b4.log
ls -la gpurun_out/*.ncu-rep
