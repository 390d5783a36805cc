#!/bin/bash
# Round-2 (late) ncu evidence: launch list of the bench command's step with the final kernels (eager, 8 + 8 FCOS) and --set full
# of the 2-D patch 3x3 kernel on the res3 conv2 shape.
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02b_launches_fcos_8x8.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph --no-extras > gpurun_out/r02b_bench_under_ncu.log 2>&1
tail -1 gpurun_out/r02b_bench_under_ncu.log | cut -c1-120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo_kernel -s 6 -c 1 -o gpurun_out/prof_r02b_conv3x3_halo_128 \
    python tools/bench_one.py 16 100 168 128 128 3 1 none fwd > gpurun_out/r02b_ncu_halo.log 2>&1
tail -1 gpurun_out/r02b_ncu_halo.log
ls -la gpurun_out/prof_r02b_conv3x3_halo_128.ncu-rep gpurun_out/r02b_launches_fcos_8x8.csv
