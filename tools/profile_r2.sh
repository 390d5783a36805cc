#!/bin/bash
# Round-2 ncu evidence (run under gpurun): the launch list of the default bench command's step (8 + 8 FCOS, eager so that every
# launch is attributed), then --set full of the dominant conv_fwd launch shape, of a conv1 data-gradient with the fused
# block-output ReLU backward (residual + mask tiles), of the fused stem + max-pool kernels and of the GroupNorm backward.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_fcos_8x8.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -1 gpurun_out/r02_bench_under_ncu.log | cut -c1-160
timeout 600 $NCU -k regex:conv_fwd_kernel -s 6 -c 1 -o gpurun_out/prof_r02_conv3x3_256_256 \
    python tools/bench_one.py 16 100 168 256 256 3 1 none fwd > gpurun_out/r02_ncu_c1.log 2>&1
timeout 600 $NCU -k regex:conv_fwd_kernel -s 6 -c 1 -o gpurun_out/prof_r02_dgrad_256_1024_resmask \
    python tools/bench_one.py 16 50 84 256 1024 1 1 resmask fwd > gpurun_out/r02_ncu_c2.log 2>&1
timeout 600 $NCU -k regex:"stem_s2d|stem_pool_tc" -s 4 -c 2 -o gpurun_out/prof_r02_stem_pool \
    python tools/bench_stem.py > gpurun_out/r02_ncu_c3.log 2>&1
timeout 600 $NCU -k regex:"gn_bwd" -s 4 -c 2 -o gpurun_out/prof_r02_gn_bwd \
    python tools/bench_gn.py > gpurun_out/r02_ncu_c4.log 2>&1
tail -n 1 gpurun_out/r02_ncu_c1.log gpurun_out/r02_ncu_c2.log
ls -la gpurun_out/prof_r02_*.ncu-rep
