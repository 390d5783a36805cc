#!/bin/bash
# Round-1 (third pass, current kernels) ncu captures (run under gpurun): launch list of one FCOS bench step, then
# --set full of the head conv fwd / wgrad, a memory-bound 1x1 conv with residual, the batched stem and GroupNorm bwd.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r1c.csv \
    python bench.py --steps 1 --warmup 1 --label 2 --unlabel 2 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
timeout 600 $NCU -k regex:conv_fwd_kernel -s 6 -c 1 -o gpurun_out/prof_r1c_conv1x1_256_1024_res \
    python tools/bench_one.py 16 50 84 256 1024 1 1 res fwd > gpurun_out/ncu_c1.log 2>&1
timeout 600 $NCU -k regex:conv_fwd_kernel -s 6 -c 1 -o gpurun_out/prof_r1c_conv3x3_256_256 \
    python tools/bench_one.py 16 100 168 256 256 3 1 none fwd > gpurun_out/ncu_c2.log 2>&1
timeout 600 $NCU -k regex:conv_wgrad_kernel -s 6 -c 1 -o gpurun_out/prof_r1c_wgrad3x3_256_256 \
    python tools/bench_one.py 16 100 168 256 256 3 1 none wgrad > gpurun_out/ncu_c3.log 2>&1
timeout 600 $NCU -k regex:conv_fwd_kernel -s 6 -c 1 -o gpurun_out/prof_r1c_conv1x1_64_256_res \
    python tools/bench_one.py 8 200 336 64 256 1 1 res fwd > gpurun_out/ncu_c4.log 2>&1
timeout 600 $NCU -k regex:"stem_tc_kernel|gn_bwd" -s 20 -c 3 -o gpurun_out/prof_r1c_stem_gn \
    python bench.py --steps 1 --warmup 0 --label 2 --unlabel 2 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/ncu_c5.log 2>&1
tail -n 1 gpurun_out/ncu_c1.log gpurun_out/ncu_c2.log gpurun_out/ncu_c3.log gpurun_out/ncu_c4.log
ls -la gpurun_out/*.ncu-rep
