#!/bin/bash
# Round-1 profiling pass (run under gpurun): smoke, launch list of one bench step, full ncu capture of the top kernels.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
# launch list: 2+2 images keeps the serialised ncu pass short; shares are what matter
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 1 --label 2 --unlabel 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 120 -c 3 -o gpurun_out/prof_conv_fwd_r1 \
    python bench.py --steps 1 --warmup 0 --label 2 --unlabel 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 40 -c 2 -o gpurun_out/prof_conv_wgrad_r1 \
    python bench.py --steps 1 --warmup 0 --label 2 --unlabel 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
