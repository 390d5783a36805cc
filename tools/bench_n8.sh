#!/bin/bash
# 8-GPU bench lines (gpurun --gpus 8): FCOS and Faster R-CNN at 8 + 8 per GPU and FCOS at 2 + 2 per GPU with the gradient
# all-reduce overlapped with the backward pass; SERIAL=1 adds the serial all-reduce runs (A/B).
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 3 --no-extras --no-cpu-baseline $2 2>/dev/null | tail -1; }
UT2_OVERLAP_ALLREDUCE=1 run 29521 "" > gpurun_out/r02_bench_fcos_n8_overlap.json
UT2_OVERLAP_ALLREDUCE=1 run 29523 "--arch rcnn" > gpurun_out/r02_bench_rcnn_n8_overlap.json
UT2_OVERLAP_ALLREDUCE=1 run 29524 "--label 2 --unlabel 2" > gpurun_out/r02_bench_fcos_n8_2x2_overlap.json
if [ -n "$SERIAL" ]; then
  UT2_OVERLAP_ALLREDUCE=0 run 29522 "" > gpurun_out/r02_bench_fcos_n8_serial.json
  UT2_OVERLAP_ALLREDUCE=0 run 29525 "--label 2 --unlabel 2" > gpurun_out/r02_bench_fcos_n8_2x2_serial.json
fi
for f in gpurun_out/r02_bench_*n8*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1) if d.get('e2e') else None, d['clocks']['sm_mhz'])"; done
