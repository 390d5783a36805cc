#!/bin/bash
# 2-GPU check (gpurun --gpus 2): the NCCL DDP test, then 8+8 and 2+2 bench lines with programmatic dependent launch on / off.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q 2>&1 | tail -2
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --no-cpu-baseline $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1) if d.get('e2e') else None, d['clocks']['sm_mhz'])"; }
UT2_PDL=1 run 29531 "" "fcos8 pdl"
UT2_PDL=0 run 29532 "" "fcos8 nopdl"
UT2_PDL=1 run 29533 "--label 2 --unlabel 2" "fcos2 pdl"
UT2_PDL=0 run 29534 "--label 2 --unlabel 2" "fcos2 nopdl"
UT2_PDL=1 run 29535 "--arch rcnn --label 2 --unlabel 2" "rcnn2 pdl"
