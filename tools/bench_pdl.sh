#!/bin/bash
# A/B of programmatic dependent launch on one box: UT2_PDL=0 (off), default (short kernels only, UT2_PDL_US=40), everywhere.
mkdir -p gpurun_out
run() {  # $1 label, rest: env
  for arch in fcos rcnn; do for bl in 8 2; do
    env "${@:2}" timeout 300 python bench.py --arch $arch --steps 20 --warmup 5 --label $bl --unlabel $bl --no-extras --no-cpu-baseline --no-e2e 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '$arch', '$bl+$bl', round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
  done; done
}
for rep in 1 2; do
  run off UT2_PDL=0
  run us40 UT2_PDL_US=40
  run us20 UT2_PDL_US=20
  run us80 UT2_PDL_US=80
  run all UT2_PDL_US=1000000
done
