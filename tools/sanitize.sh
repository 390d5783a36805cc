#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests ("race detection" aux). Run on the GPU box:
#   gpurun --timeout 2400 -- bash tools/sanitize.sh
# Writes gpurun_out/sanitizer_{memcheck,racecheck,synccheck}.log; the summaries are committed under profiles/.
set -u
OUT=gpurun_out
mkdir -p $OUT
TESTS="tests/test_conv_gpu.py tests/test_kernels_gpu.py tests/test_layers_gpu.py"
SEL='not stage_width'    # that case spawns a child interpreter (not followed by the tool)
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
      python -m pytest $TESTS -m gpu -q -x -k "$SEL" -p no:cacheprovider > $OUT/sanitizer_$tool.log 2>&1
  echo "exit code $?" >> $OUT/sanitizer_$tool.log
  tail -3 $OUT/sanitizer_$tool.log
done
