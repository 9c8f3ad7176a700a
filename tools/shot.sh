#!/bin/bash
# The one GPU call left in round 1: strip-kernel evaluation, then the whole GPU suite and the bench through
# the fastest correct strip variant.
mkdir -p gpurun_out
rm -f gpurun_out/strip_eval.json gpurun_out/strip_best.txt
timeout 70 python tools/strip_eval.py window strip32 > gpurun_out/strip_eval.log 2>&1
timeout 40 python tools/strip_eval.py strip24 dense24 window_nocache dense24_nocache >> gpurun_out/strip_eval.log 2>&1
timeout 40 python tools/strip_eval.py X_noinsert_dense24 X_noinsert_strip32 X_walkonly_dense24 X_walkonly_strip32 >> gpurun_out/strip_eval.log 2>&1
cat gpurun_out/strip_eval.log | tail -14
BEST=$(cat gpurun_out/strip_best.txt 2>/dev/null)
if [ -n "$BEST" ]; then
  ECB_TEST_STRIP=1 ECB_STRIP_KERNEL=1 ECB_STRIP_WARPS=$BEST timeout 90 python -m pytest tests -m gpu -x -q > gpurun_out/strip_pytest_gpu.log 2>&1
  tail -3 gpurun_out/strip_pytest_gpu.log
  ECB_STRIP_KERNEL=1 ECB_STRIP_WARPS=$BEST timeout 60 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/strip_bench.json 2> gpurun_out/strip_bench.err
  cat gpurun_out/strip_bench.json | cut -c1-600
fi
