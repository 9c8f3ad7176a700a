#!/bin/bash
# First GPU call of round 2 (about 4 minutes of box time): parity of the default tree, the strip kernel
# re-evaluated, and ncu captures of both grouping kernels so that the open question of DESIGN.md 4.1b
# (why an insert batch costs 25 us in the strip kernel and 6 us in the window kernel) can be read off the
# stall reasons / local-memory traffic.   usage: bash tools/build_strip_variants.sh (here, once), then gpurun --timeout 420 -- 'bash tools/round2_first_call.sh'
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu.log
rm -f gpurun_out/strip_eval.json
# separate processes: a variant that faults takes only its own group with it
timeout 90 python tools/strip_eval.py window window_mumptx window_warpprobe window_lean window_lean2 window_lean3 > gpurun_out/r2_strip_eval.log 2>&1
timeout 90 python tools/strip_eval.py flatlog flatlog3 partlog flatlog_lean flatlog_lean2 flatlog3_lean2 flatlog3_lean3 >> gpurun_out/r2_strip_eval.log 2>&1
timeout 90 python tools/strip_eval.py strip32 strip24 dense24 dense24_lean >> gpurun_out/r2_strip_eval.log 2>&1
grep -v "^columns ready\|^best correct" gpurun_out/r2_strip_eval.log
ECB_TEST_STRIP=1 ECB_TEST_FLATLOG=1 timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "strip or flat_log" > gpurun_out/r2_pytest_strip.log 2>&1; tail -2 gpurun_out/r2_pytest_strip.log
# ncu: the third launch of each grouping kernel (tools/group_only.py pushes six times)
timeout 60 python tools/group_only.py > gpurun_out/r2_plain_window.log 2>&1 &&
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:ecb_group_insert_kernel -s 2 -c 1 \
    -o gpurun_out/r2_window python tools/group_only.py > gpurun_out/r2_ncu_window.log 2>&1
ECB_STRIP_KERNEL=1 ECB_STRIP_WARPS=124 timeout 60 python tools/group_only.py > gpurun_out/r2_plain_strip.log 2>&1 &&
  ECB_STRIP_KERNEL=1 ECB_STRIP_WARPS=124 timeout 200 ncu --set full --clock-control none --import-source on \
    -k regex:ecb_group_strip_kernel -s 2 -c 1 -o gpurun_out/r2_strip python tools/group_only.py > gpurun_out/r2_ncu_strip.log 2>&1
ls -la gpurun_out/r2_*
