#!/bin/bash
# compute-sanitizer targets for the hand-written kernels.  Run on a B200:
#   gpurun --timeout 1500 -- 'bash bench/sanitize.sh'
# The kernel numerics tests (small shapes, every hfl:: kernel incl. the tcgen05/TMA ones) are replayed under memcheck
# (out-of-bounds / misaligned global+shared accesses), racecheck (shared-memory hazards) and synccheck (illegal
# barrier / mbarrier use).  Only kernels whose name contains "hfl" are instrumented, library kernels run natively.
# Results -> gpurun_out/sanitize_<tool>.log ; summary lines on stdout.
set -u
mkdir -p gpurun_out
TESTS="${SANITIZE_TESTS:-tests/test_gpu_kernels.py}"
LIMIT="${SANITIZE_LIMIT:-420}"
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/sanitize_$tool.log
  timeout $LIMIT compute-sanitizer --tool $tool --kernel-name kns=hfl --launch-timeout 300 --print-limit 20 \
      python -m pytest $TESTS -x -q -m gpu -p no:cacheprovider > $log 2>&1
  rc=$?
  echo "$tool rc=$rc | $(grep 'ERROR SUMMARY' $log | tail -1) | $(grep -E 'passed|failed' $log | tail -1)"
done
