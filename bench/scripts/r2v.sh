#!/bin/bash
# memcheck over the kernels added after the main sanitizer run (folded shortcut operand pair, mma attention incl. fp32 row sums)
mkdir -p gpurun_out/r2v
timeout 110 compute-sanitizer --tool memcheck --kernel-name kns=hfl --launch-timeout 100 --print-limit 10 \
    python -m pytest tests/test_gpu_kernels.py tests/test_fused_transformer.py -q -m gpu -p no:cacheprovider \
    -k "shortcut_fold or (attention and not 40-50)" > gpurun_out/r2v/memcheck.log 2>&1
echo "rc=$? | $(grep 'ERROR SUMMARY' gpurun_out/r2v/memcheck.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/r2v/memcheck.log | tail -1)"
