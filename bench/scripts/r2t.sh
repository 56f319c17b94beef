#!/bin/bash
# last check of the final binary: conv tests + smoke
set -u
mkdir -p gpurun_out/r2t
python -m pytest tests/test_gpu_kernels.py tests/test_fused_resnet.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2t/t.log
python __graft_entry__.py --smoke 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r2t/smoke.log
