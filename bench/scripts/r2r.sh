#!/bin/bash
# ncu --set full of convolution kernels captured INSIDE a benchmark round of the default engine (not isolated launches)
set -u
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv_fprop_tcgen05|conv_wgrad_tcgen05" -s 30000 -c 6 -f -o $O/prof_conv_inround \
    python bench.py --steps 1 --warmup 1 > $O/ncu_conv_inround.log 2>&1
echo "rc=$?"; ls -la $O/*.ncu-rep
