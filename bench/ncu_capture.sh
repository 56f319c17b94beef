#!/bin/bash
# One ncu --set full capture per hot kernel (1 GPU, under gpurun).  Reports land in gpurun_out/; summaries are
# extracted on the CPU box with bench/ncu_summarize.py and committed under profiles/.
set -u
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
$N -k regex:conv_fprop_tcgen05 -s 6 -c 2 -f -o gpurun_out/prof_conv   python bench/kernels_once.py conv  > gpurun_out/ncu_conv.log 2>&1
$N -k regex:gemm_tcgen05       -s 3 -c 1 -f -o gpurun_out/prof_gemm   python bench/kernels_once.py gemm  > gpurun_out/ncu_gemm.log 2>&1
$N -k regex:aggregate_kernel   -s 3 -c 1 -f -o gpurun_out/prof_agg    python bench/kernels_once.py fed   > gpurun_out/ncu_agg.log 2>&1
$N -k regex:distribute_kernel  -s 3 -c 1 -f -o gpurun_out/prof_dist   python bench/kernels_once.py fed   > gpurun_out/ncu_dist.log 2>&1
$N -k regex:bn_act_            -s 4 -c 2 -f -o gpurun_out/prof_norm   python bench/kernels_once.py norm  > gpurun_out/ncu_norm.log 2>&1
# launch list of one bench round (device time per kernel; shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 40000 -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
