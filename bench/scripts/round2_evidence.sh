#!/bin/bash
# Round-2 evidence run on ONE B200: ncu --set full of the two product kernels, compute-sanitizer over the new kernels,
# SASS census.  Outputs under gpurun_out/ (summaries are copied to profiles/ on the CPU box).
set -u
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
$N -k regex:mega_kernel      -s 1 -c 1 -f -o gpurun_out/prof_mega  python bench/kernels_once.py mega 148 > gpurun_out/ncu_mega.log 2>&1
$N -k regex:aggregate_v2     -s 3 -c 1 -f -o gpurun_out/prof_agg2  python bench/kernels_once.py fed  > gpurun_out/ncu_agg2.log 2>&1
$N -k regex:distribute_kernel -s 3 -c 1 -f -o gpurun_out/prof_dist2 python bench/kernels_once.py fed  > gpurun_out/ncu_dist2.log 2>&1
$N -k regex:gemm_grouped     -s 2 -c 1 -f -o gpurun_out/prof_gg    python -m pytest tests/test_fused_transformer.py -q -k grouped > gpurun_out/ncu_gg.log 2>&1
$N -k regex:gemm_mxfp8      -s 2 -c 1 -f -o gpurun_out/prof_mxfp8 python bench/mxfp8_probe.py > gpurun_out/ncu_mxfp8.log 2>&1
$N -k regex:gn_act_bwd      -s 1 -c 1 -f -o gpurun_out/prof_gn    python -m pytest tests/test_fused_resnet.py -q -m gpu -k "groupnorm_train_step_gpu_bf16 and gn-1.0" > gpurun_out/ncu_gn.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 300 python bench/mxfp8_probe.py 2>&1 | tail -12 | tee gpurun_out/mxfp8_probe.md
timeout 300 python -m pytest tests/test_fused_transformer.py -q -m gpu -s -k "mxfp8" 2>&1 | tail -5 | tee gpurun_out/t_fp8_lm.log
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/sanitize2_$tool.log
  timeout 500 compute-sanitizer --tool $tool --kernel-name kns=hfl --launch-timeout 300 --print-limit 20 \
      python -m pytest tests/test_mega.py tests/test_gpu_kernels.py tests/test_mxfp8.py tests/test_fused_resnet.py -q -m gpu -p no:cacheprovider \
      -k "conv_fprop or conv_dgrad or conv_wgrad or norm_phases or two_steps or aggregate or jobs or stats_and_eval or sgd or masked_ce or mxfp8 or mx_quantize or gn_kernels or maxpool or nvalid" > $log 2>&1
  echo "$tool rc=$? | $(grep 'ERROR SUMMARY' $log | tail -1) | $(grep -E 'passed|failed' $log | tail -1)"
done
