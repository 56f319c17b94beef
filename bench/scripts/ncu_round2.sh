set -u
mkdir -p gpurun_out
python bench/gemm_bench.py > gpurun_out/gemm_bench.json 2> gpurun_out/gemm_bench.err; tail -2 gpurun_out/gemm_bench.err
N="ncu --set full --clock-control none --import-source on"
$N -k regex:gemm_tcgen05_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm python bench/kernels_once.py gemm > gpurun_out/ncu_gemm.log 2>&1
$N -k regex:lm_attn_bwd -s 3 -c 1 -f -o gpurun_out/prof_lm_attn_bwd python bench/kernels_once.py lm > gpurun_out/ncu_lm1.log 2>&1
$N -k regex:lm_add_ln_bwd -s 3 -c 1 -f -o gpurun_out/prof_lm_ln_bwd python bench/kernels_once.py lm > gpurun_out/ncu_lm2.log 2>&1
$N -k regex:conv_wgrad_tcgen05 -s 2 -c 1 -f -o gpurun_out/prof_wgrad python bench/conv_bwd_bench.py > gpurun_out/ncu_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep
