python bench/gemm_small_probe.py 2>&1 | tail -14
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 20 -c 3 -o gpurun_out/gemm_small python bench/gemm_small_probe.py > gpurun_out/ncu_small.log 2>&1
ncu -i gpurun_out/gemm_small.ncu-rep --page raw --csv > gpurun_out/gemm_small_raw.csv 2>/dev/null
ls -la gpurun_out/gemm_small*
