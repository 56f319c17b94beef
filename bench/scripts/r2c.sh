mkdir -p gpurun_out/r2c
timeout 600 python -m pytest tests/test_mxfp8.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/r2c/t_mx.log
timeout 900 python -m pytest tests/test_fused_resnet.py "tests/test_gpu_kernels.py::test_train_slot_partial_batch_in_graph" "tests/test_gpu_kernels.py::test_engine_round_gpu" -m gpu -q -s 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/r2c/t_new.log
timeout 300 python bench/mxfp8_probe.py 2>&1 | tail -12 | tee gpurun_out/r2c/mxfp8_probe.md
timeout 900 python bench.py --config mnist --steps 2 --warmup 3 2> gpurun_out/r2c/mnist.err | tail -1 > gpurun_out/r2c/mnist.json
python -c "
import json
l=json.load(open('gpurun_out/r2c/mnist.json')); print('mnist', l['ms_per_step'], l['value'], l.get('e2e',{}).get('value'), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()}, l['last_round'].get('test'))" || tail -5 gpurun_out/r2c/mnist.err
