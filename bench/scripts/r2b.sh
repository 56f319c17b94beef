mkdir -p gpurun_out/r2b
python -m pytest tests/test_fused_resnet.py tests/test_fused_transformer.py "tests/test_gpu_kernels.py::test_train_slot_partial_batch_in_graph" -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2b/t_new.log
python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2b/t_all.log
for c in cifar_stress cifar_homog wikitext; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 2> gpurun_out/r2b/$c.err | tail -1 > gpurun_out/r2b/$c.json
  python -c "
import json,sys
l=json.load(open('gpurun_out/r2b/$c.json')); print('$c', l['ms_per_step'], l['value'], l.get('e2e',{}).get('value'), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()})" || tail -5 gpurun_out/r2b/$c.err
done
