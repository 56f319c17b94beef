mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-300
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final1.log 2>&1; tail -1 gpurun_out/bench_final1.log | cut -c1-1200
python bench.py --config wikitext --steps 3 --warmup 3 > gpurun_out/bench_final_lm.log 2>&1; tail -1 gpurun_out/bench_final_lm.log | cut -c1-300
