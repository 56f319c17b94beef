set -x
mkdir -p gpurun_out
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/ref1.json 2> gpurun_out/ref1.err; tail -2 gpurun_out/ref1.err; cat gpurun_out/ref1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/ref2.json 2> gpurun_out/ref2.err; tail -5 gpurun_out/ref2.err; cat gpurun_out/ref2.json
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --config 3 --steps 3 --warmup 3 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 2>&1 | tail -1
