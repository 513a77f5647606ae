set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_s2b_n2.json 2> gpurun_out/bench_s2b_n2.err; tail -c 400 gpurun_out/bench_s2b_n2.err; cat gpurun_out/bench_s2b_n2.json
