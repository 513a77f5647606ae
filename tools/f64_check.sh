set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_s2b_n1.json 2> gpurun_out/bench_s2b_n1.err; tail -c 300 gpurun_out/bench_s2b_n1.err; cat gpurun_out/bench_s2b_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_s2b_ref.json 2>&1; cat gpurun_out/bench_s2b_ref.json
