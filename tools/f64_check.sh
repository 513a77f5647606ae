set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_s2d_n1.json 2> gpurun_out/bench_s2d_n1.err; tail -c 300 gpurun_out/bench_s2d_n1.err; cat gpurun_out/bench_s2d_n1.json
python -c "import __graft_entry__ as g; g.smoke()"
