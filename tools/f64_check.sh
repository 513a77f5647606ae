set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_s2c_n1.json 2> gpurun_out/bench_s2c_n1.err; tail -c 400 gpurun_out/bench_s2c_n1.err; cat gpurun_out/bench_s2c_n1.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_r1b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r1b.log 2>&1
tail -2 gpurun_out/ncu_r1b.log
