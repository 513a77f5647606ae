# Round-end evidence run (one GPU): bench lines, ncu launch list of the bench command, full captures of the top kernels.
# (gpurun copies back at most 64 MiB: part A and part B are separate calls.)
set -x
if [ "$1" != "B" ]; then
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_final.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
python tools/profile_commit.py > gpurun_out/plain_pc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"rzk_vm_kernel|sparse|hybrid" --launch-skip 2 -c 7 -o gpurun_out/prof_final -f python tools/profile_commit.py > gpurun_out/ncu_pc.log 2>&1
tail -n 2 gpurun_out/ncu_pc.log
cat gpurun_out/bench_final_n1.json | cut -c1-600
else
python tools/profile_sum.py 4096 > gpurun_out/plain_sum.log 2>&1 && ncu --set full --clock-control none -k regex:"rzk_vm_kernel" --launch-skip 12 -c 10 -o gpurun_out/prof_sum -f python tools/profile_sum.py 4096 > gpurun_out/ncu_sum.log 2>&1
tail -n 2 gpurun_out/ncu_sum.log
# the report of ten large unrolled kernels exceeds what gpurun copies back: summarise it on the box
python tools/ncu_summary.py gpurun_out/prof_sum.ncu-rep gpurun_out/sum_ncu_summary.json > /dev/null && rm -f gpurun_out/prof_sum.ncu-rep
fi
ls -la gpurun_out
