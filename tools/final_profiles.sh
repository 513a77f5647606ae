# Round-end evidence run (one GPU): bench lines, ncu launch list of the bench command, full captures of the top kernels.
# usage: bash tools/final_profiles.sh <tag> [A|B]      (gpurun copies back at most 64 MiB: part A and part B are separate calls)
set -x
TAG=${1:-r2}
if [ "$2" != "B" ]; then
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/profile_phase.py commit,open_commit,open_respond,open_verify 3 > gpurun_out/${TAG}_plain_pc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"rzk_vm_kernel|sparse" --launch-skip 6 -c 14 -o /tmp/${TAG}_prof -f python tools/profile_phase.py commit,open_commit,open_respond,open_verify 3 > gpurun_out/${TAG}_ncu_pc.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu_pc.log
# the report (14 launches with imported source, > 64 MiB) stays on the box: summarised here
python tools/ncu_summary.py /tmp/${TAG}_prof.ncu-rep gpurun_out/${TAG}_ncu_summary.json > /dev/null
cut -c1-600 gpurun_out/${TAG}_bench_n1.json
else
python tools/profile_phase.py sum 2 4096 > gpurun_out/${TAG}_plain_sum.log 2>&1 && ncu --set full --clock-control none -k regex:"rzk_vm_kernel" --launch-skip 16 -c 10 -o gpurun_out/${TAG}_prof_sum -f python tools/profile_phase.py sum 2 4096 > gpurun_out/${TAG}_ncu_sum.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu_sum.log
# the report of ten large unrolled kernels exceeds what gpurun copies back: summarise it on the box
python tools/ncu_summary.py gpurun_out/${TAG}_prof_sum.ncu-rep gpurun_out/${TAG}_sum_ncu_summary.json > /dev/null && rm -f gpurun_out/${TAG}_prof_sum.ncu-rep
fi
ls -la gpurun_out | tail -20
