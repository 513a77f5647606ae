# ncu --set full capture of ONE launch of each named kernel (regex on the template argument), summarised on the box:
# key metrics (ncu_summary.py), per-function and per-phase instruction / stall breakdowns (ncu_source_breakdown.py,
# ncu_phase_breakdown.py).  The .ncu-rep files stay on the box (a report with imported source is ~20 MB).
# usage: bash tools/profile_kernels.sh <tag> <phase for tools/profile_phase.py> <kernel regex> [<kernel regex> ...]
TAG=$1; PHASE=$2; shift 2
python tools/profile_phase.py $PHASE 2 > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
for K in "$@"; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$K" --launch-skip 1 -c 1 -o /tmp/${TAG}_$K -f python tools/profile_phase.py $PHASE 2 > gpurun_out/${TAG}_${K}_ncu.log 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_$K.ncu-rep gpurun_out/${TAG}_${K}_ncu_summary.json > /dev/null
  python tools/ncu_source_breakdown.py /tmp/${TAG}_$K.ncu-rep > gpurun_out/${TAG}_${K}_functions.txt
  python tools/ncu_phase_breakdown.py /tmp/${TAG}_$K.ncu-rep > gpurun_out/${TAG}_${K}_phases.txt
done
ls -la gpurun_out | tail -30
