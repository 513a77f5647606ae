# Development helper: what was run on the GPU box while iterating (edit freely).
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh > gpurun_out/gpu_check.log 2>&1; cat gpurun_out/gpu_check.log'
set -x
timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(commit|open_|linear|sum|flags|.*Error)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/profile_sum.py > gpurun_out/plain_sum.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"rzk_vm_kernel" --launch-skip 11 -c 9 -o gpurun_out/prof_sum -f python tools/profile_sum.py > gpurun_out/ncu_sum.log 2>&1
tail -n 2 gpurun_out/ncu_sum.log
