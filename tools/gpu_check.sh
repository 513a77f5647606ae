# Development helper: what was run on the GPU box while iterating (edit freely).
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh > gpurun_out/gpu_check.log 2>&1; cat gpurun_out/gpu_check.log'
set -x
timeout 600 python tools/soak.py 12 5 2>&1 | tail -6
for B in 1 64 512; do
timeout 300 python tools/profile_sum_time.py $B
RZK_NO_SEGMENTS=1 timeout 300 python tools/profile_sum_time.py $B
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
