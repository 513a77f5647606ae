# Development helper: what was run on the GPU box while iterating (edit freely).
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh > gpurun_out/gpu_check.log 2>&1; cat gpurun_out/gpu_check.log'
set -x
timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(commit|open_|linear|sum|flags|.*Error)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
