# Development helper: what was run on the GPU box while iterating (edit freely).
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh > gpurun_out/gpu_check.log 2>&1; cat gpurun_out/gpu_check.log'
set -x
for lib in lib_w12 lib_w14; do
RZK_LIB_PATH=$PWD/ring-zk_b200/_build/$lib.so timeout 300 python tools/quick_time.py 65536 2>&1 | grep -E "^(commit|open_|linear|sum|flags|.*Error)"
done
