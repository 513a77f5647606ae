set -x
timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(open_respond|linear_respond|sum_respond|.*Error)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
python tools/profile_respond.py 4 > gpurun_out/plain_resp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sparse -s 2 -c 1 -o gpurun_out/prof_sparse2 -f python tools/profile_respond.py 4 > gpurun_out/ncu_resp.log 2>&1
