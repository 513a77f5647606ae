# Development helper: runs whatever is being checked on the GPU box.
set -x
for m in 0 2; do
  echo "=== RZK_COMMIT_MODE=$m"
  RZK_COMMIT_MODE=$m timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(commit |open_commit|flags|.*Error)"
done
RZK_PP=2 RZK_COMMIT_MODE=0 timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(commit |.*Error)"
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
