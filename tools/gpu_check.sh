# Development helper: runs whatever is being checked on the GPU box.
set -x
timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(commit|open_|linear|sum|flags|.*Error)"
RZK_COMMIT_MODE=0 timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(commit |.*Error)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
