# Development helper: runs whatever is being checked on the GPU box.
set -x
python tools/latency.py
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
