# Development helper: GPU test suite (and whatever else is being checked) on the box.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
