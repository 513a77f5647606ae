# Development helper: runs whatever is being checked on the GPU box.
set -x
timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(commit|open_|linear|sum|flags|.*Error)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 1500 --warmup 5 --no-e2e --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('sustained', {k:d[k] for k in ['value','ms_per_step','open_verifies_per_s','clocks']}, d['roofline']['kernel_ms'])"
