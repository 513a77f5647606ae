# Development helper: runs whatever is being checked on the GPU box.
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_chk.json 2> gpurun_out/bench_chk.err; tail -c 800 gpurun_out/bench_chk.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_chk.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','open_verifies_per_s','open_proves_per_s']})
print(d['other_configs']['open_prove_e2e'])
print(d['e2e'])
PY
