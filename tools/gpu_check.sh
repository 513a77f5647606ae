# Development helper: runs whatever is being checked on the GPU box.
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 300 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_n2_ref.json 2>/dev/null; tail -c 300 gpurun_out/bench_n2_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','n_gpus','ms_per_step','open_verifies_per_s','open_proves_per_s','gpu_launches']}, d['e2e']['value'], d['config']['collective'])
print(d['other_configs']['linear']['instances_per_s'], d['other_configs']['sum64']['instances_per_s'], d['other_configs']['open_prove_e2e']['instances_per_s'])
PY
