#!/bin/bash
# call 39: two ranks under torchrun: the bench contract at N = 2 with the NVFP4 quantised leg
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c39_bench.log 2>&1
tail -1 gpurun_out/c39_bench.log > gpurun_out/r02_bench_n2_call39.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2_call39.json').read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('e2e'))
q=d['quantized']; print('nvfp4', q['value'], q['ms_per_step']); print('fp8', q['fp8']['value'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-300
