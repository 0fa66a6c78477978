#!/bin/bash
# call 52: eight ranks under torchrun: the bench contract at N = 8 in the final state
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/c52_bench.log 2>&1
echo "wall seconds: $SECONDS"
tail -1 gpurun_out/c52_bench.log > gpurun_out/r02_bench_n8_call52.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n8_call52.json').read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('e2e'))
q=d['quantized']; print('nvfp4', q['value'], q['ms_per_step']); print('fp8', q['fp8']['value'])
PY
