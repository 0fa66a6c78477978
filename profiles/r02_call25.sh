#!/bin/bash
# call 25: full bench with the NVFP4-all quantised leg
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/c25_bench.log 2>&1; tail -1 gpurun_out/c25_bench.log > gpurun_out/r02_bench_call25.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_call25.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d.get('e2e'))
q=d['quantized']; print('fp8', q['value'], q['ms_per_step'], q['ms_per_denoise_step'])
q=q['nvfp4']; print('nvfp4', q['value'], q['ms_per_step'], q['ms_per_denoise_step'], q['clocks'])
PY
