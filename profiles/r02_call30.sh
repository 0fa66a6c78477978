#!/bin/bash
# call 30: bench after the 12-epilogue-warp NVFP4 kernel and the packed-register quantiser + in-situ kernel breakdown
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/c30_bench.log 2>&1; tail -1 gpurun_out/c30_bench.log > gpurun_out/r02_bench_call30.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_call30.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d.get('e2e'))
q=d['quantized']; print('fp8', q['value'], q['ms_per_step'], q['ms_per_denoise_step'])
q=q['nvfp4']; print('nvfp4', q['value'], q['ms_per_step'], q['ms_per_denoise_step'], q['clocks'])
PY
PROF_QUANT=4 timeout 400 python profiles/prof_step_kernels.py > gpurun_out/r02_step_kernels_q4.txt 2>/dev/null; head -12 gpurun_out/r02_step_kernels_q4.txt | cut -c1-175
