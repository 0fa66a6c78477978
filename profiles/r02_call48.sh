#!/bin/bash
# call 48: whole GPU suite + smoke + bench in the final state of the NVFP4 path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/c48_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c48_tests.log); tail -5 gpurun_out/c48_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c48_smoke.log 2>&1; tail -2 gpurun_out/c48_smoke.log
timeout 1200 python bench.py > gpurun_out/c48_bench.log 2>&1; tail -1 gpurun_out/c48_bench.log > gpurun_out/r02_bench_call48.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_call48.json').read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d.get('e2e'), d.get('clocks'))
q=d['quantized']; print('nvfp4', q['value'], q['ms_per_step'], q['ms_per_denoise_step'], q['clocks'])
q=q['fp8']; print('fp8', q['value'], q['ms_per_step'], q['ms_per_denoise_step'])
print(d['roofline']['frac'], d.get('kernels',{}).get('attention'))
PY
