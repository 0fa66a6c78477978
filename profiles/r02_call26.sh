#!/bin/bash
# call 26: AdaLN row norm writing the NVFP4 operand directly: parity + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fp4.py tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/c26_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c26_tests.log); tail -4 gpurun_out/c26_tests.log
timeout 1200 python bench.py > gpurun_out/c26_bench.log 2>&1; tail -1 gpurun_out/c26_bench.log > gpurun_out/r02_bench_call26.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_call26.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d.get('e2e'))
q=d['quantized']; print('fp8', q['value'], q['ms_per_step'], q['ms_per_denoise_step'])
q=q['nvfp4']; print('nvfp4', q['value'], q['ms_per_step'], q['ms_per_denoise_step'], q['clocks'])
PY
