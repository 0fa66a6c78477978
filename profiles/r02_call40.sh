#!/bin/bash
# call 40: attention epilogue emits the proj / linear2 operand chunks; producer-emitted NVFP4 operands (GELU epilogue -> e2m1, attention-output chunk quantiser, finalise): parity + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fp4.py -m gpu -q --timeout 300 -p no:cacheprovider -x -s > gpurun_out/c40_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c40_tests.log); grep -i "nvfp4 (\|chunked\|epilogue-quantised\|passed\|failed\|error\|rc=\|assert" gpurun_out/c40_tests.log | tail -24
timeout 1200 python bench.py > gpurun_out/c40_bench.log 2>&1; tail -1 gpurun_out/c40_bench.log > gpurun_out/r02_bench_call40.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_call40.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d.get('e2e'))
q=d['quantized']; print('nvfp4', q['value'], q['ms_per_step'], q['ms_per_denoise_step'], q['clocks'])
q=q['fp8']; print('fp8', q['value'], q['ms_per_step'], q['ms_per_denoise_step'])
PY
PROF_QUANT=4 timeout 400 python profiles/prof_step_kernels.py > gpurun_out/r02_step_kernels_q4.txt 2>/dev/null; head -14 gpurun_out/r02_step_kernels_q4.txt | cut -c1-175
