#!/bin/bash
# call 41: parity of the attention-emitted operands: fp4 / fp8 / kernel suites + full-size accuracy
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests/test_gpu_fp4.py tests/test_gpu_fp8.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -q --timeout 900 -p no:cacheprovider -s > gpurun_out/c41_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c41_tests.log); grep -i "attention-emitted\|nvfp4 (\|passed\|failed\|rc=" gpurun_out/c41_tests.log | tail -14
python - <<'PY'
import json
d=json.load(open('gpurun_out/fullsize_parity.json'))
for k,v in d.get('fp8_full_depth_4_steps',{}).items():
    if 'nvfp4' in k: print(k, v)
PY
