#!/bin/bash
# call 36: full-size accuracy of the fused NVFP4 mode + the rest of the full-size suite; GEMM microbench of the emitting epilogue
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py tests/test_gpu_parity.py -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/c36_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c36_tests.log); tail -4 gpurun_out/c36_tests.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/fullsize_parity.json'))
for k,v in d.get('fp8_full_depth_4_steps',{}).items():
    if 'nvfp4' in k: print(k, v)
PY
