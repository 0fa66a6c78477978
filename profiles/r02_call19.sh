#!/bin/bash
# call 19: uniform-control-flow issuers (no ELECT/R2UR broadcast loops around TMA / tcgen05 ops): parity + microbench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp8.py tests/test_gpu_fp4.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/c19_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c19_tests.log); tail -5 gpurun_out/c19_tests.log
timeout 600 python tests/gpu_microbench.py linear1 linear2 fc1 fc2 proj linear1_f8 linear2_f8 fc1_f8 fc2_f8 linear2_f4 fc2_f4 > gpurun_out/c19_mb.log 2>&1; cat gpurun_out/c19_mb.log
FX_GEMM4_NCTA=1 timeout 200 python tests/gpu_microbench.py linear2_f4 fc2_f4 >> gpurun_out/c19_mb.log 2>&1; tail -2 gpurun_out/c19_mb.log
