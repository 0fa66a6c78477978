#!/bin/bash
# round 2, GPU call 15: clusters of two CTA pairs with W-tile multicast (FX_GEMM_CL=2): correctness, then sustained time
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(FX_GEMM_CL=2 FX_GEMM_CL_MIN_TILES=1 timeout 240 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp8.py -m gpu -q --timeout 60 -p no:cacheprovider -k "gemm" > gpurun_out/c15_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c15_tests.log)
tail -25 gpurun_out/c15_tests.log
for cl in 1 2; do echo "FX_GEMM_CL=$cl"; FX_GEMM_CL=$cl timeout 300 python tests/gpu_microbench.py linear1 linear2 fc1 fc2 proj linear1_f8 linear2_f8 fc1_f8 2>&1 | grep -v "^$"; done > gpurun_out/c15_mb.log 2>&1
cat gpurun_out/c15_mb.log
