#!/bin/bash
# call 53: W scale atoms multicast inside the CTA pair (one L2 read instead of two): parity + A/B microbench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fp4.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/c53_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c53_tests.log); tail -3 gpurun_out/c53_tests.log
for v in 1 0 1 0; do echo "--- FX_GEMM4_SFB_MCAST=$v" >> gpurun_out/c53_mb.log; FX_GEMM4_SFB_MCAST=$v timeout 300 python tests/gpu_microbench.py linear2_f4 fc2_f4 qkv1_f4 fc1_f4 >> gpurun_out/c53_mb.log 2>&1; done
cat gpurun_out/c53_mb.log
