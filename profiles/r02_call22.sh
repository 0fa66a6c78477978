#!/bin/bash
# call 22: shared-address-space staging (LDS/STS instead of generic LD.E/ST.E) in every tcgen05 kernel: parity + microbench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp8.py tests/test_gpu_fp4.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/c22_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c22_tests.log); tail -4 gpurun_out/c22_tests.log
timeout 900 python tests/gpu_microbench.py linear1 linear2 fc1 fc2 proj attn attn_f8 linear1_f8 linear2_f8 fc1_f8 fc2_f8 qkv1_f4 mlp1_f4 qkv_img_f4 fc1_f4 proj_f4 linear2_f4 fc2_f4 > gpurun_out/c22_mb.log 2>&1; cat gpurun_out/c22_mb.log
