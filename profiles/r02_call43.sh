#!/bin/bash
# call 43: compute-sanitizer memcheck over the NVFP4 producers / kernels (small cases)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_fp4.py -m gpu -q --timeout 1200 -p no:cacheprovider -x \
  -k "chunk_quantiser or emits or rownorm_nvfp4 or (gemm_fp4_vs and (1-128 or 1-200 or 3-384)) or (qkv_epilogue and (1-128 or 1-200))" > gpurun_out/r02_sanitizer_nvfp4.txt 2>&1
echo "rc=$?" >> gpurun_out/r02_sanitizer_nvfp4.txt
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r02_sanitizer_nvfp4.txt; tail -6 gpurun_out/r02_sanitizer_nvfp4.txt
