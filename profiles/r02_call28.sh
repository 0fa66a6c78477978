#!/bin/bash
# call 28: NVFP4 kernel with 12 epilogue warps / 3 accumulators (QKV), packed dequant, carried tile coordinates; packed-register quantiser
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp8.py tests/test_gpu_fp4.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/c28_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c28_tests.log); tail -4 gpurun_out/c28_tests.log
timeout 900 python tests/gpu_microbench.py quant_cat_f4 quant_x_f4 linear1_f8 linear2_f8 qkv1_f4 qkv1_bf16out_f4 mlp1_f4 qkv_img_f4 fc1_f4 proj_f4 linear2_f4 fc2_f4 > gpurun_out/c28_mb.log 2>&1; cat gpurun_out/c28_mb.log
