#!/bin/bash
# call 20: NVFP4 for every block Linear: QKV-epilogue kernel (128-column tiles), parity + microbench + full-size accuracy
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fp4.py tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -x -s > gpurun_out/c20_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c20_tests.log); grep -i "nvfp4 (\|passed\|failed\|error\|rc=" gpurun_out/c20_tests.log | tail -12
timeout 600 python tests/gpu_microbench.py quant_x_f4 quant_cat_f4 qkv1_f4 qkv1_bf16out_f4 mlp1_f4 qkv_img_f4 fc1_f4 proj_f4 linear2_f4 fc2_f4 rownorm rownorm_f8 > gpurun_out/c20_mb.log 2>&1; cat gpurun_out/c20_mb.log
(timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "fp8_full_depth" > gpurun_out/c20_full.log 2>&1; echo "rc=$?" >> gpurun_out/c20_full.log); tail -5 gpurun_out/c20_full.log
