#!/bin/bash
# call 33: HBM-bound NVFP4 producers: quantiser at 4 blocks / SM, row norm at 4 / 5 / 6 blocks per SM
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tests/gpu_microbench.py quant_cat_f4 quant_mlp_f4 quant_x_f4 rownorm_f4 rownorm qkv1_f4 > gpurun_out/c33_mb.log 2>&1
for n in 5 6; do echo "--- FX_ROWNORM_BLOCKS_PER_SM=$n" >> gpurun_out/c33_mb.log; FX_ROWNORM_BLOCKS_PER_SM=$n timeout 300 python tests/gpu_microbench.py rownorm_f4 rownorm >> gpurun_out/c33_mb.log 2>&1; done
cat gpurun_out/c33_mb.log
(timeout 600 python -m pytest tests/test_gpu_fp4.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/c33_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c33_tests.log); tail -3 gpurun_out/c33_tests.log
