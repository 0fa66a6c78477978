#!/bin/bash
# call 31: TMA-store epilogue of the NVFP4 generic GEMM: parity + microbench (FX_GEMM4_TMA_OUT=0 = LDS/STG path)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fp4.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/c31_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c31_tests.log); tail -12 gpurun_out/c31_tests.log
timeout 600 python tests/gpu_microbench.py mlp1_f4 fc1_f4 proj_f4 linear2_f4 fc2_f4 > gpurun_out/c31_mb.log 2>&1
echo "--- FX_GEMM4_TMA_OUT=0" >> gpurun_out/c31_mb.log
FX_GEMM4_TMA_OUT=0 timeout 600 python tests/gpu_microbench.py mlp1_f4 fc1_f4 proj_f4 linear2_f4 fc2_f4 >> gpurun_out/c31_mb.log 2>&1
cat gpurun_out/c31_mb.log
