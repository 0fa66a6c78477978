#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 python tests/gpu_bs_probe.py pair > gpurun_out/c17_probe.log 2>&1; cat gpurun_out/c17_probe.log
(timeout 300 python -m pytest tests/test_gpu_fp4.py -m gpu -q --timeout 120 -p no:cacheprovider -x > gpurun_out/c17_fp4.log 2>&1; echo "rc=$?" >> gpurun_out/c17_fp4.log); tail -6 gpurun_out/c17_fp4.log
timeout 300 python tests/gpu_microbench.py quant_cat_f8 quant_cat_f4 linear2_f8 linear2_f4 fc2_f8 fc2_f4 > gpurun_out/c17_mb.log 2>&1; cat gpurun_out/c17_mb.log
FX_GEMM4_NCTA=1 timeout 200 python tests/gpu_microbench.py linear2_f4 fc2_f4 >> gpurun_out/c17_mb.log 2>&1; tail -2 gpurun_out/c17_mb.log
