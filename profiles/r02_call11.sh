#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python tests/gpu_microbench.py linear2_f8 linear2_f4 fc2_f8 fc2_f4 quant_cat_f8 quant_cat_f4 linear2 > gpurun_out/c11_mb.log 2>&1
cat gpurun_out/c11_mb.log
