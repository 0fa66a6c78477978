#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for sk in 0 1; do echo "FX_GEMM_DBG_SKIP_W=$sk"; FX_GEMM_DBG_SKIP_W=$sk timeout 300 python tests/gpu_microbench.py linear2 linear2_f8 linear1 fc1 2>&1 | grep -v "^$"; done > gpurun_out/c12_mb.log 2>&1
cat gpurun_out/c12_mb.log
