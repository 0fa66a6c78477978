#!/bin/bash
# call 29: where the epilogue's cost comes from: no GELU / no global stores / no epilogue
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tests/gpu_microbench.py mlp1_f4 mlp1_noact_f4 proj_f4 qkv1_f4 > gpurun_out/c29_mb.log 2>&1
echo "--- FX_GEMM4_DBG_NOEPI=2 (no global stores)" >> gpurun_out/c29_mb.log
FX_GEMM4_DBG_NOEPI=2 timeout 300 python tests/gpu_microbench.py mlp1_f4 mlp1_noact_f4 proj_f4 qkv1_f4 >> gpurun_out/c29_mb.log 2>&1
echo "--- FX_GEMM4_DBG_NOEPI=1 (no epilogue)" >> gpurun_out/c29_mb.log
FX_GEMM4_DBG_NOEPI=1 timeout 300 python tests/gpu_microbench.py mlp1_f4 proj_f4 qkv1_f4 >> gpurun_out/c29_mb.log 2>&1
cat gpurun_out/c29_mb.log
