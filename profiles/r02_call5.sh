#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for bps in 4 5 6; do echo "rownorm blocks/SM $bps"; FX_ROWNORM_BLOCKS_PER_SM=$bps timeout 120 python tests/gpu_microbench.py rownorm rownorm_f8 2>&1 | grep -v "^$"; done > gpurun_out/c5_mb.log 2>&1
cat gpurun_out/c5_mb.log
