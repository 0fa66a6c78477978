#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tests/gpu_bs_probe.py > gpurun_out/c8_probe.log 2>&1
cat gpurun_out/c8_probe.log
