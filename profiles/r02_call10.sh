#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_fp4.py -m gpu -q -s --timeout 120 -p no:cacheprovider -x > gpurun_out/c10_fp4.log 2>&1; echo "rc=$?" >> gpurun_out/c10_fp4.log)
tail -40 gpurun_out/c10_fp4.log
