#!/bin/bash
# round 2, GPU call 6: VAE encoder + training-loss forward, request coalescing, rownorm default; full suite + smoke
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log)
tail -30 gpurun_out/c6_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
