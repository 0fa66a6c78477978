#!/bin/bash
# Runs the GPU bring-up probes stage by stage, each under its own timeout (a hung kernel must not
# take the box down).  Usage: bash tests/run_bringup.sh [stage ...]
mkdir -p gpurun_out
STAGES=${@:-"umma elem gemm qkv conv attn0 attn1"}
for s in $STAGES; do
  timeout 240 python tests/gpu_bringup.py $s 2>&1 | tail -n 80
  echo "== stage $s exit ${PIPESTATUS[0]}"
done 2>&1 | tee gpurun_out/bringup.log
