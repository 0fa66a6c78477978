#!/bin/bash
# call 49: compute-sanitizer memcheck + racecheck over the text-encoder attention kernel (attn_small_mma_kernel) and the NVFP4 finalise
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 800 -p no:cacheprovider -x -k "attention_small" > gpurun_out/r02_sanitizer_attn_small_$tool.txt 2>&1
  echo "rc=$?" >> gpurun_out/r02_sanitizer_attn_small_$tool.txt
  tail -4 gpurun_out/r02_sanitizer_attn_small_$tool.txt
done
