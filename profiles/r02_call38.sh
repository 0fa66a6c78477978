#!/bin/bash
# call 38: FP8 attention: share of exponentials on the FMA pipe (quadratic software exp2) 0 / 25 / 50 / 75 %
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/c38_mb.log
for v in "" _emucubic _emu0x00 _emu0x33 _emu0x55 _emu0x77; do
  echo "--- libflux_b200$v.so" >> gpurun_out/c38_mb.log
  FLUX_B200_LIB=$PWD/flux-generator_b200/flux/libflux_b200$v.so MB_SECONDS=2 timeout 200 python tests/gpu_microbench.py attn_f8 >> gpurun_out/c38_mb.log 2>&1
done
cat gpurun_out/c38_mb.log
