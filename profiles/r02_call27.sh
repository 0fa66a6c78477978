#!/bin/bash
# call 27: in-situ (sustained, CUPTI) kernel breakdown of one denoise step: NVFP4, FP8, bf16
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for m in 4 1 0; do PROF_QUANT=$m timeout 400 python profiles/prof_step_kernels.py > gpurun_out/r02_step_kernels_q$m.txt 2> gpurun_out/c27_q$m.err; head -24 gpurun_out/r02_step_kernels_q$m.txt | cut -c1-175; tail -2 gpurun_out/c27_q$m.err; done
