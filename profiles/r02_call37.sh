#!/bin/bash
# call 37: ncu --set full of every kernel of one NVFP4 single-stream block (final state of the round)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k 'regex:(rownorm_block_kernel|gemm_nvfp4|attn_pkernel|quantize_chunks|fp4_finalize)' -c 7 -o gpurun_out/f4blk python profiles/prof_nvfp4_block.py > gpurun_out/c37.log 2>&1
tail -2 gpurun_out/c37.log
ncu -i gpurun_out/f4blk.ncu-rep --page raw --csv > gpurun_out/f4blk.raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/f4blk.raw.csv > gpurun_out/r02_ncu_nvfp4_block.txt; cut -c1-250 gpurun_out/r02_ncu_nvfp4_block.txt
rm -f gpurun_out/f4blk.ncu-rep gpurun_out/f4blk.raw.csv
timeout 300 python tests/gpu_microbench.py qkv1_f4 mlp1_f4 linear2_f4 fc2_f4 rownorm_f4 > gpurun_out/c37_mb.log 2>&1; cat gpurun_out/c37_mb.log
