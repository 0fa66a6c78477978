#!/bin/bash
# call 24: source-level stall samples of the NVFP4 GELU-epilogue GEMM (mlp rows of linear1) after the latency-hiding changes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:gemm_nvfp4' -s 1 -c 1 -o gpurun_out/f4m python profiles/prof_f4_k3072.py > gpurun_out/c24.log 2>&1
tail -2 gpurun_out/c24.log
ncu -i gpurun_out/f4m.ncu-rep --page source --csv --print-source sass > gpurun_out/f4m_src.csv 2>/dev/null
ncu -i gpurun_out/f4m.ncu-rep --page raw --csv > gpurun_out/f4m.raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/f4m.raw.csv
rm -f gpurun_out/f4m.ncu-rep gpurun_out/f4m.raw.csv
