#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k 'regex:(rownorm_block_kernel|gemm_kernel|attn_pkernel|quantize_rows_block_kernel<8>|quantize_rows_fp4_kernel<8>|gemm_nvfp4)' -s 2 -c 7 -o gpurun_out/quant python profiles/prof_quant_kernels.py > gpurun_out/c18.log 2>&1
tail -3 gpurun_out/c18.log
ncu -i gpurun_out/quant.ncu-rep --page raw --csv > gpurun_out/quant.raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/quant.raw.csv > gpurun_out/r02_ncu_quant.txt; cat gpurun_out/r02_ncu_quant.txt | cut -c1-230
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/quant.raw.csv')))
h=rows[0]; ix={k:i for i,k in enumerate(h)}
keys=[k for k in h if any(s in k for s in ('lts__t_bytes.sum','lts__t_sectors_srcunit_tex','l1tex__m_xbar2l1tex_read_bytes.sum','smsp__average_warps_issue_stalled','lts__throughput.avg.pct','sm__inst_executed_pipe_uniform'))]
for r in rows[2:]:
    if 'nvfp4' in r[ix['Kernel Name']]:
        for k in keys: print(k, r[ix[k]], rows[1][ix[k]])
PY
rm -f gpurun_out/quant.ncu-rep gpurun_out/quant.raw.csv
