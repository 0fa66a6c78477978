#!/bin/bash
# call 21: what bounds the K = 3072 NVFP4 GEMMs: main loop alone (FX_GEMM4_DBG_NOEPI) vs with epilogue; ncu --set full + source page
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tests/gpu_microbench.py qkv1_f4 mlp1_f4 fc1_f4 proj_f4 linear2_f4 > gpurun_out/c21_mb.log 2>&1
echo "--- FX_GEMM4_DBG_NOEPI=1" >> gpurun_out/c21_mb.log
FX_GEMM4_DBG_NOEPI=1 timeout 300 python tests/gpu_microbench.py qkv1_f4 mlp1_f4 fc1_f4 proj_f4 linear2_f4 >> gpurun_out/c21_mb.log 2>&1
echo "--- FX_GEMM4_NCTA=1" >> gpurun_out/c21_mb.log
FX_GEMM4_NCTA=1 timeout 300 python tests/gpu_microbench.py qkv1_f4 mlp1_f4 >> gpurun_out/c21_mb.log 2>&1
cat gpurun_out/c21_mb.log
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:gemm_nvfp4' -c 3 -o gpurun_out/f4k python profiles/prof_f4_k3072.py > gpurun_out/c21.log 2>&1
tail -3 gpurun_out/c21.log
ncu -i gpurun_out/f4k.ncu-rep --page raw --csv > gpurun_out/f4k.raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/f4k.raw.csv > gpurun_out/r02_ncu_f4_k3072.txt; cut -c1-230 gpurun_out/r02_ncu_f4_k3072.txt
python - <<'PY' >> gpurun_out/r02_ncu_f4_k3072.txt
import csv
rows=list(csv.reader(open('gpurun_out/f4k.raw.csv')))
h=rows[0]; ix={k:i for i,k in enumerate(h)}
keys=[k for k in h if any(s in k for s in ('l1tex__m_xbar2l1tex_read_bytes.sum','smsp__average_warps_issue_stalled','lts__throughput.avg.pct','sm__inst_executed_pipe_uniform','smsp__inst_executed.sum','sm__cycles_elapsed.max','smsp__cycles_active.avg'))]
for r in rows[2:]:
    print('##', r[ix['Kernel Name']][:60])
    for k in keys: print('  ', k, r[ix[k]], rows[1][ix[k]])
PY
for i in 0 1; do ncu -i gpurun_out/f4k.ncu-rep --page source --csv --print-source sass --launch-skip $i --launch-count 1 > gpurun_out/f4k_src$i.csv 2>/dev/null; done
python - <<'PY' >> gpurun_out/r02_ncu_f4_k3072.txt
import csv
for i in (0,1):
    rows=list(csv.reader(open(f'gpurun_out/f4k_src{i}.csv')))
    hi=[j for j,r in enumerate(rows) if 'Source' in r and any('Sampl' in c for c in r)]
    if not hi: print('no source table', i); continue
    h=rows[hi[0]]; ix={k:j for j,k in enumerate(h)}
    sc=[k for k in h if k.startswith('# Samples') or k=='Warp Stall Sampling (All Samples)' or 'Sampling (All' in k]
    col=ix[sc[0]] if sc else None
    print(f'## launch {i}: top SASS lines by stall samples ({sc[0] if sc else None})')
    body=[r for r in rows[hi[0]+1:] if len(r)>col and r[col].replace(',','').isdigit()]
    tot=sum(int(r[col].replace(',','')) for r in body)
    body.sort(key=lambda r:-int(r[col].replace(',','')))
    for r in body[:40]:
        print(f"  {int(r[col].replace(',','')):7d} {100*int(r[col].replace(',',''))/max(tot,1):5.1f}%  {r[ix['Source']][:110]}")
PY
tail -90 gpurun_out/r02_ncu_f4_k3072.txt | cut -c1-200
rm -f gpurun_out/f4k.raw.csv
