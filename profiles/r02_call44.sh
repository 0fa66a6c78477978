#!/bin/bash
# call 44: ncu launch lists (gpu__time_duration.sum) of one full-depth step in the final build: bf16 (+ VAE) and --quantize (NVFP4)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum --clock-control none --csv"
PROF_DEPTH=19 PROF_VAE=1 timeout 600 ncu $M --log-file gpurun_out/r02_launches_step_final.csv python profiles/prof_step.py > /dev/null 2>&1
PROF_DEPTH=19 PROF_VAE=0 PROF_QUANT=4 timeout 600 ncu $M --log-file gpurun_out/r02_launches_step_nvfp4.csv python profiles/prof_step.py > /dev/null 2>&1
wc -l gpurun_out/r02_launches_step_final.csv gpurun_out/r02_launches_step_nvfp4.csv
python - <<'PY'
import csv, collections, re
for f in ('gpurun_out/r02_launches_step_final.csv','gpurun_out/r02_launches_step_nvfp4.csv'):
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
    ix={k:i for i,k in enumerate(rows[h])}
    agg=collections.defaultdict(lambda:[0,0.0])
    for r in rows[h+1:]:
        try: v=float(r[ix['Metric Value']].replace(',',''))
        except: continue
        u=r[ix['Metric Unit']]; v*= {'ns':1e-6,'us':1e-3,'ms':1,'nsecond':1e-6,'usecond':1e-3,'msecond':1}.get(u,1e-6)
        n=re.sub(r'\(.*','',r[ix['Kernel Name']])[:60]; agg[n][0]+=1; agg[n][1]+=v
    tot=sum(v[1] for v in agg.values()); print(f, 'total ms', round(tot,2))
    for n,(c,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]: print(f"   {n:62s} {c:5d} {ms:9.3f} {100*ms/tot:5.1f}%")
PY
