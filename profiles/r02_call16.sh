#!/bin/bash
# round 2, GPU call 16: ncu evidence for every kernel family (launch lists of a full step, --set full summaries)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum --clock-control none --csv"
PROF_DEPTH=19 PROF_VAE=1 timeout 600 ncu $M --log-file gpurun_out/r02_launches_step.csv python profiles/prof_step.py > /dev/null 2>&1
PROF_DEPTH=19 PROF_VAE=0 PROF_QUANT=1 timeout 600 ncu $M --log-file gpurun_out/r02_launches_step_fp8.csv python profiles/prof_step.py > /dev/null 2>&1
F="--set full --clock-control none"
PROF_DEPTH=1 PROF_VAE=0 timeout 600 ncu $F -c 60 -o gpurun_out/flow python profiles/prof_step.py > gpurun_out/c16_a.log 2>&1
PROF_DEPTH=1 PROF_VAE=0 PROF_QUANT=4 timeout 600 ncu $F -c 80 -o gpurun_out/flow4 python profiles/prof_step.py > gpurun_out/c16_b.log 2>&1
PROF_FLOW=0 PROF_DEPTH=1 timeout 900 ncu $F -c 140 -o gpurun_out/vae python profiles/prof_step.py > gpurun_out/c16_c.log 2>&1
for n in flow flow4 vae; do
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$n.raw.csv 2>/dev/null
  python profiles/ncu_summary.py gpurun_out/r02_ncu_$n.raw.csv > gpurun_out/r02_ncu_$n.txt
  rm -f gpurun_out/$n.ncu-rep gpurun_out/r02_ncu_$n.raw.csv
  wc -l gpurun_out/r02_ncu_$n.txt
done
tail -2 gpurun_out/c16_a.log gpurun_out/c16_c.log
ls -la gpurun_out | tail -12
