#!/bin/bash
# round 2, GPU call 3: persistent rownorm, rewritten GroupNorm apply, persistent attention as the default; ncu of attention
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log)
tail -8 gpurun_out/c3_pytest.log
for bps in 4 8 12 16; do echo "rownorm blocks/SM $bps"; FX_ROWNORM_BLOCKS_PER_SM=$bps timeout 120 python tests/gpu_microbench.py rownorm rownorm_f8 2>&1 | grep -v "^$"; done > gpurun_out/c3_mb.log 2>&1
cat gpurun_out/c3_mb.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c3_bench.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], 'vae', d['ms_vae_decode_batch'], 'clk', d['clocks']['sm_mhz'], 'q', d['quantized']['value'], d['quantized']['ms_per_denoise_step'])
print({k:(round(v['ms_per_step'],1), round(v['frac_of_peak'],3)) for k,v in d['kernels'].items() if v['ms_per_step']>1})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_pkernel -s 2 -c 1 -o gpurun_out/c3_attn_p python tests/gpu_attn_one.py 7 4 > gpurun_out/c3_ncu.log 2>&1
tail -3 gpurun_out/c3_ncu.log
ls -la gpurun_out/*.ncu-rep
