#!/bin/bash
# round 2, GPU call 2: new kernels (block-per-row rownorm, GroupNorm statistics in the conv epilogue, persistent attention)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -p no:cacheprovider -k "persistent or partials or wide_rows or test_attention" > gpurun_out/c2_new.log 2>&1; echo "rc=$?" >> gpurun_out/c2_new.log)
tail -25 gpurun_out/c2_new.log
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log)
tail -8 gpurun_out/c2_pytest.log
timeout 300 python tests/gpu_microbench.py rownorm rownorm_f8 attn attn_p attn attn_p > gpurun_out/c2_mb.log 2>&1
cat gpurun_out/c2_mb.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c2_bench.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], 'vae', d['ms_vae_decode_batch'], 'clk', d['clocks']['sm_mhz'], 'q', d['quantized']['value'])
print({k:(round(v['ms_per_step'],1), round(v['frac_of_peak'],3)) for k,v in d['kernels'].items() if v['ms_per_step']>1})
PY
FX_ATTN_PERSISTENT=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/c2_bench_p.json 2> gpurun_out/c2_bench_p.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c2_bench_p.json').read().strip().splitlines()[-1])
print('bench persistent attn', d['value'], d['ms_per_step'], 'clk', d['clocks']['sm_mhz'], 'q', d['quantized']['value'])
print({k:(round(v['ms_per_step'],1), round(v['frac_of_peak'],3)) for k,v in d['kernels'].items() if v['ms_per_step']>1})
PY
