#!/bin/bash
# round 2, GPU call 7: fused upsample convolution (parity-wise 2x2 kernels); full suite; bench A/B on one box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log)
tail -30 gpurun_out/c7_pytest.log
for up in 0 1; do
FLUX_B200_UPCONV=$up timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-quantized > gpurun_out/c7_bench_$up.json 2> gpurun_out/c7_bench_$up.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c7_bench_$up.json').read().strip().splitlines()[-1])
print('UPCONV=$up bench', d['value'], d['ms_per_step'], 'vae', d['ms_vae_decode_batch'], 'clk', d['clocks']['sm_mhz'])
print({k:(round(v['ms_per_step'],1), round(v['frac_of_peak'],3)) for k,v in d['kernels'].items() if v['ms_per_step']>1})
PY
done
