#!/bin/bash
# round 2, GPU call 9: FP8 attention (e4m3 q/k/v from the QKV epilogue, kind::f8f6f4 for both products)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fp8.py -m gpu -q -s --timeout 300 -p no:cacheprovider -k "attention_fp8 or e4m3_q_k_v or full_width" > gpurun_out/c9_new.log 2>&1; echo "rc=$?" >> gpurun_out/c9_new.log)
grep -E "fp8|passed|failed|Error|assert|rc=" gpurun_out/c9_new.log | head -40
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log)
tail -15 gpurun_out/c9_pytest.log
timeout 300 python tests/gpu_microbench.py attn attn_f8 attn attn_f8 > gpurun_out/c9_mb.log 2>&1
cat gpurun_out/c9_mb.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c9_bench.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], 'vae', d['ms_vae_decode_batch'], 'clk', d['clocks']['sm_mhz'], 'q', d['quantized']['value'], d['quantized']['ms_per_denoise_step'], d['quantized']['clocks']['sm_mhz'])
PY
cat gpurun_out/fullsize_parity.json | python -c "import json,sys; d=json.load(sys.stdin); print(json.dumps(d['fp8_full_depth_4_steps'], indent=0))"
