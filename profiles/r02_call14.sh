#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fp4.py -m gpu -q -s --timeout 300 -p no:cacheprovider > gpurun_out/c14_fp4.log 2>&1; echo "rc=$?" >> gpurun_out/c14_fp4.log)
grep -E "nvfp4|NVFP4|passed|failed|Error|assert|rc=" gpurun_out/c14_fp4.log | head -30
(timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 500 -p no:cacheprovider -k "fp8_full_depth" > gpurun_out/c14_full.log 2>&1; echo "rc=$?" >> gpurun_out/c14_full.log)
tail -5 gpurun_out/c14_full.log
python -c "import json; d=json.load(open('gpurun_out/fullsize_parity.json'))['fp8_full_depth_4_steps']; print({k:v for k,v in d.items() if 'nvfp4' in k or k.startswith('fp8_vs_fp32') or 'image' in k})"
timeout 300 python tests/gpu_microbench.py quant_cat_f8 quant_cat_f4 linear2_f8 linear2_f4 > gpurun_out/c14_mb.log 2>&1; cat gpurun_out/c14_mb.log
