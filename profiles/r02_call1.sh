#!/bin/bash
# round 2, GPU call 1: full GPU test suite (incl. the new full-size parity tests), isolated kernel numbers, the GEMM
# column-band raster sweep (sustained time + DRAM traffic under ncu), one short bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_env.log 2>&1
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log)
tail -15 gpurun_out/c1_pytest.log
timeout 400 python tests/gpu_microbench.py rownorm rownorm_f8 quant_cat_f8 attn linear1 linear2 fc2 proj fc1 > gpurun_out/c1_mb.log 2>&1
cat gpurun_out/c1_mb.log
timeout 500 python tests/gpu_microbench.py --raster-sweep 0:0,6:12,6:24,4:18,4:36,3:24,3:48,2:37,6:6 linear2 fc2 > gpurun_out/c1_sweep.log 2>&1
cat gpurun_out/c1_sweep.log
for cfg in "0 0" "6 12" "4 18" "3 24"; do
  set -- $cfg
  FX_GEMM_GROUP_N=$1 FX_GEMM_GROUP_M_BAND=$2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    -k regex:gemm_kernel --csv --log-file gpurun_out/c1_traffic_$1_$2.csv python tests/gpu_l2_probe.py > /dev/null 2>&1
  python profiles/parse_traffic.py gpurun_out/c1_traffic_$1_$2.csv gpurun_out/c1_traffic_$1_$2.json | grep -E '"(linear1|linear2|fc1)"|amplification|ncu_ms' | tr -d '\n'; echo
done
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -c 3000 gpurun_out/c1_bench.json; tail -5 gpurun_out/c1_bench.err
