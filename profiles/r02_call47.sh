#!/bin/bash
# call 47: text-encoder attention on mma.sync (flash loop, K/V chunks in shared memory): parity + text-encode breakdown
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -p no:cacheprovider -k "small or t5 or clip or text or pipeline or T5" > gpurun_out/c47_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c47_tests.log); tail -5 gpurun_out/c47_tests.log
timeout 600 python profiles/prof_text_kernels.py > gpurun_out/r02_text_kernels_mma.txt 2>/dev/null; head -8 gpurun_out/r02_text_kernels_mma.txt | cut -c1-170
