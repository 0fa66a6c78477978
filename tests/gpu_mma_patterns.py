"""Tensor-pipe issue-pattern micro-benchmark (fx_dbg_mma_pattern): SM clocks per 32-MMA step for each pattern."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flux-generator_b200"))
from flux import _native  # noqa: E402

NAMES = {0: "attention order + commits", 1: "attention order, no commits", 2: "32 x SS, one accumulator",
         3: "32 x TS, one accumulator", 4: "SS alternating accumulators", 5: "TS alternating accumulators",
         6: "QK0 QK1 PV0 PV1", 7: "16 x SS N=256", 8: "attention order, commit every 4", 9: "two issuing threads (QK | PV)",
         10: "QK only (16 MMAs)", 11: "PV only (16 MMAs)"}
lib = _native.dbg_lib()
out = torch.zeros(1, device="cuda", dtype=torch.int64)
iters = 400
for pat in sorted(NAMES):
    for _ in range(2):
        _native.check(lib.fx_dbg_mma_pattern(pat, iters, out.data_ptr(), _native.stream()))
    torch.cuda.synchronize()
    print(f"pattern {pat:2d} {NAMES[pat]:36s} {out.item() / iters:8.1f} clks/step", flush=True)
