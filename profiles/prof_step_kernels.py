"""In-situ kernel breakdown of one denoise step (batch 8, 1024x1024, 19 + 38 blocks) in the sustained, power-capped regime:
warm forwards, then torch.profiler (CUPTI activity records: no replay, no serialisation) over PROF_N forwards, summed by kernel.
PROF_QUANT = 0 (bf16) | 1 (FP8) | 4 (NVFP4).  Usage: python profiles/prof_step_kernels.py > profiles/r02_step_kernels_<mode>.txt"""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flux-generator_b200"))
from flux import specs  # noqa: E402
from flux.model import Flux  # noqa: E402

dev, bf = "cuda", torch.bfloat16
B, L, S = 8, 4096, 256
mode = os.environ.get("PROF_QUANT", "0")
n_prof = int(os.environ.get("PROF_N", "3"))
model = Flux(specs.FluxParams(depth=19, depth_single_blocks=38), device=dev)
model.arena.buffer.normal_(0, 0.02)
if mode == "1":
    model.quantize()
elif mode == "4":
    model.quantize(bits=4)
img = torch.randn(B, L, 64, device=dev, dtype=bf)
txt = torch.randn(B, S, 4096, device=dev, dtype=bf)
y = torch.randn(B, 768, device=dev, dtype=bf)
ids = torch.zeros(B, L, 3, dtype=torch.int32, device=dev)
ids[:, :, 1] = torch.arange(L, device=dev) // 64
ids[:, :, 2] = torch.arange(L, device=dev) % 64
tids = torch.zeros(B, S, 3, dtype=torch.int32, device=dev)
ts = torch.full((B,), 0.75, dtype=bf, device=dev)
for _ in range(8):
    model.forward(img, ids, txt, tids, ts, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(n_prof):
        model.forward(img, ids, txt, tids, ts, y)
    e1.record()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type.name != "CUDA" or ev.device_time <= 0:
        continue
    name = re.sub(r"^void ", "", ev.name)
    name = re.sub(r"\(.*", "", name)
    agg[name][0] += 1
    agg[name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"# mode {mode}: {e0.elapsed_time(e1) / n_prof:.2f} ms per forward (events), kernel time {tot / 1e3 / n_prof:.2f} ms per forward")
print("# kernel | launches per forward | ms per forward | share | us per launch")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:100]:100s} | {n / n_prof:7.1f} | {us / 1e3 / n_prof:8.3f} | {100 * us / tot:5.1f}% | {us / n:8.1f}")
