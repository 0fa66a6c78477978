"""In-situ (CUPTI) kernel breakdown of the text encoders for one fresh prompt (T5-XXL + CLIP-L shapes, synthetic weights).
Usage: python profiles/prof_text_kernels.py > profiles/r02_text_kernels.txt"""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flux-generator_b200"))
from flux import FluxPipeline  # noqa: E402

pipe = FluxPipeline("flux-schnell", synthetic=True)
pipe.ensure_models_are_loaded()
t5_tok, clip_tok = pipe.tokenize("a photograph of an astronaut riding a horse through a field of sunflowers")
for _ in range(3):
    pipe.t5(t5_tok)
    pipe.clip(clip_tok)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n_prof = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(n_prof):
        pipe.t5(t5_tok)
        pipe.clip(clip_tok)
    e1.record()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type.name != "CUDA" or ev.device_time <= 0:
        continue
    name = re.sub(r"\(.*", "", re.sub(r"^void ", "", ev.name))
    agg[name][0] += 1
    agg[name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"# T5-XXL (S = {t5_tok.shape[1]}) + CLIP-L (S = {clip_tok.shape[1]}): {e0.elapsed_time(e1) / n_prof:.2f} ms per prompt (events), kernel time {tot / 1e3 / n_prof:.2f} ms")
print("# kernel | launches per prompt | ms per prompt | share | us per launch")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:100]:100s} | {n / n_prof:7.1f} | {us / 1e3 / n_prof:8.3f} | {100 * us / tot:5.1f}% | {us / n:8.1f}")
