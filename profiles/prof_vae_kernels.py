"""In-situ (CUPTI) kernel breakdown of the VAE decode of 8 images at 1024x1024 (latent 128x128), sustained regime.
Usage: python profiles/prof_vae_kernels.py > profiles/r02_vae_kernels.txt"""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flux-generator_b200"))
from flux import specs  # noqa: E402
from flux.autoencoder import AutoEncoder  # noqa: E402

dev, bf = "cuda", torch.bfloat16
B = int(os.environ.get("PROF_BATCH", "8"))
ae = AutoEncoder(specs.AutoEncoderParams(), device=dev)
ae.arena.buffer.normal_(0, 0.02)
z = torch.randn(B, 4096, 64, device=dev, dtype=bf)
for _ in range(6):
    ae.decode_packed(z, (128, 128))
torch.cuda.synchronize()
n_prof = 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(n_prof):
        ae.decode_packed(z, (128, 128))
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
print(f"# VAE decode, batch {B}: {e0.elapsed_time(e1) / n_prof:.2f} ms per decode (events), kernel time {tot / 1e3 / n_prof:.2f} ms")
print("# kernel | launches per decode | ms per decode | share | us per launch")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:100]:100s} | {n / n_prof:7.1f} | {us / 1e3 / n_prof:8.3f} | {100 * us / tot:5.1f}% | {us / n:8.1f}")
