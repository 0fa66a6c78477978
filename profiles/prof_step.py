"""One denoise step of Flux-schnell 1024x1024 (batch 8) + one VAE decode (batch 1), random weights filled
in place (no checkpoint generation), for ncu captures:

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      -k regex:KERNELS python profiles/prof_step.py
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 3 -o gpurun_out/gemm python profiles/prof_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flux-generator_b200"))
from flux import specs  # noqa: E402
from flux.autoencoder import AutoEncoder  # noqa: E402
from flux.model import Flux  # noqa: E402

dev = "cuda"
bf = torch.bfloat16
B = int(os.environ.get("PROF_BATCH", "8"))
depth = int(os.environ.get("PROF_DEPTH", "19"))
p = specs.FluxParams(depth=depth, depth_single_blocks=2 * depth)
model = Flux(p, device=dev)
model.arena.buffer.normal_(0, 0.02)
if os.environ.get("PROF_QUANT", "0") == "1":  # the --quantize (FP8) path
    model.quantize()
if os.environ.get("PROF_QUANT", "0") == "4":  # --quantize --quantize-bits 4 (NVFP4 proj / mlp.2 / linear2)
    model.quantize(bits=4)
L, S = 4096, 256
img = torch.randn(B, L, 64, device=dev, dtype=bf)
txt = torch.randn(B, S, 4096, device=dev, dtype=bf)
y = torch.randn(B, 768, device=dev, dtype=bf)
ids = torch.zeros(B, L, 3, dtype=torch.int32, device=dev)
ids[:, :, 1] = torch.arange(L, device=dev) // 64
ids[:, :, 2] = torch.arange(L, device=dev) % 64
tids = torch.zeros(B, S, 3, dtype=torch.int32, device=dev)
ts = torch.full((B,), 0.75, dtype=bf, device=dev)
if os.environ.get("PROF_FLOW", "1") == "1":
    model.forward(img, ids, txt, tids, ts, y)
if os.environ.get("PROF_VAE", "1") == "1":
    ae = AutoEncoder(specs.AutoEncoderParams(), device=dev)
    ae.arena.buffer.normal_(0, 0.02)
    ae.decode_packed(torch.randn(1, L, 64, device=dev, dtype=bf), (128, 128))
torch.cuda.synchronize()
print("done")
