"""A few launches of the attention kernel at the benchmark shape (B=8, H=24, N=4352) for ncu captures.
Usage: python tests/gpu_attn_one.py [variant] [launches]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flux-generator_b200"))
from flux import ops  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B, H, N = 8, 24, 4352
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(B, H, N, 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
out = torch.empty(B, N, H * 128, device="cuda", dtype=torch.bfloat16)
for _ in range(n):
    ops.attention(q, k, v, out, 128 ** -0.5, variant=variant)
torch.cuda.synchronize()
