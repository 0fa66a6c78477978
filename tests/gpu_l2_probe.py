"""Launches linear1 / linear2 / fc1 (bf16 and FP8) of the BASELINE shape twice each: run under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_kernel` to read the DRAM
traffic per launch for a given L2 policy (FX_GEMM_HINT_A / FX_GEMM_HINT_W / FX_GEMM_STCS)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flux-generator_b200"))
from flux import ops  # noqa: E402

dev, bf = "cuda", torch.bfloat16
B, L, S, D, H, M = 8, 4096, 256, 3072, 24, 12288
N = L + S
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).to(bf)  # noqa: E731
x, xm, cat = r(B, N, D), r(B, N, D), r(B, N, D + M)
q, k, v = r(B, H, N, 128), r(B, H, N, 128), r(B, H, N, 128)
w1, b1 = r(3 * D + M, D, sc=D ** -0.5), r(3 * D + M, sc=0.1)
w2, b2 = r(D, D + M, sc=(D + M) ** -0.5), r(D, sc=0.1)
wf, bfc = r(M, D, sc=D ** -0.5), r(M, sc=0.1)
qs, ks, pe, gate = r(128), r(128), r(N, 64, 2), r(B, D, sc=0.1)
xm8, xs = ops.quantize_rows(xm)
cat8, cs = ops.quantize_rows(cat)
w1q, w1s = ops.quantize_rows(w1)
w2q, w2s = ops.quantize_rows(w2)
for _ in range(2):
    ops.gemm_qkv(xm, w1, b1, qs, ks, pe, q, k, v, 0, mlp_out=cat[:, :, D:])                       # linear1
    ops.gemm(cat, w2, b2, gate=gate, resid=x, out=x)                                              # linear2
    ops.gemm(xm[:, S:], wf, bfc, act="gelu_tanh", out=cat[:, S:, D:])                             # fc1
    ops.gemm_qkv(xm8, w1q, b1, qs, ks, pe, q, k, v, 0, mlp_out=cat[:, :, D:], a_scale=xs, w_scale=w1s)   # linear1 fp8
    ops.gemm(cat8, w2q, b2, gate=gate, resid=x, out=x, a_scale=cs, w_scale=w2s)                   # linear2 fp8
torch.cuda.synchronize()
