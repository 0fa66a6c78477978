"""One launch of each --quantize kernel at the benchmark shapes (B=8, N=4352) for `ncu --set full -k regex:...`:
FP8 linear1 (e4m3 q/k/v out), FP8 attention, FP8 linear2, NVFP4 linear2 (CTA pairs), the two `cat` quantisers, e4m3 rownorm."""
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
w1, b1 = r(3 * D + M, D, sc=D ** -0.5), r(3 * D + M, sc=0.1)
w2, b2 = r(D, D + M, sc=(D + M) ** -0.5), r(D, sc=0.1)
qs, ks, pe, gate = r(128), r(128), r(N, 64, 2), r(B, D, sc=0.1)
shift, scale = r(B, D, sc=0.1), r(B, D, sc=0.1)
q8, k8, v8 = (torch.empty(B, H, N, 128, device=dev, dtype=ops.fp8) for _ in range(3))
xm8, xs = ops.quantize_rows(xm)
w1q, w1s = ops.quantize_rows(w1)
w2q, w2s = ops.quantize_rows(w2)
w24, w2sf, w2s4 = ops.fp4_weight(w2)
torch.cuda.synchronize()
print("setup done", flush=True)
for _ in range(int(os.environ.get("PROF_REPEAT", "1"))):
    ops.rownorm(x, 0, shift, scale, 1e-6, out=xm8, out_scale=xs)
    ops.gemm_qkv(xm8, w1q, b1, qs, ks, pe, q8, k8, v8, 0, mlp_out=cat[:, :, D:], a_scale=xs, w_scale=w1s)
    ops.attention(q8, k8, v8, cat[:, :, :D], 128 ** -0.5)
    cat8, cs = ops.quantize_rows(cat)
    ops.gemm(cat8, w2q, b2, gate=gate, resid=x, out=x, a_scale=cs, w_scale=w2s)
    cat4, csf4, cs4 = ops.quantize_rows_fp4(cat)
    ops.gemm_fp4(cat4, csf4, cs4, w24, w2sf, w2s4, B, bias=b2, gate=gate, resid=x, out=x)
torch.cuda.synchronize()
