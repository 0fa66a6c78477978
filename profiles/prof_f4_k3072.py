"""One launch each of the K = 3072 NVFP4 GEMMs at the benchmark shapes (B=8, N=4352) for ncu: linear1's QKV rows (QKV epilogue,
e4m3 q/k/v), linear1's MLP rows (GELU epilogue), proj (gate + residual)."""
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
wproj, b2 = r(D, D, sc=D ** -0.5), r(D, sc=0.1)
qs, ks, pe, gate = r(128), r(128), r(N, 64, 2), r(B, D, sc=0.1)
q8, k8, v8 = (torch.empty(B, H, N, 128, device=dev, dtype=ops.fp8) for _ in range(3))
xm4, xmsf4, xms4 = ops.quantize_rows_fp4(xm)
wq4, wqsf, wqs4 = ops.fp4_weight(w1[:3 * D], ops.FP4_TILE_N_QKV)
wm4, wmsf, wms4 = ops.fp4_weight(w1[3 * D:])
wp4, wpsf, wps4 = ops.fp4_weight(wproj)
ca4, casf4, cas4 = ops.quantize_rows_fp4(cat[:, :, :D])
torch.cuda.synchronize()
print("setup done", flush=True)
for _ in range(int(os.environ.get("PROF_REPEAT", "1"))):
    ops.gemm_fp4_qkv(xm4, xmsf4, xms4, wq4, wqsf, wqs4, B, b1[:3 * D], qs, ks, pe, q8, k8, v8, 0)
    ops.gemm_fp4(xm4, xmsf4, xms4, wm4, wmsf, wms4, B, bias=b1[3 * D:], act="gelu_tanh", out=cat[:, :, D:])
    ops.gemm_fp4(ca4, casf4, cas4, wp4, wpsf, wps4, B, bias=b2, gate=gate, resid=x, out=x)
torch.cuda.synchronize()
