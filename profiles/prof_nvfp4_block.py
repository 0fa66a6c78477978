"""One single-stream block of the `--quantize` (NVFP4) step at the benchmark shapes (B=8, N=4352), kernel by kernel, for
`ncu --set full`: AdaLN row norm -> NVFP4 operand, QKV GEMM (QK-RMSNorm + RoPE, e4m3 q/k/v), MLP GEMM whose GELU epilogue emits
linear2's operand, FP8 attention, chunk quantiser of the attention output, finalise, linear2 (gate + residual)."""
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
x, cat = r(B, N, D), r(B, N, D + M)
w1, b1 = r(3 * D + M, D, sc=D ** -0.5), r(3 * D + M, sc=0.1)
w2, b2 = r(D, D + M, sc=(D + M) ** -0.5), r(D, sc=0.1)
qs, ks, pe, gate = r(128), r(128), r(N, 64, 2), r(B, D, sc=0.1)
shift, scale = r(B, D, sc=0.1), r(B, D, sc=0.1)
q8, k8, v8 = (torch.empty(B, H, N, 128, device=dev, dtype=ops.fp8) for _ in range(3))
wq4, wqsf, wqs4 = ops.fp4_weight(w1[:3 * D], ops.FP4_TILE_N_QKV)
wm4, wmsf, wms4 = ops.fp4_weight(w1[3 * D:])
w24, w2sf, w2s4 = ops.fp4_weight(w2)
a4buf = (torch.empty(B * N * D // 2, device=dev, dtype=torch.uint8), torch.empty(B * N * D // 16, device=dev, dtype=torch.uint8),
         torch.empty(B * N, device=dev))
c4 = ops.Fp4Operand(B * N, D + M, dev)
torch.cuda.synchronize()
print("setup done", flush=True)
for _ in range(int(os.environ.get("PROF_REPEAT", "1"))):
    a4, sfa, sa = ops.rownorm(x, 0, shift, scale, 1e-6, out_fp4=a4buf)
    ops.gemm_fp4_qkv(a4, sfa, sa, wq4, wqsf, wqs4, B, b1[:3 * D], qs, ks, pe, q8, k8, v8, 0)
    dst = c4.view(B * N, D + M)
    ops.gemm_fp4(a4, sfa, sa, wm4, wmsf, wms4, B, bias=b1[3 * D:], act="gelu_tanh", out4=dst, out4_col0=D)
    ops.attention(q8, k8, v8, cat[:, :, :D], 128 ** -0.5)
    ops.quantize_chunks_fp4(cat[:, :, :D], dst, 0)
    c, sfc, sc = ops.fp4_finalize(dst)
    ops.gemm_fp4(c, sfc, sc, w24, w2sf, w2s4, B, bias=b2, gate=gate, resid=x, out=x)
torch.cuda.synchronize()
