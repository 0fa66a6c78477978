"""Pipeline timeline of one attention CTA (profiling builds only: csrc built with -DFX_ATTN_PROBE into the
library FLUX_B200_LIB points at).  Writes gpurun_out/attn_probe.npy [role 3][step 64][event 8] (SM clocks)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flux-generator_b200"))
from flux import _native, ops  # noqa: E402

B, H, N = 8, 24, 4352
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(B, H, N, 128, device="cuda", generator=g).bfloat16() for _ in range(3))
out = torch.empty(B, N, H * 128, device="cuda", dtype=torch.bfloat16)
lib = _native.lib()
fn = lib.fx_dbg_attn_probe
fn.restype, fn.argtypes = C.c_int, [C.c_void_p]
buf = torch.zeros(3 * 64 * 8, device="cuda", dtype=torch.int64)
_native.check(fn(buf.data_ptr()))
VARIANT = int(os.environ.get("PROBE_VARIANT", "0"))  # 4: no exp-phase turn taking between the softmax warpgroups
for _ in range(5):
    ops.attention(q, k, v, out, 128 ** -0.5, variant=VARIANT)
torch.cuda.synchronize()
buf.zero_()
ops.attention(q, k, v, out, 128 ** -0.5, variant=VARIANT)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(3, 64, 8)
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/attn_probe.npy", t)
T = (N + 127) // 128
t0 = t[t > 0].min()
for j in list(range(0, 4)) + list(range(16, 20)) + [T - 1]:
    print(f"step {j}")
    for role, name in ((0, "mma "), (1, "sm0 "), (2, "sm1 ")):
        print("   ", name, " ".join(f"{int(x - t0):7d}" if x > 0 else "      -" for x in t[role, j, :7]))
d = np.diff(t[1, 4:T - 1, 0])
if (d > 0).all():
    print("sm0 step period (clks): mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))
d = np.diff(t[0, 4:T - 1, 0])
print("mma step period (clks): mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))
c0, c1, g0, g1 = t[0, 63, :4]
print("cta: %d clks in %d ns -> %.0f MHz; prologue %d clks, epilogue %d clks" % (
    c1 - c0, g1 - g0, (c1 - c0) / max(g1 - g0, 1) * 1e3, t[0, 0, 0] - c0, c1 - t[0, T - 1, 6]))
