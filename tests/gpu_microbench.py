"""Sustained-regime micro-benchmarks of the hot kernels at the BASELINE shapes (batch 8, 1024x1024).
Each op loops for >= SECONDS so the clocks settle under the power cap; prints TFLOP/s.
Usage: python tests/gpu_microbench.py [ops...]   (FLUX_B200_LIB selects an alternative build of the library)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flux-generator_b200"))
from flux import _native, ops  # noqa: E402

try:
    import pynvml
    pynvml.nvmlInit()
    _nv = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:  # noqa: BLE001
    _nv = None

dev, bf = "cuda", torch.bfloat16
SECONDS = float(os.environ.get("MB_SECONDS", "1.5"))
B, L, S, D, H, M = 8, 4096, 256, 3072, 24, 12288
N = L + S


def sustained(fn, flops):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    n = 0
    while time.time() - t0 < SECONDS:  # launches are async: queue depth stays bounded by the sync below
        for _ in range(10):
            fn()
        n += 10
        if n % 50 == 0:
            torch.cuda.current_stream().synchronize()
            for _ in range(30):
                fn()
            n += 30
    e1.record()
    global last_clock
    last_clock = ""
    if _nv is not None:  # sampled while the queue is still draining: clocks / power UNDER load
        mhz = pynvml.nvmlDeviceGetClockInfo(_nv, pynvml.NVML_CLOCK_SM)
        watts = pynvml.nvmlDeviceGetPowerUsage(_nv) / 1000.0
        last_clock = f"  [{mhz} MHz, {watts:.0f} W]"
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return ms, flops / ms / 1e9


def main(which, sweep=None):
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).to(bf)  # noqa: E731
    x = r(B, N, D)
    xm = r(B, N, D)
    cat = r(B, N, D + M)
    q, k, v = r(B, H, N, 128), r(B, H, N, 128), r(B, H, N, 128)
    w1, b1 = r(3 * D + M, D, sc=D ** -0.5), r(3 * D + M, sc=0.1)
    w2, b2 = r(D, D + M, sc=(D + M) ** -0.5), r(D, sc=0.1)
    wfc1, bfc1 = r(M, D, sc=D ** -0.5), r(M, sc=0.1)
    wfc2, wproj = r(D, M, sc=M ** -0.5), r(D, D, sc=D ** -0.5)
    qs, ks = r(128), r(128)
    pe = r(N, 64, 2)
    gate = r(B, D, sc=0.1)
    shift, scale = r(B, D, sc=0.1), r(B, D, sc=0.1)
    f32buf = torch.empty(B, L, M, device=dev, dtype=torch.float32) if 'fc1_f32out' in which else None
    q8, k8, v8 = (t.to(ops.fp8) for t in (q, k, v)) if "attn_f8" in which else (None, None, None)
    f8 = any(n.endswith("_f8") and n != "attn_f8" for n in which)
    if f8:  # --quantize operands
        xm8, xs = ops.quantize_rows(xm)
        cat8, cs = ops.quantize_rows(cat)
        w1q, w1s = ops.quantize_rows(w1)
        w2q, w2s = ops.quantize_rows(w2)
        wfq, wfs = ops.quantize_rows(wfc1)
        catm8, cms = ops.quantize_rows(cat[:, S:, D:].contiguous())
        wf2q, wf2s = ops.quantize_rows(wfc2)
    f4 = any(n.endswith("_f4") for n in which)
    if f4:  # --quantize 4 operands (NVFP4)
        cat4, csf4, cs4 = ops.quantize_rows_fp4(cat)
        w24, w2sf, w2s4 = ops.fp4_weight(w2)
        catm4, cmsf4, cms4 = ops.quantize_rows_fp4(cat[:, S:, D:])
        wf24, wf2sf, wf2s4 = ops.fp4_weight(wfc2)
        xm4, xmsf4, xms4 = ops.quantize_rows_fp4(xm)
        xi4, xisf4, xis4 = ops.quantize_rows_fp4(xm[:, S:])
        wq4, wqsf, wqs4 = ops.fp4_weight(w1[:3 * D], ops.FP4_TILE_N_QKV)
        wm4, wmsf, wms4 = ops.fp4_weight(w1[3 * D:])
        wp4, wpsf, wps4 = ops.fp4_weight(wproj)
        ca4, casf4, cas4 = ops.quantize_rows_fp4(cat[:, S:, :D])
        q8o, k8o, v8o = (torch.empty(B, H, N, 128, device=dev, dtype=ops.fp8) for _ in range(3))
        f4bufs = (torch.empty(B * N * D // 2, device=dev, dtype=torch.uint8), torch.empty(B * N * D // 16, device=dev, dtype=torch.uint8),
                  torch.empty(B * N, device=dev))
    tests = {
        "linear1": (lambda: ops.gemm_qkv(xm, w1, b1, qs, ks, pe, q, k, v, 0, mlp_out=cat[:, :, D:]), 2.0 * B * N * (3 * D + M) * D),
        "linear2": (lambda: ops.gemm(cat, w2, b2, gate=gate, resid=x, out=x), 2.0 * B * N * D * (D + M)),
        "fc2": (lambda: ops.gemm(cat[:, S:, D:], wfc2, b2, gate=gate, resid=x[:, S:], out=x[:, S:]), 2.0 * B * L * D * M),
        "proj": (lambda: ops.gemm(cat[:, S:, :D], wproj, b2, gate=gate, resid=x[:, S:], out=x[:, S:]), 2.0 * B * L * D * D),
        "fc1": (lambda: ops.gemm(xm[:, S:], wfc1, bfc1, act="gelu_tanh", out=cat[:, S:, D:]), 2.0 * B * L * M * D),
        "fc1_bias": (lambda: ops.gemm(xm[:, S:], wfc1, bfc1, out=cat[:, S:, D:]), 2.0 * B * L * M * D),
        "fc1_gelu": (lambda: ops.gemm(xm[:, S:], wfc1, act="gelu_tanh", out=cat[:, S:, D:]), 2.0 * B * L * M * D),
        "fc1_f32out": (lambda: ops.gemm(xm[:, S:], wfc1, out=f32buf), 2.0 * B * L * M * D),
        "fc1_plain": (lambda: ops.gemm(xm[:, S:], wfc1, out=cat[:, S:, D:]), 2.0 * B * L * M * D),
        "attn": (lambda: ops.attention(q, k, v, cat[:, :, :D], 128 ** -0.5), 4.0 * B * H * N * N * 128),
        "attn_p": (lambda: ops.attention(q, k, v, cat[:, :, :D], 128 ** -0.5, variant=7), 4.0 * B * H * N * N * 128),
        "attn_f8": (lambda: ops.attention(q8, k8, v8, cat[:, :, :D], 128 ** -0.5), 4.0 * B * H * N * N * 128),
        "attn_seq": (lambda: ops.attention(q, k, v, cat[:, :, :D], 128 ** -0.5, variant=4), 4.0 * B * H * N * N * 128),
        "attn3": (lambda: ops.attention(q, k, v, cat[:, :, :D], 128 ** -0.5, variant=5), 4.0 * B * H * N * N * 128),
        "attn3_noseq": (lambda: ops.attention(q, k, v, cat[:, :, :D], 128 ** -0.5, variant=6), 4.0 * B * H * N * N * 128),
        "rownorm": (lambda: ops.rownorm(x, 0, shift, scale, 1e-6, out=xm), 0.0),
        "linear1_f8": (lambda: ops.gemm_qkv(xm8, w1q, b1, qs, ks, pe, q, k, v, 0, mlp_out=cat[:, :, D:], a_scale=xs, w_scale=w1s),
                       2.0 * B * N * (3 * D + M) * D),
        "linear2_f8": (lambda: ops.gemm(cat8, w2q, b2, gate=gate, resid=x, out=x, a_scale=cs, w_scale=w2s), 2.0 * B * N * D * (D + M)),
        "fc1_f8": (lambda: ops.gemm(xm8[:, S:], wfq, bfc1, act="gelu_tanh", out=cat[:, S:, D:], a_scale=xs[:, S:], w_scale=wfs),
                   2.0 * B * L * M * D),
        "linear2_f4": (lambda: ops.gemm_fp4(cat4, csf4, cs4, w24, w2sf, w2s4, B, bias=b2, gate=gate, resid=x, out=x), 2.0 * B * N * D * (D + M)),
        "fc2_f4": (lambda: ops.gemm_fp4(catm4, cmsf4, cms4, wf24, wf2sf, wf2s4, B, bias=b2, gate=gate, resid=x[:, S:], out=x[:, S:]), 2.0 * B * L * D * M),
        "quant_cat_f4": (lambda: ops.quantize_rows_fp4(cat), 0.0),
        "quant_x_f4": (lambda: ops.quantize_rows_fp4(xm), 0.0),
        "quant_mlp_f4": (lambda: ops.quantize_rows_fp4(cat[:, S:, D:]), 0.0),
        "rownorm_f4": (lambda: ops.rownorm(x, 0, shift, scale, 1e-6, out_fp4=f4bufs), 0.0),
        "qkv1_f4": (lambda: ops.gemm_fp4_qkv(xm4, xmsf4, xms4, wq4, wqsf, wqs4, B, b1[:3 * D], qs, ks, pe, q8o, k8o, v8o, 0), 2.0 * B * N * 3 * D * D),
        "qkv1_bf16out_f4": (lambda: ops.gemm_fp4_qkv(xm4, xmsf4, xms4, wq4, wqsf, wqs4, B, b1[:3 * D], qs, ks, pe, q, k, v, 0), 2.0 * B * N * 3 * D * D),
        "mlp1_f4": (lambda: ops.gemm_fp4(xm4, xmsf4, xms4, wm4, wmsf, wms4, B, bias=b1[3 * D:], act="gelu_tanh", out=cat[:, :, D:]), 2.0 * B * N * M * D),
        "mlp1_noact_f4": (lambda: ops.gemm_fp4(xm4, xmsf4, xms4, wm4, wmsf, wms4, B, bias=b1[3 * D:], out=cat[:, :, D:]), 2.0 * B * N * M * D),
        "qkv_img_f4": (lambda: ops.gemm_fp4_qkv(xi4, xisf4, xis4, wq4, wqsf, wqs4, B, b1[:3 * D], qs, ks, pe, q8o, k8o, v8o, S), 2.0 * B * L * 3 * D * D),
        "fc1_f4": (lambda: ops.gemm_fp4(xi4, xisf4, xis4, wm4, wmsf, wms4, B, bias=b1[3 * D:], act="gelu_tanh", out=cat[:, S:, D:]), 2.0 * B * L * M * D),
        "proj_f4": (lambda: ops.gemm_fp4(ca4, casf4, cas4, wp4, wpsf, wps4, B, bias=b2, gate=gate, resid=x[:, S:], out=x[:, S:]), 2.0 * B * L * D * D),
        "fc2_f8": (lambda: ops.gemm(catm8, wf2q, b2, gate=gate, resid=x[:, S:], out=x[:, S:], a_scale=cms, w_scale=wf2s), 2.0 * B * L * D * M),
        "quant_cat_f8": (lambda: ops.quantize_rows(cat, out=cat8, out_scale=cs), 0.0),
        "rownorm_f8": (lambda: ops.rownorm(x, 0, shift, scale, 1e-6, out=xm8, out_scale=xs), 0.0),
        "cublas_l1": (lambda: torch.matmul(xm.view(-1, D), w1.T), 2.0 * B * N * (3 * D + M) * D),
    }
    if sweep:  # raster sweep of the long-K members: FX_GEMM_GROUP_N (band width) x FX_GEMM_GROUP_M_BAND, one process
        for gn, gm in sweep:
            os.environ["FX_GEMM_GROUP_N"], os.environ["FX_GEMM_GROUP_M_BAND"] = str(gn), str(gm)
            line = f"raster group_n={gn} group_m={gm}:"
            for name in which:
                ms, tf = sustained(*tests[name])
                line += f"  {name} {ms:.3f} ms {tf:.0f} TF{last_clock}"
            print(line, flush=True)
        return
    for name in which or list(tests):
        fn, fl = tests[name]
        ms, tf = sustained(fn, fl)
        nbytes = {"rownorm": 4 * x.numel(), "rownorm_f8": 3 * x.numel(), "quant_cat_f8": 3 * cat.numel(),
                  "quant_cat_f4": 2.5625 * cat.numel(), "quant_x_f4": 2.5625 * xm.numel(),
                  "quant_mlp_f4": 2.5625 * B * L * M, "rownorm_f4": 2.5625 * x.numel()}.get(name)
        extra = f" = {nbytes / ms / 1e6:.0f} GB/s" if nbytes else f" = {tf:.0f} TFLOP/s"
        print(f"{name:10s} {ms:8.3f} ms{extra}{last_clock}", flush=True)


if __name__ == "__main__":
    argv = sys.argv[1:]
    if argv and argv[0] == "--raster-sweep":  # --raster-sweep gn:gm,gn:gm,... op op ...
        pairs = [tuple(int(v) for v in t.split(":")) for t in argv[1].split(",")]
        main(argv[2:], pairs)
    else:
        main(argv)
