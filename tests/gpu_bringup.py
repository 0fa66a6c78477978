"""GPU bring-up probes (run on the B200 box: `bash tests/run_bringup.sh`).  Not collected by pytest.

Each stage runs in its own process under `timeout` so that a hung kernel cannot take the box down.
Prints compact diagnostics; the real parity tests live in tests/test_gpu_*.py.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flux-generator_b200"))
from flux import ops  # noqa: E402

dev = "cuda"
bf = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def tm(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def stage_umma():
    g = torch.Generator(device="cpu").manual_seed(0)
    for K in (64, 128):
        for N in (64, 128):
            a = torch.randn(128, K, generator=g).to(bf).to(dev)
            bk = torch.randn(N, K, generator=g).to(bf).to(dev)       # K-major B
            ref = a.float() @ bk.float().T
            d = ops.dbg_umma_tile(a, bk, N, False, False, 16, 1024, 0)
            print(f"umma K={K} N={N} SS k-major: rel {rel(d, ref):.2e}")
            d = ops.dbg_umma_tile(a, bk, N, False, True, 16, 1024, 0)
            print(f"umma K={K} N={N} TS(A in TMEM) k-major B: rel {rel(d, ref):.2e}")
            bmn = bk.T.contiguous()                                    # [K][N] MN-major
            for (lbo, sbo, ks) in ((K * 128, 1024, 2048),):
                d = ops.dbg_umma_tile(a, bmn, N, True, False, lbo, sbo, ks)
                print(f"umma K={K} N={N} MN-major lbo={lbo} sbo={sbo} kstep={ks}: rel {rel(d, ref):.2e}")


def stage_gemm():
    g = torch.Generator(device="cpu").manual_seed(1)
    for (M, N, K) in ((128, 256, 64), (128, 256, 256), (256, 512, 512), (300, 520, 200), (1000, 72, 328), (4096, 3072, 3072),
                      (512, 64, 3072), (777, 3, 1152)):
        a = (torch.randn(M, K, generator=g)).to(bf).to(dev)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(bf).to(dev)
        ref = a.float() @ w.float().T
        out = ops.gemm(a, w, out_dtype=torch.float32)
        print(f"gemm {M}x{N}x{K} f32 out: rel {rel(out, ref):.2e}")
        out = ops.gemm(a, w)
        print(f"gemm {M}x{N}x{K} bf16 out: rel {rel(out, ref):.2e}")
    # epilogue: bias + gelu, gate + resid, batched strided views
    B, R, K, N = 3, 200, 256, 384
    a_full = torch.randn(B, R + 56, K + 64, generator=g).to(bf).to(dev)
    a = a_full[:, 56:, 64:]
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(bf).to(dev)
    bias = torch.randn(N, generator=g).to(bf).to(dev)
    gate = torch.randn(B, N, generator=g).to(bf).to(dev)
    resid = torch.randn(B, R, N, generator=g).to(bf).to(dev)
    lin = a.float() @ w.float().T + bias.float()
    out = ops.gemm(a, w, bias, act="gelu_tanh")
    print(f"gemm bias+gelu (strided batched A): rel {rel(out, torch.nn.functional.gelu(lin, approximate='tanh')):.2e}")
    ref = resid.float() + gate.float()[:, None, :] * lin
    out = ops.gemm(a, w, bias, gate=gate, resid=resid)
    print(f"gemm gate+resid: rel {rel(out, ref):.2e}")
    x = resid.clone()
    ops.gemm(a, w, bias, gate=gate, resid=x, out=x)
    print(f"gemm gate+resid in place: rel {rel(x, ref):.2e}")
    for act, fn in (("quick_gelu", lambda t: t * torch.sigmoid(1.702 * t)), ("gelu", torch.nn.functional.gelu)):
        out = ops.gemm(a, w, bias, act=act)
        print(f"gemm act {act}: rel {rel(out, fn(lin)):.2e}")
    # timing on the cfg-4 shapes
    for (M, N, K) in ((34816, 21504, 3072), (34816, 3072, 15360), (32768, 12288, 3072), (32768, 3072, 12288), (4352, 21504, 3072)):
        a = torch.randn(M, K, device=dev, dtype=bf)
        w = torch.randn(N, K, device=dev, dtype=bf) * K ** -0.5
        out = torch.empty(M, N, device=dev, dtype=bf)
        ms = tm(lambda: ops.gemm(a, w, out=out), iters=3, warm=1)
        ms_t = tm(lambda: torch.matmul(a, w.T, out=out), iters=3, warm=1)
        fl = 2.0 * M * N * K
        r = rel(ops.gemm(a[:2048], w), a[:2048].float() @ w.float().T)
        print(f"gemm {M}x{N}x{K}: {ms:.3f} ms = {fl / ms / 1e9:.0f} TFLOP/s (cuBLAS {ms_t:.3f} ms = {fl / ms_t / 1e9:.0f}); rel {r:.2e}")
        del a, w, out


def qkv_ref(a, w, bias, qs, ks, pe, H, eps=1e-5):
    y = a.float() @ w.float().T + bias.float()
    B, R, _ = y.shape
    D = H * 128
    q, k, v, mlp = y[..., :D], y[..., D:2 * D], y[..., 2 * D:3 * D], y[..., 3 * D:]

    def heads(t):
        return t.reshape(B, R, H, 128).transpose(1, 2)

    def rms(t, s):
        return t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + eps) * s.float()

    def rope(t):
        cs = pe.float()  # [R, 64, 2]
        t2 = t.reshape(B, H, R, 64, 2)
        c, s = cs[..., 0], cs[..., 1]
        o0 = t2[..., 0] * c - t2[..., 1] * s
        o1 = t2[..., 0] * s + t2[..., 1] * c
        return torch.stack([o0, o1], -1).reshape(B, H, R, 128)

    return rope(rms(heads(q), qs)), rope(rms(heads(k), ks)), heads(v), torch.nn.functional.gelu(mlp, approximate="tanh")


def stage_qkv():
    g = torch.Generator(device="cpu").manual_seed(2)
    for (B, R, H, K, mlp) in ((2, 200, 2, 256, 1024), (1, 300, 4, 512, 0), (2, 128, 24, 3072, 0)):
        D = H * 128
        N = 3 * D + mlp
        S_off, S_tot = 40, R + 40
        a = torch.randn(B, R, K, generator=g).to(bf).to(dev)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(bf).to(dev)
        bias = (0.1 * torch.randn(N, generator=g)).to(bf).to(dev)
        qs = (1 + 0.1 * torch.randn(128, generator=g)).to(bf).to(dev)
        ks = (1 + 0.1 * torch.randn(128, generator=g)).to(bf).to(dev)
        ang = torch.rand(S_tot, 64, generator=g) * 6.28
        pe = torch.stack([torch.cos(ang), torch.sin(ang)], -1).to(bf).to(dev)
        q = torch.zeros(B, H, S_tot, 128, device=dev, dtype=bf)
        k = torch.zeros_like(q)
        v = torch.zeros_like(q)
        mo = torch.zeros(B, S_tot, D + mlp, device=dev, dtype=bf) if mlp else None
        ops.gemm_qkv(a, w, bias, qs, ks, pe, q, k, v, S_off, mlp_out=(mo[:, :, D:] if mlp else None))
        rq, rk, rv, rm = qkv_ref(a, w, bias, qs, ks, pe[S_off:], H)
        print(f"qkv B{B} R{R} H{H} K{K} mlp{mlp}: q {rel(q[:, :, S_off:], rq):.2e} k {rel(k[:, :, S_off:], rk):.2e} "
              f"v {rel(v[:, :, S_off:], rv):.2e}" + (f" mlp {rel(mo[:, S_off:, D:], rm):.2e}" if mlp else "") +
              f" untouched {q[:, :, :S_off].abs().max().item():.1f}")


def stage_conv():
    g = torch.Generator(device="cpu").manual_seed(3)
    for (B, H, W, Cin, Cout) in ((1, 8, 16, 64, 64), (2, 12, 20, 128, 256), (1, 64, 64, 512, 512), (1, 33, 47, 64, 3)):
        x = torch.randn(B, H, W, Cin, generator=g).to(bf).to(dev)
        w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(bf).to(dev)
        bias = (0.1 * torch.randn(Cout, generator=g)).to(bf).to(dev)
        ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias.float(), padding=1).permute(0, 2, 3, 1)
        w_ohwi = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
        out = ops.conv3x3(x, w_ohwi, bias, out_dtype=torch.float32)
        print(f"conv B{B} {H}x{W} {Cin}->{Cout}: rel {rel(out, ref):.2e}")
        if Cout % 8 == 0:
            res = torch.randn(B, H, W, Cout, generator=g).to(bf).to(dev)
            out = ops.conv3x3(x, w_ohwi, bias, resid=res)
            print(f"conv+resid: rel {rel(out, ref + res.float()):.2e}")
    x = torch.randn(1, 512, 512, 256, device=dev, dtype=bf)
    w = torch.randn(256, 9 * 256, device=dev, dtype=bf) * 0.02
    b = torch.zeros(256, device=dev, dtype=bf)
    ms = tm(lambda: ops.conv3x3(x, w, b), iters=3, warm=1)
    print(f"conv 512x512 256->256: {ms:.3f} ms = {2.0 * 512 * 512 * 256 * 256 * 9 / ms / 1e9:.0f} TFLOP/s")
    x = torch.randn(1, 1024, 1024, 128, device=dev, dtype=bf)
    w = torch.randn(128, 9 * 128, device=dev, dtype=bf) * 0.02
    b = torch.zeros(128, device=dev, dtype=bf)
    ms = tm(lambda: ops.conv3x3(x, w, b), iters=3, warm=1)
    print(f"conv 1024x1024 128->128: {ms:.3f} ms = {2.0 * 1024 * 1024 * 128 * 128 * 9 / ms / 1e9:.0f} TFLOP/s")


def stage_attn(variant):
    g = torch.Generator(device="cpu").manual_seed(4)
    for (B, H, S) in ((1, 1, 128), (1, 2, 256), (2, 3, 512), (1, 2, 320), (1, 2, 1280), (1, 1, 77), (1, 24, 4352)):
        q = torch.randn(B, H, S, 128, generator=g).to(bf).to(dev)
        k = torch.randn(B, H, S, 128, generator=g).to(bf).to(dev)
        v = torch.randn(B, H, S, 128, generator=g).to(bf).to(dev)
        out = torch.zeros(B, S, H * 128 + 64, device=dev, dtype=bf)
        ops.attention(q, k, v, out[:, :, :H * 128], 128 ** -0.5, variant=variant)
        ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).transpose(1, 2).reshape(B, S, H * 128)
        print(f"attn v{variant} B{B} H{H} S{S}: rel {rel(out[:, :, :H * 128], ref):.2e} pad {out[:, :, H * 128:].abs().max().item():.1f}")
    # large logits (sharp softmax) exercise the lazy rescale
    q = (4 * torch.randn(1, 2, 512, 128, generator=g)).to(bf).to(dev)
    k = (4 * torch.randn(1, 2, 512, 128, generator=g)).to(bf).to(dev)
    v = torch.randn(1, 2, 512, 128, generator=g).to(bf).to(dev)
    out = torch.zeros(1, 512, 256, device=dev, dtype=bf)
    ops.attention(q, k, v, out, 128 ** -0.5, variant=variant)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).transpose(1, 2).reshape(1, 512, 256)
    print(f"attn v{variant} sharp: rel {rel(out, ref):.2e}")
    for (B, H, S) in ((8, 24, 4352), (1, 24, 4608), (1, 24, 9728)):
        q = torch.randn(B, H, S, 128, device=dev, dtype=bf)
        k = torch.randn(B, H, S, 128, device=dev, dtype=bf)
        v = torch.randn(B, H, S, 128, device=dev, dtype=bf)
        out = torch.empty(B, S, H * 128, device=dev, dtype=bf)
        ms = tm(lambda: ops.attention(q, k, v, out, 128 ** -0.5, variant=variant), iters=3, warm=1)
        fl = 4.0 * B * H * S * S * 128
        ms_t = tm(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), iters=3, warm=1)
        print(f"attn v{variant} B{B} H{H} S{S}: {ms:.3f} ms = {fl / ms / 1e9:.0f} TFLOP/s (torch sdpa {ms_t:.3f} ms = {fl / ms_t / 1e9:.0f})")


def stage_elem():
    g = torch.Generator(device="cpu").manual_seed(5)
    F = torch.nn.functional
    B, R, D = 2, 100, 3072
    x = torch.randn(B, R, D, generator=g).to(bf).to(dev)
    sh = (0.1 * torch.randn(B, D, generator=g)).to(bf).to(dev)
    sc = (0.1 * torch.randn(B, D, generator=g)).to(bf).to(dev)
    ref = (1 + sc.float()[:, None]) * F.layer_norm(x.float(), (D,), eps=1e-6) + sh.float()[:, None]
    print(f"rownorm mode0: rel {rel(ops.rownorm(x, 0, sh, sc, 1e-6), ref):.2e}")
    wgt = (1 + 0.1 * torch.randn(D, generator=g)).to(bf).to(dev)
    bia = (0.1 * torch.randn(D, generator=g)).to(bf).to(dev)
    print(f"rownorm mode1: rel {rel(ops.rownorm(x, 1, wgt, bia, 1e-5), F.layer_norm(x.float(), (D,), wgt.float(), bia.float(), 1e-5)):.2e}")
    ref = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-6) * wgt.float()
    print(f"rownorm mode2: rel {rel(ops.rownorm(x, 2, wgt, None, 1e-6), ref):.2e}")
    # gemv
    Bv, K, N = 5, 3072, 1000
    xin = torch.randn(Bv, K, generator=g).to(bf).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(bf).to(dev)
    b = (0.1 * torch.randn(N, generator=g)).to(bf).to(dev)
    add = torch.randn(Bv, N, generator=g).to(bf).to(dev)
    ref = F.silu(xin.float()).to(bf).float() @ w.float().T + b.float()
    print(f"gemv silu_in: rel {rel(ops.gemv(xin, w, b, silu_in=True), ref):.2e}")
    print(f"gemv add: rel {rel(ops.gemv(xin, w, b, add=add), xin.float() @ w.float().T + b.float() + add.float()):.2e}")
    big_w = torch.randn(1056768 // 4, 3072, device=dev, dtype=bf) * 0.01
    xin8 = torch.randn(8, 3072, device=dev, dtype=bf)
    o = torch.empty(8, big_w.shape[0], device=dev, dtype=bf)
    ms = tm(lambda: ops.gemv(xin8, big_w, None, silu_in=True, out=o), iters=3, warm=1)
    print(f"gemv 8x{big_w.shape[0]}x3072: {ms:.3f} ms = {big_w.numel() * 2 / ms / 1e6:.0f} GB/s")
    # timestep embedding
    t = torch.tensor([1.0, 0.75, 0.5, 0.0078125, 0.99357950687], dtype=bf).to(dev)
    e = ops.timestep_embedding(t, 256)
    tt = (1000.0 * t.cpu()).float()
    fr = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(128, dtype=torch.float32) / 128)
    xx = tt[:, None] * fr[None]
    ref = torch.cat([torch.cos(xx), torch.sin(xx)], -1).to(bf)
    print(f"timestep_embedding: max abs diff {(e.cpu().float() - ref.float()).abs().max().item():.3e}, mismatching {(e.cpu() != ref).sum().item()}/{ref.numel()}")
    # euler
    xt = torch.randn(2, 64, 64, generator=g).to(bf).to(dev)
    pr = torch.randn(2, 64, 64, generator=g).to(bf).to(dev)
    ref = (xt.float() + (-0.25 * pr.float()).to(bf).float()).to(bf)
    got = ops.euler_step(xt.clone(), pr, -0.25)
    print(f"euler: exact {torch.equal(got, ref)}")
    # patchify / unpatchify
    lat = torch.randn(2, 8, 12, 16, generator=g).to(bf).to(dev)
    p = ops.patchify(lat)
    refp = lat.reshape(2, 4, 2, 6, 2, 16).permute(0, 1, 3, 5, 2, 4).reshape(2, 24, 64)
    print(f"patchify exact {torch.equal(p, refp)}")
    z = ops.unpatchify_scale(p, (8, 12), 64, 0.3611, 0.1159)
    refz = ((lat.float() / 0.3611).to(bf).float() + 0.1159).to(bf)
    print(f"unpatchify_scale: max diff {(z[..., :16].float() - refz.float()).abs().max().item():.3e} pad {z[..., 16:].abs().max().item()}")
    # groupnorm
    for C in (64, 128, 256, 512):
        xg = (torch.randn(2, 24, 40, C, generator=g) * 2 + 0.5).to(bf).to(dev)
        gw = (1 + 0.1 * torch.randn(C, generator=g)).to(bf).to(dev)
        gb = (0.1 * torch.randn(C, generator=g)).to(bf).to(dev)
        ref = F.group_norm(xg.float().permute(0, 3, 1, 2), 32, gw.float(), gb.float(), 1e-6).permute(0, 2, 3, 1)
        print(f"groupnorm C{C}: rel {rel(ops.groupnorm(xg, gw, gb, 1e-6, False), ref):.2e} +silu {rel(ops.groupnorm(xg, gw, gb, 1e-6, True), F.silu(ref)):.2e}")
    xu = torch.randn(2, 5, 7, 64, generator=g).to(bf).to(dev)
    print(f"upsample exact {torch.equal(ops.upsample2x(xu), xu.repeat_interleave(2, 1).repeat_interleave(2, 2))}")
    s = torch.randn(300, 1024, generator=g).to(dev) * 5
    print(f"softmax_rows: rel {rel(ops.softmax_rows(s, 0.3), torch.softmax(s * 0.3, -1)):.2e}")
    xt2 = torch.randn(100, 70, generator=g).to(bf).to(dev)
    print(f"transpose exact {torch.equal(ops.transpose(xt2), xt2.T.contiguous())}")
    xi = torch.randn(1000, generator=g).to(dev) * 1.5
    img, u8 = ops.finish_image(xi)
    refi = torch.clip(xi + 1, 0, 2) * 0.5
    print(f"finish_image: {torch.equal(img, refi)} u8 {torch.equal(u8, (refi * 255).to(torch.uint8))}")
    ids = torch.randint(0, 100, (2, 9), generator=g, dtype=torch.int32).to(dev)
    tab = torch.randn(100, 256, generator=g).to(bf).to(dev)
    pos = torch.randn(77, 256, generator=g).to(bf).to(dev)
    print(f"embedding exact {torch.equal(ops.embedding(ids, tab), tab[ids.long()])} +pos rel {rel(ops.embedding(ids, tab, pos), tab[ids.long()].float() + pos[:9].float()):.2e}")
    a1 = torch.randn(1000, generator=g).to(bf).to(dev)
    b1 = torch.randn(1000, generator=g).to(bf).to(dev)
    print(f"act_mul: rel {rel(ops.act_mul(a1, b1, 'gelu'), F.gelu(a1.float()).to(bf).float() * b1.float()):.2e}")
    # small attention
    Bq, S, H = 2, 77, 4
    qkv = torch.randn(Bq, S, 3 * H * 64, generator=g).to(bf).to(dev)
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    bias = torch.randn(H, S, S, generator=g).to(dev)

    def hd(t):
        return t.float().reshape(Bq, S, H, 64).transpose(1, 2)

    ref = F.scaled_dot_product_attention(hd(q), hd(k), hd(v), attn_mask=bias[None], scale=1.0).transpose(1, 2).reshape(Bq, S, H * 64)
    print(f"attn_small bias: rel {rel(ops.attention_small(q, k, v, H, 1.0, bias=bias), ref):.2e}")
    ref = F.scaled_dot_product_attention(hd(q), hd(k), hd(v), is_causal=True).transpose(1, 2).reshape(Bq, S, H * 64)
    print(f"attn_small causal: rel {rel(ops.attention_small(q, k, v, H, 0.125, causal=True), ref):.2e}")


if __name__ == "__main__":
    st = sys.argv[1]
    t0 = time.time()
    print(f"== stage {st} on {torch.cuda.get_device_name(0)}", flush=True)
    if st == "umma":
        stage_umma()
    elif st == "gemm":
        stage_gemm()
    elif st == "qkv":
        stage_qkv()
    elif st == "conv":
        stage_conv()
    elif st.startswith("attn"):
        stage_attn(int(st[4:]))
    elif st == "elem":
        stage_elem()
    torch.cuda.synchronize()
    print(f"== stage {st} done in {time.time() - t0:.1f}s", flush=True)
