"""GPU kernel-level parity through the C ABI: tcgen05 GEMM family, attention, conv, and the HBM-bound
kernels, against plain fp32 torch references of the same op (tolerance: bf16 output rounding,
rel-L2 <= 5e-3; fp32 outputs <= 2e-5) plus edge cases (ragged tiles, tails, invalid arguments) and
size-independent properties at the BASELINE sizes."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from flux import ops  # noqa: E402
from helpers import rel_l2  # noqa: E402

dev = "cuda"
bf = torch.bfloat16
F = torch.nn.functional


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(bf).to(dev)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 520, 200), (1000, 72, 328), (777, 3, 1152), (1, 8, 8),
                                   (4096, 3072, 3072), (129, 257, 72)])
def test_gemm_shapes(M, N, K):
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    ref = a.float() @ w.float().T
    assert rel_l2(ops.gemm(a, w, out_dtype=torch.float32), ref) <= 2e-5
    assert rel_l2(ops.gemm(a, w), ref) <= 5e-3


def test_gemm_epilogues_and_views():
    B, R, K, N = 3, 200, 256, 384
    a = rnd(B, R + 56, K + 64, seed=3)[:, 56:, 64:]
    w, bias = rnd(N, K, seed=4, scale=K ** -0.5), rnd(N, seed=5)
    gate, resid = rnd(B, N, seed=6), rnd(B, R, N, seed=7)
    lin = a.float() @ w.float().T + bias.float()
    assert rel_l2(ops.gemm(a, w, bias, act="gelu_tanh"), F.gelu(lin, approximate="tanh")) <= 5e-3
    assert rel_l2(ops.gemm(a, w, bias, act="quick_gelu"), lin * torch.sigmoid(1.702 * lin)) <= 5e-3
    assert rel_l2(ops.gemm(a, w, bias, act="gelu"), F.gelu(lin)) <= 5e-3
    ref = resid.float() + gate.float()[:, None] * lin
    assert rel_l2(ops.gemm(a, w, bias, gate=gate, resid=resid), ref) <= 5e-3
    x = resid.clone()
    ops.gemm(a, w, bias, gate=gate, resid=x, out=x)  # in place on the residual stream
    assert rel_l2(x, ref) <= 5e-3


def test_gemm_rejects_bad_arguments():
    a, w = rnd(16, 12), rnd(8, 12)
    with pytest.raises(ValueError):  # K not a multiple of 8
        ops.gemm(a, w)
    with pytest.raises(ValueError):  # K mismatch
        ops.gemm(rnd(16, 16), rnd(8, 24))


def test_gemm_matches_cuda_core_reference_at_full_size():
    """linear1 of a single block at the BASELINE shape (rows subsampled for the check)."""
    a, w = rnd(4352, 3072, seed=8), rnd(21504, 3072, seed=9, scale=3072 ** -0.5)
    out = ops.gemm(a, w, out_dtype=torch.float32)
    ref = ops.dbg_gemm_ref(a[:256], w)
    assert rel_l2(out[:256], ref) <= 2e-5
    # linearity: gemm(a1 + a2) == gemm(a1) + gemm(a2) up to fp32 accumulation order
    a2 = rnd(4352, 3072, seed=10)
    s = ops.gemm((a.float() + a2.float()).to(bf), w, out_dtype=torch.float32)
    assert rel_l2(s, out + ops.gemm(a2, w, out_dtype=torch.float32)) <= 5e-3


def _qkv_ref(a, w, bias, qs, ks, pe, H, eps=1e-5):
    y = a.float() @ w.float().T + bias.float()
    B, R, _ = y.shape
    D = H * 128
    hd = lambda t_: t_.reshape(B, R, H, 128).transpose(1, 2)  # noqa: E731
    rms = lambda t_, s: t_ * torch.rsqrt(t_.pow(2).mean(-1, keepdim=True) + eps) * s.float()  # noqa: E731

    def rope(t_):
        t2 = t_.reshape(B, H, R, 64, 2)
        c, s = pe.float()[..., 0], pe.float()[..., 1]
        return torch.stack([t2[..., 0] * c - t2[..., 1] * s, t2[..., 0] * s + t2[..., 1] * c], -1).reshape(B, H, R, 128)

    return (rope(rms(hd(y[..., :D]), qs)), rope(rms(hd(y[..., D:2 * D]), ks)), hd(y[..., 2 * D:3 * D]),
            F.gelu(y[..., 3 * D:], approximate="tanh"))


@pytest.mark.parametrize("B,R,H,K,mlp", [(2, 200, 2, 256, 1024), (1, 300, 4, 512, 0), (1, 130, 24, 3072, 12288)])
def test_gemm_qkv_epilogue(B, R, H, K, mlp):
    D, off = H * 128, 40
    a, w = rnd(B, R, K, seed=11), rnd(3 * D + mlp, K, seed=12, scale=K ** -0.5)
    bias, qs, ks = rnd(3 * D + mlp, seed=13, scale=0.1), (1 + rnd(128, seed=14, scale=0.1).float()).to(bf), \
        (1 + rnd(128, seed=15, scale=0.1).float()).to(bf)
    ang = torch.rand(R + off, 64, generator=torch.Generator().manual_seed(16)) * 6.28
    pe = torch.stack([torch.cos(ang), torch.sin(ang)], -1).to(bf).to(dev)
    q = torch.zeros(B, H, R + off, 128, device=dev, dtype=bf)
    k, v = torch.zeros_like(q), torch.zeros_like(q)
    mo = torch.zeros(B, R + off, D + mlp, device=dev, dtype=bf) if mlp else None
    ops.gemm_qkv(a, w, bias, qs, ks, pe, q, k, v, off, mlp_out=(mo[:, :, D:] if mlp else None))
    rq, rk, rv, rm = _qkv_ref(a, w, bias, qs, ks, pe[off:], H)
    assert rel_l2(q[:, :, off:], rq) <= 5e-3 and rel_l2(k[:, :, off:], rk) <= 5e-3 and rel_l2(v[:, :, off:], rv) <= 5e-3
    assert q[:, :, :off].abs().max().item() == 0  # rows before seq_off untouched
    if mlp:
        assert rel_l2(mo[:, off:, D:], rm) <= 5e-3 and mo[:, :, :D].abs().max().item() == 0


def test_gemm_qkv_blocked_rope_table_is_bit_identical():
    """The coalesced (blocked) RoPE table layout gives the same bits as the plain [seq, 64, 2] table."""
    B, R, H, K, off = 2, 200, 2, 256, 64
    D = H * 128
    a, w = rnd(B, R, K, seed=41), rnd(3 * D, K, seed=42, scale=K ** -0.5)
    bias, qs, ks = rnd(3 * D, seed=43, scale=0.1), rnd(128, seed=44), rnd(128, seed=45)
    ang = torch.rand(R + off, 64, generator=torch.Generator().manual_seed(46)) * 6.28
    pe = torch.stack([torch.cos(ang), torch.sin(ang)], -1).to(bf).to(dev)
    outs = []
    for blocked in (False, True):
        q = torch.zeros(B, H, R + off, 128, device=dev, dtype=bf)
        k, v = torch.zeros_like(q), torch.zeros_like(q)
        ops.gemm_qkv(a, w, bias, qs, ks, ops.block_pe(pe) if blocked else pe, q, k, v, off, pe_blocked=blocked)
        outs.append((q, k, v))
    for x, y in zip(*outs):
        assert torch.equal(x, y)
    with pytest.raises(ValueError):  # the blocked layout needs seq_off % 32 == 0
        ops.gemm_qkv(a, w, bias, qs, ks, ops.block_pe(pe), q, k, v, 40, pe_blocked=True)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 16, 64, 64), (2, 12, 20, 128, 256), (1, 33, 47, 64, 3), (1, 5, 3, 64, 128)])
def test_conv3x3(B, H, W, Cin, Cout):
    x, w = rnd(B, H, W, Cin, seed=17), rnd(Cout, Cin, 3, 3, seed=18, scale=(9 * Cin) ** -0.5)
    bias = rnd(Cout, seed=19, scale=0.1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias.float(), padding=1).permute(0, 2, 3, 1)
    wf = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    assert rel_l2(ops.conv3x3(x, wf, bias, out_dtype=torch.float32), ref) <= 2e-5
    if Cout % 8 == 0:
        res = rnd(B, H, W, Cout, seed=20)
        assert rel_l2(ops.conv3x3(x, wf, bias, resid=res), ref + res.float()) <= 5e-3
    with pytest.raises(ValueError):
        ops.conv3x3(rnd(1, 4, 4, 16), rnd(8, 144), None)  # Cin must be a multiple of 64


@pytest.mark.parametrize("variant", [0, 1, 4, 5, 6, 7])
@pytest.mark.parametrize("B,H,S", [(1, 1, 128), (2, 3, 512), (1, 2, 320), (1, 1, 77), (1, 2, 1280), (1, 1, 1)])
def test_attention(variant, B, H, S):
    q, k, v = rnd(B, H, S, 128, seed=21), rnd(B, H, S, 128, seed=22), rnd(B, H, S, 128, seed=23)
    out = torch.zeros(B, S, H * 128 + 64, device=dev, dtype=bf)
    ops.attention(q, k, v, out[:, :, :H * 128], 128 ** -0.5, variant=variant)
    ref = F.scaled_dot_product_attention(q.float(), k.float(), v.float()).transpose(1, 2).reshape(B, S, -1)
    assert rel_l2(out[:, :, :H * 128], ref) <= 5e-3
    assert out[:, :, H * 128:].abs().max().item() == 0


@pytest.mark.parametrize("variant", [0, 5, 6, 7])
def test_attention_sharp_softmax_and_properties(variant):
    q, k, v = rnd(1, 2, 512, 128, seed=24, scale=4), rnd(1, 2, 512, 128, seed=25, scale=4), rnd(1, 2, 512, 128, seed=26)
    out = torch.empty(1, 512, 256, device=dev, dtype=bf)
    ops.attention(q, k, v, out, 128 ** -0.5, variant=variant)
    ref = F.scaled_dot_product_attention(q.float(), k.float(), v.float()).transpose(1, 2).reshape(1, 512, 256)
    assert rel_l2(out, ref) <= 5e-3
    # BASELINE size (one image, N = 4352): rows are convex combinations of V rows, and permuting the
    # keys/values together leaves the output unchanged
    q, k, v = rnd(1, 24, 4352, 128, seed=27), rnd(1, 24, 4352, 128, seed=28), rnd(1, 24, 4352, 128, seed=29)
    o1 = torch.empty(1, 4352, 3072, device=dev, dtype=bf)
    ops.attention(q, k, v, o1, 128 ** -0.5, variant=variant)
    assert o1.float().abs().max() <= v.float().abs().max() + 1e-2
    perm = torch.randperm(4352, generator=torch.Generator().manual_seed(30)).to(dev)
    o2 = torch.empty_like(o1)
    ops.attention(q, k[:, :, perm].contiguous(), v[:, :, perm].contiguous(), o2, 128 ** -0.5, variant=variant)
    assert rel_l2(o2, o1) <= 5e-3
    ref = F.scaled_dot_product_attention(q[:, :2].float(), k[:, :2].float(), v[:, :2].float()).transpose(1, 2).reshape(1, 4352, 256)
    assert rel_l2(o1[:, :, :256], ref) <= 5e-3


def test_persistent_attention_is_bit_identical_to_one_cta_per_item():
    """variant 7 (attn_pkernel: one CTA per SM walking the (q-block, head, batch) items) runs the same arithmetic per
    query row as the one-CTA-per-item kernel: same bits, with more items than SMs (several items per CTA), a ragged
    last key tile, an odd number of key tiles and a single item."""
    for B, H, S in [(2, 24, 4352), (3, 5, 1100), (1, 7, 640), (1, 1, 200)]:
        q, k, v = rnd(B, H, S, 128, seed=51), rnd(B, H, S, 128, seed=52), rnd(B, H, S, 128, seed=53)
        a = torch.zeros(B, S, H * 128, device=dev, dtype=bf)
        b = torch.zeros_like(a)
        ops.attention(q, k, v, a, 128 ** -0.5, variant=0)
        ops.attention(q, k, v, b, 128 ** -0.5, variant=7)
        assert torch.equal(a, b), (B, H, S)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 32, 128, 128), (1, 24, 40, 256, 256), (2, 13, 21, 512, 512), (1, 64, 64, 128, 256)])
def test_conv3x3_groupnorm_partials(B, H, W, Cin, Cout):
    """fx_conv3x3's gn_partials (GroupNorm statistics of the output accumulated in the epilogue) against the
    standalone statistics pass over the stored output; ragged image sizes leave partly empty pixel blocks."""
    x, w = rnd(B, H, W, Cin, seed=41), rnd(Cout, 9 * Cin, seed=42, scale=(9 * Cin) ** -0.5)
    bias, res = rnd(Cout, seed=43, scale=0.5), rnd(B, H, W, Cout, seed=44)
    gw, gb = (1 + rnd(Cout, seed=45, scale=0.1).float()).to(bf), rnd(Cout, seed=46, scale=0.1)
    for resid in (None, res):
        out, part = ops.conv3x3(x, w, bias, resid=resid, gn_stats=True)
        assert torch.equal(out, ops.conv3x3(x, w, bias, resid=resid))          # the output itself is untouched
        ref = F.group_norm(out.float().permute(0, 3, 1, 2), 32, gw.float(), gb.float(), eps=1e-6).permute(0, 2, 3, 1)
        fused = ops.groupnorm(out, gw, gb, 1e-6, False, partials=part)
        plain = ops.groupnorm(out, gw, gb, 1e-6, False)
        assert rel_l2(fused, ref) <= 5e-3 and rel_l2(fused, plain) <= 1e-3
        sums = part[0].double().sum(1)                                          # [B, 32, 2]
        o = out.double().view(B, H * W, 32, Cout // 32)
        assert torch.allclose(sums[..., 0], o.sum((1, 3)), rtol=1e-4, atol=1e-2)
        assert torch.allclose(sums[..., 1], (o * o).sum((1, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("B,H,W,C,Co", [(2, 8, 16, 128, 128), (1, 13, 21, 256, 256), (1, 32, 32, 512, 512), (2, 5, 40, 64, 256)])
def test_conv3x3_fused_upsample(B, H, W, C, Co):
    """fx_conv3x3(upsample2x): conv3x3(upsample_nearest(x, 2)) as four parity-wise 2x2 convolutions of the low-resolution
    input (ops.upconv_weights) vs the fp32 convolution of the materialised upsampled tensor, vs the unfused kernel pair,
    and its GroupNorm partial sums; ragged sizes leave partly empty tiles."""
    x, w = rnd(B, H, W, C, seed=71), rnd(Co, 9 * C, seed=72, scale=(9 * C) ** -0.5)
    bias = rnd(Co, seed=73, scale=0.5)
    w4 = ops.upconv_weights(w)
    assert w4.shape == (4 * Co, 4 * C)
    up = x.float().repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    ref = F.conv2d(up.permute(0, 3, 1, 2), w.float().view(Co, 3, 3, C).permute(0, 3, 1, 2), bias.float(), padding=1).permute(0, 2, 3, 1)
    out, part = ops.conv3x3(x, w4, bias, gn_stats=True, upsample=True)
    assert out.shape == (B, 2 * H, 2 * W, Co)
    assert rel_l2(out, ref) <= 5e-3, rel_l2(out, ref)
    assert rel_l2(out, ops.conv3x3(ops.upsample2x(x), w, bias)) <= 5e-3
    assert torch.equal(out, ops.conv3x3(x, w4, bias, upsample=True))
    sums = part[0].double().sum(1)
    o = out.double().view(B, 4 * H * W, 32, Co // 32)
    assert torch.allclose(sums[..., 0], o.sum((1, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(sums[..., 1], (o * o).sum((1, 3)), rtol=1e-4, atol=1e-2)
    with pytest.raises(ValueError):
        ops.conv3x3(x, w4, bias, resid=out, upsample=True)     # no residual on the fused upsample path


def test_rownorm_wide_rows_block_kernel():
    """D >= 1024 runs one block per row (rownorm_block_kernel): all three modes, D = 1024 / 3072 / 4096 and a D that
    leaves the last vector slot partly empty, strided views, and the e4m3 output against quantize_rows(bf16 output)."""
    for D in (1024, 3072, 4096, 1536):
        B, R = 2, 37
        x, sh, sc = rnd(B, R, D, seed=61), rnd(B, D, seed=62, scale=0.1), rnd(B, D, seed=63, scale=0.1)
        ref = (1 + sc.float()[:, None]) * F.layer_norm(x.float(), (D,), eps=1e-6) + sh.float()[:, None]
        y = ops.rownorm(x, 0, sh, sc, 1e-6)
        assert rel_l2(y, ref) <= 5e-3
        w, b = (1 + rnd(D, seed=64, scale=0.1).float()).to(bf), rnd(D, seed=65, scale=0.1)
        assert rel_l2(ops.rownorm(x, 1, w, b, 1e-5), F.layer_norm(x.float(), (D,), w.float(), b.float(), 1e-5)) <= 5e-3
        ref2 = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()
        assert rel_l2(ops.rownorm(x, 2, w, None, 1e-6), ref2) <= 5e-3
        wide = rnd(B, R + 5, D + 64, seed=66)                                   # row / column slice of a wider buffer
        view = wide[:, 3:3 + R, :D]
        assert torch.equal(ops.rownorm(view, 0, sh, sc, 1e-6), ops.rownorm(view.contiguous(), 0, sh, sc, 1e-6))
        q8 = torch.empty(B, R, D, device=dev, dtype=ops.fp8)
        s8 = torch.empty(B, R, device=dev, dtype=torch.float32)
        ops.rownorm(x, 0, sh, sc, 1e-6, out=q8, out_scale=s8)
        qr, sr = ops.quantize_rows(y)
        assert torch.equal(q8.view(torch.uint8), qr.view(torch.uint8)) and torch.equal(s8, sr)


def test_rownorm_gemv_and_friends():
    B, R, D = 2, 100, 3072
    x, sh, sc = rnd(B, R, D, seed=31), rnd(B, D, seed=32, scale=0.1), rnd(B, D, seed=33, scale=0.1)
    ref = (1 + sc.float()[:, None]) * F.layer_norm(x.float(), (D,), eps=1e-6) + sh.float()[:, None]
    assert rel_l2(ops.rownorm(x, 0, sh, sc, 1e-6), ref) <= 5e-3
    w, b = (1 + rnd(D, seed=34, scale=0.1).float()).to(bf), rnd(D, seed=35, scale=0.1)
    assert rel_l2(ops.rownorm(x, 1, w, b, 1e-5), F.layer_norm(x.float(), (D,), w.float(), b.float(), 1e-5)) <= 5e-3
    ref = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()
    assert rel_l2(ops.rownorm(x, 2, w, None, 1e-6), ref) <= 5e-3
    x2 = rnd(3, 7, 128, seed=36)  # D < 256 (small CLIP)
    assert rel_l2(ops.rownorm(x2, 2, w[:128].contiguous(), None, 1e-6),
                  x2.float() * torch.rsqrt(x2.float().pow(2).mean(-1, keepdim=True) + 1e-6) * w[:128].float()) <= 5e-3
    xin, wv, bv, add = rnd(11, 3072, seed=37), rnd(1000, 3072, seed=38, scale=3072 ** -0.5), rnd(1000, seed=39), rnd(11, 1000, seed=40)
    assert rel_l2(ops.gemv(xin, wv, bv, silu_in=True), F.silu(xin.float()).to(bf).float() @ wv.float().T + bv.float()) <= 5e-3
    assert rel_l2(ops.gemv(xin, wv, bv, add=add), xin.float() @ wv.float().T + bv.float() + add.float()) <= 5e-3


def test_integer_and_indexing_kernels_bit_exact():
    g = torch.Generator().manual_seed(41)
    lat = torch.randn(2, 8, 12, 16, generator=g).to(bf).to(dev)
    p = ops.patchify(lat)
    assert torch.equal(p, lat.reshape(2, 4, 2, 6, 2, 16).permute(0, 1, 3, 5, 2, 4).reshape(2, 24, 64))
    z = ops.unpatchify_scale(p, (8, 12), 64, 1.0, 0.0)  # identity affine -> exact inverse of patchify
    assert torch.equal(z[..., :16], lat) and z[..., 16:].abs().max().item() == 0
    xu = rnd(2, 5, 7, 64, seed=42)
    assert torch.equal(ops.upsample2x(xu), xu.repeat_interleave(2, 1).repeat_interleave(2, 2))
    xt = rnd(100, 70, seed=43)
    assert torch.equal(ops.transpose(xt), xt.T.contiguous())
    xi = torch.randn(1000, generator=g).to(dev) * 1.5
    img, u8 = ops.finish_image(xi)
    refi = torch.clip(xi + 1, 0, 2) * 0.5
    assert torch.equal(img, refi) and torch.equal(u8, (refi * 255).to(torch.uint8))
    ids = torch.randint(0, 100, (2, 9), generator=g, dtype=torch.int32).to(dev)
    tab = rnd(100, 256, seed=44)
    assert torch.equal(ops.embedding(ids, tab), tab[ids.long()])
    xt_, pr = rnd(2, 64, 64, seed=45), rnd(2, 64, 64, seed=46)
    assert torch.equal(ops.euler_step(xt_.clone(), pr, -0.25), (xt_.float() + (-0.25 * pr.float()).to(bf).float()).to(bf))
    with pytest.raises(ValueError):
        ops.patchify(rnd(1, 7, 8, 16))  # odd latent height


def test_groupnorm_softmax_attention_small():
    for C in (64, 128, 256, 512):
        x = (rnd(2, 24, 40, C, seed=47).float() * 2 + 0.5).to(bf)
        w, b = (1 + rnd(C, seed=48, scale=0.1).float()).to(bf), rnd(C, seed=49, scale=0.1)
        ref = F.group_norm(x.float().permute(0, 3, 1, 2), 32, w.float(), b.float(), 1e-6).permute(0, 2, 3, 1)
        assert rel_l2(ops.groupnorm(x, w, b, 1e-6, False), ref) <= 5e-3
        assert rel_l2(ops.groupnorm(x, w, b, 1e-6, True), F.silu(ref)) <= 6e-3
    s = torch.randn(300, 1024, generator=torch.Generator().manual_seed(50)).to(dev) * 5
    assert rel_l2(ops.softmax_rows(s, 0.3), torch.softmax(s * 0.3, -1)) <= 5e-3
    Bq, S, H = 2, 77, 4
    qkv = rnd(Bq, S, 3 * H * 64, seed=51)
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    bias = torch.randn(H, S, S, generator=torch.Generator().manual_seed(52)).to(dev)
    hd = lambda t_: t_.float().reshape(Bq, S, H, 64).transpose(1, 2)  # noqa: E731
    ref = F.scaled_dot_product_attention(hd(q), hd(k), hd(v), attn_mask=bias[None], scale=1.0).transpose(1, 2).reshape(Bq, S, -1)
    assert rel_l2(ops.attention_small(q, k, v, H, 1.0, bias=bias), ref) <= 5e-3
    ref = F.scaled_dot_product_attention(hd(q), hd(k), hd(v), is_causal=True).transpose(1, 2).reshape(Bq, S, -1)
    assert rel_l2(ops.attention_small(q, k, v, H, 0.125, causal=True), ref) <= 5e-3


@pytest.mark.parametrize("S,causal", [(512, False), (700, False), (1300, True), (1030, False)])
def test_attention_small_long_sequences(S, causal):
    """fx_attention_small beyond 512 keys (online softmax over 512-key chunks): the reference's T5 tokenizer never
    truncates (flux/tokenizers.py:160-173), so --no-t5-padding prompts longer than 512 tokens must encode."""
    Bq, H = 1, 3
    qkv = rnd(Bq, S, 3 * H * 64, seed=54)
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    hd = lambda t_: t_.float().reshape(Bq, S, H, 64).transpose(1, 2)  # noqa: E731
    if causal:
        ref = F.scaled_dot_product_attention(hd(q), hd(k), hd(v), is_causal=True)
        out = ops.attention_small(q, k, v, H, 0.125, causal=True)
    else:
        bias = torch.randn(H, S, S, generator=torch.Generator().manual_seed(55)).to(dev)
        ref = F.scaled_dot_product_attention(hd(q), hd(k), hd(v), attn_mask=bias[None], scale=1.0)
        out = ops.attention_small(q, k, v, H, 1.0, bias=bias)
    assert rel_l2(out, ref.transpose(1, 2).reshape(Bq, S, -1)) <= 5e-3


def _philox4x32_10(ctr, key):
    import numpy as np
    c = [np.uint32(v) for v in ctr]
    k = [np.uint32(v) for v in key]
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * np.uint64(c[0])
        p1 = np.uint64(0xCD9E8D57) * np.uint64(c[2])
        hi0, lo0 = np.uint32(p0 >> np.uint64(32)), np.uint32(p0 & np.uint64(0xFFFFFFFF))
        hi1, lo1 = np.uint32(p1 >> np.uint64(32)), np.uint32(p1 & np.uint64(0xFFFFFFFF))
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [np.uint32((int(k[0]) + 0x9E3779B9) & 0xFFFFFFFF), np.uint32((int(k[1]) + 0xBB67AE85) & 0xFFFFFFFF)]
    return [int(v) for v in c]


def test_prior_kernel_philox_known_answers_and_sharding():
    """Device prior (sample_prior + patchify fused): counter-based, so (a) values follow from a CPU Philox4x32-10,
    (b) an image's noise does not depend on batch size or on which rank generates it, (c) ~N(0, 1)."""
    import math
    assert _philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]  # Random123 KAT
    h, w, seed = 8, 12, 1234
    full = ops.prior_packed(5, (h, w), 16, seed, 0)
    z = ops.unpatchify_scale(full, (h, w), 16, 1.0, 0.0)           # back to NHWC [5, h, w, 16]
    for (bi, e) in ((0, 0), (3, 100), (4, h * w * 16 - 4)):         # element groups of 4 = one Philox call
        r = _philox4x32_10([e // 4, 0, bi, 0], [seed, 0])
        exp = []
        for kk in range(2):
            u1 = ((r[2 * kk] >> 8) + 1) / 16777216.0
            u2 = (r[2 * kk + 1] >> 8) / 16777216.0
            rad = math.sqrt(-2.0 * math.log(u1))
            exp += [rad * math.cos(2 * math.pi * u2), rad * math.sin(2 * math.pi * u2)]
        got = z[bi].flatten()[e:e + 4].float().cpu()
        assert torch.allclose(got, torch.tensor(exp).to(bf).float(), atol=2e-2, rtol=2e-2), (bi, e, got, exp)
    assert torch.equal(ops.prior_packed(2, (h, w), 16, seed, 3), full[3:5])      # rank holding images 3..4
    assert torch.equal(ops.prior_packed(5, (h, w), 16, seed, 0), full)           # deterministic
    assert not torch.equal(ops.prior_packed(1, (h, w), 16, seed + 1, 0), full[:1])
    big = ops.prior_packed(4, (128, 128), 16, 7, 0).float()
    assert abs(big.mean().item()) < 5e-3 and abs(big.std().item() - 1.0) < 5e-3
    with pytest.raises(ValueError):
        ops.prior_packed(1, (7, 8), 16, 0, 0)
