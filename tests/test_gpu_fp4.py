"""GPU parity of the NVFP4 (W4A4) path -- `--quantize 4` -- through the C ABI (include/flux_b200.h: fx_quantize_rows_fp4,
fx_gemm_fp4; csrc/gemm4.cu).

The reference's --quantize is MLX 4-bit group quantisation of the Linear weights (txt2image.py:28-29,79-82), not
restatable without MLX's packed format; this mode is pinned like the FP8 one:
  * quantiser: e2m1 bytes, UE4M3 block scales and fp32 row scales BIT-EXACT against the oracle's restatement
    (oracle.flux_oracle.nvfp4_quant_rows) -- integer / byte work;
  * GEMM (tcgen05.mma.kind::mxf4nvf4.block_scale) vs an fp32 matmul of the SAME dequantised operands: rel-L2 <= 2e-5
    (fp32 out) / 5e-3 (bf16 out) -- products of e2m1 x ue4m3 values are exact, only the summation order differs;
  * what 4 bits cost against the unquantised product is printed (about 1e-1 on Gaussian data) -- a property of the
    format, stated, not a kernel tolerance.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from flux import ops  # noqa: E402
from helpers import rel_l2  # noqa: E402
from oracle import flux_oracle as O  # noqa: E402

dev = "cuda"
bf = torch.bfloat16
E2M1 = torch.tensor([0, .5, 1, 1.5, 2, 3, 4, 6, -0., -.5, -1, -1.5, -2, -3, -4, -6])


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(bf).to(dev)


def decode(q, sf, scale, K):
    """(bytes, atoms, row scales) -> float matrix [rows, K]"""
    rows = q.shape[0]
    lut = E2M1.to(q.device)
    vals = torch.stack([lut[(q & 15).long()], lut[(q >> 4).long()]], dim=-1).reshape(rows, K)
    sfm = ops.sf_atoms_to_matrix(sf, K)[:rows].view(torch.float8_e4m3fn).float()
    return vals * sfm.repeat_interleave(16, dim=1) * scale[:, None]


@pytest.mark.parametrize("B,R,K", [(1, 128, 256), (2, 200, 3072), (1, 77, 15360), (3, 128, 1024)])
def test_quantize_rows_fp4_bit_exact_vs_oracle(B, R, K):
    x = rnd(B, R, K, seed=1)
    x[0, 3] = 0                                   # an all-zero row
    x[0, 5, 32:48] = 0                            # an all-zero block
    x[0, 7, 100] = 300.0                          # an outlier: most other blocks of that row then quantise coarsely
    q, sf, scale = ops.quantize_rows_fp4(x)
    oq, osf, og = O.nvfp4_quant_rows(x.float().cpu().view(B * R, K))
    lut = E2M1.to(dev)
    vals = torch.stack([lut[(q & 15).long()], lut[(q >> 4).long()]], dim=-1).reshape(B * R, K)
    assert torch.equal(scale.cpu(), og.view(-1))
    sfm = ops.sf_atoms_to_matrix(sf, K)[:B * R].view(torch.float8_e4m3fn).float()
    assert torch.equal(sfm.cpu(), osf)
    assert torch.equal(vals.cpu(), oq)
    # strided view (row / column slice of a wider buffer) gives the same bytes
    wide = torch.zeros(B, R + 3, K + 64, device=dev, dtype=bf)
    wide[:, 2:2 + R, 64:] = x
    q2, sf2, s2 = ops.quantize_rows_fp4(wide[:, 2:2 + R, 64:])
    assert torch.equal(q2, q) and torch.equal(sf2, sf) and torch.equal(s2, scale)
    deq = decode(q, sf, scale, K)
    print(f"NVFP4 quantisation error B={B} R={R} K={K}: rel-L2 {rel_l2(deq, x.view(B * R, K)):.3e}")
    with pytest.raises(ValueError):
        ops.quantize_rows_fp4(rnd(4, 100))        # K must be a multiple of 64


@pytest.mark.parametrize("B,R,N,K", [(1, 128, 192, 256), (2, 256, 3072, 3072), (1, 200, 500, 1024), (8, 128, 384, 15360),
                                     (1, 200, 384, 512), (3, 384, 200, 256)])
def test_gemm_fp4_vs_dequantised_matmul(B, R, N, K):
    a, w = rnd(B, R, K, seed=2), rnd(N, K, seed=3, scale=K ** -0.5)
    a4, sfa, sa = ops.quantize_rows_fp4(a)
    w4, sfw, sw = ops.fp4_weight(w)
    wq, wsf, _ = ops.quantize_rows_fp4(w)
    ref = decode(a4, sfa, sa, K) @ decode(wq, wsf, sw, K).T
    out = ops.gemm_fp4(a4, sfa, sa, w4, sfw, sw, B, out_dtype=torch.float32)
    assert out.shape == (B, R, N)
    e = rel_l2(out.view(B * R, N), ref)
    print(f"NVFP4 GEMM B={B} R={R} N={N} K={K}: vs dequantised fp32 matmul {e:.2e}; vs the unquantised product "
          f"{rel_l2(out.view(B * R, N), a.float().view(B * R, K) @ w.float().T):.2e}")
    assert e <= 2e-5
    # the generic epilogue: bias, GELU, gate, residual, bf16 output into a strided view
    bias, gate = rnd(N, seed=4, scale=0.1), rnd(B, N, seed=5, scale=0.5)
    res = rnd(B, R, N + 32, seed=6)
    buf = res.clone()
    ops.gemm_fp4(a4, sfa, sa, w4, sfw, sw, B, bias=bias, gate=gate, resid=buf[:, :, :N], out=buf[:, :, :N])
    want = res[:, :, :N].float() + gate.float()[:, None] * (ref.view(B, R, N) + bias.float())
    assert rel_l2(buf[:, :, :N], want) <= 5e-3 and torch.equal(buf[:, :, N:], res[:, :, N:])
    act = ops.gemm_fp4(a4, sfa, sa, w4, sfw, sw, B, bias=bias, act="gelu_tanh")
    assert rel_l2(act, torch.nn.functional.gelu(ref.view(B, R, N) + bias.float(), approximate="tanh")) <= 5e-3
    # bf16 output into a column window of a wider, row-padded buffer (the GELU -> `cat` case; TMA stores clip rows / columns)
    wide = torch.full((B, R + 3, N + 64), 7.0, device=dev, dtype=bf)
    ops.gemm_fp4(a4, sfa, sa, w4, sfw, sw, B, bias=bias, out=wide[:, :R, 32:32 + N])
    assert rel_l2(wide[:, :R, 32:32 + N], ref.view(B, R, N) + bias.float()) <= 5e-3
    assert (wide[:, R:] == 7.0).all() and (wide[:, :, :32] == 7.0).all() and (wide[:, :, 32 + N:] == 7.0).all()


@pytest.mark.parametrize("B,R,H,K,off,f8out", [(1, 128, 1, 256, 0, False), (2, 256, 3, 3072, 32, False), (1, 200, 2, 1024, 40, False),
                                               (2, 256, 24, 3072, 64, True)])
def test_gemm_fp4_qkv_epilogue(B, R, H, K, off, f8out):
    """fx_gemm_fp4_qkv (128-column tiles, one head each; the two epilogue warp groups take alternate accumulators) vs the fp32
    restatement of Linear -> QK-RMSNorm -> RoPE (flux/layers.py:195-214) on the SAME dequantised operands: rel-L2 <= 5e-3 for
    bf16 outputs, one e4m3 step (4e-2) for e4m3 outputs; rows before seq_off stay untouched."""
    from test_gpu_kernels import _qkv_ref
    D = H * 128
    a, w = rnd(B, R, K, seed=11), rnd(3 * D, K, seed=12, scale=K ** -0.5)
    bias, qs, ks = rnd(3 * D, seed=13, scale=0.1), (1 + rnd(128, seed=14, scale=0.1).float()).to(bf), \
        (1 + rnd(128, seed=15, scale=0.1).float()).to(bf)
    ang = torch.rand(R + off, 64, generator=torch.Generator().manual_seed(16)) * 6.28
    pe = torch.stack([torch.cos(ang), torch.sin(ang)], -1).to(bf).to(dev)
    a4, sfa, sa = ops.quantize_rows_fp4(a)
    w4, sfw, sw = ops.fp4_weight(w, ops.FP4_TILE_N_QKV)
    wq, wsf, _ = ops.quantize_rows_fp4(w)
    ad, wd = decode(a4, sfa, sa, K).view(B, R, K), decode(wq, wsf, sw, K)
    want = _qkv_ref(ad, wd, bias, qs, ks, pe[off:], H)[:3]
    dt = ops.fp8 if f8out else bf
    got = [torch.zeros(B, H, R + off, 128, device=dev, dtype=dt) for _ in range(3)]
    ops.gemm_fp4_qkv(a4, sfa, sa, w4, sfw, sw, B, bias, qs, ks, pe, *got, off, rms_eps=1e-5)
    for g_, r_ in zip(got, want):
        assert rel_l2(g_.float()[:, :, off:], r_) <= (4e-2 if f8out else 5e-3)
        assert g_.view(torch.uint8 if f8out else torch.int16)[:, :, :off].abs().max().item() == 0 if off else True
    if off % 32 == 0 and not f8out:   # the blocked RoPE table layout gives the same bits
        got2 = [torch.zeros_like(t_) for t_ in got]
        ops.gemm_fp4_qkv(a4, sfa, sa, w4, sfw, sw, B, bias, qs, ks, ops.block_pe(pe), *got2, off, rms_eps=1e-5, pe_blocked=True)
        assert all(torch.equal(x_, y_) for x_, y_ in zip(got, got2))


def _decode_operand(q, sf, scale, K):
    return decode(q, sf.view(-1, K // 64, 512), scale, K)


@pytest.mark.parametrize("B,R,C,kc,col0", [(1, 128, 256, 256, 0), (2, 256, 3072, 15360, 0), (3, 128, 1024, 2048, 512)])
def test_chunk_quantiser_and_finalise_bit_exact_vs_oracle(B, R, C, kc, col0):
    """fx_quantize_chunks_fp4 + fx_fp4_finalize (the producer-side NVFP4 format: power-of-two chunk scales lifted to the row's)
    against oracle.nvfp4_quant_rows_chunked: e2m1 bytes, UE4M3 block scales and row scales bit-exact, also when the chunks fill
    only a column window of a wider operand whose other columns come from elsewhere."""
    rows = B * R
    full = rnd(B, R, kc, seed=51) * torch.logspace(-3, 2, kc, device=dev).to(bf)[None, None]   # magnitudes spread over 1e5
    full[0, 0] = 0                                                                               # a zero row
    op = ops.Fp4Operand(rows, kc, dev)
    dst = op.view(rows, kc)
    if col0 or C != kc:   # the other columns first (two calls: left and right of the window)
        if col0:
            ops.quantize_chunks_fp4(full[:, :, :col0], dst, 0)
        if col0 + C < kc:
            ops.quantize_chunks_fp4(full[:, :, col0 + C:], dst, col0 + C)
    ops.quantize_chunks_fp4(full[:, :, col0:col0 + C], dst, col0)
    q, sf, scale = ops.fp4_finalize(dst)
    oq, osf, og = O.nvfp4_quant_rows_chunked(full.view(rows, kc).float().cpu())
    lut = {float(v): i for i, v in enumerate(E2M1.tolist()[:8])}
    code = torch.tensor([[lut[abs(float(v))] + (8 if (v < 0 or (v == 0 and torch.signbit(v))) else 0) for v in row] for row in oq[:4]])
    got = torch.stack([(q[:4] & 15), (q[:4] >> 4)], dim=-1).reshape(4, kc).cpu()
    assert torch.equal(got & 7, code & 7)                                       # magnitudes of the first rows, nibble by nibble
    deq = _decode_operand(q, sf, scale, kc).cpu()
    want = oq * osf.repeat_interleave(16, dim=-1) * og
    assert torch.equal(deq, want)                                               # every dequantised value identical
    assert torch.equal(scale.cpu(), og.flatten())
    sfm = ops.sf_atoms_to_matrix(sf.view(-1, kc // 64, 512), kc)[:rows].view(torch.float8_e4m3fn).float().cpu()
    assert torch.equal(sfm, osf)
    print(f"chunked NVFP4 error rows={rows} kc={kc}: rel-L2 {rel_l2(deq, full.view(rows, kc).float().cpu()):.3e}")


@pytest.mark.parametrize("B,R,N,K,kc,col0", [(1, 128, 192, 256, 256, 64), (2, 256, 3072, 3072, 3072, 0), (8, 128, 384, 1024, 1536, 1152)])
def test_gemm_fp4_epilogue_emits_the_next_operand(B, R, N, K, kc, col0):
    """fx_gemm_fp4 with q_out: act(A W^T + bias) leaves the epilogue as columns [col0, col0 + N) of the next GEMM's NVFP4 operand.
    Against the oracle's chunked quantisation of the fp32 reference result: dequantised rel-L2 <= 2e-2 and >= 97 % identical
    nibbles (the two differ only where the fp32 summation order moves a value across a rounding boundary); the remaining columns of
    the operand, produced by the chunk quantiser, are untouched by the GEMM."""
    a, w = rnd(B, R, K, seed=2), rnd(N, K, seed=3, scale=K ** -0.5)
    bias = rnd(N, seed=4, scale=0.1)
    a4, sfa, sa = ops.quantize_rows_fp4(a)
    w4, sfw, sw = ops.fp4_weight(w)
    wq, wsf, _ = ops.quantize_rows_fp4(w)
    ref = torch.nn.functional.gelu(decode(a4, sfa, sa, K) @ decode(wq, wsf, sw, K).T + bias.float(), approximate="tanh")
    rows = B * R
    other = rnd(B, R, kc, seed=9)
    op = ops.Fp4Operand(rows, kc, dev)
    dst = op.view(rows, kc)
    if col0:
        ops.quantize_chunks_fp4(other[:, :, :col0], dst, 0)
    if col0 + N < kc:
        ops.quantize_chunks_fp4(other[:, :, col0 + N:], dst, col0 + N)
    assert ops.gemm_fp4(a4, sfa, sa, w4, sfw, sw, B, bias=bias, act="gelu_tanh", out4=dst, out4_col0=col0) is None
    q, sf, scale = ops.fp4_finalize(dst)
    deq = _decode_operand(q, sf, scale, kc)
    full = other.view(rows, kc).float().clone()
    full[:, col0:col0 + N] = ref
    oq, osf, og = O.nvfp4_quant_rows_chunked(full.cpu())
    want = (oq * osf.repeat_interleave(16, dim=-1) * og).to(dev)
    win = slice(col0, col0 + N)
    assert rel_l2(deq[:, win], want[:, win]) <= 2e-2 and rel_l2(deq[:, win], ref) <= 1.2e-1
    same = (deq[:, win] == want[:, win]).float().mean().item()
    print(f"epilogue-quantised operand B={B} R={R} N={N}: identical values {same:.4f}, vs fp32 result {rel_l2(deq[:, win], ref):.3e}")
    assert same >= 0.97
    keep = torch.ones(kc, dtype=torch.bool, device=dev)
    keep[win] = False
    # the other columns: same nibbles and block scales as the oracle's (their chunk exponents do not depend on the GEMM's columns;
    # the row scale does, through the max -- already inside `want`)
    if keep.any():
        assert (deq[:, keep] == want[:, keep]).float().mean().item() >= 0.999


@pytest.mark.parametrize("f8,split,col0,kc", [(True, 0, 0, 384), (True, 128, 0, 384), (False, 0, 256, 1024)])
def test_attention_epilogue_emits_nvfp4_chunks(f8, split, col0, kc):
    """fx_attention with q_out: O / l leaves the epilogue as NVFP4 chunks of the next GEMM's operand(s).  Against the oracle's
    chunked quantisation of the SAME kernel's bf16 output: >= 95 % identical values and rel-L2 <= 5e-2 (the epilogue quantises the
    fp32 O / l, the comparison its bf16 rounding -- a value moves, by one e2m1 step, only when that rounding crosses an e2m1
    boundary: measured 3.5 % of the values, rel-L2 3.6e-2, against 9e-2 for the 4-bit format itself); with
    split > 0 the rows below it land in the second operand; other columns of the operand are untouched."""
    B, H, S = 2, 3, 384
    q, k, v = (rnd(B, H, S, 128, seed=60 + i) for i in range(3))
    if f8:
        q, k, v = (t_.to(ops.fp8) for t_ in (q, k, v))
    ref = torch.zeros(B, S, H * 128, device=dev, dtype=bf)
    ops.attention(q, k, v, ref, 128 ** -0.5)
    hi_rows, lo_rows = B * (S - split), B * split
    op_hi, op_lo = ops.Fp4Operand(hi_rows, kc, dev), ops.Fp4Operand(max(lo_rows, 128), 384, dev)
    dst = op_hi.view(hi_rows, kc)
    other = rnd(B, S - split, kc, seed=70)
    if col0:
        ops.quantize_chunks_fp4(other[:, :, :col0], dst, 0)
    if col0 + H * 128 < kc:
        ops.quantize_chunks_fp4(other[:, :, col0 + H * 128:], dst, col0 + H * 128)
    low = op_lo.view(lo_rows, 384) if split else None
    assert ops.attention(q, k, v, None, 128 ** -0.5, out4=dst, out4_col0=col0, out4_low=low, split=split) is None
    qh, sfh, sh = ops.fp4_finalize(dst)
    full = other.reshape(hi_rows, kc).float().clone()
    full[:, col0:col0 + H * 128] = ref[:, split:].reshape(hi_rows, H * 128).float()
    oq, osf, og = O.nvfp4_quant_rows_chunked(full.cpu())
    want = (oq * osf.repeat_interleave(16, dim=-1) * og).to(dev)
    got = _decode_operand(qh, sfh, sh, kc)
    win = slice(col0, col0 + H * 128)
    same = (got[:, win] == want[:, win]).float().mean().item()
    print(f"attention-emitted operand f8={f8} split={split}: identical {same:.4f}, rel-L2 vs oracle-quantised bf16 output "
          f"{rel_l2(got[:, win], want[:, win]):.3e}, vs the bf16 output {rel_l2(got[:, win], full[:, win].to(dev)):.3e}")
    assert same >= 0.95 and rel_l2(got[:, win], want[:, win]) <= 5e-2
    assert rel_l2(got[:, win], full[:, win].to(dev)) <= 1.2e-1   # and it IS the attention output, to 4-bit accuracy
    keep = torch.ones(kc, dtype=torch.bool, device=dev)
    keep[win] = False
    if keep.any():
        assert (got[:, keep] == want[:, keep]).float().mean().item() >= 0.999
    if split:
        ql, sfl, sl = ops.fp4_finalize(low)
        lo_ref = ref[:, :split].reshape(lo_rows, H * 128).float()
        oq, osf, og = O.nvfp4_quant_rows_chunked(lo_ref.cpu())
        want_lo = (oq * osf.repeat_interleave(16, dim=-1) * og).to(dev)
        got_lo = _decode_operand(ql, sfl, sl, 384)
        assert (got_lo == want_lo).float().mean().item() >= 0.95 and rel_l2(got_lo, want_lo) <= 5e-2


@pytest.mark.parametrize("B,R,D,mode", [(2, 256, 3072, 0), (1, 128, 4096, 2), (3, 128, 1024, 1)])
def test_rownorm_nvfp4_output_bit_identical_to_two_kernels(B, R, D, mode):
    """fx_rownorm with out_fp8 == 2 (the AdaLN / LayerNorm / RMSNorm row kernel writing the NVFP4 operand directly) produces
    the bytes, scale atoms and row scales of fx_quantize_rows_fp4 applied to its bf16 output -- integer work: bit-exact."""
    x = rnd(B, R + 8, D, seed=41)[:, 8:]                       # a strided view: rows start inside the buffer
    if mode == 0:
        p0, p1 = rnd(B, D, seed=42, scale=0.1), rnd(B, D, seed=43, scale=0.1)
    else:
        p0, p1 = (1 + rnd(D, seed=42, scale=0.1).float()).to(bf), (rnd(D, seed=43, scale=0.1) if mode == 1 else None)
    want = ops.quantize_rows_fp4(ops.rownorm(x, mode, p0, p1, 1e-6))
    rows = B * R
    bufs = (torch.full((rows * D // 2 + 64,), 0xAA, device=dev, dtype=torch.uint8),
            torch.zeros((rows // 128) * (D // 64) * 512, device=dev, dtype=torch.uint8), torch.zeros(rows, device=dev))
    got = ops.rownorm(x, mode, p0, p1, 1e-6, out_fp4=bufs)
    for g_, w_ in zip(got, want):
        assert g_.shape == w_.shape and torch.equal(g_, w_)
    assert (bufs[0][rows * D // 2:] == 0xAA).all()              # nothing written past the operand
    with pytest.raises(ValueError):
        ops.rownorm(rnd(1, 100, D, seed=1), mode, p0, p1, 1e-6, out_fp4=bufs)   # rows must fill whole 128-row scale atoms


@pytest.mark.parametrize("scope", ["all", "cat"])
def test_flow_nvfp4_full_width_vs_quantised_oracle(scope):
    """hidden 3072 / 24 heads, depth 1+1, batch 2, N = 128 + 384: Flux.quantize(bits=4) -- NVFP4 for every block Linear
    (scope "all") or for proj / mlp.2 / linear2 with FP8 elsewhere ("cat") -- against the oracle's restatement of the same
    formats -- rel-L2 <= 3e-2 -- and against the fp32 oracle: what 4-bit operands cost (stated bound for this mode:
    rel-L2 <= 2.5e-1, cosine >= 0.97; printed)."""
    from flux import specs, synthetic
    from flux.model import Flux
    from helpers import cosine
    p = specs.FluxParams(depth=1, depth_single_blocks=1, guidance_embed=True)
    sd = synthetic.synthetic_state_dict(specs.flow_manifest(p))
    model = Flux(p, device=dev).load_weights(list(sd.items()))
    model.quantize(bits=4, fp4_scope=scope)
    assert len(model._q4) == (2 * 4 + 3 if scope == "all" else 2 * 2 + 1) and model.quantized and model._q4_fused
    g = torch.Generator().manual_seed(3)
    B, h, w, S = 2, 16, 96, 128                      # L = 384, S = 128: row counts the NVFP4 kernel tiles
    x = torch.randn(B, h, w, 16, generator=g).to(bf)
    img, ids = O.prepare_latent_images(x)
    txt = torch.randn(B, S, 4096, generator=g).to(bf)
    y = torch.randn(B, 768, generator=g).to(bf)
    tids = torch.zeros(B, S, 3, dtype=torch.int32)
    ts, gd = torch.full((B,), 0.5, dtype=bf), torch.full((B,), 4.0, dtype=bf)
    op = O.FluxParams(depth=1, depth_single_blocks=1, guidance_embed=True)
    args = (img.float(), ids, txt.float(), tids, ts, y.float(), gd)
    ref = O.flux_forward(sd, op, *args)
    ref4 = O.flux_forward(sd, op, *args, mode=O.Mode("fp32", quantize=True, bits=4, fp4_scope=scope))
    out = model(*(t_.to(dev) for t_ in (img, ids, txt, tids, ts, y, gd)))
    assert "a4" in next(iter(model._ws.values()))    # the NVFP4 path did run
    assert ("c4" in next(iter(model._ws.values()))) == (scope == "all")   # ... with producer-emitted operands under scope "all"
    if scope == "all":   # the unfused form (bf16 `cat` + row quantiser) stays available and agrees with its own oracle restatement
        model.quantize(bits=4, fp4_fused=False)
        ref4u = O.flux_forward(sd, op, *args, mode=O.Mode("fp32", quantize=True, bits=4, fp4_fused=False))
        outu = model(*(t_.to(dev) for t_ in (img, ids, txt, tids, ts, y, gd)))
        print("nvfp4 (all, unfused) vs its oracle", rel_l2(outu, ref4u), "| fused vs unfused", rel_l2(out, outu))
        assert rel_l2(outu, ref4u) <= 3e-2 and rel_l2(out, outu) <= 3e-2
        model.quantize(bits=4, fp4_scope=scope)
    print(f"nvfp4 ({scope}) vs quantised oracle", rel_l2(out, ref4), cosine(out, ref4), "| vs fp32 oracle", rel_l2(out, ref), cosine(out, ref),
          "| oracle nvfp4 vs fp32", rel_l2(ref4, ref))
    assert rel_l2(out, ref4) <= 3e-2 and cosine(out, ref4) >= 0.999
    assert rel_l2(out, ref) <= 2.5e-1 and cosine(out, ref) >= 0.97
    a = [t_.to(dev) for t_ in (img, ids, txt, tids, ts, y, gd)]
    assert torch.equal(model.forward(*a).clone(), model.forward_graphed(*a))     # graph replay: same bits
    # a shape whose row counts are not multiples of 128 falls back to the FP8 kernels for those Linears
    img2, ids2 = O.prepare_latent_images(torch.randn(1, 8, 12, 16, generator=g).to(bf))
    out2 = model(img2.to(dev), ids2.to(dev), txt[:1, :40].contiguous().to(dev), tids[:1, :40].contiguous().to(dev), ts[:1].to(dev),
                 y[:1].to(dev), gd[:1].to(dev))
    assert torch.isfinite(out2.float()).all()


def test_pipeline_nvfp4_end_to_end():
    """FluxPipeline with Flux.quantize(bits=4) -- what `txt2image.py --quantize` runs -- at a shape the NVFP4 kernels tile
    (reduced-width model of the golden fixtures, 256 image + 128 text tokens, 2 steps, 2 images): every block Linear in NVFP4 with
    producer-emitted operands, e4m3 attention, CUDA-graph replay.  Against the same pipeline in bf16: latents rel-L2 <= 5e-2,
    cosine >= 0.998, image mean |diff| <= 3/255 (measured 9.2e-3, 0.99996, 0.82/255; the full-size figures are in
    tests/test_gpu_fullsize.py); the unfused form (bf16 `cat` + row quantiser) lands within the same bound of the fused one
    (measured 5.7e-3)."""
    from flux import FluxPipeline, specs, synthetic
    from helpers import FixedTokenizer, cosine, small_configs
    fcfg, acfg, t5c, clc = small_configs()
    pipe = FluxPipeline("flux-schnell", synthetic=True, device=dev, flow_params=specs.FluxParams(**fcfg, guidance_embed=False),
                        ae_params=specs.AutoEncoderParams(**acfg), t5_config=specs.T5Config(**t5c),
                        clip_config=specs.CLIPTextModelConfig(**clc))
    for mod, man in ((pipe.flow, specs.flow_manifest(pipe.flow.params)), (pipe.ae, specs.ae_decoder_manifest(pipe.ae.params)),
                     (pipe.t5, specs.t5_manifest(pipe.t5.config)), (pipe.clip, specs.clip_manifest(pipe.clip.config))):
        sd = synthetic.synthetic_state_dict(man)
        mod.load_weights(list(mod.sanitize(sd).items()) if mod is pipe.ae else list(sd.items()))
    gen_ = torch.Generator().manual_seed(5)
    pipe.t5_tokenizer = FixedTokenizer(torch.randint(3, t5c["vocab_size"], (1, 128), generator=gen_))
    pipe.clip_tokenizer = FixedTokenizer(torch.randint(3, clc["vocab_size"] - 2, (1, 16), generator=gen_))

    def run():
        gen = pipe.generate_latents("a prompt", n_images=2, num_steps=2, latent_size=(32, 32), seed=11)
        next(gen)
        lat = list(gen)[-1]
        return lat.float(), pipe.decode(lat, (32, 32)).float()

    l16, i16 = run()
    pipe.flow.quantize(bits=4)
    l4, i4 = run()
    ws = next(iter(pipe.flow._ws.values()))
    assert pipe.flow.quantized and "c4" in ws and "p4" in ws            # NVFP4 with producer-emitted operands did run
    pipe.flow.quantize(bits=4, fp4_fused=False)
    l4u, _ = run()
    pipe.flow.dequantize()
    d = (i4 - i16).abs().mean().item() * 255
    print(f"NVFP4 pipeline vs bf16: latents rel-L2 {rel_l2(l4, l16):.3e} cosine {cosine(l4, l16):.5f} image mean |diff| {d:.2f}/255; "
          f"fused vs unfused latents {rel_l2(l4, l4u):.3e}")
    assert rel_l2(l4, l16) <= 5e-2 and cosine(l4, l16) >= 0.998 and d <= 3.0
    assert rel_l2(l4, l4u) <= 5e-2


@pytest.mark.parametrize("B,h,w,S", [(1, 64, 64, 256), (3, 32, 32, 128), (2, 48, 96, 128), (1, 128, 128, 512)])
def test_flow_nvfp4_shapes(B, h, w, S):
    """Flux.quantize(bits=4) at full width (depth 1+1) over the token counts of other image sizes / batch sizes / the dev model's
    512 text tokens -- CTA pairs (rows % 256 == 0) and single CTAs (L = 1152), batch 1 and odd batches: against the bf16 forward
    of the same model rel-L2 <= 4e-2, cosine >= 0.999; graph replay bit-identical."""
    from flux import specs, synthetic
    from flux.model import Flux
    from helpers import cosine
    p = specs.FluxParams(depth=1, depth_single_blocks=1, guidance_embed=True)
    sd = synthetic.synthetic_state_dict(specs.flow_manifest(p))
    model = Flux(p, device=dev).load_weights(list(sd.items()))
    g = torch.Generator().manual_seed(B * 1000 + S)
    img, ids = O.prepare_latent_images(torch.randn(B, h, w, 16, generator=g).to(bf))
    txt = torch.randn(B, S, 4096, generator=g).to(bf)
    y = torch.randn(B, 768, generator=g).to(bf)
    tids = torch.zeros(B, S, 3, dtype=torch.int32)
    ts, gd = torch.full((B,), 0.5, dtype=bf), torch.full((B,), 4.0, dtype=bf)
    a = [t_.to(dev) for t_ in (img, ids, txt, tids, ts, y, gd)]
    ref = model(*a).float().clone()
    model.quantize(bits=4)
    out = model(*a).float().clone()
    assert "c4" in next(iter(model._ws.values()))
    print(f"nvfp4 B={B} L={img.shape[1]} S={S}: vs bf16 rel-L2 {rel_l2(out, ref):.3e} cosine {cosine(out, ref):.5f}")
    assert torch.isfinite(out).all() and rel_l2(out, ref) <= 4e-2 and cosine(out, ref) >= 0.999
    assert torch.equal(model.forward(*a).clone(), model.forward_graphed(*a))
