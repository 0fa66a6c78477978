"""Thin torch-tensor wrappers over the C ABI (include/flux_b200.h).

torch is used for device memory and streams only; every function enqueues one library call on the
current CUDA stream.  Tensors may be strided views (column / row slices of wider buffers) as long
as the innermost stride is 1.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _native as N

ACT = {"none": 0, None: 0, "gelu_tanh": 1, "quick_gelu": 2, "gelu": 3, "gelu_erf": 3}
bf16 = torch.bfloat16
fp8 = torch.float8_e4m3fn  # storage dtype of --quantize operands (1 byte per element)


def _as3(t: torch.Tensor) -> Tuple[int, int, int, int, int]:
    """(batch, rows, cols, ld, batch_stride) of a 2-D/3-D view with unit inner stride."""
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3:
        raise ValueError(f"expected a 2-D or 3-D tensor, got {tuple(t.shape)}")
    if t.stride(2) != 1 and t.shape[2] > 1:
        raise ValueError("innermost dimension must be contiguous")
    b, r, c = t.shape
    bs = t.stride(0) if b > 1 else r * t.stride(1)
    return b, r, c, t.stride(1), bs


def _chk(t: torch.Tensor, dtype=bf16) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("flux_b200 ops need CUDA tensors; there is no CPU fallback")
    if t.dtype != dtype:
        raise ValueError(f"expected {dtype}, got {t.dtype}")
    return t


def _fp8_operands(args, a: torch.Tensor, w: torch.Tensor, a_scale: Optional[torch.Tensor],
                   w_scale: Optional[torch.Tensor], B: int, R: int) -> None:
    """Fill the fp8 fields of GemmArgs / QkvArgs: a_scale fp32 [B, R] (or [R]), w_scale fp32 [N]."""
    if a_scale is None or w_scale is None:
        raise ValueError("fp8 operands need a_scale and w_scale")
    _chk(a_scale, torch.float32), _chk(w_scale, torch.float32)
    if a_scale.numel() != B * R or w_scale.numel() != w.shape[0] or not w_scale.is_contiguous():
        raise ValueError("fp8 scale shapes do not match the operands")
    if a_scale.stride(-1) != 1:
        raise ValueError("a_scale rows must be contiguous")
    args.fp8 = 1
    args.a_scale, args.w_scale = a_scale.data_ptr(), w_scale.data_ptr()
    args.a_scale_bs = a_scale.stride(0) if (a_scale.dim() == 2 and B > 1) else R


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         act=None, gate: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
         out_dtype: torch.dtype = bf16, a_scale: Optional[torch.Tensor] = None,
         w_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = resid + gate * act(a @ w.T + bias).  a [B,R,K] | [R,K]; w [N,K]; gate [B,N]; resid like out.
    a and w may both be float8_e4m3fn with per-row scales (a_scale [B,R], w_scale [N]): the --quantize path."""
    is8 = a.dtype == fp8
    _chk(a, fp8 if is8 else bf16), _chk(w, fp8 if is8 else bf16)
    B, R, K, lda, abs_ = _as3(a)
    Nn = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"weight {tuple(w.shape)} does not match K={K}")
    if out is None:
        out = torch.empty((B, R, Nn) if a.dim() == 3 else (R, Nn), device=a.device, dtype=out_dtype)
    _chk(out, out.dtype)
    _, _, _, ldo, obs = _as3(out)
    args = N.GemmArgs()
    args.A, args.lda, args.a_bs = a.data_ptr(), lda, abs_
    args.W, args.ldw = w.data_ptr(), w.stride(0)
    args.bias = N.ptr(bias)
    args.out, args.ldo, args.out_bs = out.data_ptr(), ldo, obs
    args.out_f32 = 1 if out.dtype == torch.float32 else 0
    args.act = ACT[act]
    if gate is not None:
        _chk(gate)
        args.gate, args.gate_bs = gate.data_ptr(), (gate.stride(0) if gate.dim() == 2 else 0)
    if resid is not None:
        _chk(resid)
        _, _, _, ldr, rbs = _as3(resid)
        args.resid, args.ldr, args.resid_bs = resid.data_ptr(), ldr, rbs
    args.batch, args.rows, args.N, args.K = B, R, Nn, K
    if is8:
        _fp8_operands(args, a, w, a_scale, w_scale, B, R)
    N.check(N.lib().fx_gemm(C.byref(args), N.stream()))
    return out


def gemm_qkv(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], q_scale: torch.Tensor,
             k_scale: torch.Tensor, pe: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
             seq_off: int, mlp_out: Optional[torch.Tensor] = None, rms_eps: float = 1e-5,
             a_scale: Optional[torch.Tensor] = None, w_scale: Optional[torch.Tensor] = None,
             pe_blocked: bool = False) -> None:
    """Fused QKV(+MLP-in) projection; q/k/v are [B, H, seq_total, 128]; pe [seq_total, 64, 2] bf16, or the blocked
    layout of block_pe() with pe_blocked=True.  a / w may be float8_e4m3fn with a_scale / w_scale (see gemm)."""
    is8 = a.dtype == fp8
    out8 = q.dtype == fp8  # e4m3 q / k / v for the FP8 attention (any operand precision of the projection itself)
    _chk(a, fp8 if is8 else bf16), _chk(w, fp8 if is8 else bf16), _chk(pe)
    _chk(q, fp8 if out8 else bf16), _chk(k, fp8 if out8 else bf16), _chk(v, fp8 if out8 else bf16)
    B, R, K, lda, abs_ = _as3(a)
    H, seq_total = q.shape[1], q.shape[2]
    args = N.QkvArgs()
    args.A, args.lda, args.a_bs = a.data_ptr(), lda, abs_
    args.W, args.ldw, args.bias = w.data_ptr(), w.stride(0), N.ptr(bias)
    args.q_scale, args.k_scale, args.pe = q_scale.data_ptr(), k_scale.data_ptr(), pe.data_ptr()
    args.q, args.k, args.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    if mlp_out is not None:
        _chk(mlp_out)
        _, _, _, ldm, mbs = _as3(mlp_out)
        args.mlp_out, args.ld_mlp, args.mlp_bs = mlp_out.data_ptr(), ldm, mbs
    args.rms_eps = rms_eps
    args.batch, args.rows, args.N, args.K = B, R, w.shape[0], K
    args.heads, args.seq_total, args.seq_off = H, seq_total, seq_off
    args.pe_blocked = int(pe_blocked)
    args.qkv_fp8 = int(out8)
    if is8:
        _fp8_operands(args, a, w, a_scale, w_scale, B, R)
    N.check(N.lib().fx_gemm_qkv(C.byref(args), N.stream()))


def block_pe(pe: torch.Tensor) -> torch.Tensor:
    """RoPE table [N, 64, 2] bf16 -> blocked layout [ceil(N/32), 16 pieces, 32 rows, 8 bf16] for gemm_qkv(pe_blocked=True):
    the QKV epilogue owns one row per thread, and in this layout a warp's 32 rows read each 16-byte piece contiguously."""
    n = pe.shape[0]
    nb = (n + 31) // 32
    flat = torch.zeros((nb * 32, 16, 8), device=pe.device, dtype=pe.dtype)
    flat[:n] = pe.reshape(n, 16, 8)
    return flat.view(nb, 32, 16, 8).permute(0, 2, 1, 3).contiguous()


def conv3x3(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], resid: Optional[torch.Tensor] = None,
            out_dtype: torch.dtype = bf16, out: Optional[torch.Tensor] = None, gn_stats: bool = False,
            upsample: bool = False):
    """x [B,H,W,Cin] NHWC contiguous; w [Cout, 9*Cin] (OHWI flattened).
    gn_stats=True: the epilogue also accumulates the GroupNorm(32) partial sums of the output; returns
    (out, (partials, blocks per image)) for groupnorm(..., partials=...).
    upsample=True: out = conv3x3(upsample_nearest(x, 2)) [B,2H,2W,Cout] without the upsampled tensor; w is the
    parity-decomposed kernel [4*Cout, 4*Cin] of upconv_weights()."""
    _chk(x), _chk(w)
    B, H, W, Cin = x.shape
    Cout = w.shape[0] // 4 if upsample else w.shape[0]
    if w.shape[1] != (4 if upsample else 9) * Cin:
        raise ValueError(f"weight {tuple(w.shape)} does not match Cin={Cin} (upsample={upsample})")
    if not x.is_contiguous():
        raise ValueError("conv3x3 input must be contiguous NHWC")
    s = 2 if upsample else 1
    if out is None:
        out = torch.empty((B, s * H, s * W, Cout), device=x.device, dtype=out_dtype)
    args = N.ConvArgs()
    args.x, args.W, args.bias, args.out = x.data_ptr(), w.data_ptr(), N.ptr(bias), out.data_ptr()
    args.out_f32 = 1 if out.dtype == torch.float32 else 0
    args.resid = N.ptr(resid)
    args.batch, args.H, args.Wd, args.Cin, args.Cout = B, H, W, Cin, Cout
    args.upsample2x = int(upsample)
    part = None
    if gn_stats:
        nblk = int(N.lib().fx_conv3x3_gn_blocks(H, W, Cout, int(upsample)))
        part = (torch.empty((B, nblk, 32, 2), device=x.device, dtype=torch.float32), nblk)
        args.gn_partials = part[0].data_ptr()
    N.check(N.lib().fx_conv3x3(C.byref(args), N.stream()))
    return (out, part) if gn_stats else out


def upconv_weights(w: torch.Tensor) -> torch.Tensor:
    """3x3 kernel [Cout, 9*Cin] (OHWI flattened) -> the four 2x2 kernels of conv3x3(upsample_nearest(x, 2)), one per output
    parity (py, px), stacked [4*Cout, 4*Cin] (parity = 2*py + px, tap = 2*a + b).  Output pixel (2y+py, 2x+px) reads the
    source rows {y-1+py, y+py}: for py = 0 tap row a=0 is kernel row 0 and a=1 is kernel rows 1+2, for py = 1 a=0 is rows
    0+1 and a=1 is row 2; columns likewise.  Sums in fp32, one rounding to bf16."""
    Cout, k9 = w.shape
    Cin = k9 // 9
    w9 = w.float().view(Cout, 3, 3, Cin)
    groups = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    out = torch.empty((4, Cout, 2, 2, Cin), device=w.device, dtype=torch.float32)
    for py in (0, 1):
        for px in (0, 1):
            for a in (0, 1):
                for b in (0, 1):
                    out[2 * py + px, :, a, b] = w9[:, list(groups[py][a])][:, :, list(groups[px][b])].sum((1, 2))
    return out.view(4 * Cout, 4 * Cin).to(w.dtype).contiguous()


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: Optional[torch.Tensor], scale: float,
              variant: int = 0, out4=None, out4_col0: int = 0, out4_low=None, split: int = 0) -> Optional[torch.Tensor]:
    """q,k,v [B,H,S,128] contiguous (bf16, or all three float8_e4m3fn: both products then run in FP8);
    out [B,S,>=H*128] bf16 view (head h -> columns h*128..).
    out4 = Fp4Operand.view(B * (S - split), kc): instead of `out`, the epilogue writes O / l as columns [out4_col0 + h * 128, ...)
    of that NVFP4 operand (chunked form: fp4_finalize completes it); with split > 0 the rows [0, split) of every batch element go to
    out4_low = Fp4Operand.view(B * split, kc2) at column 0 (text / image streams of a double block)."""
    is8 = q.dtype == fp8
    _chk(q, fp8 if is8 else bf16), _chk(k, fp8 if is8 else bf16), _chk(v, fp8 if is8 else bf16)
    if not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
        raise ValueError("attention operands must be contiguous [B, H, S, 128]")
    B, H, S, D = q.shape
    if D != 128:
        raise ValueError("attention kernel is specialised for head_dim 128")
    args = N.AttnArgs()
    args.q, args.k, args.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    if out4 is not None:
        q4, sf4, e4, _ = out4
        if q4.shape[0] != B * (S - split) or (split and (out4_low is None or out4_low[0].shape[0] != B * split)):
            raise ValueError("attention(out4=...): operand rows do not match")
        args.q_out, args.sf_out, args.e_out, args.out_kc, args.out_col0 = q4.data_ptr(), sf4.data_ptr(), e4.data_ptr(), 2 * q4.shape[1], out4_col0
        if split:
            q2, sf2, e2, _ = out4_low
            args.q_out2, args.sf_out2, args.e_out2, args.out_kc2, args.out_split = q2.data_ptr(), sf2.data_ptr(), e2.data_ptr(), 2 * q2.shape[1], split
        out = None
    else:
        _chk(out)
        _, _, _, ldo, obs = _as3(out)
        args.out, args.ld_out, args.out_bs = out.data_ptr(), ldo, obs
    args.scale = scale
    args.batch, args.heads, args.seq, args.variant = B, H, S, variant
    args.fp8 = int(is8)
    N.check(N.lib().fx_attention(C.byref(args), N.stream()))
    return out


def attention_small(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: float,
                    bias: Optional[torch.Tensor] = None, causal: bool = False,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q,k,v [B,S,heads*64] views sharing ld / batch stride; bias fp32 [heads,S,S]."""
    _chk(q), _chk(k), _chk(v)
    B, S, _, ld, bs = _as3(q)
    if out is None:
        out = torch.empty((B, S, heads * 64), device=q.device, dtype=bf16)
    _, _, _, ldo, obs = _as3(out)
    args = N.AttnSmallArgs()
    args.q, args.k, args.v, args.ld, args.bs = q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, bs
    args.bias = N.ptr(bias)
    args.out, args.ld_out, args.out_bs, args.scale = out.data_ptr(), ldo, obs, scale
    args.batch, args.heads, args.seq, args.causal = B, heads, S, int(causal)
    N.check(N.lib().fx_attention_small(C.byref(args), N.stream()))
    return out


def rownorm(x: torch.Tensor, mode: int, p0: torch.Tensor, p1: Optional[torch.Tensor], eps: float,
            out: Optional[torch.Tensor] = None, out_scale: Optional[torch.Tensor] = None, out_fp4=None):
    """mode 0: (1+p1[b])*LN(x)+p0[b]; mode 1: LN(x)*p0+p1; mode 2: RMSNorm(x)*p0.
    A float8_e4m3fn `out` (with out_scale fp32 [B,R]) gets the row-quantised result (--quantize).
    out_fp4 = (q, sf, scale) flat uint8 / uint8 / fp32 buffers (as quantize_rows_fp4(out=...)): the NVFP4 operand of the
    result, bit-identical to quantize_rows_fp4(rownorm(x)); returns (q [rows, D/2], sf atoms, scale [rows])."""
    _chk(x)
    B, R, D, ldx, xbs = _as3(x)
    if out_fp4 is not None:
        rows = B * R
        if rows % 128 or D % 64 or D < 1024:
            raise ValueError("rownorm(out_fp4=...) needs a multiple of 128 rows and D % 64 == 0, D >= 1024")
        q = out_fp4[0][:rows * (D // 2)].view(rows, D // 2)
        sf = out_fp4[1][:(rows // 128) * (D // 64) * 512].view(rows // 128, D // 64, 512)
        scale = out_fp4[2][:rows]
        args = N.RowNormArgs()
        args.x, args.ldx, args.x_bs = x.data_ptr(), ldx, xbs
        args.out, args.ldo, args.out_bs = q.data_ptr(), 0, 0
        args.p0, args.p1 = p0.data_ptr(), N.ptr(p1)
        args.p_bs = p0.stride(0) if (mode == 0 and p0.dim() == 2) else 0
        args.eps, args.mode, args.batch, args.rows, args.D = eps, mode, B, R, D
        args.out_fp8, args.scale_out, args.scale_bs, args.sf_out = 2, scale.data_ptr(), R, sf.data_ptr()
        N.check(N.lib().fx_rownorm(C.byref(args), N.stream()))
        return q, sf, scale
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=bf16)
    _chk(out, out.dtype)
    _, _, _, ldo, obs = _as3(out)
    args = N.RowNormArgs()
    args.x, args.ldx, args.x_bs = x.data_ptr(), ldx, xbs
    args.out, args.ldo, args.out_bs = out.data_ptr(), ldo, obs
    args.p0, args.p1 = p0.data_ptr(), N.ptr(p1)
    args.p_bs = p0.stride(0) if (mode == 0 and p0.dim() == 2) else 0
    args.eps, args.mode, args.batch, args.rows, args.D = eps, mode, B, R, D
    if out.dtype == fp8:
        if out_scale is None:
            raise ValueError("fp8 rownorm output needs out_scale")
        _chk(out_scale, torch.float32)
        args.out_fp8, args.scale_out = 1, out_scale.data_ptr()
        args.scale_bs = out_scale.stride(0) if (out_scale.dim() == 2 and B > 1) else R
    N.check(N.lib().fx_rownorm(C.byref(args), N.stream()))
    return out


def quantize_rows(x: torch.Tensor, out: Optional[torch.Tensor] = None,
                  out_scale: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row-wise FP8 quantisation: x bf16 [B,R,K] | [R,K] -> (q float8_e4m3fn like x, scale fp32 [B,R] | [R]);
    scale = absmax / 448, q = e4m3(x / scale); x ~= q * scale."""
    _chk(x)
    B, R, K, ldx, xbs = _as3(x)
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=fp8)
    if out_scale is None:
        out_scale = torch.empty(x.shape[:-1], device=x.device, dtype=torch.float32)
    _chk(out, fp8), _chk(out_scale, torch.float32)
    _, _, _, ldq, qbs = _as3(out)
    args = N.QuantArgs()
    args.x, args.ldx, args.x_bs = x.data_ptr(), ldx, xbs
    args.q, args.ldq, args.q_bs = out.data_ptr(), ldq, qbs
    args.scale = out_scale.data_ptr()
    args.scale_bs = out_scale.stride(0) if (out_scale.dim() == 2 and B > 1) else R
    args.batch, args.rows, args.K = B, R, K
    N.check(N.lib().fx_quantize_rows(C.byref(args), N.stream()))
    return out, out_scale


# ---------------------------------------------------------------------------------------------
# NVFP4 (W4A4): e2m1 + UE4M3 block scales (per 16) + fp32 row scales  (include/flux_b200.h: fx_quantize_rows_fp4 / fx_gemm_fp4)
# ---------------------------------------------------------------------------------------------
def sf_atoms_to_matrix(sf: torch.Tensor, K: int) -> torch.Tensor:
    """Scale-factor atoms (512 bytes = 128 rows x 4 scales, byte (r % 32) * 16 + (r / 32) * 4 + s; [row block][K / 64]) ->
    the plain [row blocks * 128, K / 16] byte matrix."""
    G = K // 64
    rb = sf.numel() // (G * 512)
    return sf.view(rb, G, 32, 4, 4).permute(0, 3, 2, 1, 4).reshape(rb * 128, G * 4)


def sf_matrix_to_atoms(m: torch.Tensor) -> torch.Tensor:
    """Inverse of sf_atoms_to_matrix: [R (multiple of 128), nsf (multiple of 4)] bytes -> atoms [R / 128][nsf / 4][512]."""
    R, nsf = m.shape
    return m.view(R // 128, 4, 32, nsf // 4, 4).permute(0, 3, 2, 1, 4).contiguous().view(R // 128, nsf // 4, 512)


def quantize_rows_fp4(x: torch.Tensor, out=None):
    """x bf16 [B,R,K] | [R,K] -> (q uint8 [rows, K/2], sf uint8 atoms [ceil(rows/128), K/64, 512], scale fp32 [rows]);
    rows = B*R flattened.  x ~= e2m1 * ue4m3 * scale (two-level NVFP4 quantisation, see the header).
    out: (q, sf, scale) flat uint8 / uint8 / fp32 buffers to carve the results from (no allocation: hot path);
    the scale-atom buffer then must not need padding rows (rows % 128 == 0)."""
    _chk(x)
    B, R, K, ldx, xbs = _as3(x)
    rows = B * R
    if out is not None:
        if rows % 128:
            raise ValueError("quantize_rows_fp4(out=...) needs a multiple of 128 rows")
        q = out[0][:rows * (K // 2)].view(rows, K // 2)
        sf = out[1][:(rows // 128) * (K // 64) * 512].view(rows // 128, K // 64, 512)
        scale = out[2][:rows]
    else:
        q = torch.empty((rows, K // 2), device=x.device, dtype=torch.uint8)
        sf = torch.zeros(((rows + 127) // 128, K // 64, 512), device=x.device, dtype=torch.uint8)
        scale = torch.empty((rows,), device=x.device, dtype=torch.float32)
    args = N.Quant4Args()
    args.x, args.ldx, args.x_bs = x.data_ptr(), ldx, xbs
    args.q, args.sf, args.scale = q.data_ptr(), sf.data_ptr(), scale.data_ptr()
    args.batch, args.rows, args.K = B, R, K
    N.check(N.lib().fx_quantize_rows_fp4(C.byref(args), N.stream()))
    return q, sf, scale


FP4_TILE_N = 192      # column tile of the NVFP4 GEMM with the generic epilogue (csrc/gemm4.cu)
FP4_TILE_N_QKV = 128  # ... with the QKV epilogue: one head per tile


def fp4_weight(w: torch.Tensor, tile_n: int = FP4_TILE_N):
    """Linear weight bf16 [N, K] -> (w4 uint8 [N, K/2], scale atoms regrouped per `tile_n`-row column tile
    [ceil(N/tile_n), K/64, ceil(tile_n/128), 512], w_scale fp32 [N]) for gemm_fp4 (tile_n = 192) / gemm_fp4_qkv (128);
    done once, at Flux.quantize."""
    Nn, K = w.shape
    q, sf, scale = quantize_rows_fp4(w)
    m = sf_atoms_to_matrix(sf, K)[:Nn]                                   # [N, K/16]
    tiles = (Nn + tile_n - 1) // tile_n
    na = (tile_n + 127) // 128                                           # atoms per tile and K-group
    pad = torch.zeros((tiles, na * 128, K // 16), device=w.device, dtype=torch.uint8)
    full = torch.zeros((tiles * tile_n, K // 16), device=w.device, dtype=torch.uint8)
    full[:Nn] = m
    pad[:, :tile_n] = full.view(tiles, tile_n, K // 16)
    atoms = sf_matrix_to_atoms(pad.view(tiles * na * 128, K // 16))     # [tiles * na, K/64, 512]
    atoms = atoms.view(tiles, na, K // 64, 512).permute(0, 2, 1, 3).contiguous()
    return q, atoms, scale


def gemm_fp4_qkv(a4: torch.Tensor, sfa: torch.Tensor, a_scale: torch.Tensor, w4: torch.Tensor, sfw: torch.Tensor,
                 w_scale: torch.Tensor, batch: int, bias: Optional[torch.Tensor], q_scale: torch.Tensor, k_scale: torch.Tensor,
                 pe: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, seq_off: int, rms_eps: float = 1e-6,
                 pe_blocked: bool = False) -> None:
    """gemm_qkv on NVFP4 operands: qkv = (a4 @ w4.T) * a_scale * w_scale + bias; q, k = rope(rmsnorm(.)) per head;
    q/k/v [B, H, seq_total, 128] (bf16 or e4m3) filled at rows seq_off + r.  w4 / sfw from fp4_weight(w, FP4_TILE_N_QKV)."""
    _chk(a4, torch.uint8), _chk(w4, torch.uint8), _chk(sfa, torch.uint8), _chk(sfw, torch.uint8)
    _chk(a_scale, torch.float32), _chk(w_scale, torch.float32)
    rows_total, kb = a4.shape
    Nn = w4.shape[0]
    B, H, St, hd = q.shape
    if w4.shape[1] != kb or rows_total % batch or batch != B or hd != 128 or Nn != 3 * H * 128:
        raise ValueError("fp4 qkv operands do not match")
    f8 = q.dtype == fp8
    for t in (q, k, v):
        _chk(t, fp8 if f8 else bf16)
        if not t.is_contiguous() or t.shape != q.shape:
            raise ValueError("q/k/v must be contiguous [B,H,S,128]")
    args = N.Gemm4QkvArgs()
    args.A, args.sfa, args.a_scale = a4.data_ptr(), sfa.data_ptr(), a_scale.data_ptr()
    args.W, args.sfw, args.w_scale = w4.data_ptr(), sfw.data_ptr(), w_scale.data_ptr()
    args.bias = N.ptr(bias)
    args.q_scale, args.k_scale, args.pe, args.pe_blocked = q_scale.data_ptr(), k_scale.data_ptr(), pe.data_ptr(), int(pe_blocked)
    args.q, args.k, args.v, args.qkv_fp8 = q.data_ptr(), k.data_ptr(), v.data_ptr(), int(f8)
    args.rms_eps = rms_eps
    args.batch, args.rows, args.K, args.heads, args.seq_total, args.seq_off = B, rows_total // batch, 2 * kb, H, St, seq_off
    N.check(N.lib().fx_gemm_fp4_qkv(C.byref(args), N.stream()))


class Fp4Operand:
    """Buffers of an NVFP4 operand that its PRODUCERS fill chunk by chunk (gemm_fp4(out4=...), quantize_chunks_fp4) and
    fp4_finalize completes: e2m1 rows, UE4M3 scale atoms, per-chunk exponents, fp32 row scales.  `view(rows, kc)` carves the
    tensors of one use from the flat buffers (no allocation on the hot path)."""

    def __init__(self, max_rows: int, max_kc: int, device):
        self.q = torch.empty((max_rows * max_kc // 2,), device=device, dtype=torch.uint8)
        self.sf = torch.empty((max_rows * max_kc // 16,), device=device, dtype=torch.uint8)
        self.e = torch.empty((max_rows * max_kc // 32,), device=device, dtype=torch.int8)
        self.scale = torch.empty((max_rows,), device=device, dtype=torch.float32)

    def view(self, rows: int, kc: int):
        if rows % 128 or kc % 128:
            raise ValueError("an NVFP4 operand needs multiples of 128 rows and columns")
        return (self.q[:rows * kc // 2].view(rows, kc // 2), self.sf[:rows * kc // 16].view(rows // 128, kc // 64, 512),
                self.e[:rows * kc // 32].view(rows, kc // 32), self.scale[:rows])


def quantize_chunks_fp4(x: torch.Tensor, dst, col0: int = 0) -> None:
    """x bf16 [B,R,C] -> columns [col0, col0 + C) of the operand `dst` = Fp4Operand.view(B * R, kc) (chunked form: call
    fp4_finalize once every column of the operand has been produced)."""
    _chk(x)
    B, R, Cc, ldx, xbs = _as3(x)
    q, sf, e, _ = dst
    if q.shape[0] != B * R:
        raise ValueError("operand rows do not match")
    args = N.Quant4cArgs()
    args.x, args.ldx, args.x_bs, args.batch, args.rows, args.C = x.data_ptr(), ldx, xbs, B, R, Cc
    args.q, args.sf, args.e, args.kc, args.col0 = q.data_ptr(), sf.data_ptr(), e.data_ptr(), 2 * q.shape[1], col0
    N.check(N.lib().fx_quantize_chunks_fp4(C.byref(args), N.stream()))


def fp4_finalize(dst):
    """Completes a chunk-produced operand (row scale = 2^max chunk exponent, block scales shifted to it); returns
    (a4, sfa, a_scale) for gemm_fp4."""
    q, sf, e, scale = dst
    N.check(N.lib().fx_fp4_finalize(sf.data_ptr(), e.data_ptr(), scale.data_ptr(), q.shape[0], 2 * q.shape[1], N.stream()))
    return q, sf, scale


def gemm_fp4(a4: torch.Tensor, sfa: torch.Tensor, a_scale: torch.Tensor, w4: torch.Tensor, sfw: torch.Tensor,
             w_scale: torch.Tensor, batch: int, bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
             act=None, gate: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
             out_dtype: torch.dtype = bf16, out4=None, out4_col0: int = 0) -> Optional[torch.Tensor]:
    """out = resid + gate * act((a4 @ w4.T) * a_scale[:, None] * w_scale + bias) on tcgen05 kind::mxf4nvf4.block_scale.
    a4 / sfa / a_scale from quantize_rows_fp4 (rows = batch * rows_per_batch), w4 / sfw / w_scale from fp4_weight.
    out4 = Fp4Operand.view(rows, kc): instead of `out`, act(. + bias) is written as columns [out4_col0, out4_col0 + N) of that
    NVFP4 operand (the epilogue quantises; the result never exists in bf16); returns None."""
    _chk(a4, torch.uint8), _chk(w4, torch.uint8), _chk(sfa, torch.uint8), _chk(sfw, torch.uint8)
    _chk(a_scale, torch.float32), _chk(w_scale, torch.float32)
    rows_total, kb = a4.shape
    K, Nn = 2 * kb, w4.shape[0]
    if w4.shape[1] != kb or rows_total % batch:
        raise ValueError("fp4 operands do not match")
    R = rows_total // batch
    args = N.Gemm4Args()
    args.A, args.sfa, args.a_scale = a4.data_ptr(), sfa.data_ptr(), a_scale.data_ptr()
    args.W, args.sfw, args.w_scale = w4.data_ptr(), sfw.data_ptr(), w_scale.data_ptr()
    args.bias = N.ptr(bias)
    if out4 is not None:
        q4, sf4, e4, _ = out4
        if q4.shape[0] != rows_total or gate is not None or resid is not None:
            raise ValueError("gemm_fp4(out4=...): operand rows must match and no gate / residual")
        args.q_out, args.sf_out, args.e_out = q4.data_ptr(), sf4.data_ptr(), e4.data_ptr()
        args.out_kc, args.out_col0 = 2 * q4.shape[1], out4_col0
        out = None
    else:
        if out is None:
            out = torch.empty((batch, R, Nn), device=a4.device, dtype=out_dtype)
        _chk(out, out.dtype)
        _, _, _, ldo, obs = _as3(out)
        args.out, args.ldo, args.out_bs = out.data_ptr(), ldo, obs
        args.out_f32 = 1 if out.dtype == torch.float32 else 0
    args.act = ACT[act]
    if gate is not None:
        _chk(gate)
        args.gate, args.gate_bs = gate.data_ptr(), (gate.stride(0) if gate.dim() == 2 else 0)
    if resid is not None:
        _chk(resid)
        _, _, _, ldr, rbs = _as3(resid)
        args.resid, args.ldr, args.resid_bs = resid.data_ptr(), ldr, rbs
    args.batch, args.rows, args.N, args.K = batch, R, Nn, K
    N.check(N.lib().fx_gemm_fp4(C.byref(args), N.stream()))
    return out


def gemv(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None,
         silu_in: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b] = W @ f(x[b]) + bias (+ add[b]);  x [B,K] (row stride arbitrary), w [N,K]."""
    _chk(x), _chk(w)
    B, K = x.shape
    Nn = w.shape[0]
    if out is None:
        out = torch.empty((B, Nn), device=x.device, dtype=bf16)
    args = N.GemvArgs()
    args.in_, args.ld_in = x.data_ptr(), x.stride(0)
    args.W, args.ldw, args.bias = w.data_ptr(), w.stride(0), N.ptr(bias)
    if add is not None:
        args.add, args.ld_add = add.data_ptr(), add.stride(0)
    args.out, args.ld_out = out.data_ptr(), out.stride(0)
    args.batch, args.N, args.K, args.silu_in, args.silu_out = B, Nn, K, int(silu_in), 0
    N.check(N.lib().fx_gemv(C.byref(args), N.stream()))
    return out


def timestep_embedding(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    _chk(t)
    out = torch.empty((t.shape[0], dim), device=t.device, dtype=bf16)
    N.check(N.lib().fx_timestep_embedding(t.data_ptr(), out.data_ptr(), t.shape[0], dim, N.stream()))
    return out


def euler_step(x: torch.Tensor, pred: torch.Tensor, dt: float) -> torch.Tensor:
    _chk(x), _chk(pred)
    if not (x.is_contiguous() and pred.is_contiguous()):
        raise ValueError("euler_step needs contiguous tensors")
    N.check(N.lib().fx_euler_step(x.data_ptr(), pred.data_ptr(), float(dt), x.numel(), N.stream()))
    return x


def patchify(x: torch.Tensor) -> torch.Tensor:
    _chk(x)
    b, h, w, c = x.shape
    out = torch.empty((b, h * w // 4, 4 * c), device=x.device, dtype=bf16)
    N.check(N.lib().fx_patchify(x.contiguous().data_ptr(), out.data_ptr(), b, h, w, c, N.stream()))
    return out


def prior_packed(n_images: int, latent_size, channels: int = 16, seed: int = 0, first_index: int = 0,
                 device="cuda") -> torch.Tensor:
    """Standard-normal prior written directly in the packed [B, L, 4c] layout (Philox keyed by global image index)."""
    h, w = latent_size
    out = torch.empty((n_images, h * w // 4, 4 * channels), device=device, dtype=bf16)
    N.check(N.lib().fx_prior_packed(out.data_ptr(), n_images, h, w, channels, int(seed) & 0xFFFFFFFFFFFFFFFF, first_index,
                                    N.stream()))
    return out


def unpatchify_scale(packed: torch.Tensor, latent_size, c_pad: int, scale_factor: float,
                     shift_factor: float) -> torch.Tensor:
    _chk(packed)
    h, w = latent_size
    b, L, f = packed.shape
    if L != h * w // 4:
        raise ValueError(f"packed latents have {L} tokens, latent_size {latent_size} needs {h * w // 4}")
    c = f // 4
    z = torch.empty((b, h, w, c_pad), device=packed.device, dtype=bf16)
    N.check(N.lib().fx_unpatchify_scale(packed.contiguous().data_ptr(), z.data_ptr(), b, h, w, c, c_pad,
                                        scale_factor, shift_factor, N.stream()))
    return z


def groupnorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float, silu: bool,
              out: Optional[torch.Tensor] = None, partials=None) -> torch.Tensor:
    """GroupNorm(32) on NHWC-like [B, ..., C] contiguous (+ optional SiLU).
    partials: (tensor, blocks per image) written by the conv3x3 that produced x (gn_stats=True): the statistics pass
    over x is skipped."""
    _chk(x)
    B, Cc = x.shape[0], x.shape[-1]
    hw = x.numel() // (B * Cc)
    l = N.lib()
    if out is None:
        out = torch.empty_like(x)
    stats = torch.empty((B, 32, 2), device=x.device, dtype=torch.float32)
    if partials is not None:
        N.check(l.fx_groupnorm_finalize_blocks(partials[0].data_ptr(), stats.data_ptr(), B, partials[1], hw, Cc, eps, N.stream()))
    else:
        sums = torch.empty((int(l.fx_groupnorm_partials_count(B, hw)),), device=x.device, dtype=torch.float32)
        N.check(l.fx_groupnorm_stats(x.data_ptr(), sums.data_ptr(), B, hw, Cc, N.stream()))
        N.check(l.fx_groupnorm_finalize(sums.data_ptr(), stats.data_ptr(), B, hw, Cc, eps, N.stream()))
    N.check(l.fx_groupnorm_apply(x.data_ptr(), stats.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                 B, hw, Cc, int(silu), N.stream()))
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    _chk(x)
    B, H, W, Cc = x.shape
    out = torch.empty((B, 2 * H, 2 * W, Cc), device=x.device, dtype=bf16)
    N.check(N.lib().fx_upsample2x(x.data_ptr(), out.data_ptr(), B, H, W, Cc, N.stream()))
    return out


def softmax_rows(s: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk(s, torch.float32)
    rows, cols = s.shape
    if out is None:
        out = torch.empty((rows, cols), device=s.device, dtype=bf16)
    N.check(N.lib().fx_softmax_rows(s.data_ptr(), s.stride(0), out.data_ptr(), out.stride(0), rows, cols, scale,
                                    N.stream()))
    return out


def transpose(x: torch.Tensor) -> torch.Tensor:
    _chk(x)
    rows, cols = x.shape
    out = torch.empty((cols, rows), device=x.device, dtype=bf16)
    N.check(N.lib().fx_transpose(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, cols, N.stream()))
    return out


def finish_image(x: torch.Tensor, want_u8: bool = True):
    _chk(x, torch.float32)
    img = torch.empty_like(x)
    u8 = torch.empty(x.shape, device=x.device, dtype=torch.uint8) if want_u8 else None
    N.check(N.lib().fx_finish_image(x.data_ptr(), img.data_ptr(), N.ptr(u8), x.numel(), N.stream()))
    return img, u8


def embedding(ids: torch.Tensor, table: torch.Tensor, pos_table: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk(table)
    if ids.dtype != torch.int32:
        raise ValueError("token ids must be int32")
    B, S = ids.shape
    D = table.shape[1]
    out = torch.empty((B, S, D), device=table.device, dtype=bf16)
    N.check(N.lib().fx_embedding(ids.contiguous().data_ptr(), table.data_ptr(), N.ptr(pos_table), out.data_ptr(),
                                 B * S, S, D, N.stream()))
    return out


def act_mul(a: torch.Tensor, b: torch.Tensor, act: str) -> torch.Tensor:
    _chk(a), _chk(b)
    out = torch.empty_like(a)
    N.check(N.lib().fx_act_mul(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), ACT[act], N.stream()))
    return out


def dbg_gemm_ref(a: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    M, K = a.shape
    Nn = w.shape[0]
    out = torch.empty((M, Nn), device=a.device, dtype=torch.float32)
    N.check(N.dbg_lib().fx_dbg_gemm_ref(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), Nn, M, Nn,
                                    K, N.stream()))
    return out


def dbg_umma_tile(a: torch.Tensor, b: torch.Tensor, n: int, b_mn_major: bool, a_tmem: bool, lbo: int, sbo: int,
                  kstep: int) -> torch.Tensor:
    K = a.shape[1]
    d = torch.empty((128, n), device=a.device, dtype=torch.float32)
    N.check(N.dbg_lib().fx_dbg_umma_tile(a.data_ptr(), b.data_ptr(), d.data_ptr(), K, n, int(b_mn_major), int(a_tmem),
                                     lbo, sbo, kstep, N.stream()))
    return d
