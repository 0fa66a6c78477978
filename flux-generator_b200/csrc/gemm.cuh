// Persistent warp-specialised bf16 (or FP8 e4m3, template F8) GEMM for sm_100a:  out = epilogue(A . W^T)
//   A  [batch][rows][K]   bf16, K contiguous (activations; or NHWC image read through a 4-D TMA box
//                         for the 3x3 convolution mode -- implicit GEMM, no im2col buffer)
//   W  [N][K]             bf16, K contiguous (nn.Linear layout / OHWI conv weights flattened)
// One CTA per SM; warps 0..7 = epilogue, warp 8 = TMA producer, warp 9 = tcgen05.mma issuer (+TMEM owner);
// epilogue (TMEM -> registers -> global; two warps per TMEM lane quarter, each owning half the columns).  smem ring of 128B-swizzled K-major tiles filled by TMA;
// accumulators double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include <cuda_fp8.h>

#include "sm100.cuh"

namespace fx {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle span (FP8: 128 e4m3 elements in the same 128 B)
constexpr int GEMM_THREADS = 384;  // warps 0..7 epilogue (two per TMEM lane quarter), warp 8 TMA, warp 9 MMA, (10, 11 idle)
// The warp scheduler favours higher warp ids: the single-thread TMA / MMA issuers sit ABOVE the epilogue
// warps so that a compute-heavy epilogue (GELU, RoPE) can never starve the tensor pipe of instructions.
#ifdef FX_ROLES_LOW  // A/B builds only: issuers below the epilogue warps
constexpr int GEMM_WARP_TMA = 0, GEMM_WARP_MMA = 1, GEMM_CTRL0 = 0, GEMM_EPI0 = 4;
#else
constexpr int GEMM_WARP_TMA = 8, GEMM_WARP_MMA = 9, GEMM_CTRL0 = 8, GEMM_EPI0 = 0;
#endif

enum : int { EPI_GENERIC = 0, EPI_QKV = 1 };

// ------------------------------------------------------------------ wave lockstep (L2 reuse across CTA pairs)
// The CTA pairs that run concurrently share operand tiles (a wave of 74 pairs = ~6 row tiles x 12 column tiles for the
// narrow-output GEMMs), but nothing keeps them at the same K position: with K = 12288 / 15360 a wave's operand set is
// ~140 MB > L2, pairs drift apart by whole tiles, and ncu showed linear2 reading 11 GB from DRAM for 1.2 GB of
// operands.  Under the 1 kW cap that traffic is time (measured ~0.09 ms per GB).  A sliding-window barrier keeps the
// producers of a wave within `ls_slack` K-groups of each other, so a tile fetched by one pair is still in L2 when the
// others ask for it.  Split-phase, sense-reversing, self-cleaning slots in global memory; a bounded wait (the barrier
// is an optimisation: on timeout the pair just stops synchronising) makes a hang impossible.
constexpr int LS_RING = 256;
struct LockstepSlots {
  unsigned int count[LS_RING];
  unsigned int gen[LS_RING];
};
static __device__ LockstepSlots g_lockstep;  // (static: the header is included by several translation units)

struct GemmParams {
  int batch, rows, N, K;
  int tiles_m_per_batch, tiles_m, tiles_n, num_tiles, k_blocks, group_m;
  int group_n;  // raster super-columns: n-tiles per column band (0 = one band of all n-tiles)
  // ---- generic epilogue: v = acc + bias; act; v *= gate[b][n]; v += resid[b][r][n]; store
  const __nv_bfloat16* bias;
  void* out;
  long long ldo, out_bs;
  int out_f32, act;
  const __nv_bfloat16* gate;
  long long gate_bs;
  const __nv_bfloat16* resid;
  long long ldr, resid_bs;
  // ---- qkv epilogue (columns [q | k | v | mlp]); per 128-column head: QK-RMSNorm, RoPE, head scatter
  int heads, seq_total, seq_off;
  float rms_eps;
  const __nv_bfloat16 *qnorm_w, *knorm_w;
  const uint32_t* pe;  // [seq_total][64] (cos, sin) bf16 pairs; or, pe_blocked, [seq/32][16 pieces][32 rows][16 B]
  int pe_blocked;      // blocked layout: the 32 rows of a warp read each 16-byte piece as ONE 512-byte coalesced request
  __nv_bfloat16 *q, *k, *v;  // [batch][heads][seq_total][128]
  int qkv_f8;                // q, k, v are e4m3 bytes in the same [batch][heads][seq_total][128] order (FP8 attention)
  // ---- FP8 mode (F8): A, W are e4m3 with per-row / per-output-channel dequantisation scales,
  //      acc * a_scale[b][row] * w_scale[n] enters the epilogue in place of the raw accumulator
  const float* a_scale;
  long long a_scale_bs;
  const float* w_scale;
  // ---- L2 management: eviction-priority policies of the A / W tile loads and streaming (evict-first) output stores
  unsigned long long hint_a, hint_w;
  int stream_out;
  int ls_group, ls_slack;  // wave lockstep: k-blocks per group (0 = off), groups a producer may run ahead
  // ---- 3x3 conv mode (A is [batch][H][W][C] NHWC, pad 1, stride 1)
  int conv_H, conv_W, conv_tiles_x, conv_tiles_y, cin_blocks;
  // ---- conv mode: GroupNorm(32) statistics of the bf16 output, per 32-pixel block (one epilogue warp's rows):
  //      gn_partials[batch][gn_nblk][32 groups][sum, sum of squares]; gn_gs = channels per group (4 / 8 / 16)
  float* gn_partials;
  int gn_gs, gn_nblk;
  // ---- conv mode, conv_up = 1: nearest-2x upsample + 3x3 convolution (flux/autoencoder.py:121-124) as FOUR 2x2
  //      convolutions of the LOW-resolution input, one per output-pixel parity (py, px).  On the upsampled image the
  //      3x3 taps of an output pixel (2y + py, 2x + px) fall on only 2 x 2 source pixels -- rows {y - 1 + py, y + py},
  //      columns likewise -- so taps that share a source pixel are pre-summed into W4[parity][Cout][4 * Cin] (host:
  //      ops.upconv_weights).  16 instead of 36 MACs per output element and channel pair, and the 4x larger upsampled
  //      tensor is never written or read.  The parity rides in the tile's batch index (b' = 4 * image + parity);
  //      conv_line = elements between output image lines of one parity (ldo = elements between its pixels).
  int conv_up;
  long long conv_line;
  // ---- timing experiment only (FX_GEMM_DBG_SKIP_W=1, wrong results): the W tile load of every other k-block is skipped,
  //      i.e. 25 % less L2 -> SM traffic at the same MMA work: what sharing operand tiles across CTAs could buy at most
  int dbg_skip_w;
  // ---- CL = 2 (two CTA pairs per cluster): tiles_m counts SUPER row tiles (two row tiles that share their W tile);
  //      tiles_m_real = the row tiles that exist (the second tile of the last super tile may not)
  int tiles_m_real;
};

// (sum, sum of squares) of one 32-column chunk of a warp's 32 output rows for the GS-channel GroupNorm groups it
// covers: per-thread sums over the row's channels, then a halving butterfly across the 32 lanes (2G values are reduced
// with 2G-1+log2(32/2G) shuffles instead of 10 G) -- a fixed order, so the statistics are bit-reproducible.  The values
// are the bf16-ROUNDED outputs: exactly what a separate statistics pass over the stored tensor would read.
template <int GS>
__device__ __forceinline__ void gn_chunk_partials_t(float* dst, const float* f, bool valid, int lane) {
  constexpr int G = 32 / GS, NV = 2 * G;
  float vals[NV];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < GS; ++j) {
      const float v = valid ? __bfloat162float(__float2bfloat16(f[g * GS + j])) : 0.f;
      s += v;
      q += v * v;
    }
    vals[2 * g] = s;
    vals[2 * g + 1] = q;
  }
  int n = NV, idx = 0;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (n > 1) {
      const int h = n >> 1;
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        if (i < h) {
          const float mine = up ? vals[h + i] : vals[i];
          const float other = up ? vals[i] : vals[h + i];
          vals[i] = mine + __shfl_xor_sync(0xffffffffu, other, off);
        }
      }
      n = h;
      if (up) idx += h;
    } else {
      vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], off);
    }
  }
  if ((lane & (32 / NV - 1)) == 0) dst[idx] = vals[0];  // this lane ended up with value `idx` = 2 * group + statistic
}
__device__ __forceinline__ void gn_chunk_partials(float* dst, const float* f, bool valid, int lane, int gs) {
  if (gs == 4) gn_chunk_partials_t<4>(dst, f, valid, lane);
  else if (gs == 8) gn_chunk_partials_t<8>(dst, f, valid, lane);
  else gn_chunk_partials_t<16>(dst, f, valid, lane);
}

template <int BN, int NCTA = 1>
struct GemmCfg {
  // NCTA = 2: a CTA pair computes a 256 x BN tile with cta_group::2 MMAs; each CTA stages its own 128 A
  // rows and HALF of the B tile, so the same 192 KB of shared memory buffers 50% more MMA work.
  static constexpr int STAGES = NCTA == 2 ? 6 : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / NCTA) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES + 256;  // after the barriers
  static constexpr int STORE_OFF = EPI_OFF + 8 * 384 * 4;    // per-warp 32 x 64 B transposition buffers for coalesced stores
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 8 * 384 * 4 /*epilogue staging*/ +
                                    8 * 2048 /*store transposition*/ + 1024 /*align*/;
};

// Tile raster.  The n-tiles are cut into bands of `group_n` columns (one band = all columns when group_n == 0); a band
// is walked in groups of `group_m` row tiles x the band's columns, rows fastest.  A wave of ~74 CTA pairs then works on
// group_m A row-blocks and one band of W: for the long-K, narrow-N members (linear2 / fc2: W = 94 / 75 MB, more than
// the L2 can keep next to the streaming A) a half-width band makes the resident part (the band's W rows, 47 MB) fit,
// at the price of streaming A once per band.
__device__ __forceinline__ void gemm_tile_coords(const GemmParams& p, int tile, int& tm, int& tn) {
  const int gn = p.group_n > 0 ? p.group_n : p.tiles_n;
  const int per_band = p.tiles_m * gn;  // every band before the last is full
  const int band = tile / per_band;
  const int n0 = band * gn;
  const int bw = min(gn, p.tiles_n - n0);
  const int t = tile - band * per_band;
  const int per_group = p.group_m * bw;
  const int g = t / per_group;
  const int first_m = g * p.group_m;
  const int gsize = min(p.tiles_m - first_m, p.group_m);
  const int rem = t - g * per_group;
  tm = first_m + rem % gsize;
  tn = n0 + rem / gsize;
}

// 8 bf16 (one 16-byte vector) -> 8 floats
__device__ __forceinline__ void ld_bf16x8(const __nv_bfloat16* p, float* f) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void st_bf16x8(__nv_bfloat16* p, const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
  u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// One 32-column chunk of the generic epilogue for one accumulator row.  Tile-uniform vectors (bias, gate)
// come from this warp's shared-memory staging area `sb` / `sg` (floats, already bounds-checked: bias 0 /
// gate 1 past N); the row's residual values may have been prefetched into `rr` before the accumulator
// was ready (rr_ok), so that no long-latency load sits between the TMEM read and the store.
template <bool F8 = false>
__device__ __forceinline__ void epi_generic_chunk(const GemmParams& p, float* f, const float* sb, const float* sg,
                                                  const uint4* rr, bool rr_ok, long long out_off, long long res_off,
                                                  int n0, bool fast, const float* sw = nullptr, float rs = 1.f,
                                                  uint8_t* wst = nullptr, int lane = 0, uint32_t vmask = 0, bool valid = true,
                                                  long long lane_off = -1, long long ld_hi = -1, const void* tmap_out = nullptr,
                                                  int t_row0 = 0, int t_b = 0, bool math_only = false) {
  if (F8) {  // dequantise: acc * a_scale[row] * w_scale[n], then + bias -- on packed fp32 pairs (FMUL2 / FFMA2: same roundings)
    const uint64_t rs2 = pack2f(rs, rs);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(sb + i);
      const float4 w = *reinterpret_cast<const float4*>(sw + i);
      const float2 lo = unpack2f(ffma2(pack2f(f[i], f[i + 1]), fmul2(rs2, pack2f(w.x, w.y)), pack2f(t.x, t.y)));
      const float2 hi = unpack2f(ffma2(pack2f(f[i + 2], f[i + 3]), fmul2(rs2, pack2f(w.z, w.w)), pack2f(t.z, t.w)));
      f[i] = lo.x; f[i + 1] = lo.y; f[i + 2] = hi.x; f[i + 3] = hi.y;
    }
  } else {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(sb + i);
    f[i] += t.x; f[i + 1] += t.y; f[i + 2] += t.z; f[i + 3] += t.w;
  }
  }
  if (p.act == 1) {
#ifdef FX_GELU_SCALAR
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = gelu_tanh(f[i]);
#else
    // GELU(tanh) on packed fp32 pairs (FMUL2 / FFMA2): half the issue slots of the scalar form
    const uint64_t c0 = pack2f(0.7978845608028654f, 0.7978845608028654f);
    const uint64_t c1 = pack2f(0.0356774081363001f, 0.0356774081363001f);
    const uint64_t hf = pack2f(0.5f, 0.5f);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const uint64_t x = pack2f(f[i], f[i + 1]);
      const uint64_t x2 = fmul2(x, x);
      const uint64_t u = fmul2(x, ffma2(c1, x2, c0));
      const float2 uu = unpack2f(u);
      const uint64_t t = pack2f(tanh_approx(uu.x), tanh_approx(uu.y));
      const uint64_t h = fmul2(x, hf);
      const float2 r = unpack2f(ffma2(h, t, h));
      f[i] = r.x;
      f[i + 1] = r.y;
    }
#endif
  } else if (p.act != 0) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = apply_act(f[i], p.act);
  }
  if (p.gate) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(sg + i);
      const float2 lo = unpack2f(fmul2(pack2f(f[i], f[i + 1]), pack2f(t.x, t.y)));
      const float2 hi = unpack2f(fmul2(pack2f(f[i + 2], f[i + 3]), pack2f(t.z, t.w)));
      f[i] = lo.x; f[i + 1] = lo.y; f[i + 2] = hi.x; f[i + 3] = hi.y;
    }
  }
  if (math_only) return;  // f[] = act(acc * scales + bias) (* gate): the caller consumes it (NVFP4 operand producer)
  if (fast) {
    if (p.resid && valid) {
      const __nv_bfloat16* r = p.resid + res_off + n0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = rr_ok ? rr[i] : *reinterpret_cast<const uint4*>(r + i * 8);  // plain load: may alias `out`
        float2 a = unpack_bf16(u.x), bb = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        f[i * 8] += a.x; f[i * 8 + 1] += a.y; f[i * 8 + 2] += bb.x; f[i * 8 + 3] += bb.y;
        f[i * 8 + 4] += c.x; f[i * 8 + 5] += c.y; f[i * 8 + 6] += d.x; f[i * 8 + 7] += d.y;
      }
    }
    if (p.out_f32) {
      if (valid) {
        float* o = reinterpret_cast<float*>(p.out) + out_off + n0;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
      }
    } else if (tmap_out != nullptr) {  // warp-collective: staged, then one TMA store (rows / columns clipped by the tensor map)
      store_chunk32_tma(wst, lane, f, tmap_out, n0, t_row0, t_b);
    } else if (wst != nullptr) {  // warp-collective: every lane takes part, rows are masked
      // lane_off: distance (elements) from the warp's row 0 to this lane's row (lane * ldo unless the rows are image patches)
      store_chunk32_coalesced(wst, lane, f,
                              reinterpret_cast<__nv_bfloat16*>(p.out) + out_off - (lane_off < 0 ? (long long)lane * p.ldo : lane_off) + n0,
                              p.ldo, vmask, p.stream_out != 0 && p.resid == nullptr, ld_hi);
    } else if (valid) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_off + n0;
#pragma unroll
      for (int i = 0; i < 32; i += 8) st_bf16x8(o + i, f + i);
    }
  } else if (valid) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {  // static indices keep f[] in registers
      const int n = n0 + i;
      if (n >= p.N) continue;
      float v = f[i];
      if (p.resid) v += __bfloat162float(p.resid[res_off + n]);
      if (p.out_f32) reinterpret_cast<float*>(p.out)[out_off + n] = v;
      else reinterpret_cast<__nv_bfloat16*>(p.out)[out_off + n] = __float2bfloat16(v);
    }
  }
}

// Stage `n` tile-uniform bf16 values src[n0 .. n0+n) as floats into this warp's shared-memory slot
// (fill value past `limit` or when src is null).  One coalesced load per warp, no cross-warp barrier.
__device__ __forceinline__ void stage_vec(float* dst, const __nv_bfloat16* src, int n0, int n, int limit, float fill,
                                          int lane) {
  for (int i = lane; i < n; i += 32) dst[i] = (src != nullptr && n0 + i < limit) ? __bfloat162float(__ldg(src + n0 + i)) : fill;
}
__device__ __forceinline__ void stage_vec_f32(float* dst, const float* src, int n0, int n, int limit, float fill, int lane) {
  for (int i = lane; i < n; i += 32) dst[i] = (src != nullptr && n0 + i < limit) ? __ldg(src + n0 + i) : fill;
}

// CL = 2: a cluster of TWO CTA pairs works on two row tiles of the same column tile.  Their W tile is identical, so
// each of the four CTAs fetches a quarter of it and TMA-multicasts the quarter to the CTA that holds the same half in the
// other pair: 25 % fewer bytes leave the L2 per FLOP.  (Measured with FX_GEMM_DBG_SKIP_W: 25 % less L2 -> SM traffic is
// worth 8-17 % on these GEMMs -- the chip is power-limited and moving operands costs as much as multiplying them;
// profiles/r02_gemm_l2_traffic_sensitivity.txt.)  The price: a slot of the ring is free only when BOTH pairs have consumed
// it (empty barriers count two commits), and only 33 clusters of 4 are co-resident on the 148 SMs (132 CTAs).
// QKV(+MLP) epilogue of one warp for one 128-column group (= one head of q, k or v, or 128 MLP columns) of one accumulator:
// bias (and, DEQ, the row / column scales), per-head RMSNorm and RoPE for q and k, split stores (bf16 or e4m3); MLP columns take
// the generic chunk (+bias, GELU).  `ta` = TMEM address of the group for this warp's lane quarter, `g0` its first output column.
//
// Latency structure (it matters once the main loop is short -- FP8 / NVFP4 at K = 3072 -- and the epilogue sets the pace):
//  * epi_qkv_prefetch issues every tile-uniform global load of a group (bias, column scales, norm weights, row scale) back to
//    back into registers (QkvPre); epi_qkv_group commits them to the warp's staging slots.  A caller with a tile queue
//    prefetches the NEXT group between commit and processing (`nxt`), so the round trip hides behind this group's work;
//  * both TMEM passes keep the tcgen05.ld of the following chunk in flight while a chunk is processed.
struct QkvPre {
  float b[4], w[4], g[4], rs;
};
struct QkvNext {   // the group a queue-driven caller will process next (has == false: none)
  bool has;
  int b;
  long long row;
  bool valid;
  int g0;
};
template <bool DEQ>
__device__ __forceinline__ void epi_qkv_prefetch(const GemmParams& p, int b, long long row, bool valid, int lane, int g0, QkvPre& pr) {
  const int which = (g0 >= 3 * p.heads * 128) ? 3 : (g0 >> 7) / p.heads;  // 0 q, 1 k, 2 v, 3 mlp
  const __nv_bfloat16* nw = which == 0 ? p.qnorm_w : p.knorm_w;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = i * 32 + lane;
    const bool in = g0 + j < p.N;
    pr.b[i] = (p.bias != nullptr && in) ? __bfloat162float(__ldg(p.bias + g0 + j)) : 0.f;
    pr.w[i] = (DEQ && in) ? __ldg(p.w_scale + g0 + j) : 0.f;
    pr.g[i] = which < 2 ? __bfloat162float(__ldg(nw + j)) : 1.f;
  }
  pr.rs = (DEQ && valid) ? __ldg(p.a_scale + (long long)b * p.a_scale_bs + row) : 1.f;
}

template <bool DEQ>
__device__ __forceinline__ void epi_qkv_group(const GemmParams& p, int b, long long row, bool valid, uint32_t vmask, int lane, uint32_t ta,
                                              int g0, float* sb, float* sg, float* sw, uint8_t* wst, uint64_t* tfull, uint32_t tfull_phase,
                                              QkvPre& pr, const QkvNext& nxt) {
  constexpr bool F8 = DEQ;
  const int D3 = 3 * p.heads * 128;
  const long long pos = (long long)p.seq_off + row;
  const bool active = g0 < p.N;
  const int hidx = g0 >> 7;
  const int which = (g0 >= D3) ? 3 : hidx / p.heads;  // 0 q, 1 k, 2 v, 3 mlp
  // piece k (4 (cos, sin) pairs) of this row: pe4[k * pe_st]
  const uint4* pe4 = reinterpret_cast<const uint4*>(p.pe) + (p.pe_blocked ? (pos >> 5) * 512 + (pos & 31) : pos * 16);
  const int pe_st = p.pe_blocked ? 32 : 1;
  uint4 pcur[4], pnxt[4];
  __syncwarp();  // the previous group is done with the staging slots
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sb[i * 32 + lane] = pr.b[i];
    sw[i * 32 + lane] = pr.w[i];
    sg[i * 32 + lane] = pr.g[i];
  }
  const float rs = pr.rs;
  if (active && which < 2 && valid) {  // first RoPE chunk to registers before the accumulator wait
#pragma unroll
    for (int i = 0; i < 4; ++i) pcur[i] = __ldg(pe4 + i * pe_st);
  }
  if (nxt.has) epi_qkv_prefetch<DEQ>(p, nxt.b, nxt.row, nxt.valid, lane, nxt.g0, pr);
  __syncwarp();
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  if (!active) return;
  if (which == 3) {
    // mlp region -> +bias, gelu -> `out` at column (g0 - 3D)
    const long long out_off = (long long)b * p.out_bs + pos * p.ldo - D3;
    uint32_t v[32];
    tmem_ld_x32(ta, v);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float f[32];
      tmem_ld_wait_x32(v);
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
      if (c < 3) tmem_ld_x32(ta + (c + 1) * 32, v);
      epi_generic_chunk<F8>(p, f, sb + c * 32, sg, nullptr, false, out_off, 0, g0 + c * 32, true, sw + c * 32, rs, wst, lane, vmask,
                            valid);
    }
    return;
  }
  const int head = hidx - which * p.heads;
  __nv_bfloat16* dst = (which == 0 ? p.q : (which == 1 ? p.k : p.v)) + (((long long)b * p.heads + head) * p.seq_total + pos) * 128;
  // dequantise (or just add the bias to) one chunk: x = acc * rs * w + bias
  const uint64_t rs2 = pack2f(rs, rs);
  auto load_chunk = [&](const uint32_t* v, int c, float* x) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(sb + c * 32 + i);
      if (F8) {  // packed fp32 pairs (FMUL2 / FFMA2): the same roundings as fmaf(v, rs * w, t), half the issue slots
        const float4 w = *reinterpret_cast<const float4*>(sw + c * 32 + i);
        const float2 lo = unpack2f(ffma2(pack2f(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), fmul2(rs2, pack2f(w.x, w.y)), pack2f(t.x, t.y)));
        const float2 hi = unpack2f(ffma2(pack2f(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), fmul2(rs2, pack2f(w.z, w.w)), pack2f(t.z, t.w)));
        x[i] = lo.x; x[i + 1] = lo.y; x[i + 2] = hi.x; x[i + 3] = hi.y;
      } else {
        x[i] = __uint_as_float(v[i]) + t.x; x[i + 1] = __uint_as_float(v[i + 1]) + t.y;
        x[i + 2] = __uint_as_float(v[i + 2]) + t.z; x[i + 3] = __uint_as_float(v[i + 3]) + t.w;
      }
    }
  };
  uint32_t v[32];
  tmem_ld_x32(ta, v);
  float rr = 1.f;
  if (which < 2) {  // pass 1: sum of squares of (acc + bias) over the head
    float ss = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float x[32];
      tmem_ld_wait_x32(v);
      load_chunk(v, c, x);
      tmem_ld_x32(ta + ((c + 1) & 3) * 32, v);  // next chunk of this pass; after the last one, chunk 0 again for pass 2
#pragma unroll
      for (int i = 0; i < 32; i += 4) ss += x[i] * x[i] + x[i + 1] * x[i + 1] + x[i + 2] * x[i + 2] + x[i + 3] * x[i + 3];
    }
    rr = rsqrtf(ss * (1.0f / 128.0f) + p.rms_eps);
  }
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {  // pass 2: normalise, rotate, store
    if (which < 2 && valid && c < 3) {
#pragma unroll
      for (int i = 0; i < 4; ++i) pnxt[i] = __ldg(pe4 + ((c + 1) * 4 + i) * pe_st);
    }
    // every lane computes (rows past the end hold zeros / stale table values and are masked at the store)
    float f[32];
    tmem_ld_wait_x32(v);
    load_chunk(v, c, f);
    if (c < 3) tmem_ld_x32(ta + (c + 1) * 32, v);
    if (which < 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w0 = *reinterpret_cast<const float4*>(sg + c * 32 + i * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(sg + c * 32 + i * 8 + 4);
        const float wt[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const uint32_t pw[4] = {pcur[i].x, pcur[i].y, pcur[i].z, pcur[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 cs = unpack_bf16(pw[j]);  // (cos, sin)
          const float x0 = f[i * 8 + 2 * j] * rr * wt[2 * j];
          const float x1 = f[i * 8 + 2 * j + 1] * rr * wt[2 * j + 1];
          f[i * 8 + 2 * j] = x0 * cs.x - x1 * cs.y;
          f[i * 8 + 2 * j + 1] = x0 * cs.y + x1 * cs.x;
        }
      }
    }
    if (p.qkv_f8) {  // 32 e4m3 bytes of this row: two full 16-byte vectors (rows are 128 bytes apart)
      if (valid) {
        uint32_t w8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(make_float2(f[4 * i], f[4 * i + 1]), __NV_SATFINITE, __NV_E4M3);
          const __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(make_float2(f[4 * i + 2], f[4 * i + 3]), __NV_SATFINITE, __NV_E4M3);
          w8[i] = uint32_t(lo) | (uint32_t(hi) << 16);
        }
        uint8_t* d8 = reinterpret_cast<uint8_t*>(which == 0 ? p.q : (which == 1 ? p.k : p.v)) +
                      (((long long)b * p.heads + head) * p.seq_total + pos) * 128 + c * 32;
        *reinterpret_cast<uint4*>(d8) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
        *reinterpret_cast<uint4*>(d8 + 16) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
      }
    } else {
      store_chunk32_coalesced(wst, lane, f, dst - (long long)lane * 128 + c * 32, 128, vmask, p.stream_out != 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) pcur[i] = pnxt[i];
  }
}

template <int BN, int EPI, bool CONV, int NCTA = 1, bool F8 = false, int CL = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
            const GemmParams p) {
  using Cfg = GemmCfg<BN, NCTA>;
  static_assert(CL == 1 || (NCTA == 2 && !CONV), "pair clusters are for the CTA-pair GEMMs");
  const uint32_t cl_rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const uint32_t cta_rank = cl_rank & 1u;   // rank inside the CTA pair
  const int pair_idx = int(cl_rank >> 1);   // which pair of the cluster (CL = 2)
  const int first_tile = blockIdx.x / (NCTA * CL), tile_stride = gridDim.x / (NCTA * CL);
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile uint32_t* ls_issued = reinterpret_cast<volatile uint32_t*>(smem + STAGES * Cfg::STAGE_BYTES + 192);   // producer -> sync warp
  volatile uint32_t* ls_allowed = ls_issued + 1;                                                               // sync warp -> producer
  const bool lockstep = (NCTA == 2) && CL == 1 && !CONV && p.ls_group > 0 && cta_rank == 0;

  // the warp index goes through a shuffle so that ptxas knows it is warp-uniform: the producer / issuer loops below then run
  // with every lane (uniform control flow) and their descriptors, coordinates and barrier addresses live in uniform registers
  // -- an `if (lane == 0)` region instead costs a broadcast loop (ELECT / R2UR / BRA.U.ANY) around every TMA / tcgen05 op
  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  constexpr int KE = F8 ? 2 * GEMM_BK : GEMM_BK;  // K elements per 128-byte stage row

  if (warp == GEMM_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);  // one tcgen05.commit per pair that reads (or writes into) this CTA's slot
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8 * NCTA);
    }
    *ls_issued = 0;
    *ls_allowed = uint32_t(p.ls_slack);
    fence_barrier_init();
  }
  if (warp == GEMM_WARP_MMA) {
    if (NCTA == 2) {
      tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync();  // peer barriers must be initialised before any remote arrive / TMA credit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The register file is partitioned per SM sub-partition (16K registers each, 3 warps here): the QKV
  // epilogue keeps a whole 128-column head in registers, so warpgroup 0 hands registers to the epilogue.
  if (warp >= GEMM_CTRL0 && warp < GEMM_CTRL0 + 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == GEMM_WARP_TMA) {
    // ================= TMA producer =================
    {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t lsg = 0;  // global K-group index of this pair (monotonic across its tiles)
      for (int tile = first_tile; tile < p.num_tiles; tile += tile_stride) {
        int tm, tn;
        gemm_tile_coords(p, tile, tm, tn);
        if (CL == 2) tm = 2 * tm + pair_idx;  // (a row tile past the end loads zeros: TMA out-of-bounds fill)
        const int b = tm / p.tiles_m_per_batch;
        const int tmb = tm - b * p.tiles_m_per_batch;
        int cy = 0, cx = 0;
        if (CONV) {  // conv_tiles_x counts tiles of the CTA (pair): 16 * NCTA pixels wide
          cy = (tmb / p.conv_tiles_x) * 8;
          cx = (tmb % p.conv_tiles_x) * (16 * NCTA) + int(cta_rank) * 16;
        }
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (lockstep && kb % p.ls_group == 0) {  // entering K-group `lsg`: stay within ls_slack groups of the slowest pair
            if ((kb != 0 || lsg != 0) && lane == 0) *ls_issued = lsg;  // groups < lsg are fully issued
            while (*ls_allowed <= lsg) {
            }
            ++lsg;
            __syncwarp();
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          // conv: which source pixel offset this k-block's tap reads, which image, which weight rows
          int dx = 0, dy = 0, bimg = b, wrow = 0, c0 = 0;
          if (CONV) {
            const int tap = kb / p.cin_blocks;
            c0 = (kb - tap * p.cin_blocks) * GEMM_BK;
            if (p.conv_up) {  // b = 4 * image + parity; 2 x 2 taps at rows {py - 1, py}, columns {px - 1, px}
              const int par = b & 3;
              bimg = b >> 2;
              dx = (par & 1) - 1 + (tap & 1);
              dy = (par >> 1) - 1 + (tap >> 1);
              wrow = par * p.N;
            } else {
              dx = tap % 3 - 1;
              dy = tap / 3 - 1;
            }
          }
          if (elect_one()) {
          if (NCTA == 2) {
            // the leader's barrier collects both CTAs' bytes; only the leader arrives on it
            const bool skip_w = CL == 1 && p.dbg_skip_w && (kb & 1);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], skip_w ? 2 * Cfg::A_BYTES : 2 * Cfg::STAGE_BYTES);
            if (CONV) {
              tma2_load_4d(sa, &tmap_a, &full_bar[stage], c0, cx + dx, cy + dy, bimg);
            } else {
              tma2_load_3d_hint(sa, &tmap_a, &full_bar[stage], kb * KE, tmb * (2 * GEMM_BM) + int(cta_rank) * GEMM_BM, b, p.hint_a);
            }
            if (CL == 2) {
              // this CTA's quarter of the W tile, multicast to the CTA of the other pair that holds the same half
              tma2_load_2d_mcast(sb + pair_idx * (Cfg::B_BYTES / 2), &tmap_w, &full_bar[stage], kb * KE,
                                 tn * BN + int(cta_rank) * (BN / 2) + pair_idx * (BN / 4), uint16_t(5u << cta_rank), p.hint_w);
            } else if (!skip_w) {
              tma2_load_2d_hint(sb, &tmap_w, &full_bar[stage], kb * KE, wrow + tn * BN + int(cta_rank) * (BN / 2), p.hint_w);
            }
          } else {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (CONV) {
            tma_load_4d(sa, &tmap_a, &full_bar[stage], c0, cx + dx, cy + dy, bimg);
          } else {
            tma_load_3d(sa, &tmap_a, &full_bar[stage], kb * KE, tmb * GEMM_BM, b);
          }
          tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * KE, wrow + tn * BN);
          }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (lockstep && lane == 0) *ls_issued = lsg;
    }
  } else if (warp == GEMM_CTRL0 + 2) {
    // ================= wave-lockstep sync warp (leader CTA of the pair) =================
    if (lockstep && lane == 0) {
      const int groups_per_tile = (p.k_blocks + p.ls_group - 1) / p.ls_group;
      int rounds = 0;
      for (int tile = first_tile; tile < p.num_tiles; tile += tile_stride) ++rounds;
      const uint32_t total = uint32_t(rounds) * groups_per_tile;
      bool alive = true;
      for (uint32_t g = 0; g < total; ++g) {
        if (alive) {
          while (*ls_issued <= g) {  // the producer has issued every load of group g
          }
          const int round = int(g) / groups_per_tile;
          const int members = min(tile_stride, p.num_tiles - round * tile_stride);  // pairs with a tile in this round
          const int slot = int(g % LS_RING);
          const unsigned int gen0 = *reinterpret_cast<volatile unsigned int*>(&g_lockstep.gen[slot]);
          const unsigned int old = atomicAdd(&g_lockstep.count[slot], 1u);
          if (old == unsigned(members - 1)) {
            g_lockstep.count[slot] = 0;
            __threadfence();
            atomicAdd(&g_lockstep.gen[slot], 1u);
          } else {
            int spins = 0;
            while (*reinterpret_cast<volatile unsigned int*>(&g_lockstep.gen[slot]) == gen0) {
              if (++spins > (1 << 18)) {  // ~0.2 s: something else is wrong; run unsynchronised from here on
                alive = false;
                break;
              }
              __nanosleep(200);
            }
          }
        }
        *ls_allowed = alive ? g + 1 + uint32_t(p.ls_slack) : 0xffffffffu;
      }
    }
  } else if (warp == GEMM_WARP_MMA) {
    // ================= MMA issuer =================
    if (cta_rank == 0) {
      constexpr uint32_t idesc = F8 ? make_idesc_f8(GEMM_BM * NCTA, BN) : make_idesc_bf16(GEMM_BM * NCTA, BN, 0, 0);
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < p.num_tiles; tile += tile_stride) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tbase + acc * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem0 + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          if (elect_one()) {
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * 32, 16, 1024);
            if (F8) {
              if (NCTA == 2) umma2_ss_f8(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_ss_f8(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            } else {
              if (NCTA == 2) umma2_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          if (NCTA == 2 && CL == 2) tc_commit2_mask(&empty_bar[stage], 0xF);  // both pairs wait for both pairs
          else if (NCTA == 2) tc_commit2(&empty_bar[stage]);
          else tc_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
          if (NCTA == 2 && CL == 2) tc_commit2_mask(&tfull_bar[acc], uint16_t(3u << (2 * pair_idx)));
          else if (NCTA == 2) tc_commit2(&tfull_bar[acc]);
          else tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }
  } else {
    // ================= epilogue warps =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int half = (warp - GEMM_EPI0) >> 2;  // which half of the tile's columns this warp owns
    const int r = quarter * 32 + lane;
    float* sb = reinterpret_cast<float*>(smem + Cfg::EPI_OFF) + (warp - GEMM_EPI0) * 384;  // bias (<= 128 floats)
    float* sg = sb + 128;                                                           // gate / norm weight
    float* sw = sb + 256;                                                           // FP8: per-column weight scales
    // conv tiles map accumulator rows to 8x16 pixel patches: a warp's 32 rows are two image lines of 16 pixels
    uint8_t* wst = smem + Cfg::STORE_OFF + (warp - GEMM_EPI0) * 2048;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < p.num_tiles; tile += tile_stride) {
      int tm, tn;
      gemm_tile_coords(p, tile, tm, tn);
      if (CL == 2) tm = 2 * tm + pair_idx;
      const int b = tm / p.tiles_m_per_batch;
      const int tmb = tm - b * p.tiles_m_per_batch;
      bool valid;
      long long row;  // row index inside the batch (pixel index for conv)
      long long conv_off = 0;  // conv: element offset of this thread's output pixel inside its image
      int bimg = b, par = 0;   // conv_up: b = 4 * image + parity
      if (CONV) {
        const int y = (tmb / p.conv_tiles_x) * 8 + (r >> 4);
        const int x = (tmb % p.conv_tiles_x) * (16 * NCTA) + int(cta_rank) * 16 + (r & 15);
        valid = (y < p.conv_H) && (x < p.conv_W);
        row = (long long)y * p.conv_W + x;
        if (p.conv_up) {  // output pixel (2y + py, 2x + px) of the 2H x 2W image: ldo = 2 Cout, conv_line = 4 W Cout
          par = b & 3;
          bimg = b >> 2;
          conv_off = (long long)y * p.conv_line + (long long)x * p.ldo + ((long long)(par >> 1) * 2 * p.conv_W + (par & 1)) * (p.ldo >> 1);
        } else {
          conv_off = row * p.ldo;
        }
      } else {
        row = (long long)tmb * (GEMM_BM * NCTA) + cta_rank * GEMM_BM + r;
        valid = row < p.rows && (CL == 1 || tm < p.tiles_m_real);
      }
      const uint32_t taddr = tmem_base + acc * BN + (uint32_t(quarter * 32) << 16);
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);

      // NOTE on code size: the chunk loops below are deliberately NOT unrolled.  A fully unrolled epilogue is
      // ~200 KB of SASS that eight warps stream through once per tile -> instruction-cache misses
      // (stall_no_inst) dominated the epilogue; one 32-column chunk body fits the L0/L1 instruction caches.
      if (EPI == EPI_GENERIC) {
        constexpr int CH = BN / 64;        // 32-column chunks per warp
        constexpr int WN = BN / 2;         // columns per warp
        const int nw0 = tn * BN + half * WN;
        const long long out_off = CONV ? (long long)bimg * p.out_bs + conv_off : (long long)b * p.out_bs + row * p.ldo;
        const long long res_off = (long long)bimg * p.resid_bs + row * p.ldr;
        const bool vec_ok = ((p.ldo | p.ldr) & 7) == 0;
        // ---- everything that does not depend on the accumulator happens BEFORE the wait
        __syncwarp();
        stage_vec(sb, p.bias, nw0, WN, p.N, 0.f, lane);
        if (p.gate) stage_vec(sg, p.gate + (long long)b * p.gate_bs, nw0, WN, p.N, 1.f, lane);
        float rs = 1.f;
        if (F8) {
          stage_vec_f32(sw, p.w_scale, nw0, WN, p.N, 0.f, lane);
          if (valid) rs = __ldg(p.a_scale + (long long)b * p.a_scale_bs + row);
        }
        const bool rr_ok = p.resid != nullptr && valid && vec_ok && (nw0 + WN <= p.N);
        uint4 rcur[4], rnxt[4];
        if (rr_ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i) rcur[i] = *reinterpret_cast<const uint4*>(p.resid + res_off + nw0 + i * 8);
        }
        __syncwarp();
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < CH; ++c) {
          const int n0 = nw0 + c * 32;
          if (n0 >= p.N) break;
          if (rr_ok && c + 1 < CH) {  // residual of the next chunk travels while this chunk is processed
#pragma unroll
            for (int i = 0; i < 4; ++i) rnxt[i] = *reinterpret_cast<const uint4*>(p.resid + res_off + n0 + 32 + i * 8);
          }
          uint32_t v[32];
          __syncwarp();
          tmem_ld_x32(taddr + half * WN + c * 32, v);
          tmem_ld_wait();
          {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
            epi_generic_chunk<F8>(p, f, sb + c * 32, sg + c * 32, rcur, rr_ok, out_off, res_off, n0,
                                  vec_ok && (n0 + 32 <= p.N), sw + c * 32, rs, wst, lane, vmask, valid,
                                  CONV ? (long long)(lane & 15) * p.ldo + (long long)(lane >> 4) * p.conv_line : -1,
                                  CONV ? p.conv_line : -1);
            if (CONV && p.gn_partials != nullptr)  // f[] now holds the final values (bias, residual) of 32 channels
              gn_chunk_partials(p.gn_partials + (((long long)bimg * p.gn_nblk +
                                                  (long long)((par * p.tiles_m_per_batch + tmb) * NCTA + int(cta_rank)) * 4 + quarter) * 32 +
                                                 n0 / p.gn_gs) * 2,
                                f, valid, lane, p.gn_gs);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) rcur[i] = rnxt[i];
        }
      } else {
        // ---- QKV(+MLP) epilogue: this warp owns one 128-column group (= one head of q, k or v, or 128 MLP columns)
        QkvPre pre;
        epi_qkv_prefetch<F8>(p, b, row, valid, lane, tn * BN + half * 128, pre);
        epi_qkv_group<F8>(p, b, row, valid, vmask, lane, taddr + half * 128, tn * BN + half * 128, sb, sg, sw, wst, &tfull_bar[acc],
                          acc_phase, pre, QkvNext{false, 0, 0, false, 0});
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_leader(&tempty_bar[acc]);
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync();  // neither CTA may exit (or free TMEM) while its peer still uses it
  else __syncthreads();
  if (warp == GEMM_WARP_MMA) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace fx
