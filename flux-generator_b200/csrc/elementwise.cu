// HBM-bound kernels of the hot path: row norms (+AdaLN modulation), conditioning GEMVs, timestep
// embedding, Euler step, latent packing, GroupNorm(+SiLU), nearest upsample, row softmax, transpose,
// image finish, embedding gather, gated activation.  All vectorised to 16-byte accesses, bf16 in HBM,
// fp32 in registers.
#include <cuda_fp4.h>
#include <cuda_fp8.h>
#include <stdlib.h>

#include "api_common.cuh"
#include "sm100.cuh"

namespace fx {

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
  u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// FP8 row quantisation (--quantize): q = e4m3(x * inv), inv = 448 / absmax (round-to-nearest-even, saturating);
// dequantisation scale = absmax * (1/448); an all-zero row gets inv = scale = 1.  Every operation is a single
// IEEE fp32 operation, so the CPU oracle reproduces bytes and scales exactly.
__device__ __forceinline__ float fp8_row_scale(float amax) { return amax > 0.f ? amax * (1.0f / 448.0f) : 1.0f; }
__device__ __forceinline__ float fp8_row_inv(float amax) { return amax > 0.f ? __fdiv_rn(448.0f, amax) : 1.0f; }
__device__ __forceinline__ uint2 quant8(const float* f, float inv) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = __nv_cvt_float2_to_fp8x2(make_float2(__fmul_rn(f[2 * i], inv), __fmul_rn(f[2 * i + 1], inv)), __NV_SATFINITE, __NV_E4M3);
  return make_uint2(w[0] | (w[1] << 16), w[2] | (w[3] << 16));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ row norm: one warp per row
struct RowNormParams {
  const __nv_bfloat16* x; long long ldx, x_bs;
  __nv_bfloat16* out; long long ldo, out_bs;
  const __nv_bfloat16 *p0, *p1; long long p_bs;
  float eps; int mode, batch, rows, D;
  float* scale_out; long long scale_bs;  // F8OUT: out holds e4m3 bytes, scale_out[b][r] the row's dequantisation scale
  uint8_t* sf_out;                       // NVFP4 output (block kernel, OUT == 2): UE4M3 block-scale atoms; out = compact e2m1 rows
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// The row stays packed (bf16) in registers and is unpacked in each of the three passes: 48 instead of 96
// data registers for D = 3072, so two 8-row blocks fit per SM instead of one (the kernel is a pure
// HBM stream: more rows in flight = more bandwidth).
template <int ITERS, bool F8OUT = false>  // ITERS = ceil(D / 256); D % 8 == 0
__global__ void __launch_bounds__(256, 2) rownorm_kernel(const RowNormParams p) {
  const int lane = threadIdx.x & 31;
  // grid-stride over rows: with a persistent grid (a few blocks per SM) a warp walks many rows and the block
  // launch / retire churn of one-row-per-warp blocks disappears
  for (long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); gw < (long long)p.batch * p.rows;
       gw += (long long)gridDim.x * 8) {
  const int b = int(gw / p.rows);
  const long long r = gw - (long long)b * p.rows;
  const __nv_bfloat16* xr = p.x + b * p.x_bs + r * p.ldx;
  uint4 raw[ITERS];
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = i * 256 + lane * 8;
    raw[i] = (c < p.D) ? *reinterpret_cast<const uint4*>(xr + c) : make_uint4(0, 0, 0, 0);
  }
  float mean = 0.f;
  if (p.mode != 2) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      float v[8];
      unpack8(raw[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
    }
    mean = warp_sum(s) / float(p.D);
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    if (i * 256 + lane * 8 < p.D) {
      float v[8];
      unpack8(raw[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[j] - mean;
        ss += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / float(p.D) + p.eps);
  const long long pb = (p.mode == 0) ? b * p.p_bs : 0;
  __nv_bfloat16* orow = p.out + b * p.out_bs + r * p.ldo;
  uint8_t* qrow = reinterpret_cast<uint8_t*>(p.out) + b * p.out_bs + r * p.ldo;
  // F8OUT: the modulated row is kept packed (bf16, 48 more registers for D = 3072) while its absmax is reduced,
  // then quantised -- exactly quantize_rows(rownorm(x)) in one kernel: one HBM read, a half-size write
  uint4 keep[F8OUT ? ITERS : 1];
  float amax = 0.f;
  __nv_bfloat162 amax2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = i * 256 + lane * 8;
    if (c >= p.D) continue;
    float v[8], a[8], s[8], o[8];
    unpack8(raw[i], v);
    load8(p.p0 + pb + c, a);
    if (p.mode != 2) load8(p.p1 + pb + c, s);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float y = (v[j] - mean) * rstd;
      if (p.mode == 0) o[j] = (1.0f + s[j]) * y + a[j];  // p0 = shift, p1 = scale
      else if (p.mode == 1) o[j] = y * a[j] + s[j];      // p0 = weight, p1 = bias
      else o[j] = y * a[j];                              // p0 = weight
    }
    if (!F8OUT) {
      store8(orow + c, o);
    } else {
      uint4 u;
      u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]);
      u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
      keep[i] = u;
      // absmax of the ROUNDED values, on packed bf16 pairs (|x| and max are exact in bf16: 8 instead of 16 instructions)
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        amax2 = __hmax2(amax2, __habs2(*reinterpret_cast<const __nv_bfloat162*>(&w4[j])));
    }
  }
  if (F8OUT) {
    amax = warp_max(fmaxf(__low2float(amax2), __high2float(amax2)));
    if (lane == 0) p.scale_out[b * p.scale_bs + r] = fp8_row_scale(amax);
    const float inv = fp8_row_inv(amax);
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int c = i * 256 + lane * 8;
      if (c >= p.D) continue;
      float o[8];
      unpack8(keep[i], o);
      *reinterpret_cast<uint2*>(qrow + c) = quant8(o, inv);
    }
  }
  }  // row loop
}

// Wide rows (D >= 1024: the MMDiT's 3072-wide residual stream, T5's 4096): one 128-thread block per row, the row
// AND its modulation vectors requested up front (one exposed latency instead of two), statistics through two
// block reductions.  Same arithmetic per element as rownorm_kernel.  Measured on the benchmark's [8 x 4352, 3072]
// tensor: the warp-per-row kernel above sustains 3.9 TB/s (bf16 out) / 1.6 TB/s (e4m3 out, 3 B per element); the
// block-per-row quantiser below, which has this structure, 6.5 TB/s.
__device__ __forceinline__ float block_sum4(float v, float* red, int lane, int warp) {
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  return (red[0] + red[1]) + (red[2] + red[3]);
}
__device__ __forceinline__ uint64_t bf2_to_f2(uint32_t u) {  // packed bf16 pair -> packed fp32 pair (2 ALU ops)
  return pack2f(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
// OUT: 0 bf16, 1 e4m3 + row scale (= fx_quantize_rows of the bf16 result), 2 NVFP4 (= fx_quantize_rows_fp4 of the bf16 result:
// compact e2m1 rows, UE4M3 scale atoms, fp32 row scale; bit-identical to the two-kernel sequence, csrc/gemm4.cu)
template <int VPT, int OUT, int MINB>  // VPT = ceil(D / 1024) 16-byte vectors per thread; MINB resident blocks per SM
__global__ void __launch_bounds__(128, MINB) rownorm_block_kernel(const RowNormParams p) {
  constexpr bool F8OUT = OUT == 1;
  // Persistent blocks walk the rows with a grid stride; the next row's loads are issued before the current row is
  // reduced.  The kernel was INSTRUCTION-bound, not HBM-bound (~350 instructions per thread and row: the packed row was
  // unpacked three times and the modulation vectors once per row; 3.9 TB/s whatever the launch geometry).  Now the row
  // is unpacked once into packed fp32 pairs, all arithmetic runs on FADD2 / FFMA2, the centred row is reused by the
  // output pass, and the per-column affine (P, A) -- (1 + scale, shift) / (weight, bias) / (weight, 0) -- is converted
  // once per batch element and kept in registers: ~140 instructions per thread and row.
  __shared__ float red[2][3][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long total = (long long)p.batch * p.rows;
  long long gr = blockIdx.x;
  if (gr >= total) return;
  constexpr int NP = VPT * 4;  // fp32 pairs per thread
  uint4 nxt[VPT];
  uint64_t P2[NP], A2[NP];
  bool in[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) in[i] = (i * 128 + tid) * 8 < p.D;
  int pb_batch = -1;
  auto load_row = [&](long long g, uint4* dst) {
    const int b = int(g / p.rows);
    const long long r = g - (long long)b * p.rows;
    const __nv_bfloat16* xr = p.x + b * p.x_bs + r * p.ldx;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      dst[i] = in[i] ? *reinterpret_cast<const uint4*>(xr + (i * 128 + tid) * 8) : make_uint4(0, 0, 0, 0);
  };
  load_row(gr, nxt);
  const uint64_t one2 = pack2f(1.0f, 1.0f);
  const float inv_d = 1.0f / float(p.D);
  int flip = 0;
  for (; gr < total; gr += gridDim.x, flip ^= 1) {
    const int b = int(gr / p.rows);
    const long long r = gr - (long long)b * p.rows;
    uint64_t v2[NP];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      v2[4 * i] = bf2_to_f2(nxt[i].x); v2[4 * i + 1] = bf2_to_f2(nxt[i].y);
      v2[4 * i + 2] = bf2_to_f2(nxt[i].z); v2[4 * i + 3] = bf2_to_f2(nxt[i].w);
    }
    const int pbb = (p.mode == 0) ? b : 0;
    if (pbb != pb_batch) {  // (1 + scale, shift) of this batch element / (weight, bias) / (weight, 0)
      pb_batch = pbb;
      const long long pb = (long long)pbb * p.p_bs;
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int c = (i * 128 + tid) * 8;
        const uint4 z = make_uint4(0, 0, 0, 0);
        const uint4 u0 = in[i] ? __ldg(reinterpret_cast<const uint4*>(p.p0 + pb + c)) : z;
        const uint4 u1 = (in[i] && p.mode != 2) ? __ldg(reinterpret_cast<const uint4*>(p.p1 + pb + c)) : z;
        const uint32_t w0[4] = {u0.x, u0.y, u0.z, u0.w}, w1[4] = {u1.x, u1.y, u1.z, u1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (p.mode == 0) { P2[4 * i + j] = fadd2(one2, bf2_to_f2(w1[j])); A2[4 * i + j] = bf2_to_f2(w0[j]); }
          else { P2[4 * i + j] = bf2_to_f2(w0[j]); A2[4 * i + j] = bf2_to_f2(w1[j]); }
        }
      }
    }
    if (gr + gridDim.x < total) load_row(gr + gridDim.x, nxt);  // in flight while this row is reduced
    float (*rd)[4] = red[flip];  // double-buffered: the next iteration's first reduction cannot overtake a reader
    float mean = 0.f;
    if (p.mode != 2) {
      uint64_t s0 = pack2f(0.f, 0.f), s1 = s0;
#pragma unroll
      for (int k = 0; k < NP; k += 2) { s0 = fadd2(s0, v2[k]); s1 = fadd2(s1, v2[k + 1]); }
      const float2 t = unpack2f(fadd2(s0, s1));
      mean = block_sum4(t.x + t.y, rd[0], lane, warp) * inv_d;
    }
    const uint64_t nm2 = pack2f(-mean, -mean);
    uint64_t q0 = pack2f(0.f, 0.f), q1 = q0;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t d = in[i] ? fadd2(v2[4 * i + j], nm2) : pack2f(0.f, 0.f);  // centred row, reused below
        v2[4 * i + j] = d;
        if (j & 1) q1 = ffma2(d, d, q1);
        else q0 = ffma2(d, d, q0);
      }
    }
    const float2 tq = unpack2f(fadd2(q0, q1));
    const float rstd = rsqrtf(block_sum4(tq.x + tq.y, rd[1], lane, warp) * inv_d + p.eps);
    const uint64_t k2 = pack2f(rstd, rstd);
    __nv_bfloat16* orow = p.out + b * p.out_bs + r * p.ldo;
    uint8_t* qrow = reinterpret_cast<uint8_t*>(p.out) + b * p.out_bs + r * p.ldo;
    __nv_bfloat162 amax2 = __floats2bfloat162_rn(0.f, 0.f);
    uint4 keep[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 o = unpack2f(ffma2(fmul2(v2[4 * i + j], k2), P2[4 * i + j], A2[4 * i + j]));
        w[j] = pack_bf16(o.x, o.y);
        if (F8OUT) amax2 = __hmax2(amax2, __habs2(*reinterpret_cast<const __nv_bfloat162*>(&w[j])));
      }
      keep[i] = make_uint4(w[0], w[1], w[2], w[3]);
      if (OUT == 0 && in[i]) *reinterpret_cast<uint4*>(orow + (i * 128 + tid) * 8) = keep[i];
    }
    if (OUT == 2) {  // exactly quantize_rows_fp4(rownorm(x)): two-level NVFP4 of the ROUNDED values (see csrc/gemm4.cu)
      float bmax[VPT], amax = 0.f;
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        float o[8];
        unpack8(keep[i], o);
        float mx = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) mx = fmaxf(mx, fabsf(o[j]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));  // a block of 16 = two neighbouring threads
        bmax[i] = mx;
        amax = fmaxf(amax, mx);
      }
      amax = warp_max(amax);
      if (lane == 0) rd[2][warp] = amax;
      __syncthreads();
      amax = fmaxf(fmaxf(rd[2][0], rd[2][1]), fmaxf(rd[2][2], rd[2][3]));
      const float g = amax > 0.f ? __fmul_rn(amax, 1.0f / 2688.0f) : 1.0f;
      const float rg = __frcp_rn(g);
      if (tid == 0) p.scale_out[gr] = g;
      uint8_t* sf_row = p.sf_out + (gr >> 7) * (long long)(p.D / 64) * 512 + (int(gr) & 31) * 16 + ((int(gr) & 127) >> 5) * 4;
      uint8_t* q4 = reinterpret_cast<uint8_t*>(p.out) + gr * (long long)(p.D / 2);
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int c = (i * 128 + tid) * 8;
        const float u = __fmul_rn(__fmul_rn(bmax[i], 1.0f / 6.0f), rg);
        const __nv_fp8_storage_t sf8 = __nv_cvt_float_to_fp8(u, __NV_SATFINITE, __NV_E4M3);
        const float d = __fmul_rn(__half2float(__half(__nv_cvt_fp8_to_halfraw(sf8, __NV_E4M3))), g);
        uint32_t sfw = uint32_t(sf8);  // the four scales of a K-group sit in lanes 8m, 8m+2, 8m+4, 8m+6
        sfw |= __shfl_down_sync(0xffffffffu, sfw, 2) << 8;
        sfw |= __shfl_down_sync(0xffffffffu, sfw, 4) << 16;
        if (in[i] && (tid & 7) == 0) *reinterpret_cast<uint32_t*>(sf_row + (c >> 6) * 512) = sfw;
        const float rdv = d > 0.f ? __frcp_rn(d) : 0.f;
        float o[8];
        unpack8(keep[i], o);
        uint32_t w = 0;
#pragma unroll
        for (int e = 0; e < 8; e += 2)
          w |= uint32_t(__nv_cvt_float2_to_fp4x2(make_float2(__fmul_rn(o[e], rdv), __fmul_rn(o[e + 1], rdv)), __NV_E2M1, cudaRoundNearest))
               << (4 * e);
        if (in[i]) *reinterpret_cast<uint32_t*>(q4 + (c >> 1)) = w;
      }
    }
    if (F8OUT) {  // exactly quantize_rows(rownorm(x)): absmax of the ROUNDED values, then e4m3 + the row's scale
      float amax = warp_max(fmaxf(__low2float(amax2), __high2float(amax2)));
      if (lane == 0) rd[2][warp] = amax;
      __syncthreads();
      amax = fmaxf(fmaxf(rd[2][0], rd[2][1]), fmaxf(rd[2][2], rd[2][3]));
      if (tid == 0) p.scale_out[b * p.scale_bs + r] = fp8_row_scale(amax);
      const float inv = fp8_row_inv(amax);
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        if (!in[i]) continue;
        float o[8];
        unpack8(keep[i], o);
        *reinterpret_cast<uint2*>(qrow + (i * 128 + tid) * 8) = quant8(o, inv);
      }
    }
  }
}

// ------------------------------------------------------------------ FP8 row quantisation: one warp per row
// x bf16 [batch][rows][K] -> q e4m3 [batch][rows][K] + scale fp32 [batch][rows].  Two passes over the row; the
// second one hits L1 / L2 (a row is at most 30 KB), so HBM sees one bf16 read and one byte-wide write.
struct QuantParams {
  const __nv_bfloat16* x; long long ldx, x_bs;
  uint8_t* q; long long ldq, q_bs;
  float* scale; long long scale_bs;
  int batch, rows, K;
};
__global__ void __launch_bounds__(256) quantize_rows_kernel(const QuantParams p) {
  const int lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (gw >= (long long)p.batch * p.rows) return;
  const int b = int(gw / p.rows);
  const long long r = gw - (long long)b * p.rows;
  const __nv_bfloat16* xr = p.x + b * p.x_bs + r * p.ldx;
  float amax = 0.f;
  for (int c = lane * 8; c < p.K; c += 256) {
    float v[8];
    load8(xr + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(v[j]));
  }
  amax = warp_max(amax);
  if (lane == 0) p.scale[b * p.scale_bs + r] = fp8_row_scale(amax);
  const float inv = fp8_row_inv(amax);
  uint8_t* qr = p.q + b * p.q_bs + r * p.ldq;
  for (int c = lane * 8; c < p.K; c += 256) {
    float v[8];
    load8(xr + c, v);
    *reinterpret_cast<uint2*>(qr + c) = quant8(v, inv);
  }
}

// Long rows (the attention | GELU(mlp) operand of linear2 / mlp.2, K = 12288 / 15360): one 256-thread block per
// row, the row held packed in registers between the absmax reduction and the conversion -> a single HBM pass.
template <int ITERS>  // ITERS = ceil(K / 2048)
__global__ void __launch_bounds__(256) quantize_rows_block_kernel(const QuantParams p) {
  __shared__ float s_max[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / p.rows;
  const long long r = blockIdx.x - (long long)b * p.rows;
  const __nv_bfloat16* xr = p.x + b * p.x_bs + r * p.ldx;
  uint4 raw[ITERS];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = i * 2048 + tid * 8;
    raw[i] = (c < p.K) ? *reinterpret_cast<const uint4*>(xr + c) : make_uint4(0, 0, 0, 0);
    float v[8];
    unpack8(raw[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(v[j]));
  }
  amax = warp_max(amax);
  if (lane == 0) s_max[warp] = amax;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) amax = fmaxf(amax, s_max[w]);
  if (tid == 0) p.scale[b * p.scale_bs + r] = fp8_row_scale(amax);
  const float inv = fp8_row_inv(amax);
  uint8_t* qr = p.q + b * p.q_bs + r * p.ldq;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = i * 2048 + tid * 8;
    if (c >= p.K) continue;
    float v[8];
    unpack8(raw[i], v);
    *reinterpret_cast<uint2*>(qr + c) = quant8(v, inv);
  }
}

// ------------------------------------------------------------------ conditioning GEMV
// out[b][n] = f_out( sum_k f_in(in[b][k]) W[n][k] + bias[n] + add[b][n] ), batch <= 8 per launch pass.
struct GemvParams {
  const __nv_bfloat16* in; long long ld_in;
  const __nv_bfloat16* W; long long ldw;
  const __nv_bfloat16 *bias, *add; long long ld_add;
  __nv_bfloat16* out; long long ld_out;
  int batch, N, K, silu_in, silu_out;
};

__global__ void __launch_bounds__(256) gemv_kernel(const GemvParams p) {
  extern __shared__ __nv_bfloat16 s_in[];  // [batch][K]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int idx = threadIdx.x; idx < p.batch * p.K; idx += blockDim.x) {
    const int b = idx / p.K, k = idx - b * p.K;
    float x = __bfloat162float(p.in[b * p.ld_in + k]);
    if (p.silu_in) x = silu(x);
    s_in[idx] = __float2bfloat16(x);  // bf16 like the reference's nn.silu output
  }
  __syncthreads();
  // Each warp owns GEMV_R consecutive output rows: one shared-memory read of the activations feeds
  // GEMV_R weight rows (the kernel streams weights from HBM; smem traffic was the co-limiter at R = 1).
  constexpr int R = 4;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int n0 = (blockIdx.x * (blockDim.x >> 5) + warp) * R; n0 < p.N; n0 += warps_total * R) {
    float acc[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[r][b] = 0.f;
    for (int k = lane * 8; k < p.K; k += 256) {
      uint4 u[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        u[r] = (n0 + r < p.N) ? __ldg(reinterpret_cast<const uint4*>(p.W + (long long)(n0 + r) * p.ldw + k)) : make_uint4(0, 0, 0, 0);
      float w[R][8];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float2 a = unpack_bf16(u[r].x), bb = unpack_bf16(u[r].y), c = unpack_bf16(u[r].z), d = unpack_bf16(u[r].w);
        w[r][0] = a.x; w[r][1] = a.y; w[r][2] = bb.x; w[r][3] = bb.y; w[r][4] = c.x; w[r][5] = c.y; w[r][6] = d.x; w[r][7] = d.y;
      }
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if (b < p.batch) {
          float x[8];
          load8(s_in + b * p.K + k, x);
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r][b] += w[r][j] * x[j];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[r][b] = warp_sum(acc[r][b]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int n = n0 + r;
        if (n >= p.N) break;
        const float bias = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if (b >= p.batch) continue;
          float v = acc[r][b] + bias;
          if (p.silu_out) v = silu(__bfloat162float(__float2bfloat16(v)));
          if (p.add) v = __bfloat162float(__float2bfloat16(v)) + __bfloat162float(p.add[b * p.ld_add + n]);
          p.out[b * p.ld_out + n] = __float2bfloat16(v);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ timestep embedding
__global__ void timestep_embedding_kernel(const __nv_bfloat16* t, __nv_bfloat16* out, int batch, int dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= batch * half) return;
  const int b = idx / half, j = idx - b * half;
  // bf16(1000 * bf16(t)) -- the multiply happens in bf16 in the reference (flux/layers.py:54)
  const float tt = __bfloat162float(__float2bfloat16(1000.0f * __bfloat162float(t[b])));
  const float freq = expf(-9.210340371976184f * (float(j) / float(half)));
  const float x = tt * freq;
  out[b * dim + j] = __float2bfloat16(cosf(x));
  out[b * dim + half + j] = __float2bfloat16(sinf(x));
}

// ------------------------------------------------------------------ Euler step
__global__ void euler_kernel(__nv_bfloat16* x, const __nv_bfloat16* pred, float dt, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float a[8], b[8];
    load8(x + i, a);
    load8(pred + i, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = a[j] + __bfloat162float(__float2bfloat16(dt * b[j]));
    store8(x + i, a);
  } else {
    for (long long k = i; k < n; ++k)
      x[k] = __float2bfloat16(__bfloat162float(x[k]) + __bfloat162float(__float2bfloat16(dt * __bfloat162float(pred[k]))));
  }
}

// ------------------------------------------------------------------ latent packing
__global__ void patchify_kernel(const __nv_bfloat16* x, __nv_bfloat16* out, int b, int h, int w, int c) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)b * h * w * c;
  if (idx >= total) return;
  // out index: [b][row*(w/2)+col][ch*4 + dy*2 + dx]
  const int f = int(idx % (4 * c));
  const long long tok = idx / (4 * c);
  const int col = int(tok % (w / 2));
  const int row = int((tok / (w / 2)) % (h / 2));
  const int bi = int(tok / ((long long)(w / 2) * (h / 2)));
  const int ch = f >> 2, dy = (f >> 1) & 1, dx = f & 1;
  out[idx] = x[(((long long)bi * h + (row * 2 + dy)) * w + (col * 2 + dx)) * c + ch];
}

__global__ void unpatchify_scale_kernel(const __nv_bfloat16* packed, __nv_bfloat16* z, int b, int h, int w, int c,
                                        int c_pad, float inv_scale, float shift) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)b * h * w * c_pad;
  if (idx >= total) return;
  const int ch = int(idx % c_pad);
  const long long pix = idx / c_pad;
  const int x = int(pix % w), y = int((pix / w) % h), bi = int(pix / ((long long)w * h));
  float v = 0.f;
  if (ch < c) {
    const long long tok = ((long long)bi * (h / 2) + y / 2) * (w / 2) + x / 2;
    const float pv = __bfloat162float(packed[tok * (4 * c) + ch * 4 + (y & 1) * 2 + (x & 1)]);
    // z / scale_factor then + shift_factor, each rounded like the reference's two bf16 ops
    v = __bfloat162float(__float2bfloat16(pv * inv_scale)) + shift;
  }
  z[idx] = __float2bfloat16(v);
}

// ------------------------------------------------------------------ Gaussian prior, written packed
// FluxSampler.sample_prior + _prepare_latent_images in one pass (flux/sampler.py:44-45, flux/flux.py:53-58):
// counter-based Philox4x32-10 keyed by (seed, GLOBAL image index) and counted by the NHWC element index, so an
// image's noise does not depend on the batch it is generated in nor on how the batch is sharded over GPUs.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void prior_kernel(__nv_bfloat16* out, int b, int h, int w, int c, uint32_t seed_lo, uint32_t seed_hi, int first_index) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one Philox call = 4 consecutive NHWC elements
  const long long per_image = (long long)h * w * c / 4;
  if (q >= (long long)b * per_image) return;
  const int bi = int(q / per_image);
  const long long qi = q - (long long)bi * per_image;
  uint32_t r[4];
  philox4x32_10((uint32_t)qi, (uint32_t)(qi >> 32), (uint32_t)(first_index + bi), 0u, seed_lo, seed_hi, r);
  float n[4];
#pragma unroll
  for (int k = 0; k < 2; ++k) {  // Box-Muller on two uniform pairs
    const float u1 = ((r[2 * k] >> 8) + 1) * (1.0f / 16777216.0f);    // (0, 1]
    const float u2 = (r[2 * k + 1] >> 8) * (1.0f / 16777216.0f);      // [0, 1)
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    n[2 * k] = rad * cs;
    n[2 * k + 1] = rad * sn;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long e = qi * 4 + k;                  // NHWC element ((y*w + x)*c + ch)
    const int ch = int(e % c);
    const long long pix = e / c;
    const int x = int(pix % w), y = int(pix / w);
    const long long tok = (long long)(y >> 1) * (w >> 1) + (x >> 1);
    out[((long long)bi * (h / 2) * (w / 2) + tok) * (4 * c) + ch * 4 + (y & 1) * 2 + (x & 1)] = __float2bfloat16(n[k]);
  }
}

// ------------------------------------------------------------------ GroupNorm (32 groups) on NHWC
// stats: each block reduces a slab of 256 rows; thread owns 8 consecutive channels.  No atomics anywhere:
// per-block partial sums are combined in a fixed order, so results are bit-reproducible run to run.
constexpr int GN_ROWS_PER_BLOCK = 256;

__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const __nv_bfloat16* x, float* partials, long long hw, int C) {
  __shared__ float sm[256][17];  // [thread][8 sums | 8 sums of squares], padded
  const int b = blockIdx.y;
  const int tpr = C / 8;                 // threads per row
  const int rpi = 256 / tpr;             // rows per iteration
  const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
  const long long r0 = (long long)blockIdx.x * GN_ROWS_PER_BLOCK;
  const long long r1 = min(hw, r0 + GN_ROWS_PER_BLOCK);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  const __nv_bfloat16* base = x + (long long)b * hw * C + tc * 8;
  for (long long r = r0 + tr; r < r1; r += rpi) {
    float v[8];
    load8(base + r * C, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] += v[j] * v[j]; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sm[threadIdx.x][j] = s[j]; sm[threadIdx.x][8 + j] = q[j]; }
  __syncthreads();
  if (threadIdx.x < 64) {  // (group, statistic): fixed summation order over rows then channels
    const int g = threadIdx.x >> 1, st = threadIdx.x & 1;
    const int gs = C / 32;
    float acc = 0.f;
    for (int t = 0; t < rpi; ++t)
      for (int c = g * gs; c < (g + 1) * gs; ++c) acc += sm[t * tpr + (c >> 3)][st * 8 + (c & 7)];
    partials[(((long long)b * gridDim.x + blockIdx.x) * 32 + g) * 2 + st] = acc;
  }
}

// per-block partials -> (mean, rstd) in float; one warp per (batch, group), fixed order, double accumulation
__global__ void groupnorm_finalize_kernel(const float* partials, float2* stats, int n_groups, int nblk, double n, float eps) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_groups) return;
  const int b = wid >> 5, g = wid & 31;
  double su = 0.0, sq = 0.0;
  for (int k = lane; k < nblk; k += 32) {
    const float* pp = partials + (((long long)b * nblk + k) * 32 + g) * 2;
    su += (double)pp[0];
    sq += (double)pp[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    su += __shfl_xor_sync(0xffffffffu, su, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if (lane == 0) {
    const double mean = su / n;
    const double var = fmax(sq / n - mean * mean, 0.0);
    stats[wid] = make_float2((float)mean, rsqrtf((float)var + eps));
  }
}

// apply: each thread owns ONE 8-channel slot and walks a slab of pixels: the per-channel affine (rstd * weight,
// bias - mean * rstd * weight) is folded once into registers, so the inner loop is one 16-byte load, 8 FMAs (+ SiLU)
// and one 16-byte store per pixel, four pixels in flight.  (The first version re-read statistics, weight and bias and
// divided by the group size for every element: 1.65 TB/s in the bench.)
constexpr int GN_APPLY_ROWS = 512;  // pixels per block
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const __nv_bfloat16* x, const float2* stats,
                                                              const __nv_bfloat16* weight, const __nv_bfloat16* bias,
                                                              __nv_bfloat16* out, long long hw, int C, int do_silu) {
  const int b = blockIdx.y;
  const int tpr = C / 8;                 // threads per pixel
  const int rpi = 256 / tpr;             // pixels per iteration of the block
  const int tr = threadIdx.x / tpr, c0 = (threadIdx.x % tpr) * 8;
  const int gs = C / 32;
  float A[8], Bc[8];
  {
    float w[8], bb[8];
    load8(weight + c0, w);
    load8(bias + c0, bb);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 st = __ldg(&stats[b * 32 + (c0 + j) / gs]);
      A[j] = st.y * w[j];
      Bc[j] = bb[j] - st.x * A[j];
    }
  }
  const long long r0 = (long long)blockIdx.x * GN_APPLY_ROWS;
  const long long r1 = min(hw, r0 + GN_APPLY_ROWS);
  const __nv_bfloat16* xb = x + (long long)b * hw * C + c0;
  __nv_bfloat16* ob = out + (long long)b * hw * C + c0;
  for (long long r = r0 + tr; r < r1; r += 4 * rpi) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long rr = r + (long long)k * rpi;
      if (rr < r1) u[k] = *reinterpret_cast<const uint4*>(xb + rr * C);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long rr = r + (long long)k * rpi;
      if (rr >= r1) break;
      float v[8], o[8];
      unpack8(u[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = fmaf(v[j], A[j], Bc[j]);
        if (do_silu) y = silu(__bfloat162float(__float2bfloat16(y)));
        o[j] = y;
      }
      store8(ob + rr * C, o);
    }
  }
}

// ------------------------------------------------------------------ nearest 2x upsample (NHWC)
__global__ void upsample2x_kernel(const uint4* x, uint4* out, int batch, int H, int W, int cv /* C/8 */) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)batch * (2 * H) * (2 * W) * cv;
  if (idx >= total) return;
  const int c = int(idx % cv);
  const long long pix = idx / cv;
  const int ox = int(pix % (2 * W)), oy = int((pix / (2 * W)) % (2 * H)), b = int(pix / ((long long)4 * W * H));
  out[idx] = x[(((long long)b * H + oy / 2) * W + ox / 2) * cv + c];
}

// ------------------------------------------------------------------ row softmax fp32 -> bf16
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* S, long long ld_s, __nv_bfloat16* P,
                                                           long long ld_p, int cols, float scale) {
  __shared__ float red[8];
  __shared__ float bcast;
  const float* row = S + (long long)blockIdx.x * ld_s;
  float mx = -INFINITY;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bcast = m;
  }
  __syncthreads();
  mx = bcast * scale;
  float sum = 0.f;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    sum += __expf(v.x * scale - mx) + __expf(v.y * scale - mx) + __expf(v.z * scale - mx) + __expf(v.w * scale - mx);
  }
  sum = warp_sum(sum);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    bcast = 1.0f / s;
  }
  __syncthreads();
  const float inv = bcast;
  __nv_bfloat16* prow = P + (long long)blockIdx.x * ld_p;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    uint2 u;
    u.x = pack_bf16(__expf(v.x * scale - mx) * inv, __expf(v.y * scale - mx) * inv);
    u.y = pack_bf16(__expf(v.z * scale - mx) * inv, __expf(v.w * scale - mx) * inv);
    *reinterpret_cast<uint2*>(prow + c) = u;
  }
}

// ------------------------------------------------------------------ bf16 transpose (32x32 smem tiles)
__global__ void transpose_kernel(const __nv_bfloat16* x, long long ldx, __nv_bfloat16* out, long long ldo, int rows,
                                 int cols) {
  __shared__ __nv_bfloat16 t[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = blockIdx.y * 32 + i;
    if (r < rows && c < cols) t[i][threadIdx.x] = x[(long long)r * ldx + c];
  }
  __syncthreads();
  const int orow_c = blockIdx.y * 32 + threadIdx.x;  // original row -> output column
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int ocol = blockIdx.x * 32 + i;  // original column -> output row
    if (ocol < cols && orow_c < rows) out[(long long)ocol * ldo + orow_c] = t[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------ image finish
__global__ void finish_image_kernel(const float* x, float* img, uint8_t* u8, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = fminf(fmaxf(x[i] + 1.0f, 0.0f), 2.0f) * 0.5f;
  if (img) img[i] = v;
  if (u8) u8[i] = (uint8_t)(v * 255.0f);  // truncation, like astype(uint8)
}

// ------------------------------------------------------------------ embedding gather
__global__ void embedding_kernel(const int32_t* ids, const uint4* table, const __nv_bfloat16* pos_table, uint4* out,
                                 long long n_ids, int seq, int dv /* D/8 */) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_ids * dv) return;
  const long long t = idx / dv;
  const int c = int(idx % dv);
  uint4 v = table[(long long)ids[t] * dv + c];
  if (pos_table) {
    float a[8], b[8];
    float2 f;
    f = unpack_bf16(v.x); a[0] = f.x; a[1] = f.y; f = unpack_bf16(v.y); a[2] = f.x; a[3] = f.y;
    f = unpack_bf16(v.z); a[4] = f.x; a[5] = f.y; f = unpack_bf16(v.w); a[6] = f.x; a[7] = f.y;
    load8(pos_table + ((long long)(t % seq) * dv + c) * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    v.x = pack_bf16(a[0], a[1]); v.y = pack_bf16(a[2], a[3]); v.z = pack_bf16(a[4], a[5]); v.w = pack_bf16(a[6], a[7]);
  }
  out[idx] = v;
}

__global__ void act_mul_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* out, long long n, int act) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float x[8], y[8];
    load8(a + i, x);
    load8(b + i, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __bfloat162float(__float2bfloat16(apply_act(x[j], act))) * y[j];
    store8(out + i, x);
  } else {
    for (long long k = i; k < n; ++k)
      out[k] = __float2bfloat16(__bfloat162float(__float2bfloat16(apply_act(__bfloat162float(a[k]), act))) *
                                __bfloat162float(b[k]));
  }
}

}  // namespace fx

using namespace fx;

extern "C" int fx_rownorm(const fx_rownorm_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->x && a->out && a->p0, "fx_rownorm: null pointer");
  FX_REQUIRE(a->mode >= 0 && a->mode <= 2, "fx_rownorm: bad mode %d", a->mode);
  FX_REQUIRE(a->mode == 2 || a->p1, "fx_rownorm: p1 required for mode %d", a->mode);
  FX_REQUIRE(a->D % 8 == 0 && a->D >= 8 && a->D <= 4096, "fx_rownorm: D (%d) must be a multiple of 8 in [8, 4096]", a->D);
  FX_REQUIRE(a->ldx % 8 == 0 && a->ldo % 8 == 0 && a->x_bs % 8 == 0 && a->out_bs % 8 == 0 && a->p_bs % 8 == 0,
             "fx_rownorm: strides must be multiples of 8 elements");
  if (a->batch <= 0 || a->rows <= 0) return FX_OK;
  FX_REQUIRE(a->out_fp8 != 1 || (a->scale_out && a->ldo % 16 == 0 && a->out_bs % 16 == 0),
             "fx_rownorm: out_fp8 needs scale_out and 16-byte aligned output rows");
  RowNormParams p{(const __nv_bfloat16*)a->x, a->ldx, a->x_bs, (__nv_bfloat16*)a->out, a->ldo, a->out_bs,
                  (const __nv_bfloat16*)a->p0, (const __nv_bfloat16*)a->p1, a->p_bs, a->eps, a->mode, a->batch, a->rows, a->D,
                  a->scale_out, a->scale_bs, (uint8_t*)a->sf_out};
  const long long rows = (long long)a->batch * a->rows;
  if (a->out_fp8 == 2)
    FX_REQUIRE(a->sf_out && a->scale_out && a->D % 64 == 0 && a->D >= 1024 && rows < (1ll << 31),
               "fx_rownorm: NVFP4 output needs sf_out, scale_out and D %% 64 == 0, D >= 1024");
  int blocks = int((rows + 7) / 8);
  {  // FX_ROWNORM_PERSIST = blocks per SM of a persistent grid-stride grid (0 = one row per warp).  Default 2 = the
     // kernel's occupancy: +5 % (3.69 -> 3.88 TB/s); more blocks than are resident is slower (3.0 TB/s)
    static int persist = -1;
    if (persist < 0) {
      const char* e = getenv("FX_ROWNORM_PERSIST");
      persist = e ? atoi(e) : 2;
    }
    if (persist > 0 && blocks > persist * num_sms()) blocks = persist * num_sms();
  }
  cudaStream_t st = (cudaStream_t)stream;
  static int block_rows = -1;  // FX_ROWNORM_BLOCK=0: the warp-per-row kernel for every D (A/B)
  if (block_rows < 0) {
    const char* e = getenv("FX_ROWNORM_BLOCK");
    block_rows = e ? atoi(e) : 1;
  }
  if ((block_rows || a->out_fp8 == 2) && a->D >= 1024 && rows < (1ll << 31)) {  // wide rows: one block per row
    static int per_sm = -1;  // FX_ROWNORM_BLOCKS_PER_SM: resident 128-thread blocks per SM of the persistent grid
    if (per_sm < 0) {
      const char* e = getenv("FX_ROWNORM_BLOCKS_PER_SM");
      per_sm = e ? atoi(e) : 4;  // rows in flight per SM scale the bandwidth (latency-bound): 4 / 5 / 6 resident blocks
      per_sm = per_sm < 5 ? 4 : (per_sm > 5 ? 6 : 5);
    }
    const long long cap = (long long)per_sm * num_sms();
    const unsigned g = (unsigned)(rows < cap ? rows : cap);
    const int vpt = (a->D + 1023) / 1024;
#define FX_RB(V, M)                                                                  \
  case V:                                                                            \
    if (a->out_fp8 == 2) rownorm_block_kernel<V, 2, M><<<g, 128, 0, st>>>(p);        \
    else if (a->out_fp8) rownorm_block_kernel<V, 1, M><<<g, 128, 0, st>>>(p);        \
    else rownorm_block_kernel<V, 0, M><<<g, 128, 0, st>>>(p);                        \
    break;
    if (per_sm == 4) { switch (vpt) { FX_RB(1, 4) FX_RB(2, 4) FX_RB(3, 4) FX_RB(4, 4) } }
    else if (per_sm == 5) { switch (vpt) { FX_RB(1, 5) FX_RB(2, 5) FX_RB(3, 5) FX_RB(4, 5) } }
    else { switch (vpt) { FX_RB(1, 6) FX_RB(2, 6) FX_RB(3, 6) FX_RB(4, 6) } }
#undef FX_RB
    return launched("rownorm_block_kernel");
  }
  if (a->out_fp8) {
    switch ((a->D + 255) / 256) {
#define FX_RN(I) case I: rownorm_kernel<I, true><<<blocks, 256, 0, st>>>(p); break;
      FX_RN(1) FX_RN(2) FX_RN(3) FX_RN(4) FX_RN(5) FX_RN(6) FX_RN(7) FX_RN(8)
      FX_RN(9) FX_RN(10) FX_RN(11) FX_RN(12) FX_RN(13) FX_RN(14) FX_RN(15) FX_RN(16)
#undef FX_RN
    }
    return launched("rownorm_kernel");
  }
  switch ((a->D + 255) / 256) {
#define FX_RN(I) case I: rownorm_kernel<I><<<blocks, 256, 0, st>>>(p); break;
    FX_RN(1) FX_RN(2) FX_RN(3) FX_RN(4) FX_RN(5) FX_RN(6) FX_RN(7) FX_RN(8)
    FX_RN(9) FX_RN(10) FX_RN(11) FX_RN(12) FX_RN(13) FX_RN(14) FX_RN(15) FX_RN(16)
#undef FX_RN
  }
  return launched("rownorm_kernel");
}

extern "C" int fx_quantize_rows(const fx_quant_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->x && a->q && a->scale, "fx_quantize_rows: null pointer");
  FX_REQUIRE(a->K > 0 && a->K % 8 == 0 && a->ldx % 8 == 0 && a->x_bs % 8 == 0 && a->ldq % 8 == 0 && a->q_bs % 8 == 0,
             "fx_quantize_rows: K and strides must be multiples of 8 elements");
  FX_REQUIRE(aligned16(a->x) && (reinterpret_cast<uintptr_t>(a->q) & 7) == 0, "fx_quantize_rows: unaligned pointers");
  if (a->batch <= 0 || a->rows <= 0) return FX_OK;
  QuantParams p{(const __nv_bfloat16*)a->x, a->ldx, a->x_bs, (uint8_t*)a->q, a->ldq, a->q_bs, a->scale, a->scale_bs,
                a->batch, a->rows, a->K};
  const long long rows = (long long)a->batch * a->rows;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->K > 2048 && a->K <= 16384 && rows < (1ll << 31)) {
    switch ((a->K + 2047) / 2048) {
#define FX_QB(I) case I: quantize_rows_block_kernel<I><<<(unsigned)rows, 256, 0, st>>>(p); break;
      FX_QB(2) FX_QB(3) FX_QB(4) FX_QB(5) FX_QB(6) FX_QB(7) FX_QB(8)
#undef FX_QB
    }
    return launched("quantize_rows_block_kernel");
  }
  quantize_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(p);
  return launched("quantize_rows_kernel");
}

extern "C" int fx_gemv(const fx_gemv_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->in && a->W && a->out, "fx_gemv: null pointer");
  FX_REQUIRE(a->K % 8 == 0 && a->ldw % 8 == 0 && a->K > 0 && a->N > 0 && a->batch > 0, "fx_gemv: K, ldw multiples of 8");
  FX_REQUIRE(a->K <= 8192, "fx_gemv: K too large");
  static bool done = false;
  if (!done) {
    FX_CUDA(cudaFuncSetAttribute(gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192 * 2));
    done = true;
  }
  for (int b0 = 0; b0 < a->batch; b0 += 8) {
    const int nb = a->batch - b0 < 8 ? a->batch - b0 : 8;
    GemvParams p{(const __nv_bfloat16*)a->in + b0 * a->ld_in, a->ld_in, (const __nv_bfloat16*)a->W, a->ldw,
                 (const __nv_bfloat16*)a->bias, a->add ? (const __nv_bfloat16*)a->add + b0 * a->ld_add : nullptr, a->ld_add,
                 (__nv_bfloat16*)a->out + b0 * a->ld_out, a->ld_out, nb, a->N, a->K, a->silu_in, a->silu_out};
    const int warps_needed = (a->N + 3) / 4;
    int blocks = (warps_needed + 7) / 8;
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    gemv_kernel<<<blocks, 256, (size_t)nb * a->K * 2, (cudaStream_t)stream>>>(p);
    int rc = launched("gemv_kernel");
    if (rc) return rc;
  }
  return FX_OK;
}

extern "C" int fx_timestep_embedding(const void* t, void* out, int32_t batch, int32_t dim, fx_stream stream) {
  FX_REQUIRE(t && out && batch > 0 && dim > 0 && dim % 2 == 0, "fx_timestep_embedding: bad arguments");
  const int n = batch * dim / 2;
  timestep_embedding_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)t, (__nv_bfloat16*)out, batch, dim);
  return launched("timestep_embedding_kernel");
}

extern "C" int fx_euler_step(void* x, const void* pred, float dt, int64_t n, fx_stream stream) {
  FX_REQUIRE(x && pred && n > 0, "fx_euler_step: bad arguments");
  FX_REQUIRE(aligned16(x) && aligned16(pred), "fx_euler_step: unaligned");
  const long long vecs = (n + 7) / 8;
  euler_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)x, (const __nv_bfloat16*)pred, dt, n);
  return launched("euler_kernel");
}

extern "C" int fx_patchify(const void* x, void* out, int32_t b, int32_t h, int32_t w, int32_t c, fx_stream stream) {
  FX_REQUIRE(x && out && b > 0 && h > 0 && w > 0 && c > 0, "fx_patchify: bad arguments");
  FX_REQUIRE(h % 2 == 0 && w % 2 == 0, "fx_patchify: latent size (%d, %d) must be even", h, w);
  const long long n = (long long)b * h * w * c;
  patchify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, b, h, w, c);
  return launched("patchify_kernel");
}

extern "C" int fx_prior_packed(void* out, int32_t b, int32_t h, int32_t w, int32_t c, uint64_t seed, int32_t first_index,
                               fx_stream stream) {
  FX_REQUIRE(out && b > 0 && h > 0 && w > 0 && c > 0, "fx_prior_packed: bad arguments");
  FX_REQUIRE(h % 2 == 0 && w % 2 == 0 && c % 4 == 0, "fx_prior_packed: latent size (%d, %d) must be even, channels %% 4 == 0", h, w);
  const long long n = (long long)b * h * w * c / 4;
  prior_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out, b, h, w, c, (uint32_t)seed,
                                                                          (uint32_t)(seed >> 32), first_index);
  return launched("prior_kernel");
}

extern "C" int fx_unpatchify_scale(const void* packed, void* z, int32_t b, int32_t h, int32_t w, int32_t c, int32_t c_pad,
                                   float scale_factor, float shift_factor, fx_stream stream) {
  FX_REQUIRE(packed && z && b > 0 && h > 0 && w > 0 && c > 0 && c_pad >= c, "fx_unpatchify_scale: bad arguments");
  FX_REQUIRE(h % 2 == 0 && w % 2 == 0, "fx_unpatchify_scale: latent size (%d, %d) must be even", h, w);
  const long long n = (long long)b * h * w * c_pad;
  unpatchify_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)packed, (__nv_bfloat16*)z, b, h, w, c, c_pad, 1.0f / scale_factor, shift_factor);
  return launched("unpatchify_scale_kernel");
}

extern "C" int64_t fx_groupnorm_partials_count(int32_t batch, int64_t hw) {
  return (int64_t)batch * ((hw + GN_ROWS_PER_BLOCK - 1) / GN_ROWS_PER_BLOCK) * 64;
}

extern "C" int fx_groupnorm_stats(const void* x, float* partials, int32_t batch, int64_t hw, int32_t C, fx_stream stream) {
  FX_REQUIRE(x && partials && batch > 0 && hw > 0, "fx_groupnorm_stats: bad arguments");
  FX_REQUIRE(C % 32 == 0 && C >= 64 && C <= 2048 && (C / 8) <= 256 && 256 % (C / 8) == 0,
             "fx_groupnorm_stats: unsupported channel count %d", C);
  dim3 grid((unsigned)((hw + GN_ROWS_PER_BLOCK - 1) / GN_ROWS_PER_BLOCK), batch);
  groupnorm_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, partials, hw, C);
  return launched("groupnorm_stats_kernel");
}

extern "C" int fx_groupnorm_finalize(const float* partials, float* stats, int32_t batch, int64_t hw, int32_t C, float eps,
                                     fx_stream stream) {
  FX_REQUIRE(partials && stats && batch > 0 && hw > 0 && C % 32 == 0, "fx_groupnorm_finalize: bad arguments");
  const int n = batch * 32;
  const int nblk = (int)((hw + GN_ROWS_PER_BLOCK - 1) / GN_ROWS_PER_BLOCK);
  groupnorm_finalize_kernel<<<(n * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(partials, (float2*)stats, n, nblk,
                                                                                 (double)hw * (C / 32), eps);
  return launched("groupnorm_finalize_kernel");
}

extern "C" int fx_groupnorm_finalize_blocks(const float* partials, float* stats, int32_t batch, int64_t nblk, int64_t hw,
                                            int32_t C, float eps, fx_stream stream) {
  FX_REQUIRE(partials && stats && batch > 0 && hw > 0 && nblk > 0 && nblk < (1ll << 31) && C % 32 == 0,
             "fx_groupnorm_finalize_blocks: bad arguments");
  const int n = batch * 32;
  groupnorm_finalize_kernel<<<(n * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(partials, (float2*)stats, n, (int)nblk,
                                                                                 (double)hw * (C / 32), eps);
  return launched("groupnorm_finalize_kernel");
}

extern "C" int fx_groupnorm_apply(const void* x, const float* stats, const void* weight, const void* bias, void* out,
                                  int32_t batch, int64_t hw, int32_t C, int32_t do_silu, fx_stream stream) {
  FX_REQUIRE(x && stats && weight && bias && out && batch > 0 && hw > 0 && C % 32 == 0 && C % 8 == 0, "fx_groupnorm_apply: bad arguments");
  FX_REQUIRE(C >= 64 && (C / 8) <= 256 && 256 % (C / 8) == 0, "fx_groupnorm_apply: unsupported channel count %d", C);
  dim3 grid((unsigned)((hw + GN_APPLY_ROWS - 1) / GN_APPLY_ROWS), batch);
  groupnorm_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const float2*)stats, (const __nv_bfloat16*)weight,
                                                              (const __nv_bfloat16*)bias, (__nv_bfloat16*)out, hw, C, do_silu);
  return launched("groupnorm_apply_kernel");
}

extern "C" int fx_upsample2x(const void* x, void* out, int32_t batch, int32_t H, int32_t W, int32_t C, fx_stream stream) {
  FX_REQUIRE(x && out && batch > 0 && H > 0 && W > 0 && C % 8 == 0, "fx_upsample2x: bad arguments");
  const long long n = (long long)batch * 4 * H * W * (C / 8);
  upsample2x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)out, batch, H, W, C / 8);
  return launched("upsample2x_kernel");
}

extern "C" int fx_softmax_rows(const float* S, int64_t ld_s, void* P, int64_t ld_p, int64_t rows, int32_t cols, float scale,
                               fx_stream stream) {
  FX_REQUIRE(S && P && rows > 0 && cols > 0 && cols % 4 == 0 && ld_s % 4 == 0 && ld_p % 4 == 0, "fx_softmax_rows: bad arguments");
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(S, ld_s, (__nv_bfloat16*)P, ld_p, cols, scale);
  return launched("softmax_rows_kernel");
}

extern "C" int fx_transpose(const void* x, int64_t ldx, void* out, int64_t ldo, int32_t rows, int32_t cols, fx_stream stream) {
  FX_REQUIRE(x && out && rows > 0 && cols > 0, "fx_transpose: bad arguments");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)out, ldo, rows, cols);
  return launched("transpose_kernel");
}

extern "C" int fx_finish_image(const float* x, float* img, uint8_t* u8, int64_t n, fx_stream stream) {
  FX_REQUIRE(x && n > 0 && (img || u8), "fx_finish_image: bad arguments");
  finish_image_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, img, u8, n);
  return launched("finish_image_kernel");
}

extern "C" int fx_embedding(const int32_t* ids, const void* table, const void* pos_table, void* out, int64_t n_ids,
                            int32_t seq, int32_t D, fx_stream stream) {
  FX_REQUIRE(ids && table && out && n_ids > 0 && D % 8 == 0 && seq > 0, "fx_embedding: bad arguments");
  const long long n = n_ids * (D / 8);
  embedding_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ids, (const uint4*)table, (const __nv_bfloat16*)pos_table,
                                                                               (uint4*)out, n_ids, seq, D / 8);
  return launched("embedding_kernel");
}

extern "C" int fx_act_mul(const void* a, const void* b, void* out, int64_t n, int32_t act, fx_stream stream) {
  FX_REQUIRE(a && b && out && n > 0, "fx_act_mul: bad arguments");
  const long long vecs = (n + 7) / 8;
  act_mul_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                                                (__nv_bfloat16*)out, n, act);
  return launched("act_mul_kernel");
}
