// Persistent warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A . W^T)
//   A  [batch][rows][K]   bf16, K contiguous (activations; or NHWC image read through a 4-D TMA box
//                         for the 3x3 convolution mode -- implicit GEMM, no im2col buffer)
//   W  [N][K]             bf16, K contiguous (nn.Linear layout / OHWI conv weights flattened)
// One CTA per SM; warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+TMEM owner), warps 4..11 =
// epilogue (TMEM -> registers -> global; two warps per TMEM lane quarter, each owning half the columns).  smem ring of 128B-swizzled K-major tiles filled by TMA;
// accumulators double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "sm100.cuh"

namespace fx {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int GEMM_THREADS = 384;  // warp 0 TMA, warp 1 MMA, (2,3 idle), warps 4..11 epilogue (two per TMEM lane quarter)

enum : int { EPI_GENERIC = 0, EPI_QKV = 1 };

struct GemmParams {
  int batch, rows, N, K;
  int tiles_m_per_batch, tiles_m, tiles_n, num_tiles, k_blocks, group_m;
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
  const uint32_t* pe;  // [seq_total][64] (cos, sin) bf16 pairs
  __nv_bfloat16 *q, *k, *v;  // [batch][heads][seq_total][128]
  // ---- 3x3 conv mode (A is [batch][H][W][C] NHWC, pad 1, stride 1)
  int conv_H, conv_W, conv_tiles_x, conv_tiles_y, cin_blocks;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void gemm_tile_coords(const GemmParams& p, int tile, int& tm, int& tn) {
  const int per_group = p.group_m * p.tiles_n;
  const int g = tile / per_group;
  const int first_m = g * p.group_m;
  const int gsize = min(p.tiles_m - first_m, p.group_m);
  const int rem = tile - g * per_group;
  tm = first_m + rem % gsize;
  tn = rem / gsize;
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

// One 32-column chunk of the generic epilogue for one accumulator row.
__device__ __forceinline__ void epi_generic_chunk(const GemmParams& p, float* f, int b, long long out_off,
                                                  long long res_off, int n0, bool fast) {
  if (fast) {
    if (p.bias) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float t[8];
        ld_bf16x8(p.bias + n0 + i, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[i + j] += t[j];
      }
    }
    if (p.act != 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = apply_act(f[i], p.act);
    }
    if (p.gate) {
      const __nv_bfloat16* g = p.gate + (long long)b * p.gate_bs + n0;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float t[8];
        ld_bf16x8(g + i, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[i + j] *= t[j];
      }
    }
    if (p.resid) {
      const __nv_bfloat16* r = p.resid + res_off + n0;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float t[8];
        const uint4 u = *reinterpret_cast<const uint4*>(r + i);  // plain load: may alias `out`
        float2 a = unpack_bf16(u.x), bb = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        t[0] = a.x; t[1] = a.y; t[2] = bb.x; t[3] = bb.y; t[4] = c.x; t[5] = c.y; t[6] = d.x; t[7] = d.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) f[i + j] += t[j];
      }
    }
    if (p.out_f32) {
      float* o = reinterpret_cast<float*>(p.out) + out_off + n0;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_off + n0;
#pragma unroll
      for (int i = 0; i < 32; i += 8) st_bf16x8(o + i, f + i);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) {  // static indices keep f[] in registers
      const int n = n0 + i;
      if (n >= p.N) continue;
      float v = f[i];
      if (p.bias) v += __bfloat162float(p.bias[n]);
      if (p.act != 0) v = apply_act(v, p.act);
      if (p.gate) v *= __bfloat162float(p.gate[(long long)b * p.gate_bs + n]);
      if (p.resid) v += __bfloat162float(p.resid[res_off + n]);
      if (p.out_f32) reinterpret_cast<float*>(p.out)[out_off + n] = v;
      else reinterpret_cast<__nv_bfloat16*>(p.out)[out_off + n] = __float2bfloat16(v);
    }
  }
}

template <int BN, int EPI, bool CONV>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
            const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The register file is partitioned per SM sub-partition (16K registers each, 3 warps here): the QKV
  // epilogue keeps a whole 128-column head in registers, so warpgroup 0 hands registers to the epilogue.
  if (warp < 4) {
  if (EPI == EPI_QKV) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int tm, tn;
        gemm_tile_coords(p, tile, tm, tn);
        const int b = tm / p.tiles_m_per_batch;
        const int tmb = tm - b * p.tiles_m_per_batch;
        int cy = 0, cx = 0;
        if (CONV) {
          cy = (tmb / p.conv_tiles_x) * 8;
          cx = (tmb % p.conv_tiles_x) * 16;
        }
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (CONV) {
            const int tap = kb / p.cin_blocks;
            const int c0 = (kb - tap * p.cin_blocks) * GEMM_BK;
            tma_load_4d(sa, &tmap_a, &full_bar[stage], c0, cx + tap % 3 - 1, cy + tap / 3 - 1, b);
          } else {
            tma_load_3d(sa, &tmap_a, &full_bar[stage], kb * GEMM_BK, tmb * GEMM_BM, b);
          }
          tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * GEMM_BK, tn * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * 32, 16, 1024);
            umma_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }
  } else {
    // ================= epilogue warps =================
    if (EPI == EPI_QKV) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;    // which half of the tile's columns this warp owns
    const int r = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int tm, tn;
      gemm_tile_coords(p, tile, tm, tn);
      const int b = tm / p.tiles_m_per_batch;
      const int tmb = tm - b * p.tiles_m_per_batch;
      bool valid;
      long long row;  // row index inside the batch (pixel index for conv)
      if (CONV) {
        const int y = (tmb / p.conv_tiles_x) * 8 + (r >> 4);
        const int x = (tmb % p.conv_tiles_x) * 16 + (r & 15);
        valid = (y < p.conv_H) && (x < p.conv_W);
        row = (long long)y * p.conv_W + x;
      } else {
        row = (long long)tmb * GEMM_BM + r;
        valid = row < p.rows;
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (uint32_t(quarter * 32) << 16);

      if (EPI == EPI_GENERIC) {
        const long long out_off = (long long)b * p.out_bs + row * p.ldo;
        const long long res_off = (long long)b * p.resid_bs + row * p.ldr;
        const bool vec_ok = ((p.ldo | p.ldr | p.gate_bs) & 7) == 0;
        constexpr int CH = BN / 64;  // 32-column chunks per warp
#pragma unroll 1
        for (int c = half * CH; c < (half + 1) * CH; ++c) {
          const int n0 = tn * BN + c * 32;
          if (n0 >= p.N) break;
          uint32_t v[32];
          __syncwarp();
          tmem_ld_x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (valid) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
            epi_generic_chunk(p, f, b, out_off, res_off, n0, vec_ok && (n0 + 32 <= p.N));
          }
        }
      } else {
        // ---- QKV(+MLP) epilogue: this warp owns one 128-column group; the whole group sits in registers
        const int D3 = 3 * p.heads * 128;
        const long long pos = (long long)p.seq_off + row;
        const int g0 = tn * BN + half * 128;
        if (g0 < p.N) {
          uint32_t v[128];
          __syncwarp();
          tmem_ld_x32(taddr + half * 128, v);
          tmem_ld_x32(taddr + half * 128 + 32, v + 32);
          tmem_ld_x32(taddr + half * 128 + 64, v + 64);
          tmem_ld_x32(taddr + half * 128 + 96, v + 96);
          tmem_ld_wait();
          if (valid) {
            if (g0 >= D3) {
              // mlp region -> +bias, gelu -> `out` at column (g0 - 3D)
              const long long out_off = (long long)b * p.out_bs + pos * p.ldo - D3;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[c * 32 + i]);
                epi_generic_chunk(p, f, b, out_off, 0, g0 + c * 32, true);
              }
            } else {
              const int hidx = g0 >> 7;
              const int which = hidx / p.heads;  // 0 q, 1 k, 2 v
              const int head = hidx - which * p.heads;
              __nv_bfloat16* dst = (which == 0 ? p.q : (which == 1 ? p.k : p.v)) +
                                   (((long long)b * p.heads + head) * p.seq_total + pos) * 128;
              float ss = 0.f;
              if (p.bias) {
#pragma unroll
                for (int i = 0; i < 128; i += 8) {
                  float t[8];
                  ld_bf16x8(p.bias + g0 + i, t);
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float x = __uint_as_float(v[i + j]) + t[j];
                    v[i + j] = __float_as_uint(x);
                    ss += x * x;
                  }
                }
              } else {
#pragma unroll
                for (int i = 0; i < 128; ++i) ss += __uint_as_float(v[i]) * __uint_as_float(v[i]);
              }
              if (which < 2) {
                const float rr = rsqrtf(ss * (1.0f / 128.0f) + p.rms_eps);
                const __nv_bfloat16* nw = which == 0 ? p.qnorm_w : p.knorm_w;
                const uint4* pe4 = reinterpret_cast<const uint4*>(p.pe + pos * 64);
#pragma unroll
                for (int i = 0; i < 128; i += 8) {
                  float t[8], f[8];
                  ld_bf16x8(nw + i, t);
                  const uint4 u = __ldg(pe4 + (i >> 3));
                  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 cs = unpack_bf16(w[j]);  // (cos, sin)
                    const float x0 = __uint_as_float(v[i + 2 * j]) * rr * t[2 * j];
                    const float x1 = __uint_as_float(v[i + 2 * j + 1]) * rr * t[2 * j + 1];
                    f[2 * j] = x0 * cs.x - x1 * cs.y;
                    f[2 * j + 1] = x0 * cs.y + x1 * cs.x;
                  }
                  st_bf16x8(dst + i, f);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 128; i += 8) {
                  float f[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[i + j]);
                  st_bf16x8(dst + i, f);
                }
              }
            }
          }
        }
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace fx
