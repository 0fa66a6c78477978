// Bring-up probe (tests only, never on the product path): one CTA issues K/16 tcgen05.mma on operands
// it lays out in shared memory by hand, with the descriptor fields supplied by the caller.  Lets a
// single GPU run sweep UMMA descriptor encodings (MN-major B, A-from-TMEM) against a CPU matmul.
#include "../../include/flux_b200_dbg.h"
#include "api_common.cuh"
#include "sm100.cuh"

namespace fx {

struct DbgUmmaParams {
  const __nv_bfloat16* A;  // [128][K]
  const __nv_bfloat16* B;  // K-major: [N][K]; MN-major: [K][N]
  float* D;                // [128][N]
  int K, N, b_mn_major, a_tmem;
  uint32_t lbo, sbo, kstep_bytes;
};

__global__ void __launch_bounds__(128, 1) dbg_umma_kernel(const DbgUmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sa = smem;                          // K/64 halves x [128 rows x 128 B]
  uint8_t* sb = smem + (p.K / 64) * 16384;     // see below
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;

  // A: K-major SW128
  for (int idx = tid; idx < 128 * (p.K / 8); idx += 128) {
    const int r = idx / (p.K / 8), ch = idx % (p.K / 8);
    const int half = ch / 8, c = ch % 8;
    const uint4 v = *reinterpret_cast<const uint4*>(p.A + (long long)r * p.K + ch * 8);
    *reinterpret_cast<uint4*>(sa + half * 16384 + r * 128 + ((c ^ (r & 7)) * 16)) = v;
  }
  if (p.b_mn_major) {
    // B[K][N]: per 64-wide N block: [K rows x 128 B], chunk ^= (k & 7)
    for (int idx = tid; idx < p.K * (p.N / 8); idx += 128) {
      const int k = idx / (p.N / 8), ch = idx % (p.N / 8);
      const int nb = ch / 8, c = ch % 8;
      const uint4 v = *reinterpret_cast<const uint4*>(p.B + (long long)k * p.N + ch * 8);
      *reinterpret_cast<uint4*>(sb + nb * (p.K * 128) + k * 128 + ((c ^ (k & 7)) * 16)) = v;
    }
  } else {
    for (int idx = tid; idx < p.N * (p.K / 8); idx += 128) {
      const int n = idx / (p.K / 8), ch = idx % (p.K / 8);
      const int half = ch / 8, c = ch % 8;
      const uint4 v = *reinterpret_cast<const uint4*>(p.B + (long long)n * p.K + ch * 8);
      *reinterpret_cast<uint4*>(sb + half * (p.N * 128) + n * 128 + ((c ^ (n & 7)) * 16)) = v;
    }
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  if (p.a_tmem) {
    // row tid -> TMEM lane tid, columns 128 + k/2 (packed bf16 pairs)
    for (int c = 0; c < p.K / 64; ++c) {
      uint32_t v[32];
#pragma unroll
      for (int e = 0; e < 32; ++e)
        v[e] = *reinterpret_cast<const uint32_t*>(p.A + (long long)tid * p.K + c * 64 + e * 2);
      tmem_st_x32(tmem + lane_base + 128 + c * 32, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, p.N, 0, p.b_mn_major);
    const uint32_t a_base = smem_u32(sa), b_base = smem_u32(sb);
    for (int ks = 0; ks < p.K / 16; ++ks) {
      uint64_t bd;
      if (p.b_mn_major) bd = make_smem_desc_sw128(b_base + ks * p.kstep_bytes, p.lbo, p.sbo);
      else bd = make_smem_desc_sw128(b_base + (ks >> 2) * (p.N * 128) + (ks & 3) * 32, 16, 1024);
      if (p.a_tmem) umma_ts(tmem, tmem + 128 + ks * 8, bd, idesc, ks != 0);
      else umma_ss(tmem, make_smem_desc_sw128(a_base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), bd, idesc, ks != 0);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < p.N / 32; ++c) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(tmem + lane_base + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e) p.D[(long long)tid * p.N + c * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}


// ------------------------------------------------------------------------------------------
// Byte-operand probe: one CTA, 8-bit / 4-bit operands laid out by hand, one of
//   kind 0  tcgen05.mma.kind::f8f6f4                         e4m3 x e4m3 (A from shared memory or from TMEM; B K- or MN-major)
//   kind 1  tcgen05.mma.kind::mxf8f6f4.block_scale           e4m3, one UE8M0 scale per 32 elements of K
//   kind 2  tcgen05.mma.kind::mxf4nvf4.block_scale.block16   e2m1 (two per byte), one UE4M3 scale per 16 elements (NVFP4)
//   kind 3  tcgen05.mma.kind::mxf4nvf4.block_scale.block32   e2m1, one UE8M0 scale per 32 elements (MXFP4)
// Every MMA consumes 32 bytes of K per operand row.  Scale factors are staged the way the block-scaled GEMM stages
// them: 512-byte atoms of 128 rows x 4 scales in shared memory, byte (r % 32) * 16 + (r / 32) * 4 + s, copied with
// tcgen05.cp.32x128b.warpx4 into 4 TMEM columns (lane r % 32 of every sub-partition, column r / 32, byte s).
// ------------------------------------------------------------------------------------------
struct DbgBsParams {
  const uint8_t *A, *B, *SFA, *SFB;  // A [128][kbytes]; B [N][kbytes] (or, b_mn_major, [K][N]); SF [rows][nsf]
  float* D;                          // [128][N]
  int N, kbytes, kind, a_tmem, b_mn_major, nsf;
  uint32_t b_lbo, b_sbo, b_kstep, cp_lbo, cp_sbo;
};
__device__ __forceinline__ uint64_t make_smem_desc_plain(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;  // no swizzle (layout type 0), descriptor version 1
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  return d;
}
__device__ __forceinline__ void tc_cp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// block-scaled instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled bit layout)
__device__ __forceinline__ uint32_t make_idesc_bs(int M, int N, int ab_fmt, int scale_fmt, int b_major, int a_sf_id, int b_sf_id) {
  return (uint32_t(b_sf_id) << 4) | (uint32_t(ab_fmt) << 7) | (uint32_t(ab_fmt) << 10) | (uint32_t(b_major) << 16) |
         (uint32_t(N >> 3) << 17) | (uint32_t(scale_fmt) << 23) | (uint32_t(M >> 4) << 24) | (uint32_t(a_sf_id) << 29);
}
template <int KIND>
__device__ __forceinline__ void umma_bs(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc, uint32_t sfa, uint32_t sfb) {
  if (KIND == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d), "l"(ad), "l"(bd),
                 "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
  else if (KIND == 2)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf4nvf4.block_scale.block16 [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d), "l"(ad),
                 "l"(bd), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf4nvf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d), "l"(ad),
                 "l"(bd), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ void umma_ts_f8(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc)
               : "memory");
}

__global__ void __launch_bounds__(128, 1) dbg_bs_kernel(const DbgBsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  const int atoms = p.kbytes / 128;                 // 128-byte K atoms per row
  uint8_t* sa = smem;                               // atoms x [128 rows x 128 B]
  uint8_t* sb = sa + atoms * 16384;                 // K-major: atoms x [N rows x 128 B]; MN-major: [K rows x 128 B]
  uint8_t* ssfa = sb + (p.b_mn_major ? p.kbytes * 128 : atoms * p.N * 128);
  const int g4n = (p.nsf + 3) / 4, nrb = (p.N + 127) / 128;
  uint8_t* ssfb = ssfa + g4n * 512;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int idx = tid; idx < 128 * (p.kbytes / 16); idx += 128) {  // A: K-major SW128
    const int r = idx / (p.kbytes / 16), ch = idx % (p.kbytes / 16);
    const int at = ch / 8, c = ch % 8;
    *reinterpret_cast<uint4*>(sa + at * 16384 + r * 128 + ((c ^ (r & 7)) * 16)) =
        *reinterpret_cast<const uint4*>(p.A + (long long)r * p.kbytes + ch * 16);
  }
  if (p.b_mn_major) {  // B [K][N = 128] bytes: one MN atom of 128 B per k-row, chunk ^= (k & 7)
    for (int idx = tid; idx < p.kbytes * 8; idx += 128) {
      const int k = idx / 8, c = idx % 8;
      *reinterpret_cast<uint4*>(sb + k * 128 + ((c ^ (k & 7)) * 16)) = *reinterpret_cast<const uint4*>(p.B + (long long)k * p.N + c * 16);
    }
  } else {
    for (int idx = tid; idx < p.N * (p.kbytes / 16); idx += 128) {
      const int n = idx / (p.kbytes / 16), ch = idx % (p.kbytes / 16);
      const int at = ch / 8, c = ch % 8;
      *reinterpret_cast<uint4*>(sb + at * (p.N * 128) + n * 128 + ((c ^ (n & 7)) * 16)) =
          *reinterpret_cast<const uint4*>(p.B + (long long)n * p.kbytes + ch * 16);
    }
  }
  if (p.kind != 0) {  // scale-factor atoms
    for (int idx = tid; idx < 128 * p.nsf; idx += 128) {
      const int r = idx / p.nsf, s = idx % p.nsf;
      ssfa[(s / 4) * 512 + (r % 32) * 16 + (r / 32) * 4 + (s % 4)] = p.SFA[(long long)r * p.nsf + s];
    }
    for (int idx = tid; idx < nrb * 128 * p.nsf; idx += 128) {
      const int n = idx / p.nsf, s = idx % p.nsf;
      const int rb = n / 128, r = n % 128;
      ssfb[((s / 4) * nrb + rb) * 512 + (r % 32) * 16 + (r / 32) * 4 + (s % 4)] = n < p.N ? p.SFB[(long long)n * p.nsf + s] : 0;
    }
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  constexpr uint32_t A_COL = 256, SFA_COL = 384, SFB_COL = 448;
  if (p.a_tmem) {  // row tid -> TMEM lane tid, byte k at column k / 4
    for (int c = 0; c < p.kbytes / 128; ++c) {
      uint32_t v[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = *reinterpret_cast<const uint32_t*>(p.A + (long long)tid * p.kbytes + c * 128 + e * 4);
      tmem_st_x32(tmem + lane_base + A_COL + c * 32, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t a_base = smem_u32(sa), b_base = smem_u32(sb);
    if (p.kind != 0) {
      for (int g = 0; g < g4n; ++g) {
        tc_cp_32x128b_warpx4(tmem + SFA_COL + g * 4, make_smem_desc_plain(smem_u32(ssfa + g * 512), p.cp_lbo, p.cp_sbo));
        for (int rb = 0; rb < nrb; ++rb)
          tc_cp_32x128b_warpx4(tmem + SFB_COL + (g * nrb + rb) * 4, make_smem_desc_plain(smem_u32(ssfb + (g * nrb + rb) * 512), p.cp_lbo, p.cp_sbo));
      }
    }
    for (int ks = 0; ks < p.kbytes / 32; ++ks) {
      const uint64_t ad = make_smem_desc_sw128(a_base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
      uint64_t bd;
      if (p.b_mn_major) bd = make_smem_desc_sw128(b_base + ks * p.b_kstep, p.b_lbo, p.b_sbo);
      else bd = make_smem_desc_sw128(b_base + (ks >> 2) * (p.N * 128) + (ks & 3) * 32, 16, 1024);
      const uint32_t acc = ks != 0;
      if (p.kind == 0) {
        const uint32_t idesc = make_idesc_f8(128, p.N) | (uint32_t(p.b_mn_major) << 16);
        if (p.a_tmem) umma_ts_f8(tmem, tmem + A_COL + ks * 8, bd, idesc, acc);
        else umma_ss_f8(tmem, ad, bd, idesc, acc);
      } else if (p.kind == 1) {  // one scale per MMA: byte ks % 4 of the group's columns
        umma_bs<1>(tmem, ad, bd, make_idesc_bs(128, p.N, 0, 1, 0, ks & 3, ks & 3), acc, tmem + SFA_COL + (ks >> 2) * 4,
                   tmem + SFB_COL + (ks >> 2) * nrb * 4);
      } else if (p.kind == 2) {  // four scales per MMA: the whole 32-bit column
        umma_bs<2>(tmem, ad, bd, make_idesc_bs(128, p.N, 1, 0, 0, 0, 0), acc, tmem + SFA_COL + ks * 4, tmem + SFB_COL + ks * nrb * 4);
      } else {                   // two scales per MMA: bytes 0-1 or 2-3
        umma_bs<3>(tmem, ad, bd, make_idesc_bs(128, p.N, 1, 1, 0, (ks & 1) * 2, (ks & 1) * 2), acc, tmem + SFA_COL + (ks >> 1) * 4,
                   tmem + SFB_COL + (ks >> 1) * nrb * 4);
      }
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < p.N / 32; ++c) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(tmem + lane_base + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e) p.D[(long long)tid * p.N + c * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------
// The same probe for a CTA PAIR (cta_group::2): M = 256 (128 rows per CTA), the W tile split over the two CTAs (N / 2 rows
// each), NVFP4 (kind 2).  Question it answers: where do tcgen05.cp.cta_group::2 and the block-scaled MMA of a pair take
// their scale factors from?  sfb_mode 0: every CTA stages the scale atoms of ALL N columns in its own shared memory and the
// leader's tcgen05.cp.cta_group::2 copies, in each CTA, from that CTA's shared memory into that CTA's TMEM (hypothesis);
// sfb_mode 1: every CTA stages only the atoms of its own half of the W rows.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_cp2_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma2_nvf4(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc, uint32_t sfa, uint32_t sfb) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::mxf4nvf4.block_scale.block16 [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d), "l"(ad),
               "l"(bd), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
}
struct DbgBs2Params {
  const uint8_t *A, *B, *SFA, *SFB;  // A [256][kbytes]; B [N][kbytes]; SFA [256][nsf]; SFB [N][nsf]
  float* D;                          // [256][N]
  int N, kbytes, nsf, sfb_mode;
};
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) dbg_bs2_kernel(const DbgBs2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  const uint32_t rank = cluster_ctarank();
  const int atoms = p.kbytes / 128, NH = p.N / 2;
  uint8_t* sa = smem;                                   // atoms x [128 rows x 128 B]: rows 128 * rank ...
  uint8_t* sb = sa + atoms * 16384;                     // atoms x [NH rows x 128 B]: W rows NH * rank ...
  uint8_t* ssfa = sb + atoms * NH * 128;
  const int g4n = (p.nsf + 3) / 4, nrb = (p.N + 127) / 128;
  uint8_t* ssfb = ssfa + g4n * 512;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int idx = tid; idx < 128 * (p.kbytes / 16); idx += 128) {
    const int r = idx / (p.kbytes / 16), ch = idx % (p.kbytes / 16);
    const int at = ch / 8, c = ch % 8;
    *reinterpret_cast<uint4*>(sa + at * 16384 + r * 128 + ((c ^ (r & 7)) * 16)) =
        *reinterpret_cast<const uint4*>(p.A + (long long)(rank * 128 + r) * p.kbytes + ch * 16);
  }
  for (int idx = tid; idx < NH * (p.kbytes / 16); idx += 128) {
    const int n = idx / (p.kbytes / 16), ch = idx % (p.kbytes / 16);
    const int at = ch / 8, c = ch % 8;
    *reinterpret_cast<uint4*>(sb + at * (NH * 128) + n * 128 + ((c ^ (n & 7)) * 16)) =
        *reinterpret_cast<const uint4*>(p.B + (long long)(rank * NH + n) * p.kbytes + ch * 16);
  }
  for (int idx = tid; idx < 128 * p.nsf; idx += 128) {
    const int r = idx / p.nsf, s = idx % p.nsf;
    ssfa[(s / 4) * 512 + (r % 32) * 16 + (r / 32) * 4 + (s % 4)] = p.SFA[(long long)(rank * 128 + r) * p.nsf + s];
  }
  for (int idx = tid; idx < nrb * 128 * p.nsf; idx += 128) {
    const int n = idx / p.nsf, s = idx % p.nsf;   // n: slot row of the staged SFB set
    const int rb = n / 128, r = n % 128;
    int src = n;                                   // mode 0: all N columns, in order
    if (p.sfb_mode == 1) src = (n < NH) ? int(rank) * NH + n : -1;  // mode 1: only this CTA's half, packed at the front
    ssfb[((s / 4) * nrb + rb) * 512 + (r % 32) * 16 + (r / 32) * 4 + (s % 4)] = (src >= 0 && src < p.N) ? p.SFB[(long long)src * p.nsf + s] : 0;
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc2(&tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  constexpr uint32_t SFA_COL = 384, SFB_COL = 448;
  if (rank == 0 && tid == 0) {
    const uint32_t a_base = smem_u32(sa), b_base = smem_u32(sb);
    for (int g = 0; g < g4n; ++g) {
      tc_cp2_32x128b_warpx4(tmem + SFA_COL + g * 4, make_smem_desc_plain(smem_u32(ssfa + g * 512), 0, 128));
      for (int rb = 0; rb < nrb; ++rb)
        tc_cp2_32x128b_warpx4(tmem + SFB_COL + (g * nrb + rb) * 4, make_smem_desc_plain(smem_u32(ssfb + (g * nrb + rb) * 512), 0, 128));
    }
    for (int ks = 0; ks < p.kbytes / 32; ++ks) {
      const uint64_t ad = make_smem_desc_sw128(a_base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
      const uint64_t bd = make_smem_desc_sw128(b_base + (ks >> 2) * (NH * 128) + (ks & 3) * 32, 16, 1024);
      umma2_nvf4(tmem, ad, bd, make_idesc_bs(256, p.N, 1, 0, 0, 0, 0), ks != 0, tmem + SFA_COL + ks * 4, tmem + SFB_COL + ks * nrb * 4);
    }
    tc_commit2(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < p.N / 32; ++c) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(tmem + lane_base + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e) p.D[(long long)(rank * 128 + tid) * p.N + c * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc2(tmem, 512);
  }
}

}  // namespace fx

extern "C" int fx_dbg_bs_tile(const void* A, const void* B, const void* SFA, const void* SFB, float* D, int32_t N, int32_t kbytes,
                              int32_t kind, int32_t a_tmem, int32_t b_mn_major, int32_t nsf, uint32_t b_lbo, uint32_t b_sbo,
                              uint32_t b_kstep, uint32_t cp_lbo, uint32_t cp_sbo, fx_stream stream) {
  FX_REQUIRE(A && B && D && N % 32 == 0 && N >= 32 && N <= 256 && kbytes % 128 == 0 && kbytes >= 128 && kbytes <= 512 && kind >= 0 && kind <= 3,
             "fx_dbg_bs_tile: bad arguments");
  FX_REQUIRE(kind == 0 || (SFA && SFB && nsf > 0), "fx_dbg_bs_tile: block-scaled kinds need scale factors");
  FX_REQUIRE(!b_mn_major || N == 128, "fx_dbg_bs_tile: the MN-major probe is for N = 128");
  fx::DbgBsParams p{(const uint8_t*)A, (const uint8_t*)B, (const uint8_t*)SFA, (const uint8_t*)SFB, D, N, kbytes, kind, a_tmem, b_mn_major,
                    nsf, b_lbo, b_sbo, b_kstep, cp_lbo, cp_sbo};
  const int atoms = kbytes / 128;
  const int smem = atoms * 16384 + (b_mn_major ? kbytes * 128 : atoms * N * 128) + ((nsf + 3) / 4) * 512 * (1 + (N + 127) / 128) + 2048;
  static bool attr = false;
  if (!attr) {
    FX_CUDA(cudaFuncSetAttribute(fx::dbg_bs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  fx::dbg_bs_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  return fx::launched("dbg_bs_kernel");
}

extern "C" int fx_dbg_bs2_tile(const void* A, const void* B, const void* SFA, const void* SFB, float* D, int32_t N, int32_t kbytes,
                               int32_t nsf, int32_t sfb_mode, fx_stream stream) {
  FX_REQUIRE(A && B && SFA && SFB && D && N % 64 == 0 && N >= 64 && N <= 256 && kbytes % 128 == 0 && kbytes >= 128 && kbytes <= 512 && nsf > 0,
             "fx_dbg_bs2_tile: bad arguments");
  fx::DbgBs2Params p{(const uint8_t*)A, (const uint8_t*)B, (const uint8_t*)SFA, (const uint8_t*)SFB, D, N, kbytes, nsf, sfb_mode};
  const int atoms = kbytes / 128;
  const int smem = atoms * 16384 + atoms * (N / 2) * 128 + ((nsf + 3) / 4) * 512 * (1 + (N + 127) / 128) + 2048;
  static bool attr = false;
  if (!attr) {
    FX_CUDA(cudaFuncSetAttribute(fx::dbg_bs2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  fx::dbg_bs2_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(p);
  return fx::launched("dbg_bs2_kernel");
}

namespace fx {

// ------------------------------------------------------------------------------------------
// Tensor-pipe pattern micro-benchmark (profiling only): one thread per CTA issues a fixed sequence of
// 128x128x16 tcgen05.mma -- the attention kernel's QK (A, B from shared memory) and PV (A from TMEM,
// B MN-major) instructions on the attention kernel's own shared-memory / TMEM layout, without any
// producer / consumer -- and reports SM clocks per "step" (32 MMAs = one key tile for two query tiles).
// ------------------------------------------------------------------------------------------
constexpr int PAT_TILE = 128 * 128 * 2;
constexpr int PAT_SMEM = 7 * PAT_TILE + 1024 + 64;

struct PatCtx {
  uint32_t tmem, q_base, kv_base;
  uint64_t* scratch;
};
__device__ __forceinline__ void pat_qk(const PatCtx& c, int i, int stage) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
  const uint32_t k_addr = c.kv_base + stage * PAT_TILE;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
    umma_ss(c.tmem + i * 128, make_smem_desc_sw128(c.q_base + i * PAT_TILE + off, 16, 1024),
            make_smem_desc_sw128(k_addr + off, 16, 1024), idesc, ks != 0);
  }
}
__device__ __forceinline__ void pat_pv(const PatCtx& c, int i, int stage, int hf) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 1);
  const uint32_t v_addr = c.kv_base + stage * PAT_TILE;
#pragma unroll
  for (int ks = hf * 4; ks < hf * 4 + 4; ++ks)
    umma_ts(c.tmem + 256 + i * 128, c.tmem + i * 128 + ks * 8, make_smem_desc_sw128(v_addr + ks * 2048, 16384, 1024), idesc, 1u);
}
__device__ __forceinline__ void pat_qk256(const PatCtx& c, int stage) {  // N = 256 accumulator (columns 0..255)
  constexpr uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
  const uint32_t k_addr = c.kv_base + (stage % 3) * PAT_TILE;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    umma_ss(c.tmem, make_smem_desc_sw128(c.q_base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
            make_smem_desc_sw128(k_addr + (ks >> 2) * 32768 + (ks & 3) * 32, 16, 1024), idesc, ks != 0);
}

template <int PAT>
__device__ __forceinline__ void pat_step(const PatCtx& c, int ks_, int vs, int role) {
  if (PAT == 0) {  // the attention kernel's order and commits
    pat_pv(c, 0, vs, 0); pat_pv(c, 0, vs, 1); pat_qk(c, 0, ks_); tc_commit(c.scratch);
    pat_pv(c, 1, vs, 0); pat_pv(c, 1, vs, 1); tc_commit(c.scratch + 1);
    pat_qk(c, 1, ks_); tc_commit(c.scratch + 2); tc_commit(c.scratch + 3);
  } else if (PAT == 1) {  // same order, no commits
    pat_pv(c, 0, vs, 0); pat_pv(c, 0, vs, 1); pat_qk(c, 0, ks_);
    pat_pv(c, 1, vs, 0); pat_pv(c, 1, vs, 1); pat_qk(c, 1, ks_);
  } else if (PAT == 2) {  // 32 x SS into one accumulator
    pat_qk(c, 0, ks_); pat_qk(c, 0, vs); pat_qk(c, 0, ks_); pat_qk(c, 0, vs);
  } else if (PAT == 3) {  // 32 x TS into one accumulator
#pragma unroll
    for (int r = 0; r < 4; ++r) { pat_pv(c, 0, r & 1 ? vs : ks_, 0); pat_pv(c, 0, r & 1 ? vs : ks_, 1); }
  } else if (PAT == 4) {  // SS, alternating accumulators
    pat_qk(c, 0, ks_); pat_qk(c, 1, ks_); pat_qk(c, 0, vs); pat_qk(c, 1, vs);
  } else if (PAT == 5) {  // TS, alternating accumulators
    pat_pv(c, 0, vs, 0); pat_pv(c, 0, vs, 1); pat_pv(c, 1, vs, 0); pat_pv(c, 1, vs, 1);
    pat_pv(c, 0, ks_, 0); pat_pv(c, 0, ks_, 1); pat_pv(c, 1, ks_, 0); pat_pv(c, 1, ks_, 1);
  } else if (PAT == 6) {  // QK0 QK1 PV0 PV1
    pat_qk(c, 0, ks_); pat_qk(c, 1, ks_);
    pat_pv(c, 0, vs, 0); pat_pv(c, 0, vs, 1); pat_pv(c, 1, vs, 0); pat_pv(c, 1, vs, 1);
  } else if (PAT == 7) {  // same FLOPs with N = 256 SS instructions
    pat_qk256(c, ks_); pat_qk256(c, vs);
  } else if (PAT == 8) {  // the kernel's order, a commit after every 4 MMAs
    pat_pv(c, 0, vs, 0); tc_commit(c.scratch); pat_pv(c, 0, vs, 1); tc_commit(c.scratch + 1);
    pat_qk(c, 0, ks_); tc_commit(c.scratch + 2);
    pat_pv(c, 1, vs, 0); tc_commit(c.scratch + 3); pat_pv(c, 1, vs, 1); tc_commit(c.scratch);
    pat_qk(c, 1, ks_); tc_commit(c.scratch + 1);
  } else if (PAT == 9) {  // two issuing threads: role 0 issues the QKs, role 1 the PVs (no ordering between them)
    if (role == 0) { pat_qk(c, 0, ks_); pat_qk(c, 1, ks_); }
    else { pat_pv(c, 0, vs, 0); pat_pv(c, 0, vs, 1); pat_pv(c, 1, vs, 0); pat_pv(c, 1, vs, 1); }
  } else if (PAT == 10) {  // QK only (16 MMAs)
    pat_qk(c, 0, ks_); pat_qk(c, 1, ks_);
  } else if (PAT == 11) {  // PV only (16 MMAs)
    pat_pv(c, 0, vs, 0); pat_pv(c, 0, vs, 1); pat_pv(c, 1, vs, 0); pat_pv(c, 1, vs, 1);
  }
}

template <int PAT>
__global__ void __launch_bounds__(128, 1) dbg_mma_pattern_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * PAT_TILE);
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // operands: small pseudo-random bf16 values (the timing does not depend on them; the power does a little)
  for (int idx = tid; idx < 7 * PAT_TILE / 4; idx += 128) {
    uint32_t h = idx * 2654435761u;
    h ^= h >> 15;
    const float a = float(int(h & 0xffff) - 32768) * (1.0f / 32768.0f);
    const float b = float(int((h >> 16) & 0xffff) - 32768) * (1.0f / 32768.0f);
    reinterpret_cast<uint32_t*>(smem)[idx] = pack_bf16(a, b);
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bars[6], PAT == 9 ? 2 : 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  PatCtx c{tmem_slot, smem_u32(smem), smem_u32(smem + 2 * PAT_TILE), bars};
  {  // P operand region in TMEM: finite values
    uint32_t v[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = 0x3c003c00u;
    for (int col = 0; col < 512; col += 32) tmem_st_x32(c.tmem + (uint32_t(warp * 32) << 16) + col, v);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const bool issuer = (tid == 0) || (PAT == 9 && tid == 32);
  if (issuer) {
    const int role = tid == 0 ? 0 : 1;
    int stage = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int vs = stage;
      stage = (stage + 1 == 5) ? 0 : stage + 1;
      const int ks_ = stage;
      stage = (stage + 1 == 5) ? 0 : stage + 1;
      pat_step<PAT>(c, ks_, vs, role);
    }
    tc_commit(&bars[6]);
    mbar_wait(&bars[6], 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && role == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(c.tmem, 512);
  }
}

}  // namespace fx

using namespace fx;

template <int PAT>
static int launch_pattern(int iters, long long* out, cudaStream_t st) {
  auto kern = dbg_mma_pattern_kernel<PAT>;
  FX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PAT_SMEM));
  kern<<<num_sms(), 128, PAT_SMEM, st>>>(iters, out);
  return launched("dbg_mma_pattern_kernel");
}

extern "C" int fx_dbg_mma_pattern(int32_t pattern, int32_t iters, int64_t* clocks_out, fx_stream stream) {
  FX_REQUIRE(clocks_out && iters > 0, "fx_dbg_mma_pattern: bad arguments");
  long long* out = (long long*)clocks_out;
  cudaStream_t st = (cudaStream_t)stream;
  switch (pattern) {
    case 0: return launch_pattern<0>(iters, out, st);
    case 1: return launch_pattern<1>(iters, out, st);
    case 2: return launch_pattern<2>(iters, out, st);
    case 3: return launch_pattern<3>(iters, out, st);
    case 4: return launch_pattern<4>(iters, out, st);
    case 5: return launch_pattern<5>(iters, out, st);
    case 6: return launch_pattern<6>(iters, out, st);
    case 7: return launch_pattern<7>(iters, out, st);
    case 8: return launch_pattern<8>(iters, out, st);
    case 9: return launch_pattern<9>(iters, out, st);
    case 10: return launch_pattern<10>(iters, out, st);
    case 11: return launch_pattern<11>(iters, out, st);
  }
  return fail(FX_ERR_INVALID, "fx_dbg_mma_pattern: unknown pattern %d", pattern);
}


extern "C" int fx_dbg_umma_tile(const void* A, const void* B, float* D, int32_t K, int32_t N, int32_t b_mn_major,
                                int32_t a_tmem, uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes, fx_stream stream) {
  FX_REQUIRE(K % 64 == 0 && K <= 128 && N % 64 == 0 && N <= 128, "fx_dbg_umma_tile: K in {64,128}, N in {64,128}");
  DbgUmmaParams p{(const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, K, N, b_mn_major, a_tmem, lbo, sbo, kstep_bytes};
  const int smem = (K / 64) * 16384 + K * N * 2 + 2048;
  static bool done = false;
  if (!done) {
    FX_CUDA(cudaFuncSetAttribute(dbg_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    done = true;
  }
  dbg_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  return launched("dbg_umma_kernel");
}

// CUDA-core reference GEMM (tests only): out[m][n] = sum_k A[m][k] W[n][k], fp32 out
__global__ void dbg_gemm_ref_kernel(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw,
                                    float* out, long long ldo, int M, int N, int K) {
  __shared__ float sa[16][17], sw[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx;
    sa[ty][tx] = (m < M && ka < K) ? __bfloat162float(A[(long long)m * lda + ka]) : 0.f;
    const int nw = blockIdx.x * 16 + ty;
    sw[ty][tx] = (nw < N && ka < K) ? __bfloat162float(W[(long long)nw * ldw + ka]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sw[tx][k];
    __syncthreads();
  }
  if (m < M && n < N) out[(long long)m * ldo + n] = acc;
}

extern "C" int fx_dbg_gemm_ref(const void* A, int64_t lda, const void* W, int64_t ldw, float* out, int64_t ldo,
                               int32_t M, int32_t N, int32_t K, fx_stream stream) {
  dim3 grid((N + 15) / 16, (M + 15) / 16), block(16, 16);
  dbg_gemm_ref_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)W,
                                                               ldw, out, ldo, M, N, K);
  return launched("dbg_gemm_ref_kernel");
}
