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
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
