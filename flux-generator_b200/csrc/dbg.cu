// Bring-up probe (tests only, never on the product path): one CTA issues K/16 tcgen05.mma on operands
// it lays out in shared memory by hand, with the descriptor fields supplied by the caller.  Lets a
// single GPU run sweep UMMA descriptor encodings (MN-major B, A-from-TMEM) against a CPU matmul.
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

}  // namespace fx

using namespace fx;

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
