// NVFP4 operand producers shared by the GEMM epilogue (gemm4.cu), the chunk quantiser and the attention epilogue (attention.cu).
#pragma once
#include <cuda_fp4.h>
#include <cuda_fp8.h>
#include <stdint.h>

namespace fx {

// Destination of a producer that emits an NVFP4 operand chunk by chunk (32 columns of a row at a time): e2m1 rows, UE4M3
// scale atoms and one power-of-two exponent per chunk; fx_fp4_finalize then lifts the block scales to the row's scale
// (oracle: nvfp4_quant_rows_chunked).  kc = total columns (K) of the consumer's operand.  The exponents mirror the scale atoms:
// 256 bytes per (128-row block, K-group of 64): byte (r % 32) * 8 + (r / 32) * 2 + (chunk & 1) -- so the finalise pass reads the
// exponents of the four rows whose scales share a 16-byte atom piece with one 8-byte load.
struct ChunkQ {
  uint8_t* q;
  uint8_t* sf;
  int8_t* e;
  int kc;
  int col0;   // column of the operand at which this producer's column 0 lands
};
// one chunk (32 values of flattened row m, operand columns [col, col + 32)) -> 16 bytes of e2m1, two UE4M3 block scales, the
// chunk's exponent.  Every step a single IEEE fp32 operation (bit-exact against the oracle for identical inputs).
__device__ __forceinline__ void fp4_chunk_quantise(const float* f, long long m, int col, const ChunkQ& o, bool valid) {
  float bm[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    bm[0] = fmaxf(bm[0], fabsf(f[i]));
    bm[1] = fmaxf(bm[1], fabsf(f[16 + i]));
  }
  const float t = __fmul_rn(fmaxf(bm[0], bm[1]), 1.0f / 2688.0f);
  const uint32_t tb = __float_as_uint(t);
  int e = int((tb >> 23) & 255u) - 127 + ((tb & 0x7fffffu) != 0u ? 1 : 0);   // smallest e with 2^e >= t
  e = e < -100 ? -100 : e;
  const float g = __uint_as_float(uint32_t(127 + e) << 23), inv_g = __uint_as_float(uint32_t(127 - e) << 23);
  uint32_t w[4];
  uint32_t sfw = 0;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    const float u = __fmul_rn(__fmul_rn(bm[b], 1.0f / 6.0f), inv_g);
    const __nv_fp8_storage_t sf8 = __nv_cvt_float_to_fp8(u, __NV_SATFINITE, __NV_E4M3);
    const float d = __fmul_rn(__half2float(__half(__nv_cvt_fp8_to_halfraw(sf8, __NV_E4M3))), g);
    const float rd = d > 0.f ? __frcp_rn(d) : 0.f;
    sfw |= uint32_t(sf8) << (8 * b);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t x = 0;
#pragma unroll
      for (int i = 0; i < 8; i += 2)
        x |= uint32_t(__nv_cvt_float2_to_fp4x2(make_float2(__fmul_rn(f[16 * b + 8 * h + i], rd), __fmul_rn(f[16 * b + 8 * h + i + 1], rd)),
                                               __NV_E2M1, cudaRoundNearest))
             << (4 * i);
      w[2 * b + h] = x;
    }
  }
  if (valid) {
    *reinterpret_cast<uint4*>(o.q + m * (long long)(o.kc / 2) + (col >> 1)) = make_uint4(w[0], w[1], w[2], w[3]);
    const int rr = int(m & 127);
    *reinterpret_cast<uint16_t*>(o.sf + ((m >> 7) * (long long)(o.kc / 64) + (col >> 6)) * 512 + (rr & 31) * 16 + (rr >> 5) * 4 +
                                 ((col >> 4) & 3)) = uint16_t(sfw);
    o.e[((m >> 7) * (long long)(o.kc / 64) + (col >> 6)) * 256 + (rr & 31) * 8 + (rr >> 5) * 2 + ((col >> 5) & 1)] = int8_t(e);
  }
}

}  // namespace fx
