/* flux_b200_dbg.h -- TEST-ONLY companion of flux_b200.h (libflux_b200_dbg.so).
 *
 * Bring-up probes and CUDA-core reference kernels that tests/ and the profiling scripts use to check and
 * time the tcgen05 paths at sizes the CPU oracle cannot reach.  Nothing here is on the product path: the
 * product library (libflux_b200.so) neither contains nor calls these symbols; this library links against
 * the product library only for its error / launch-count plumbing. */
#ifndef FLUX_B200_DBG_H
#define FLUX_B200_DBG_H
#include "flux_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Plain CUDA-core reference GEMM: out[m][n] = sum_k A[m][k] W[n][k], fp32 out. */
int fx_dbg_gemm_ref(const void* A, int64_t lda, const void* W, int64_t ldw, float* out, int64_t ldo, int32_t M,
                    int32_t N, int32_t K, fx_stream stream);

/* One CTA, K/16 tcgen05.mma on hand-laid shared-memory operands with caller-supplied descriptor fields:
 * sweeps UMMA encodings (MN-major B, A-from-TMEM) against a CPU matmul.  A [128][K]; B [N][K] (K-major) or
 * [K][N] (MN-major); D float [128][N]. */
int fx_dbg_umma_tile(const void* A, const void* B, float* D, int32_t K, int32_t N, int32_t b_mn_major,
                     int32_t a_tmem, uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes, fx_stream stream);

/* Tensor-pipe issue-pattern micro-benchmark (profiling only): every SM's CTA has one thread issue `iters`
 * steps of 32 tcgen05.mma (128x128x16) in one of a dozen fixed orders (the attention kernel's QK / PV
 * sequence with and without commits, single-kind streams, N = 256, two issuing threads ...) on the attention
 * kernel's shared-memory / TMEM layout; writes the SM-clock count of CTA 0 to clocks_out[0]. */
int fx_dbg_mma_pattern(int32_t pattern, int32_t iters, int64_t* clocks_out, fx_stream stream);

/* Byte-operand probe for the 8-bit / 4-bit tensor-core kinds (one CTA, operands laid out by hand, 32 bytes of K per
 * MMA): kind 0 kind::f8f6f4 (A from shared memory or, a_tmem, from TMEM; B K-major or, b_mn_major, MN-major with the
 * caller's LBO / SBO / per-MMA byte step), kind 1 kind::mxf8f6f4.block_scale (UE8M0 per 32), kind 2
 * kind::mxf4nvf4.block_scale.block16 (NVFP4: e2m1, UE4M3 per 16), kind 3 ...block32 (MXFP4).  A [128][kbytes],
 * B [N][kbytes] (or [K][N]), SFA [128][nsf], SFB [N][nsf] bytes; D float [128][N].  Scale factors travel through
 * shared-memory atoms + tcgen05.cp.32x128b.warpx4 (descriptor LBO / SBO from the caller). */
int fx_dbg_bs_tile(const void* A, const void* B, const void* SFA, const void* SFB, float* D, int32_t N, int32_t kbytes,
                   int32_t kind, int32_t a_tmem, int32_t b_mn_major, int32_t nsf, uint32_t b_lbo, uint32_t b_sbo,
                   uint32_t b_kstep, uint32_t cp_lbo, uint32_t cp_sbo, fx_stream stream);

/* The NVFP4 probe for a CTA pair (cta_group::2, M = 256): A [256][kbytes], B [N][kbytes], SFA [256][nsf], SFB [N][nsf], D float
 * [256][N].  sfb_mode 0: every CTA stages the scale atoms of all N columns; 1: only those of its own half of the W rows. */
int fx_dbg_bs2_tile(const void* A, const void* B, const void* SFA, const void* SFB, float* D, int32_t N, int32_t kbytes,
                    int32_t nsf, int32_t sfb_mode, fx_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* FLUX_B200_DBG_H */
