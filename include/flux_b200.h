/* flux_b200.h -- C ABI of libflux_b200.so: the B200-native (sm_100a) kernels behind the Flux
 * denoising hot path (sampler step -> MMDiT forward -> VAE decode; T5/CLIP helpers).
 *
 * The reference (voipnuggets/flux-generator) has no native ABI: its hot path is Python calling
 * Apple-MLX primitives.  Each entry point below replaces the MLX call sites cited beside it
 * (paths relative to the reference root); INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add to route those call sites here.
 *
 * Conventions
 *  - plain C: POD structs, raw DEVICE pointers, sizes in elements unless the name says bytes.
 *  - the library never allocates, frees or synchronises device memory; every function only
 *    enqueues work on the caller's `stream` (CUDA-graph capturable) and returns 0 or a negative
 *    fx_status.  fx_last_error() returns a thread-local message for the last failure.
 *  - bf16 storage everywhere unless stated; accumulation is fp32.
 *  - "ld" = leading dimension (elements between consecutive rows); "bs" = batch stride (elements).
 */
#ifndef FLUX_B200_H
#define FLUX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fx_stream; /* cudaStream_t */

typedef enum fx_status {
  FX_OK = 0,
  FX_ERR_INVALID = -1, /* bad shape / alignment / argument (Python shim raises ValueError) */
  FX_ERR_CUDA = -2,    /* CUDA runtime / driver error (RuntimeError) */
  FX_ERR_ARCH = -3     /* device is not sm_100 */
} fx_status;

int fx_version(void);
const char* fx_last_error(void);
/* number of kernels this library has launched on any stream since load (bench.py: gpu_launches) */
uint64_t fx_launch_count(void);
/* 0 when `device` is compute capability 10.x */
int fx_check_device(int device);

/* ---------------------------------------------------------------- tcgen05 GEMM family
 * out = epilogue(A . W^T), A [batch][rows][K], W [N][K] (nn.Linear layout, K contiguous).
 * Replaces every nn.Linear on the path: flux/layers.py:82-85,104-106,134,162-166,175-179,250-252,
 * 292-296; flux/model.py:56-64; flux/autoencoder.py:36-39,84 (1x1 convs); flux/t5.py:126-129,
 * 166-170; flux/clip.py:55-58 and mlx.nn.MultiHeadAttention's projections. */
typedef enum fx_act { FX_ACT_NONE = 0, FX_ACT_GELU_TANH = 1, FX_ACT_QUICK_GELU = 2, FX_ACT_GELU_ERF = 3 } fx_act;

typedef struct fx_gemm_args {
  const void* A; int64_t lda; int64_t a_bs;
  const void* W; int64_t ldw;
  const void* bias;            /* [N] or NULL */
  void* out; int64_t ldo; int64_t out_bs;
  int32_t out_f32;             /* 1: out is float32 */
  int32_t act;                 /* fx_act, applied after bias */
  const void* gate; int64_t gate_bs;               /* [batch][N] multiplier or NULL */
  const void* resid; int64_t ldr; int64_t resid_bs; /* [batch][rows][N] addend or NULL (may alias out) */
  int32_t batch, rows, N, K;
  /* --quantize (txt2image.py:56,79-82; the reference's MLX 4-bit nn.quantize becomes FP8 here): when fp8 != 0,
   * A and W hold e4m3 bytes produced by fx_quantize_rows / fx_rownorm(out_fp8); the accumulator is
   * multiplied by a_scale[b][row] * w_scale[n] before the epilogue.  N > 128 (wide tiles only). */
  int32_t fp8;
  const float* a_scale; int64_t a_scale_bs;         /* [batch][rows] */
  const float* w_scale;                             /* [N] */
} fx_gemm_args;
/* v = A.W^T + bias; v = act(v); v *= gate[b][n]; v += resid[b][r][n]   (each optional) */
int fx_gemm(const fx_gemm_args* a, fx_stream stream);

/* Fused QKV(+MLP-in) projection: W columns are [q | k | v | mlp] (flux/layers.py:195-199,269-277).
 * Per 128-wide head: +bias, QK-RMSNorm (eps, learned scale; flux/layers.py:88-95), RoPE with the
 * (cos,sin) table `pe` (flux/layers.py:24-33), scatter to q/k/v [batch][heads][seq_total][128] at
 * sequence offset seq_off (deletes the split/transpose/concat of flux/layers.py:195-214).  Columns
 * past 3*heads*128 get +bias, GELU(tanh) and land in `mlp_out` (flux/layers.py:283's concat). */
typedef struct fx_qkv_args {
  const void* A; int64_t lda; int64_t a_bs;
  const void* W; int64_t ldw;
  const void* bias;
  const void* q_scale; const void* k_scale; /* [128] RMSNorm weights */
  const void* pe;              /* [seq_total][64][2] bf16 (cos, sin) */
  void* q; void* k; void* v;   /* [batch][heads][seq_total][128] */
  void* mlp_out; int64_t ld_mlp; int64_t mlp_bs; /* [batch][seq_total][ld_mlp] (row = seq_off + r) or NULL */
  float rms_eps;
  int32_t batch, rows, N, K, heads, seq_total, seq_off;
  int32_t fp8;                 /* as fx_gemm_args.fp8 */
  const float* a_scale; int64_t a_scale_bs;
  const float* w_scale;
  /* pe_blocked != 0: `pe` is laid out [ceil(seq_total / 32)][16 pieces][32 rows][16 bytes] (piece = 4 (cos, sin) pairs)
   * so that the 32 rows a warp owns read every piece as one coalesced 512-byte request; needs seq_off % 32 == 0. */
  int32_t pe_blocked;
  /* qkv_fp8 != 0: q, k, v are written as e4m3 bytes [batch][heads][seq_total][128] (round to nearest, saturating, no
   * scale: after QK-RMSNorm and RoPE the values are O(1)) -- the operands of fx_attention's fp8 mode */
  int32_t qkv_fp8;
} fx_qkv_args;
int fx_gemm_qkv(const fx_qkv_args* a, fx_stream stream);

/* 3x3 / pad 1 / stride 1 convolution as implicit GEMM on NHWC bf16 (nn.Conv2d at
 * flux/autoencoder.py:69-81,117-119,237-239,269).  W is OHWI flattened to [Cout][9*Cin]
 * (AutoEncoder.sanitize's layout, flux/autoencoder.py:336-345).  Cin % 64 == 0.
 * out = conv + bias (+ resid).  out may be float32. */
typedef struct fx_conv3x3_args {
  const void* x;               /* [batch][H][W][Cin] */
  const void* W; const void* bias;
  void* out; int32_t out_f32;  /* [batch][H][W][Cout] */
  const void* resid;           /* [batch][H][W][Cout] or NULL */
  int32_t batch, H, Wd, Cin, Cout;
  float* gn_partials;          /* NULL, or [batch][fx_conv3x3_gn_blocks(H, Wd, Cout)][32][2] floats: the epilogue also
                                  writes per-pixel-block (sum, sum of squares) of the bf16 output for each of the 32
                                  GroupNorm groups -- the statistics pass of the GroupNorm that consumes this output
                                  (flux/autoencoder.py:88-94) folded into its producer; reduce with
                                  fx_groupnorm_finalize_blocks.  Needs Cout % 128 == 0, bf16 output. */
  int32_t upsample2x;          /* 1: out = conv3x3(upsample_nearest(x, 2)) (Upsample, flux/autoencoder.py:121-124) without
                                  materialising the upsampled tensor: out is [batch][2H][2W][Cout] and W holds the four
                                  parity-wise 2x2 kernels [4][Cout][4*Cin] (taps that fall on the same source pixel
                                  pre-summed: ops.upconv_weights).  No resid; Cout % 128 == 0. */
} fx_conv3x3_args;
int fx_conv3x3(const fx_conv3x3_args* a, fx_stream stream);
int64_t fx_conv3x3_gn_blocks(int32_t H, int32_t Wd, int32_t Cout, int32_t upsample2x);

/* ---------------------------------------------------------------- NVFP4 (W4A4) block Linears: `--quantize` (4 bits, the default)
 * The Blackwell analogue of the reference's 4-bit `nn.quantize(group_size=64)` of the Linear layers
 * (txt2image.py:28-29,79-82): e2m1 values (two per byte) with one UE4M3 scale per 16 elements of K, consumed by
 * tcgen05.mma.kind::mxf4nvf4.block_scale, plus one fp32 scale per row (activations) / output channel (weights).
 *
 * fx_quantize_rows_fp4: x bf16 [batch][rows][K] (K % 64 == 0, K <= 16384) ->
 *   q     [batch * rows][K / 2] bytes, element 2i in the low nibble
 *   sf    UE4M3 block scales as 512-byte atoms [ceil(batch * rows / 128)][K / 64]: the scale of elements
 *         [16 j, 16 j + 16) of row m sits in atom (m / 128, j / 4) at byte (m % 32) * 16 + ((m % 128) / 32) * 4 + j % 4
 *         (the order tcgen05.cp.32x128b.warpx4 expects; the buffer must cover whole 128-row blocks)
 *   scale fp32 [batch * rows]:  x ~= e2m1 * ue4m3 * scale;  scale = absmax(row) / 2688, ue4m3 = rn(absmax(block) / 6 / scale) */
typedef struct fx_quant4_args {
  const void* x; int64_t ldx; int64_t x_bs;
  void* q; void* sf; float* scale;
  int32_t batch, rows, K;
} fx_quant4_args;
int fx_quantize_rows_fp4(const fx_quant4_args* a, fx_stream stream);

/* out = resid + gate * act((A4 . W4^T) * a_scale[row] * w_scale[col] + bias): A4 / sfa / a_scale as written by
 * fx_quantize_rows_fp4 for the flattened [batch * rows] activation rows (rows % 128 == 0 unless batch == 1); W4 [N][K / 2]
 * with its scale atoms regrouped per 192-row column tile, [ceil(N / 192)][K / 64][2][512] (rows 0-127 and 128-191 of the
 * tile; flux.ops.fp4_weight prepares both from fx_quantize_rows_fp4's output).  K % 256 == 0.  Epilogue fields as fx_gemm_args. */
typedef struct fx_gemm4_args {
  const void* A; const void* sfa; const float* a_scale;
  const void* W; const void* sfw; const float* w_scale;
  const void* bias;
  void* out; int64_t ldo; int64_t out_bs; int32_t out_f32; int32_t act;
  const void* gate; int64_t gate_bs;
  const void* resid; int64_t ldr; int64_t resid_bs;
  int32_t batch, rows, N, K;
  /* q_out != NULL: instead of `out`, the epilogue (bias, act) writes columns [out_col0, out_col0 + N) of the NEXT GEMM's NVFP4
   * operand (K = out_kc): e2m1 rows q_out [batch * rows][out_kc / 2], scale atoms sf_out, one power-of-two exponent per 32-column
   * chunk e_out int8 (256 bytes per 128-row block and K-group of 64, byte (r % 32) * 8 + (r / 32) * 2 + (chunk & 1): the scale atoms'
   * order); fx_fp4_finalize completes the operand once every column has been produced
   * (the GELU(mlp) hidden never exists in bf16).  N % 64 == 0, rows % 128 == 0, no gate / residual. */
  void* q_out; void* sf_out; void* e_out; int32_t out_kc; int32_t out_col0;
} fx_gemm4_args;
int fx_gemm_fp4(const fx_gemm4_args* a, fx_stream stream);

/* The same chunked producer for a bf16 tensor x [batch][rows][C] (stand-alone form; in the model the attention epilogue emits
 * its chunks itself, fx_attn_args.q_out): C % 64 == 0; layout of q / sf / e as fx_gemm4_args.q_out.
 * Format of a producer-emitted operand (oracle: nvfp4_quant_rows_chunked): per 32-column chunk e_c = ceil(log2(absmax(chunk) / 2688))
 * (>= -100), per block of 16 sf = e4m3_rn(absmax(block) / 6 * 2^-e_c), q = e2m1_rn(x * rcp(sf * 2^e_c)); fx_fp4_finalize then sets
 * scale[row] = 2^max_c(e_c) and sf <- e4m3_rn(sf * 2^(e_c - max e_c)), so that x ~= e2m1 * ue4m3 * scale as for fx_quantize_rows_fp4. */
typedef struct fx_quant4c_args {
  const void* x; int64_t ldx; int64_t x_bs;
  int32_t batch, rows, C;
  void* q; void* sf; void* e; int32_t kc; int32_t col0;
} fx_quant4c_args;
int fx_quantize_chunks_fp4(const fx_quant4c_args* a, fx_stream stream);
/* scale[row] = 2^max(e[row][:]); every block scale of sf shifted from its chunk's exponent to the row's.  rows % 128 == 0. */
int fx_fp4_finalize(void* sf, const void* e, float* scale, int64_t rows, int32_t kc, fx_stream stream);

/* fx_gemm_fp4_qkv: the NVFP4 form of fx_gemm_qkv (flux/layers.py:195-214: qkv Linear -> per-head QK-RMSNorm -> RoPE -> q, k, v
 * [batch][heads][seq_total][128] at rows seq_off + r, bf16 or e4m3).  N = 3 * heads * 128; tiles are one head (128 columns) wide,
 * so W's scale atoms are [N / 128][K / 64] (ops.fp4_weight(w, tile_n=128)).  Rows % 128 == 0 unless batch == 1; K % 256 == 0. */
typedef struct fx_gemm4_qkv_args {
  const void* A; const void* sfa; const float* a_scale;
  const void* W; const void* sfw; const float* w_scale;
  const void* bias;
  const void* q_scale; const void* k_scale; /* [128] RMSNorm weights */
  const void* pe; int32_t pe_blocked;       /* as fx_qkv_args */
  void* q; void* k; void* v; int32_t qkv_fp8;
  float rms_eps;
  int32_t batch, rows, K, heads, seq_total, seq_off;
} fx_gemm4_qkv_args;
int fx_gemm_fp4_qkv(const fx_gemm4_qkv_args* a, fx_stream stream);

/* ---------------------------------------------------------------- attention
 * Non-causal softmax(q k^T * scale) v over head_dim 128 on tcgen05 (flash-style, online softmax);
 * replaces mx.fast.scaled_dot_product_attention + the transpose/reshape at flux/layers.py:41-43.
 * q,k,v [batch][heads][seq][128]; out [batch][seq][ld_out], head h at columns h*128. */
typedef struct fx_attn_args {
  const void* q; const void* k; const void* v;
  void* out; int64_t ld_out; int64_t out_bs;
  float scale;
  int32_t batch, heads, seq;
  int32_t variant;             /* 0 = default; 1 = P via shared memory (coupled schedule); 4 = default with
                                  exp-phase turn taking between the softmax warpgroups; 5 / 6 = decoupled schedule with P through shared memory (attn3_kernel), with / without
                                  turn taking (tests / A-B); 7 = persistent work loop (what 0 runs unless
                                  FX_ATTN_PERSISTENT=0) */
  int32_t fp8;                 /* 1: q, k, v are e4m3 [batch][heads][seq][128] bytes (written by fx_gemm_qkv with qkv_fp8);
                                  both products run in FP8 (P is converted to e4m3), fp32 softmax, bf16 output */
  /* q_out != NULL (persistent kernel): instead of bf16 `out`, O / l is written as columns [out_col0 + h * 128, ...) of the NEXT
   * GEMM's NVFP4 operand (layout and finalisation as fx_gemm4_args.q_out).  Rows >= out_split of every batch element go to the first
   * operand (flattened row b * (seq - out_split) + row - out_split), rows below it to the second (b * out_split + row, column 0):
   * the image / text streams of a double block feed different `proj` GEMMs.  Both row counts multiples of 128. */
  void* q_out; void* sf_out; void* e_out; int32_t out_kc; int32_t out_col0;
  void* q_out2; void* sf_out2; void* e_out2; int32_t out_kc2; int32_t out_split;
} fx_attn_args;
int fx_attention(const fx_attn_args* a, fx_stream stream);

/* Small generic attention for the once-per-prompt text encoders (head_dim 64): optional additive
 * bias [heads][seq][seq] (T5, flux/t5.py:153-155) and causal mask (CLIP, flux/clip.py:90-94).
 * q,k,v are column slices of [batch][seq][ld] activations, head h at columns h*64. */
typedef struct fx_attn_small_args {
  const void* q; const void* k; const void* v; int64_t ld; int64_t bs;
  const float* bias;           /* fp32 [heads][seq][seq] or NULL */
  void* out; int64_t ld_out; int64_t out_bs;
  float scale;
  int32_t batch, heads, seq, causal;
} fx_attn_small_args;
int fx_attention_small(const fx_attn_small_args* a, fx_stream stream);

/* ---------------------------------------------------------------- row-wise normalisation
 * mode 0: LayerNorm(no affine, eps) then (1+scale[b])*y + shift[b]  (flux/layers.py:193,222,267,300)
 * mode 1: LayerNorm(affine weight/bias per column)                  (flux/clip.py:52-53,87)
 * mode 2: RMSNorm(weight per column)                                (flux/t5.py:196-197,216) */
typedef struct fx_rownorm_args {
  const void* x; int64_t ldx; int64_t x_bs;
  void* out; int64_t ldo; int64_t out_bs;
  const void* p0; const void* p1; int64_t p_bs; /* mode 0: shift, scale [batch][D]; mode 1: weight, bias; mode 2: weight */
  float eps;
  int32_t mode, batch, rows, D;
  /* --quantize: out_fp8 != 0 writes e4m3 bytes to `out` (ldo / out_bs in bytes) with one dequantisation scale per
   * row in scale_out[b * scale_bs + r]; identical to fx_quantize_rows applied to the bf16 result: the A operand of an
   * fp8 fx_gemm / fx_gemm_qkv */
  int32_t out_fp8;
  float* scale_out; int64_t scale_bs;
  /* --quantize 4: out_fp8 == 2 writes the NVFP4 operand of fx_gemm_fp4 / fx_gemm_fp4_qkv instead: `out` = compact e2m1 rows
   * [batch * rows][D / 2] (ldo / out_bs ignored), sf_out = UE4M3 scale atoms, scale_out = fp32 [batch * rows]; bit-identical
   * to fx_quantize_rows_fp4 applied to the bf16 result.  D % 64 == 0, D >= 1024. */
  void* sf_out;
} fx_rownorm_args;
int fx_rownorm(const fx_rownorm_args* a, fx_stream stream);

/* FP8 row quantisation for --quantize (txt2image.py:56,79-82 -- the reference quantises nn.Linear weights to MLX
 * 4-bit groups; here weights AND activations of the block Linears go to e4m3 for the tcgen05 f8f6f4 MMAs):
 * x bf16 [batch][rows][K] -> q e4m3 bytes + scale fp32 [batch][rows]; q = e4m3_rn_satfinite(x * (448 / absmax)),
 * scale = absmax * (1/448) (both 1 for a zero row).  Used once per weight at load and per step for GEMM inputs that no
 * norm kernel produces (attention output | GELU(mlp) -> linear2 / proj / fc2). */
typedef struct fx_quant_args {
  const void* x; int64_t ldx; int64_t x_bs;
  void* q; int64_t ldq; int64_t q_bs;
  float* scale; int64_t scale_bs;
  int32_t batch, rows, K;
} fx_quant_args;
int fx_quantize_rows(const fx_quant_args* a, fx_stream stream);

/* ---------------------------------------------------------------- conditioning vector path
 * out[b][n] = sum_k f(in[b][k]) W[n][k] + bias[n] (+ add[b][n]); f = SiLU when silu_in.
 * The M=batch GEMVs of MLPEmbedder / Modulation / LastLayer.adaLN (flux/layers.py:78-85,129-143,
 * 294-299): weight-streaming, HBM-bound. */
typedef struct fx_gemv_args {
  const void* in; int64_t ld_in;
  const void* W; int64_t ldw;
  const void* bias; const void* add; int64_t ld_add;
  void* out; int64_t ld_out;
  int32_t batch, N, K, silu_in, silu_out;
} fx_gemv_args;
int fx_gemv(const fx_gemv_args* a, fx_stream stream);

/* timestep_embedding (flux/layers.py:46-57): t [batch] bf16 -> [batch][256] bf16 = bf16([cos|sin]
 * (bf16(1000*t) * freqs)). */
int fx_timestep_embedding(const void* t_bf16, void* out, int32_t batch, int32_t dim, fx_stream stream);

/* FluxSampler.step (flux/sampler.py:56-57): x = bf16(x + dt * pred), n elements. */
int fx_euler_step(void* x, const void* pred, float dt, int64_t n, fx_stream stream);

/* ---------------------------------------------------------------- latent packing
 * _prepare_latent_images (flux/flux.py:53-58): [b][h][w][c] -> [b][hw/4][4c], feature c*4+dy*2+dx */
int fx_patchify(const void* x, void* out, int32_t b, int32_t h, int32_t w, int32_t c, fx_stream stream);
/* FluxSampler.sample_prior fused with the packing above (flux/sampler.py:44-45): standard-normal bf16 noise
 * written as [b][hw/4][4c]; Philox4x32-10 keyed by (seed, first_index + image) and counted by the NHWC element
 * index, so an image's prior is independent of batch composition and GPU sharding.  (MLX's threefry stream
 * is not reproducible offline; callers that need specific noise pass x_T instead.) */
int fx_prior_packed(void* out, int32_t b, int32_t h, int32_t w, int32_t c, uint64_t seed, int32_t first_index,
                    fx_stream stream);
/* FluxPipeline.decode's inverse + AutoEncoder.decode's affine (flux/flux.py:159-160,
 * flux/autoencoder.py:353): packed [b][hw/4][4c] -> z [b][h][w][c_pad] = x/scale + shift, channels
 * c..c_pad-1 zero (conv_in reads 64-channel blocks). */
int fx_unpatchify_scale(const void* packed, void* z, int32_t b, int32_t h, int32_t w, int32_t c,
                        int32_t c_pad, float scale_factor, float shift_factor, fx_stream stream);

/* ---------------------------------------------------------------- VAE decoder pieces
 * GroupNorm(32 groups, eps, affine) [+ SiLU] on NHWC bf16 (flux/autoencoder.py:29-35,62-78,88-94).
 * stats writes per-block partial sums (float, fx_groupnorm_partials_count(batch, hw) elements: no atomics,
 * bit-reproducible); finalize reduces them in a fixed order to float [batch][32][2] (mean, rstd); apply
 * normalises with those. */
int64_t fx_groupnorm_partials_count(int32_t batch, int64_t hw);
int fx_groupnorm_stats(const void* x, float* partials, int32_t batch, int64_t hw, int32_t C, fx_stream stream);
int fx_groupnorm_finalize(const float* partials, float* stats, int32_t batch, int64_t hw, int32_t C, float eps,
                          fx_stream stream);
/* the same reduction over `nblk` partial blocks per image (the layout fx_conv3x3's gn_partials writes) */
int fx_groupnorm_finalize_blocks(const float* partials, float* stats, int32_t batch, int64_t nblk, int64_t hw, int32_t C,
                                 float eps, fx_stream stream);
int fx_groupnorm_apply(const void* x, const float* stats, const void* weight, const void* bias, void* out,
                       int32_t batch, int64_t hw, int32_t C, int32_t silu, fx_stream stream);
/* upsample_nearest(x, (2,2)) on NHWC (flux/autoencoder.py:122) */
int fx_upsample2x(const void* x, void* out, int32_t batch, int32_t H, int32_t W, int32_t C, fx_stream stream);
/* P = softmax(scale * S) row-wise, S fp32 [rows][ld_s] -> P bf16 [rows][ld_p]  (VAE mid attention,
 * flux/autoencoder.py:49, head_dim 512 via two GEMMs) */
int fx_softmax_rows(const float* S, int64_t ld_s, void* P, int64_t ld_p, int64_t rows, int32_t cols, float scale,
                    fx_stream stream);
/* [rows][cols] -> [cols][rows] bf16 transpose (V^T for the P.V GEMM) */
int fx_transpose(const void* x, int64_t ldx, void* out, int64_t ldo, int32_t rows, int32_t cols, fx_stream stream);
/* clip(x+1,0,2)*0.5 (flux/flux.py:162) -> float32 image, and (img*255) truncated to uint8
 * (txt2image.py:133).  x is float32 [n]; either output may be NULL. */
int fx_finish_image(const float* x, float* img, uint8_t* u8, int64_t n, fx_stream stream);

/* ---------------------------------------------------------------- text-encoder helpers
 * rows of `table` [vocab][D] gathered by int32 ids (+ optional pos table row i%seq) */
int fx_embedding(const int32_t* ids, const void* table, const void* pos_table, void* out, int64_t n_ids,
                 int32_t seq, int32_t D, fx_stream stream);
/* y = act(a) * b elementwise (T5 gated FFN, flux/t5.py:178-184), act = fx_act */
int fx_act_mul(const void* a, const void* b, void* out, int64_t n, int32_t act, fx_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* FLUX_B200_H */
