// Joint (text+image) attention for the MMDiT on tcgen05:  out = softmax(q k^T * scale) v, head_dim 128,
// no mask (flux/layers.py:36-43).  Flash-style: one CTA owns 256 query rows of one (batch, head) as two
// 128-row tiles that ping-pong on the tensor pipe; K/V tiles stream through a TMA ring; S and O live
// in TMEM; two softmax warpgroups (one row per thread) run the online softmax in fp32 with exp2 and a
// lazy O-rescale (only when the running max grows by > 2^8).
//   warps 0-3: softmax/correction/epilogue for tile 0      warps 4-7: same for tile 1
//   warp 8: TMA producer      warp 9: tcgen05.mma issuer + TMEM owner   (10, 11 idle)
// (the scheduler favours higher warp ids: the single-thread issuers sit above the softmax warps)
// P (bf16 probabilities) is handed to the P.V MMA either through TMEM (aliasing S, default) or through
// 128B-swizzled shared memory (variant 1, bring-up fallback).  V is consumed MN-major straight from its
// [seq][128] layout (no transposed copy).
#include <stdlib.h>

#include <mutex>

#include "api_common.cuh"
#include "fp4.cuh"
#include "sm100.cuh"
#include "tmap.cuh"

namespace fx {

constexpr int ATT_THREADS = 384;
#ifdef FX_ROLES_LOW  // A/B builds only
constexpr int ATT_WARP_TMA = 0, ATT_WARP_MMA = 1, ATT_CTRL0 = 0, ATT_SM0 = 4;
#else
constexpr int ATT_WARP_TMA = 8, ATT_WARP_MMA = 9, ATT_CTRL0 = 8, ATT_SM0 = 0;
#endif
constexpr int ATT_TILE_BYTES = 128 * 128 * 2;  // one 128x128 bf16 tile = two 16 KB swizzled halves

struct AttnParams {
  int seq, heads, kv_tiles;
  int sequence;  // 1: the two softmax warpgroups take turns on the exp phase (MUFU is shared per sub-partition)
  float scale_log2;
  __nv_bfloat16* out;
  long long ld_out, out_bs;
  // NVFP4 output (persistent kernel): O / l leaves the epilogue as chunks of the next GEMM's operand (fp4.cuh) instead of bf16.
  // Rows >= out4_split of every batch element go to out4[0] (flattened row b * (seq - split) + row - split), rows below it to
  // out4[1] (b * split + row): the image / text streams of a double block feed different `proj` operands.
  ChunkQ out4[2];
  int out4_split;
};

__device__ __forceinline__ float fast_exp2(float x) {
#ifdef FX_ATTN_NOEXP  // timing ablation only (wrong numerics): no MUFU work
  return x * 0.0078125f;
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { return pack2f(lo, hi); }
__device__ __forceinline__ float2 unpack2(uint64_t v) { return unpack2f(v); }
// exp2 on the FMA pipe for a pair of values (Cody-Waite split + cubic minimax of 2^f on [-0.5, 0.5], max
// relative error 7.5e-5 -- far below the bf16 rounding of P).  MUFU.EX2 sustains only 4 lanes/clk per
// sub-partition, exactly as many cycles as the two MMAs of a tile take, so a share of the exponentials
// is moved off the MUFU pipe (same idea as FlashAttention-4's software exp2).
__device__ __forceinline__ void exp2_emu2(uint64_t t2, float& p0, float& p1) {
  float2 t = unpack2f(t2);
  t.x = fmaxf(t.x, -125.0f);
  t.y = fmaxf(t.y, -125.0f);
  t2 = pack2f(t.x, t.y);
  const uint64_t magic = pack2f(12582912.0f, 12582912.0f);  // 1.5 * 2^23: low mantissa bits <- round(t)
  const uint64_t r2 = fadd2(t2, magic);
  const uint64_t f2 = fadd2(r2, pack2f(-12582912.0f, -12582912.0f));
  const uint64_t x2 = ffma2(f2, pack2f(-1.0f, -1.0f), t2);  // t - round(t) in [-0.5, 0.5]
  uint64_t q2 = ffma2(pack2f(0.0551716648f, 0.0551716648f), x2, pack2f(0.2426111251f, 0.2426111251f));
  q2 = ffma2(q2, x2, pack2f(0.6932609677f, 0.6932609677f));
  q2 = ffma2(q2, x2, pack2f(0.9999280572f, 0.9999280572f));
  const float2 q = unpack2f(q2), r = unpack2f(r2);
  p0 = __int_as_float(__float_as_int(q.x) + (__float_as_int(r.x) << 23));
  p1 = __int_as_float(__float_as_int(q.y) + (__float_as_int(r.y) << 23));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

#ifndef FX_ATTN_EMU_MASK
#define FX_ATTN_EMU_MASK 0x11u  // bit i set: pair i of every 8 uses the software exp2 (25 %: 3471 -> 3303 clocks per key tile)
#endif
constexpr uint32_t EMU_MASK = FX_ATTN_EMU_MASK;
// the e4m3 kernel: P is rounded to three mantissa bits, so a QUADRATIC 2^f (max relative error 1.8e-3, three instructions fewer per
// pair) is as good as exact, and with the MMAs halved the MUFU pipe weighs more in the loop: its share is tuned separately
#ifndef FX_ATTN_EMU_MASK_F8
#define FX_ATTN_EMU_MASK_F8 0x11u
#endif
constexpr uint32_t EMU_MASK_F8 = FX_ATTN_EMU_MASK_F8;
__device__ __forceinline__ void exp2_emu2_quad(uint64_t t2, float& p0, float& p1) {
  float2 t = unpack2f(t2);
  t.x = fmaxf(t.x, -125.0f);
  t.y = fmaxf(t.y, -125.0f);
  t2 = pack2f(t.x, t.y);
  const uint64_t r2 = fadd2(t2, pack2f(12582912.0f, 12582912.0f));
  const uint64_t f2 = fadd2(r2, pack2f(-12582912.0f, -12582912.0f));
  const uint64_t x2 = ffma2(f2, pack2f(-1.0f, -1.0f), t2);
  uint64_t q2 = ffma2(pack2f(0.23783042f, 0.23783042f), x2, pack2f(0.70339652f, 0.70339652f));
  q2 = ffma2(q2, x2, pack2f(1.0005137f, 1.0005137f));
  const float2 q = unpack2f(q2), r = unpack2f(r2);
  p0 = __int_as_float(__float_as_int(q.x) + (__float_as_int(r.x) << 23));
  p1 = __int_as_float(__float_as_int(q.y) + (__float_as_int(r.y) << 23));
}

// FX_ATTN_PROBE (profiling builds only): one CTA records SM-clock timestamps of its pipeline events into
// a global buffer [role 3][step 64][event 8]; roles: 0 MMA issuer, 1/2 softmax warpgroup 0/1 (first warp).
// waits of the MMA warp: all 32 lanes poll (warp-uniform control flow).  A/B in the sustained, power-capped
// regime: one polling lane + __syncwarp was 2 % slower; FX_ATTN_MMA_ONEWAIT selects it.
#ifdef FX_ATTN_MMA_ONEWAIT
#define MMA_WAIT(bar, ph) do { if (lane == 0) mbar_wait(bar, ph); __syncwarp(); } while (0)
#else
#define MMA_WAIT(bar, ph) mbar_wait(bar, ph)
#endif
#ifdef FX_ATTN_MMAONLY  // timing ablation: tensor pipe alone (no softmax, P never waited for)
#define P_WAIT(bar, ph) do {} while (0)
#else
#define P_WAIT(bar, ph) MMA_WAIT(bar, ph)
#endif
#ifdef FX_ATTN_PROBE
__device__ long long* g_attn_probe = nullptr;
__device__ __forceinline__ long long probe_clock() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}
__device__ __forceinline__ long long probe_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PROBE(role, step, ev)                                                                         \
  do {                                                                                                \
    if (probe_on && (step) < 64) g_attn_probe[((role) * 64 + (step)) * 8 + (ev)] = probe_clock();     \
  } while (0)
#else
#define PROBE(role, step, ev) do {} while (0)
#endif

template <bool P_TMEM>
struct AttnCfg {
  static constexpr int KV_STAGES = P_TMEM ? 5 : 3;
  static constexpr int Q_OFF = 0;
  static constexpr int KV_OFF = 2 * ATT_TILE_BYTES;
  static constexpr int P_OFF = KV_OFF + KV_STAGES * ATT_TILE_BYTES;
  static constexpr int BAR_OFF = P_OFF + (P_TMEM ? 0 : 2 * ATT_TILE_BYTES);
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
};

template <bool P_TMEM>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
            const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  using Cfg = AttnCfg<P_TMEM>;
  constexpr int NS = Cfg::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* q_full = bars;             // 1
  uint64_t* kv_full = bars + 1;        // NS
  uint64_t* kv_empty = kv_full + NS;   // NS
  uint64_t* s_full = kv_empty + NS;    // 2
  uint64_t* p_full = s_full + 2;       // [tile][half] = 4: P is handed over in two 64-key halves
  uint64_t* o_full = p_full + 4;       // 2
  uint64_t* seq_bar = o_full + 2;      // 2: exp-phase turn taking between the softmax warpgroups
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(seq_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int bh = blockIdx.z * p.heads + blockIdx.y;
  const int T = p.kv_tiles;
#ifdef FX_ATTN_PROBE
  const bool probe_on = g_attn_probe && blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 1 && lane == 0;
#endif

  if (warp == ATT_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 4);
      mbar_init(&p_full[2 * i + 1], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&seq_bar[i], 4);
    }
    fence_barrier_init();
  }
#ifdef FX_ATTN_PROBE
  if (probe_on && warp == ATT_WARP_MMA) {
    g_attn_probe[(0 * 64 + 63) * 8 + 0] = probe_clock();
    g_attn_probe[(0 * 64 + 63) * 8 + 2] = probe_gtime();
  }
#endif
  if (warp == ATT_WARP_MMA) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_i aliases S_i[0,64)

  // warpgroup 0 (TMA / MMA / 2 idle warps) gives registers to the two softmax warpgroups
  if (warp >= ATT_CTRL0 && warp < ATT_CTRL0 + 4) {
  reg_dec<88>();
  if (warp == ATT_WARP_TMA) {
    if (lane == 0) {
      // ---------------- TMA producer: Q (both tiles), then K0 V0 K1 V1 ...
      mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      for (int i = 0; i < 2; ++i)
        for (int hf = 0; hf < 2; ++hf)
          tma_load_3d(smem + Cfg::Q_OFF + i * ATT_TILE_BYTES + hf * 16384, &tmap_q, q_full, hf * 64, q0 + i * 128, bh);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < 2 * T; ++t) {
        const CUtensorMap* m = (t & 1) ? &tmap_v : &tmap_k;
        const int row = (t >> 1) * 128;
        mbar_wait(&kv_empty[stage], phase ^ 1);
#ifdef FX_ATTN_NOTMA  // timing ablation: K/V tiles are loaded once, later ring slots are only re-armed
        if (t >= NS) {
          mbar_arrive(&kv_full[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
          continue;
        }
#endif
        mbar_arrive_expect_tx(&kv_full[stage], ATT_TILE_BYTES);
        uint8_t* dst = smem + Cfg::KV_OFF + stage * ATT_TILE_BYTES;
        tma_load_3d(dst, m, &kv_full[stage], 0, row, bh);
        tma_load_3d(dst + 16384, m, &kv_full[stage], 64, row, bh);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == ATT_WARP_MMA) {
    // ---------------- MMA issuer.  The whole warp runs this loop with warp-uniform control flow and one
    // elected lane issues the tcgen05 instructions: the compiler then keeps descriptors in uniform registers
    // (no per-instruction ELECT / BRA.U.ANY loop), and every descriptor is "precomputed low word + constant".
    // The issuing thread was the kernel's bottleneck before this (probe: ~105 clocks per MMA against a
    // 64-clock tensor-pipe floor, P always ready before the issuer asked for it).
    {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);  // B (= V) is MN-major
      const uint32_t q_lo = (smem_u32(smem + Cfg::Q_OFF) >> 4) | (1u << 16);      // K-major: LBO field 1
      const uint32_t k_lo0 = (smem_u32(smem + Cfg::KV_OFF) >> 4) | (1u << 16);
      const uint32_t v_lo0 = (smem_u32(smem + Cfg::KV_OFF) >> 4) | (1024u << 16);  // MN-major: LBO = 16384 B
      const uint32_t p_lo = (smem_u32(smem + Cfg::P_OFF) >> 4) | (1u << 16);
      constexpr uint32_t TILE16 = ATT_TILE_BYTES >> 4;
      int stage = 0;
      uint32_t phase = 0;
      auto issue_qk = [&](int i, int kslot) {
#ifdef FX_ATTN_NOQK
        return;
#endif
        const uint32_t a_lo = q_lo + i * TILE16, b_lo = k_lo0 + kslot * TILE16;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 1024 + (ks & 3) * 2;  // (half * 16384 + k16 * 32) >> 4
          umma_ss(tmem + i * 128, make_desc(a_lo + off, kDescHiSw128), make_desc(b_lo + off, kDescHiSw128), idesc_qk, ks != 0);
        }
      };
      // the P.V product is issued in two halves of 64 keys: the first four MMAs start as soon as the first half
      // of P exists, while the softmax warps are still exponentiating the second half
      auto issue_pv = [&](int i, int vslot, bool acc, int hf) {
#ifdef FX_ATTN_NOPV
        return;
#endif
        const uint32_t b_lo = v_lo0 + vslot * TILE16;
#pragma unroll
        for (int ks = hf * 4; ks < hf * 4 + 4; ++ks) {
          const uint64_t vd = make_desc(b_lo + ks * 128, kDescHiSw128);  // 16 keys = 2048 B
          if (P_TMEM) {
            umma_ts(tmem + 256 + i * 128, tmem + i * 128 + ks * 8, vd, idesc_pv, (acc || ks != 0) ? 1u : 0u);
          } else {
            const uint32_t off = (ks >> 2) * 1024 + (ks & 3) * 2;
            umma_ss(tmem + 256 + i * 128, make_desc(p_lo + i * TILE16 + off, kDescHiSw128), vd, idesc_pv,
                    (acc || ks != 0) ? 1u : 0u);
          }
        }
      };
      MMA_WAIT(q_full, 0);
      MMA_WAIT(&kv_full[stage], phase);  // K(0)
      tc_fence_after();
      if (elect_one()) {
        issue_qk(0, stage);
        tc_commit(&s_full[0]);
        issue_qk(1, stage);
        tc_commit(&s_full[1]);
        tc_commit(&kv_empty[stage]);
      }
      __syncwarp();
      if (++stage == NS) { stage = 0; phase ^= 1; }
      for (int j = 0; j < T; ++j) {
        const bool more = (j + 1 < T);
        const int vs = stage;
        MMA_WAIT(&kv_full[vs], phase);  // V(j)
        PROBE(0, j, 0);
        if (++stage == NS) { stage = 0; phase ^= 1; }
        const int ks_ = stage;
        if (more) MMA_WAIT(&kv_full[ks_], phase);  // K(j+1): landed long ago (it trails V(j) in the ring)
        P_WAIT(&p_full[0], j & 1);
        PROBE(0, j, 1);
        tc_fence_after();
        if (elect_one()) issue_pv(0, vs, j > 0, 0);
        __syncwarp();
        P_WAIT(&p_full[1], j & 1);
        PROBE(0, j, 2);
        PROBE(0, j, 3);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(0, vs, j > 0, 1);
          if (more) {
            issue_qk(0, ks_);
            tc_commit(&s_full[0]);
          }
        }
        __syncwarp();
        P_WAIT(&p_full[2], j & 1);
        PROBE(0, j, 4);
        tc_fence_after();
        if (elect_one()) issue_pv(1, vs, j > 0, 0);
        __syncwarp();
        P_WAIT(&p_full[3], j & 1);
        PROBE(0, j, 5);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(1, vs, j > 0, 1);
          tc_commit(&kv_empty[vs]);
          if (more) {
            issue_qk(1, ks_);
            tc_commit(&s_full[1]);
            tc_commit(&kv_empty[ks_]);
          }
        }
        __syncwarp();
        if (more) {
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        PROBE(0, j, 6);
      }
      if (elect_one()) {
        tc_commit(&o_full[0]);
        tc_commit(&o_full[1]);
      }
      __syncwarp();
    }
  }
  } else {
    // ---------------- softmax / correction / epilogue warpgroups
    reg_inc<208>();
    const int i = (warp - ATT_SM0) >> 2;  // query tile 0/1
    const int quarter = warp & 3;    // TMEM lane quarter
    const int r = quarter * 32 + lane;
    const int q_row = q0 + i * 128 + r;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    const uint32_t s_addr = tmem + lane_base + i * 128;
    const uint32_t o_addr = tmem + lane_base + 256 + i * 128;
    uint8_t* p_smem = smem + Cfg::P_OFF + i * ATT_TILE_BYTES;
    const float sl2 = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;

    const uint64_t sl2_2 = pack2(sl2, sl2);
#ifdef FX_ATTN_PROBE
    const bool probe_on_sm = probe_on && quarter == 0;
#define SPROBE(j, ev) do { if (probe_on_sm && (j) < 64) g_attn_probe[((1 + i) * 64 + (j)) * 8 + (ev)] = probe_clock(); } while (0)
#else
#define SPROBE(j, ev) do {} while (0)
#endif
#ifdef FX_ATTN_MMAONLY
#ifdef FX_ATTN_FREELD  // contention generator: unsynchronised TMEM reads of S (and writes of P) while the MMAs run
    for (int j = 0; j < T; ++j) {
      uint32_t sv[128];
      __syncwarp();
      tmem_ld_x32(s_addr, sv);
      tmem_ld_x32(s_addr + 32, sv + 32);
      tmem_ld_x32(s_addr + 64, sv + 64);
      tmem_ld_x32(s_addr + 96, sv + 96);
      tmem_ld_wait();
      uint32_t acc = 0;
#pragma unroll
      for (int e = 0; e < 128; ++e) acc ^= sv[e];
      if (acc == 0x12345u) l_run += 1.f;
#ifdef FX_ATTN_FREEST
      tmem_st_x16(s_addr, sv);
      tmem_st_x16(s_addr + 16, sv + 16);
      tmem_st_x16(s_addr + 32, sv + 32);
      tmem_st_x16(s_addr + 48, sv + 48);
      tmem_st_wait();
#endif
      for (int w = 0; w < FX_ATTN_FREELD; ++w) __nanosleep(100);
    }
#endif
    for (int j = 0; j < 0; ++j) {
#else
    for (int j = 0; j < T; ++j) {
#endif
      const int kv_valid = min(128, p.seq - j * 128);
      mbar_wait(&s_full[i], j & 1);
      SPROBE(j, 0);
      tc_fence_after();
      // the whole S row (128 fp32) lives in registers: one TMEM read per tile
      uint32_t sv[128];
      __syncwarp();
      tmem_ld_x32(s_addr, sv);
      tmem_ld_x32(s_addr + 32, sv + 32);
      tmem_ld_x32(s_addr + 64, sv + 64);
      tmem_ld_x32(s_addr + 96, sv + 96);
      tmem_ld_wait();
      SPROBE(j, 1);
      if (kv_valid < 128) {  // ragged last tile: keys beyond seq do not exist
#pragma unroll
        for (int e = 0; e < 128; ++e)
          if (e >= kv_valid) sv[e] = 0xff800000u;  // -inf
      }
      // p = exp2(s * scale_log2 - m_run) for one 32-key chunk, packed two at a time
      auto exp_chunk = [&](int c, uint32_t* dst, uint32_t* pk, uint64_t nm2) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const uint64_t t2 = ffma2(pack2(__uint_as_float(sv[c * 32 + e]), __uint_as_float(sv[c * 32 + e + 1])), sl2_2, nm2);
          float p0, p1;
          if (EMU_MASK & (1u << ((e >> 1) & 7))) {  // compile-time pattern: which pairs go to the FMA pipe
            exp2_emu2(t2, p0, p1);
          } else {
            const float2 t = unpack2(t2);
            p0 = fast_exp2(t.x);
            p1 = fast_exp2(t.y);
          }
          dst[e] = __float_as_uint(p0);
          dst[e + 1] = __float_as_uint(p1);
          pk[e >> 1] = pack_bf16(p0, p1);
        }
      };
      // The running maximum is LAZY (it only moves when a tile exceeds it by more than 2^8), so the first
      // chunk's exponentials are issued speculatively against the current m_run while the row maximum of the
      // tile is still being reduced on the ALU pipe: the ~400-clock max chain leaves the S -> P -> PV critical
      // path.  Only when some row's maximum did grow (the first tiles of a row, then almost never) is the
      // chunk recomputed after the rescale.  The two warpgroups take turns on the exp phase (the MUFU of a
      // sub-partition serves one warp of each); the row sum is accumulated AFTER P has been handed over.
      SPROBE(j, 2);
      if (p.sequence) mbar_wait(&seq_bar[i], (i == 0) ? ((j & 1) ^ 1) : (j & 1));
      SPROBE(j, 3);
      uint64_t nm2 = pack2(-m_run, -m_run);
      uint32_t pa[32], pk0[16];
#ifndef FX_ATTN_NOSPEC
      exp_chunk(0, pa, pk0, nm2);
#endif
#ifdef FX_ATTN_NOMAX  // timing ablation only (wrong numerics)
      float mx = __uint_as_float(sv[0]);
#else
      float mx = fmaxf(__uint_as_float(sv[0]), __uint_as_float(sv[1]));
      float mxb = fmaxf(__uint_as_float(sv[2]), __uint_as_float(sv[3]));
#pragma unroll
      for (int e = 4; e < 128; e += 4) {  // two independent chains
        mx = fmax3(mx, __uint_as_float(sv[e]), __uint_as_float(sv[e + 1]));
        mxb = fmax3(mxb, __uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3]));
      }
      mx = fmaxf(mx, mxb);
#endif
      const float m_new = fmaxf(m_run, mx * sl2);
      const bool need = (m_new - m_run) > 8.0f;
      bool redo = false;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? fast_exp2(m_run - m_new) : 1.0f;
        if (need) {
          m_run = m_new;
          l_run *= alpha;
        }
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld_x32(o_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
            tmem_st_x32(o_addr + c * 32, v);
          }
          tmem_st_wait();
        }
        nm2 = pack2(-m_run, -m_run);
        redo = true;
      }
#ifdef FX_ATTN_NOSPEC
      exp_chunk(0, pa, pk0, nm2);
#else
      if (redo) exp_chunk(0, pa, pk0, nm2);
#endif
#pragma unroll
      for (int e = 0; e < 32; ++e) sv[e] = pa[e];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
        if (c == 0) {
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = pk0[e];
        } else {
          exp_chunk(c, sv + c * 32, pk, nm2);
        }
        if (P_TMEM) {
          tmem_st_x16(s_addr + c * 16, pk);
        } else {
          // K-major SW128 tile: row r, keys c*32 .. c*32+31 -> half c>>1, 16-byte chunks ((c&1)*4 + q) ^ (r&7)
          uint8_t* rowp = p_smem + (c >> 1) * 16384 + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
            *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          }
        }
        if (c & 1) {  // a 64-key half of P is complete: hand it to the MMA warp
          if (c == 3 && p.sequence) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&seq_bar[i ^ 1]);  // the other warpgroup's turn on the MUFU
          }
          if (P_TMEM) {
            tmem_st_wait();
            tc_fence_before();
          } else {
            fence_proxy_async_smem();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[2 * i + (c >> 1)]);
          SPROBE(j, 4 + (c >> 1));
        }
      }
#ifdef FX_ATTN_NOSUM  // timing ablation only (wrong numerics)
      l_run += __uint_as_float(sv[0]);
#else
      {
        uint64_t ls0 = pack2(0.f, 0.f), ls1 = pack2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 128; e += 4) {
          ls0 = fadd2(ls0, pack2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])));
          ls1 = fadd2(ls1, pack2(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])));
        }
        const float2 ls = unpack2(fadd2(ls0, ls1));
        l_run += ls.x + ls.y;
      }
#endif
      SPROBE(j, 6);
    }

    // epilogue: O / l -> bf16 -> out[b][q_row][h*128 ...].  Every MMA has completed (o_full), so the K/V ring is free:
    // it serves as the per-warp transposition buffer for coalesced stores (a thread owns a row: see sm100.cuh)
    mbar_wait(&o_full[i], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    uint8_t* wst = smem + Cfg::KV_OFF + (warp - ATT_SM0) * 2048;
    const uint32_t vmask = __ballot_sync(0xffffffffu, q_row < p.seq);
    __nv_bfloat16* dst0 = p.out + (long long)blockIdx.z * p.out_bs + (long long)(q_row - lane) * p.ld_out + blockIdx.y * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      __syncwarp();
      tmem_ld_x32(o_addr + c * 32, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]) * inv_l;
      store_chunk32_coalesced(wst, lane, f, dst0 + c * 32, p.ld_out, vmask);
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef FX_ATTN_PROBE
  if (probe_on && warp == ATT_WARP_MMA) {
    g_attn_probe[(0 * 64 + 63) * 8 + 1] = probe_clock();
    g_attn_probe[(0 * 64 + 63) * 8 + 3] = probe_gtime();
  }
#endif
  if (warp == ATT_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------
// attn_pkernel (variant 7): attn_kernel's schedule as a PERSISTENT work loop.  One CTA per SM walks the work items
// (256 query rows of one (batch, head)), q-block fastest so that the CTAs running at the same time share K / V in L2.
// What the loop buys over one CTA per item (measured on attn_kernel: ~4100 clocks of prologue per ~105 000-clock item):
//   * barrier initialisation, TMEM allocation and the CTA-wide syncs happen once per SM, not once per item;
//   * the producer runs ahead across the item boundary: Q of item n+1 is loaded as soon as the last Q K^T of item n has
//     completed (q_empty), its K(0) / V(0) are already in the ring, and the issuer starts S(0) of item n+1 right behind
//     the last P.V of item n -- so when the softmax warps return from the O epilogue their first S tile is waiting.
// Differences in resources: the K/V ring has 4 stages instead of 5 (the O epilogue can no longer borrow the ring as its
// transposition buffer -- it is full of the next item's tiles -- and gets 16 KB of its own).
// Barrier phases are tracked with running counters (n = tile steps so far), so any kv_tiles parity works.
// ------------------------------------------------------------------------------------------
// F8 (--quantize): q, k, v are e4m3 (written by the FP8 QKV epilogue) and P is handed to the P.V product as e4m3:
// both products run on tcgen05.mma.kind::f8f6f4 (K = 32 per instruction: 4 instead of 8 MMAs per 128 x 128 x 128
// product), a tile is 16 KB (ONE 128-byte-wide swizzle atom: 128 e4m3 per row) and P takes 32 TMEM columns.  The per-tile
// dependency loop of a query tile is  softmax -> P.V -> Q K^T  (S and P alias in TMEM), so halving the tensor time of
// the two products shortens the loop itself, not just the tensor pipe's share.  P is scaled by 2^P_SHIFT before the
// conversion (e4m3 resolves 2^-9 .. 448; the row sum l is accumulated from the same scaled fp32 values, so O / l is
// unaffected) and the lazy running maximum may lag by at most 2^LAZY: P_SHIFT + LAZY <= 8 keeps every p representable.
template <bool F8>
struct AttnPCfg {
  static constexpr int TILE_BYTES = F8 ? ATT_TILE_BYTES / 2 : ATT_TILE_BYTES;
  static constexpr int KV_STAGES = F8 ? 8 : 4;
  static constexpr int Q_OFF = 0;
  static constexpr int KV_OFF = 2 * TILE_BYTES;
  static constexpr int ST_OFF = KV_OFF + KV_STAGES * TILE_BYTES;  // 8 warps x 2 KB store transposition buffers
  static constexpr int BAR_OFF = ST_OFF + 8 * 2048;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
};
constexpr float ATT_F8_P_SHIFT = 4.0f, ATT_F8_LAZY = 4.0f;

// four fp32 -> four e4m3 bytes (round to nearest, saturating), element 0 in the low byte
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
  return uint32_t(lo) | (uint32_t(hi) << 16);
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// A (= P, e4m3) from TMEM x B (= V, e4m3, shared memory): kind::f8f6f4
__device__ __forceinline__ void umma_ts_f8(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool F8>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_pkernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
             const __grid_constant__ CUtensorMap tmap_v, const AttnParams p, const int q_blocks, const int items) {
  using Cfg = AttnPCfg<F8>;
  constexpr int TILE = Cfg::TILE_BYTES;
  constexpr int NS = Cfg::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* q_full = bars;             // 1: Q of the current item has landed
  uint64_t* q_empty = bars + 1;        // 1: every Q K^T of the current item has completed (tcgen05.commit)
  uint64_t* kv_full = bars + 2;        // NS
  uint64_t* kv_empty = kv_full + NS;   // NS
  uint64_t* s_full = kv_empty + NS;    // 2
  uint64_t* p_full = s_full + 2;       // [tile][half] = 4
  uint64_t* o_full = p_full + 4;       // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.kv_tiles;

  if (warp == ATT_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 4);
      mbar_init(&p_full[2 * i + 1], 4);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == ATT_WARP_MMA) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_i aliases S_i[0,64)

  if (warp >= ATT_CTRL0 && warp < ATT_CTRL0 + 4) {
    reg_dec<88>();
    if (warp == ATT_WARP_TMA) {
      if (lane == 0) {
        // ---------------- TMA producer: per item Q (both tiles), then K0 V0 K1 V1 ... through ONE ring across items
        int stage = 0;
        uint32_t phase = 0;
        uint32_t it = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
          const int qb = w % q_blocks, bh = w / q_blocks;
          const int q0 = qb * 256;
          mbar_wait(q_empty, (it & 1) ^ 1);  // the previous item's Q K^T products are done with the Q buffer
          mbar_arrive_expect_tx(q_full, 2 * TILE);
          for (int i = 0; i < 2; ++i) {
            if (F8) {  // one 128-byte-wide atom per tile
              tma_load_3d(smem + Cfg::Q_OFF + i * TILE, &tmap_q, q_full, 0, q0 + i * 128, bh);
            } else {
              for (int hf = 0; hf < 2; ++hf)
                tma_load_3d(smem + Cfg::Q_OFF + i * TILE + hf * 16384, &tmap_q, q_full, hf * 64, q0 + i * 128, bh);
            }
          }
          for (int t = 0; t < 2 * T; ++t) {
            const CUtensorMap* m = (t & 1) ? &tmap_v : &tmap_k;
            const int row = (t >> 1) * 128;
            mbar_wait(&kv_empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&kv_full[stage], TILE);
            uint8_t* dst = smem + Cfg::KV_OFF + stage * TILE;
            tma_load_3d(dst, m, &kv_full[stage], 0, row, bh);
            if (!F8) tma_load_3d(dst + 16384, m, &kv_full[stage], 64, row, bh);
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == ATT_WARP_MMA) {
      // ---------------- MMA issuer (whole warp, warp-uniform control flow, one elected lane issues: see attn_kernel)
      constexpr uint32_t idesc_qk = F8 ? make_idesc_f8(128, 128) : make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = F8 ? (make_idesc_f8(128, 128) | (1u << 16)) : make_idesc_bf16(128, 128, 0, 1);  // B (= V) is MN-major
      const uint32_t q_lo = (smem_u32(smem + Cfg::Q_OFF) >> 4) | (1u << 16);
      const uint32_t k_lo0 = (smem_u32(smem + Cfg::KV_OFF) >> 4) | (1u << 16);
      const uint32_t v_lo0 = (smem_u32(smem + Cfg::KV_OFF) >> 4) | (1024u << 16);
      constexpr uint32_t TILE16 = TILE >> 4;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t n = 0;   // tile steps so far (parity of the s_full / p_full phases)
      uint32_t it = 0;  // items so far (parity of q_full / o_full)
      auto issue_qk = [&](int i, int kslot) {
        const uint32_t a_lo = q_lo + i * TILE16, b_lo = k_lo0 + kslot * TILE16;
        if (F8) {  // 4 x (K = 32 e4m3 = 32 bytes of the 128-byte rows)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss_f8(tmem + i * 128, make_desc(a_lo + ks * 2, kDescHiSw128), make_desc(b_lo + ks * 2, kDescHiSw128), idesc_qk, ks != 0);
        } else {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off = (ks >> 2) * 1024 + (ks & 3) * 2;
            umma_ss(tmem + i * 128, make_desc(a_lo + off, kDescHiSw128), make_desc(b_lo + off, kDescHiSw128), idesc_qk, ks != 0);
          }
        }
      };
      auto issue_pv = [&](int i, int vslot, bool acc, int hf) {
        const uint32_t b_lo = v_lo0 + vslot * TILE16;
        if (F8) {  // a 64-key half = 2 x (K = 32 keys): P columns ks * 8, V rows ks * 32 (4096 B)
#pragma unroll
          for (int ks = hf * 2; ks < hf * 2 + 2; ++ks)
            umma_ts_f8(tmem + 256 + i * 128, tmem + i * 128 + ks * 8, make_desc(b_lo + ks * 256, kDescHiSw128), idesc_pv,
                       (acc || ks != 0) ? 1u : 0u);
        } else {
#pragma unroll
          for (int ks = hf * 4; ks < hf * 4 + 4; ++ks)
            umma_ts(tmem + 256 + i * 128, tmem + i * 128 + ks * 8, make_desc(b_lo + ks * 128, kDescHiSw128), idesc_pv,
                    (acc || ks != 0) ? 1u : 0u);
        }
      };
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        mbar_wait(q_full, it & 1);
        mbar_wait(&kv_full[stage], phase);  // K(0)
        tc_fence_after();
        if (elect_one()) {
          issue_qk(0, stage);
          tc_commit(&s_full[0]);
          issue_qk(1, stage);
          tc_commit(&s_full[1]);
          tc_commit(&kv_empty[stage]);
          if (T == 1) tc_commit(q_empty);
        }
        __syncwarp();
        if (++stage == NS) { stage = 0; phase ^= 1; }
        for (int j = 0; j < T; ++j, ++n) {
          const bool more = (j + 1 < T);
          const uint32_t par = n & 1;
          const int vs = stage;
          mbar_wait(&kv_full[vs], phase);  // V(j)
          if (++stage == NS) { stage = 0; phase ^= 1; }
          const int ks_ = stage;
          if (more) mbar_wait(&kv_full[ks_], phase);  // K(j+1): landed long ago (it trails V(j) in the ring)
          mbar_wait(&p_full[0], par);
          tc_fence_after();
          if (elect_one()) issue_pv(0, vs, j > 0, 0);
          __syncwarp();
          mbar_wait(&p_full[1], par);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(0, vs, j > 0, 1);
            if (more) {
              issue_qk(0, ks_);
              tc_commit(&s_full[0]);
            }
          }
          __syncwarp();
          mbar_wait(&p_full[2], par);
          tc_fence_after();
          if (elect_one()) issue_pv(1, vs, j > 0, 0);
          __syncwarp();
          mbar_wait(&p_full[3], par);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(1, vs, j > 0, 1);
            tc_commit(&kv_empty[vs]);
            if (more) {
              issue_qk(1, ks_);
              tc_commit(&s_full[1]);
              tc_commit(&kv_empty[ks_]);
              if (j + 2 == T) tc_commit(q_empty);  // that was the item's last Q K^T: the Q buffer may be refilled
            }
          }
          __syncwarp();
          if (more) {
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) {
          tc_commit(&o_full[0]);
          tc_commit(&o_full[1]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue warpgroups
    reg_inc<208>();
    const int i = (warp - ATT_SM0) >> 2;  // query tile 0/1
    const int quarter = warp & 3;         // TMEM lane quarter
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    const uint32_t s_addr = tmem + lane_base + i * 128;
    const uint32_t o_addr = tmem + lane_base + 256 + i * 128;
    const float sl2 = p.scale_log2;
    const uint64_t sl2_2 = pack2(sl2, sl2);
    uint8_t* wst = smem + Cfg::ST_OFF + (warp - ATT_SM0) * 2048;
    uint32_t n = 0, it = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      const int qb = w % q_blocks, bh = w / q_blocks;
      const int bz = bh / p.heads, hy = bh - bz * p.heads;
      const int q_row = qb * 256 + i * 128 + r;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < T; ++j, ++n) {
        const int kv_valid = min(128, p.seq - j * 128);
        mbar_wait(&s_full[i], n & 1);
        tc_fence_after();
        uint32_t sv[128];
        __syncwarp();
        tmem_ld_x32(s_addr, sv);
        tmem_ld_x32(s_addr + 32, sv + 32);
        tmem_ld_x32(s_addr + 64, sv + 64);
        tmem_ld_x32(s_addr + 96, sv + 96);
        tmem_ld_wait();
        if (kv_valid < 128) {  // ragged last tile: keys beyond seq do not exist
#pragma unroll
          for (int e = 0; e < 128; ++e)
            if (e >= kv_valid) sv[e] = 0xff800000u;  // -inf
        }
        auto exp_chunk = [&](int c, uint32_t* dst, uint32_t* pk, uint64_t nm2) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const uint64_t t2 = ffma2(pack2(__uint_as_float(sv[c * 32 + e]), __uint_as_float(sv[c * 32 + e + 1])), sl2_2, nm2);
            float p0, p1;
            if ((F8 ? EMU_MASK_F8 : EMU_MASK) & (1u << ((e >> 1) & 7))) {
#ifdef FX_ATTN_EMU_F8_CUBIC
              exp2_emu2(t2, p0, p1);
#else
              if (F8) exp2_emu2_quad(t2, p0, p1);
              else exp2_emu2(t2, p0, p1);
#endif
            } else {
              const float2 t = unpack2(t2);
              p0 = fast_exp2(t.x);
              p1 = fast_exp2(t.y);
            }
            dst[e] = __float_as_uint(p0);
            dst[e + 1] = __float_as_uint(p1);
            if (!F8) pk[e >> 1] = pack_bf16(p0, p1);
          }
          if (F8) {  // 32 probabilities -> 8 words of four e4m3
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              pk[e >> 2] = pack_e4m3x4(__uint_as_float(dst[e]), __uint_as_float(dst[e + 1]), __uint_as_float(dst[e + 2]),
                                       __uint_as_float(dst[e + 3]));
          }
        };
        // F8: p = 2^(s - m_run + P_SHIFT); the running maximum may lag the true one by at most 2^LAZY
        constexpr float P_SHIFT = F8 ? ATT_F8_P_SHIFT : 0.0f, LAZY = F8 ? ATT_F8_LAZY : 8.0f;
        // speculative first chunk against the lazy running max (see attn_kernel)
        uint64_t nm2 = pack2(P_SHIFT - m_run, P_SHIFT - m_run);
        uint32_t pa[32], pk0[16];
        exp_chunk(0, pa, pk0, nm2);
        float mx = fmaxf(__uint_as_float(sv[0]), __uint_as_float(sv[1]));
        float mxb = fmaxf(__uint_as_float(sv[2]), __uint_as_float(sv[3]));
#pragma unroll
        for (int e = 4; e < 128; e += 4) {
          mx = fmax3(mx, __uint_as_float(sv[e]), __uint_as_float(sv[e + 1]));
          mxb = fmax3(mxb, __uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3]));
        }
        mx = fmaxf(mx, mxb);
        const float m_new = fmaxf(m_run, mx * sl2);
        const bool need = (m_new - m_run) > LAZY;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? fast_exp2(m_run - m_new) : 1.0f;
          if (need) {
            m_run = m_new;
            l_run *= alpha;
          }
          if (j > 0) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t v[32];
              __syncwarp();
              tmem_ld_x32(o_addr + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
              tmem_st_x32(o_addr + c * 32, v);
            }
            tmem_st_wait();
          }
          nm2 = pack2(P_SHIFT - m_run, P_SHIFT - m_run);
          exp_chunk(0, pa, pk0, nm2);
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) sv[e] = pa[e];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
          if (c == 0) {
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = pk0[e];
          } else {
            exp_chunk(c, sv + c * 32, pk, nm2);
          }
          if (F8) tmem_st_x8(s_addr + c * 8, pk);
          else tmem_st_x16(s_addr + c * 16, pk);
          if (c & 1) {  // a 64-key half of P is complete: hand it to the MMA warp
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[2 * i + (c >> 1)]);
          }
        }
        {
          uint64_t ls0 = pack2(0.f, 0.f), ls1 = pack2(0.f, 0.f);
#pragma unroll
          for (int e = 0; e < 128; e += 4) {
            ls0 = fadd2(ls0, pack2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])));
            ls1 = fadd2(ls1, pack2(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])));
          }
          const float2 ls = unpack2(fadd2(ls0, ls1));
          l_run += ls.x + ls.y;
        }
      }
      // epilogue of the item: O / l -> bf16 -> out[b][q_row][h*128 ...] (coalesced through this warp's own buffer)
      mbar_wait(&o_full[i], it & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      const uint32_t vmask = __ballot_sync(0xffffffffu, q_row < p.seq);
      __nv_bfloat16* dst0 = p.out + (long long)bz * p.out_bs + (long long)(q_row - lane) * p.ld_out + hy * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_x32(o_addr + c * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]) * inv_l;
        if (p.out4[0].q != nullptr) {
          const bool hi = q_row >= p.out4_split;
          ChunkQ o;  // (static indices: a dynamic index into a kernel parameter array forces a local-memory copy)
          o.q = hi ? p.out4[0].q : p.out4[1].q;
          o.sf = hi ? p.out4[0].sf : p.out4[1].sf;
          o.e = hi ? p.out4[0].e : p.out4[1].e;
          o.kc = hi ? p.out4[0].kc : p.out4[1].kc;
          o.col0 = hi ? p.out4[0].col0 : 0;
          const long long m = hi ? (long long)bz * (p.seq - p.out4_split) + (q_row - p.out4_split) : (long long)bz * p.out4_split + q_row;
          fp4_chunk_quantise(f, m, o.col0 + hy * 128 + c * 32, o, q_row < p.seq);
        } else {
          store_chunk32_coalesced(wst, lane, f, dst0 + c * 32, p.ld_out, vmask);
        }
      }
      // O has been read out of TMEM (tcgen05.wait::ld above): order it before the next item's first p_full arrive, which
      // is what allows the issuer to overwrite O with P.V of the next item
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ATT_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------
// attn3_kernel (variants 5 / 6; measured, NOT the default): same math and roles as attn_kernel, but the per-tile
// dependency loop is cut.
//   attn_kernel hands P to the P.V MMA through TMEM, aliasing S: S_i(j+1) = Q_i K(j+1)^T cannot be issued before
//   P_i(j) V(j) has consumed P_i(j), so every query tile runs the chain  S -> softmax -> P -> PV -> QK -> S  (probe:
//   ~2100 + ~1100 clocks per key tile against a 2048-clock tensor floor for both tiles together).
//   Here P goes through shared memory, so S_i is free again as soon as the softmax warps have READ it (~120 clocks
//   after it was ready): the issuer starts Q_i K(j+1)^T right then and S_i(j+1) is waiting when the softmax of step j
//   ends -- the warpgroups never wait for the tensor pipe, the tensor pipe only for P.
//   Shared memory: Q 64 KB + ONE 32 KB P buffer (two 64-key half slots that the two query tiles take turns on:
//   tile i writes half h after tile 1-i's P.V of that half has completed) + a 4-stage K/V ring = 224 KB.
//   TMEM: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512).
//   Result (probe, SM clocks per key tile): 3880 with / 3707 without turn taking against attn_kernel's 3300 / 3230: the
//   softmax warps no longer wait for S, but their exp phase grows from ~1730 to ~2600 clocks (P through st.shared +
//   proxy fence + half-slot waits, both warpgroups now always contending for the same sub-partition), and the in-order
//   issuer idles the tensor pipe between Q K(j+1)^T and the P it then waits for.  With two softmax warps per
//   sub-partition the kernel is bound by the serial latency of the softmax instruction stream, not by the S -> P -> PV
//   loop.  Two more variants were built, measured and removed again: attn2_kernel (decoupled 64-key steps: -15 %, N = 64
//   MMAs double the operand traffic) and attn4_kernel (two threads per query row = 16 softmax warps, partial maxima
//   exchanged through shared memory behind 64-thread named barriers: correct, 3551 clocks -- the S -> P time per tile
//   stayed at ~2100 clocks although each thread did half the work, i.e. the sub-partition, not the warp, is the limit;
//   lesson kept: setmaxnreg.inc draws from the register pool the CTA was LAUNCHED with and blocks forever beyond it).
// ------------------------------------------------------------------------------------------
struct Attn3Cfg {
  static constexpr int KV_STAGES = 4;
  static constexpr int Q_OFF = 0;
  static constexpr int P_OFF = 2 * ATT_TILE_BYTES;
  static constexpr int KV_OFF = 3 * ATT_TILE_BYTES;
  static constexpr int BAR_OFF = KV_OFF + KV_STAGES * ATT_TILE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
};

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
             const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  using Cfg = Attn3Cfg;
  constexpr int NS = Cfg::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* q_full = bars;              // 1
  uint64_t* kv_full = bars + 1;         // NS
  uint64_t* kv_empty = kv_full + NS;    // NS
  uint64_t* s_full = kv_empty + NS;     // [tile] 2: S_i(j) is in TMEM
  uint64_t* s_read = s_full + 2;        // [tile] 2: the softmax warps hold S_i(j) in registers (4 warp arrivals)
  uint64_t* p_full = s_read + 2;        // [tile][half] 4: P_i(j) half h is in shared memory (4 warp arrivals)
  uint64_t* p_free = p_full + 4;        // [tile][half] 4: P_i(j) V(j) of half h has completed (tcgen05.commit)
  uint64_t* pv_done = p_free + 4;       // [tile] 2: O_i is quiescent after step j (rare rescale path)
  uint64_t* o_full = pv_done + 2;       // [tile] 2
  uint64_t* seq_bar = o_full + 2;       // [tile] 2: exp-phase turn taking
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(seq_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int bh = blockIdx.z * p.heads + blockIdx.y;
  const int T = p.kv_tiles;
#ifdef FX_ATTN_PROBE
  const bool probe_on = g_attn_probe && blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 1 && lane == 0;
#endif

  if (warp == ATT_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_read[i], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&seq_bar[i], 4);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&p_full[i], 4);
      mbar_init(&p_free[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == ATT_WARP_MMA) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= ATT_CTRL0 && warp < ATT_CTRL0 + 4) {
    reg_dec<88>();
    if (warp == ATT_WARP_TMA) {
      if (lane == 0) {
        // ---------------- TMA producer: Q (both tiles), then K0 V0 K1 V1 ... through one ring.  Consumption order is
        // K(j+1) early in step j, V(j) late in step j; every load is still issued a full step before its use.
        mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
        for (int i = 0; i < 2; ++i)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(smem + Cfg::Q_OFF + i * ATT_TILE_BYTES + hf * 16384, &tmap_q, q_full, hf * 64, q0 + i * 128, bh);
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < 2 * T; ++t) {
          const CUtensorMap* m = (t & 1) ? &tmap_v : &tmap_k;
          const int row = (t >> 1) * 128;
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[stage], ATT_TILE_BYTES);
          uint8_t* dst = smem + Cfg::KV_OFF + stage * ATT_TILE_BYTES;
          tma_load_3d(dst, m, &kv_full[stage], 0, row, bh);
          tma_load_3d(dst + 16384, m, &kv_full[stage], 64, row, bh);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == ATT_WARP_MMA) {
      // ---------------- MMA issuer (warp-uniform control flow, one elected lane issues)
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);  // B (= V) is MN-major
      constexpr uint32_t TILE16 = ATT_TILE_BYTES >> 4;
      const uint32_t q_lo = (smem_u32(smem + Cfg::Q_OFF) >> 4) | (1u << 16);
      const uint32_t p_lo = (smem_u32(smem + Cfg::P_OFF) >> 4) | (1u << 16);
      const uint32_t k_lo0 = (smem_u32(smem + Cfg::KV_OFF) >> 4) | (1u << 16);
      const uint32_t v_lo0 = (smem_u32(smem + Cfg::KV_OFF) >> 4) | (1024u << 16);
      // ring slot / parity of load number t (K(j) is load 2j, V(j) is load 2j+1)
      auto slot_of = [](int t) { return t & (NS - 1); };
      auto par_of = [](int t) { return uint32_t(t / NS) & 1u; };
      auto issue_qk = [&](int i, int kslot) {
        const uint32_t a_lo = q_lo + i * TILE16, b_lo = k_lo0 + kslot * TILE16;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 1024 + (ks & 3) * 2;
          umma_ss(tmem + i * 128, make_desc(a_lo + off, kDescHiSw128), make_desc(b_lo + off, kDescHiSw128), idesc_qk, ks != 0);
        }
      };
      // O_i (+)= P(half hf: 64 keys, K-major 16 KB slot) . V(rows hf*64 .. of the tile in vslot)
      auto issue_pv = [&](int i, int vslot, bool acc, int hf) {
        const uint32_t a_lo = p_lo + hf * 1024, b_lo = v_lo0 + vslot * TILE16;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int ks = hf * 4 + kk;
          umma_ss(tmem + 256 + i * 128, make_desc(a_lo + kk * 2, kDescHiSw128), make_desc(b_lo + ks * 128, kDescHiSw128), idesc_pv,
                  (acc || ks != 0) ? 1u : 0u);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);  // K(0)
      tc_fence_after();
      if (elect_one()) {
        issue_qk(0, 0);
        tc_commit(&s_full[0]);
        issue_qk(1, 0);
        tc_commit(&s_full[1]);
        tc_commit(&kv_empty[0]);
      }
      __syncwarp();
      for (int j = 0; j < T; ++j) {
        const bool more = (j + 1 < T);
        const int tk = 2 * j + 2, tv = 2 * j + 1;  // load numbers of K(j+1) and V(j)
        const int ks_ = slot_of(tk), vs = slot_of(tv);
        const uint32_t ph = j & 1;
        if (more) mbar_wait(&kv_full[ks_], par_of(tk));
        PROBE(0, j, 0);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (more) {  // S_i(j) has been read: S_i(j+1) may overwrite it
            mbar_wait(&s_read[i], ph);
            tc_fence_after();
            if (elect_one()) {
              issue_qk(i, ks_);
              tc_commit(&s_full[i]);
              if (i == 1) tc_commit(&kv_empty[ks_]);
            }
            __syncwarp();
          }
          PROBE(0, j, 1 + 3 * i);
          mbar_wait(&p_full[2 * i], ph);
          if (i == 0) mbar_wait(&kv_full[vs], par_of(tv));  // V(j)
          PROBE(0, j, 2 + 3 * i);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(i, vs, j > 0, 0);
            tc_commit(&p_free[2 * i]);
          }
          __syncwarp();
          mbar_wait(&p_full[2 * i + 1], ph);
          PROBE(0, j, 3 + 3 * i);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(i, vs, j > 0, 1);
            tc_commit(&p_free[2 * i + 1]);
            tc_commit(&pv_done[i]);
            if (i == 1) tc_commit(&kv_empty[vs]);
          }
          __syncwarp();
        }
      }
      if (elect_one()) {
        tc_commit(&o_full[0]);
        tc_commit(&o_full[1]);
      }
      __syncwarp();
    }
  } else {
    // ---------------- softmax / correction / epilogue warpgroups
    reg_inc<208>();
    const int i = (warp - ATT_SM0) >> 2;  // query tile 0/1
    const int quarter = warp & 3;         // TMEM lane quarter
    const int r = quarter * 32 + lane;
    const int q_row = q0 + i * 128 + r;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    const uint32_t s_addr = tmem + lane_base + i * 128;
    const uint32_t o_addr = tmem + lane_base + 256 + i * 128;
    uint8_t* p_smem = smem + Cfg::P_OFF;
    const float sl2 = p.scale_log2;
    const uint64_t sl2_2 = pack2(sl2, sl2);
    float m_run = -INFINITY, l_run = 0.f;
#ifdef FX_ATTN_PROBE
    const bool probe_sm = probe_on && quarter == 0;
#define S3PROBE(j, ev) do { if (probe_sm && (j) < 64) g_attn_probe[((1 + i) * 64 + (j)) * 8 + (ev)] = probe_clock(); } while (0)
#else
#define S3PROBE(j, ev) do {} while (0)
#endif

    for (int j = 0; j < T; ++j) {
      const int kv_valid = min(128, p.seq - j * 128);
      mbar_wait(&s_full[i], j & 1);
      S3PROBE(j, 0);
      tc_fence_after();
      uint32_t sv[128];
      __syncwarp();
      tmem_ld_x32(s_addr, sv);
      tmem_ld_x32(s_addr + 32, sv + 32);
      tmem_ld_x32(s_addr + 64, sv + 64);
      tmem_ld_x32(s_addr + 96, sv + 96);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_read[i]);  // S_i is free: the issuer may start Q_i K(j+1)^T
      S3PROBE(j, 1);
      if (kv_valid < 128) {
#pragma unroll
        for (int e = 0; e < 128; ++e)
          if (e >= kv_valid) sv[e] = 0xff800000u;  // -inf
      }
      auto exp_chunk = [&](int c, uint32_t* dst, uint32_t* pk, uint64_t nm2) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const uint64_t t2 = ffma2(pack2(__uint_as_float(sv[c * 32 + e]), __uint_as_float(sv[c * 32 + e + 1])), sl2_2, nm2);
          float p0, p1;
          if (EMU_MASK & (1u << ((e >> 1) & 7))) {
            exp2_emu2(t2, p0, p1);
          } else {
            const float2 t = unpack2(t2);
            p0 = fast_exp2(t.x);
            p1 = fast_exp2(t.y);
          }
          dst[e] = __float_as_uint(p0);
          dst[e + 1] = __float_as_uint(p1);
          pk[e >> 1] = pack_bf16(p0, p1);
        }
      };
      if (p.sequence) mbar_wait(&seq_bar[i], (i == 0) ? ((j & 1) ^ 1) : (j & 1));
      S3PROBE(j, 2);
      uint64_t nm2 = pack2(-m_run, -m_run);
      uint32_t pa[32], pk0[16];
      exp_chunk(0, pa, pk0, nm2);  // speculative against the lazy running max (see attn_kernel)
      float mx = fmaxf(__uint_as_float(sv[0]), __uint_as_float(sv[1]));
      float mxb = fmaxf(__uint_as_float(sv[2]), __uint_as_float(sv[3]));
#pragma unroll
      for (int e = 4; e < 128; e += 4) {
        mx = fmax3(mx, __uint_as_float(sv[e]), __uint_as_float(sv[e + 1]));
        mxb = fmax3(mxb, __uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3]));
      }
      mx = fmaxf(mx, mxb);
      const float m_new = fmaxf(m_run, mx * sl2);
      const bool need = (m_new - m_run) > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? fast_exp2(m_run - m_new) : 1.0f;
        if (need) {
          m_run = m_new;
          l_run *= alpha;
        }
        if (j > 0) {
          // O_i must be quiescent: S_i(j) being ready only proves P_i(j-2) V(j-2) complete
          mbar_wait(&pv_done[i], (j - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld_x32(o_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
            tmem_st_x32(o_addr + c * 32, v);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        nm2 = pack2(-m_run, -m_run);
        exp_chunk(0, pa, pk0, nm2);
      }
#pragma unroll
      for (int e = 0; e < 32; ++e) sv[e] = pa[e];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
        if (c == 0) {
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = pk0[e];
        } else {
          exp_chunk(c, sv + c * 32, pk, nm2);
        }
        if ((c & 1) == 0) {
          // half slot c>>1 was last read by the OTHER tile's P.V: tile 0 waits for P_1(j-1) V, tile 1 for P_0(j) V
          const int hf = c >> 1;
          if (i == 0) {
            if (j > 0) mbar_wait(&p_free[2 + hf], (j - 1) & 1);
          } else {
            mbar_wait(&p_free[hf], j & 1);
          }
        }
        {
          // K-major SW128 half slot: row r, keys (c&1)*32 .. +31 -> 16-byte chunks ((c&1)*4 + q) ^ (r&7)
          uint8_t* rowp = p_smem + (c >> 1) * 16384 + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
            *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          }
        }
        if (c & 1) {
          if (c == 3 && p.sequence) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&seq_bar[i ^ 1]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[2 * i + (c >> 1)]);
          S3PROBE(j, 4 + (c >> 1));
        }
      }
      {
        uint64_t ls0 = pack2(0.f, 0.f), ls1 = pack2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 128; e += 4) {
          ls0 = fadd2(ls0, pack2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])));
          ls1 = fadd2(ls1, pack2(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])));
        }
        const float2 ls = unpack2(fadd2(ls0, ls1));
        l_run += ls.x + ls.y;
      }
      S3PROBE(j, 6);
    }

    // epilogue: O / l -> bf16 -> out[b][q_row][h*128 ...]
    mbar_wait(&o_full[i], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    __nv_bfloat16* dst = p.out + (long long)blockIdx.z * p.out_bs + (long long)q_row * p.ld_out + blockIdx.y * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      __syncwarp();
      tmem_ld_x32(o_addr + c * 32, v);
      tmem_ld_wait();
      if (q_row < p.seq) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[e]) * inv_l, __uint_as_float(v[e + 1]) * inv_l);
          u.y = pack_bf16(__uint_as_float(v[e + 2]) * inv_l, __uint_as_float(v[e + 3]) * inv_l);
          u.z = pack_bf16(__uint_as_float(v[e + 4]) * inv_l, __uint_as_float(v[e + 5]) * inv_l);
          u.w = pack_bf16(__uint_as_float(v[e + 6]) * inv_l, __uint_as_float(v[e + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c * 32 + e) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ATT_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------
// Small generic attention (text encoders; head_dim 64; any sequence length): one warp per (b, h, query).
// ------------------------------------------------------------------------------------------
struct AttnSmallParams {
  const __nv_bfloat16 *q, *k, *v;
  long long ld, bs;
  const float* bias;
  __nv_bfloat16* out;
  long long ld_out, out_bs;
  float scale;
  int batch, heads, seq, causal;
};

__global__ void __launch_bounds__(256) attn_small_kernel(const AttnSmallParams p) {
  const int lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)p.batch * p.heads * p.seq;
  if (gw >= total) return;
  const int qi = int(gw % p.seq);
  const int h = int((gw / p.seq) % p.heads);
  const int b = int(gw / ((long long)p.seq * p.heads));
  const __nv_bfloat16* qp = p.q + b * p.bs + (long long)qi * p.ld + h * 64;
  // each lane keeps the full (scaled) query: 64 floats
  float qv[64];
#pragma unroll
  for (int d = 0; d < 64; d += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(qp + d);
    float2 a = unpack_bf16(u.x), bb = unpack_bf16(u.y), c = unpack_bf16(u.z), e = unpack_bf16(u.w);
    qv[d] = a.x * p.scale; qv[d + 1] = a.y * p.scale; qv[d + 2] = bb.x * p.scale; qv[d + 3] = bb.y * p.scale;
    qv[d + 4] = c.x * p.scale; qv[d + 5] = c.y * p.scale; qv[d + 6] = e.x * p.scale; qv[d + 7] = e.y * p.scale;
  }
  // keys are walked in chunks of 512 (16 per lane) with an online softmax across chunks, so any sequence length works
  // (the reference's T5 tokenizer never truncates: flux/tokenizers.py:160-173); one chunk = the plain softmax.
  constexpr int MAXK = 16;
  const int nk = p.causal ? qi + 1 : p.seq;
  float m_run = -INFINITY, sum = 0.f, o0 = 0.f, o1 = 0.f;
  for (int k0 = 0; k0 < nk; k0 += 32 * MAXK) {
    float sc[MAXK];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < MAXK; ++t) {
      const int kj = k0 + t * 32 + lane;
      float s = -INFINITY;
      if (kj < nk) {
        const __nv_bfloat16* kp = p.k + b * p.bs + (long long)kj * p.ld + h * 64;
        s = 0.f;
#pragma unroll
        for (int d = 0; d < 64; d += 8) {
          const uint4 u = *reinterpret_cast<const uint4*>(kp + d);
          float2 a = unpack_bf16(u.x), bb = unpack_bf16(u.y), c = unpack_bf16(u.z), e = unpack_bf16(u.w);
          s += qv[d] * a.x + qv[d + 1] * a.y + qv[d + 2] * bb.x + qv[d + 3] * bb.y + qv[d + 4] * c.x + qv[d + 5] * c.y +
               qv[d + 6] * e.x + qv[d + 7] * e.y;
        }
        if (p.bias) s += p.bias[((long long)h * p.seq + qi) * p.seq + kj];
      }
      sc[t] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float m_new = fmaxf(m_run, mx);
    const float alpha = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_new);  // first chunk: nothing to rescale
    m_run = m_new;
    float csum = 0.f;
#pragma unroll
    for (int t = 0; t < MAXK; ++t) {
      sc[t] = (sc[t] == -INFINITY) ? 0.f : __expf(sc[t] - m_new);
      csum += sc[t];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
    sum = sum * alpha + csum;
    o0 *= alpha;
    o1 *= alpha;
    // out[d]: lane owns dims 2*lane, 2*lane+1; loop over the chunk's keys, probabilities broadcast by shuffle
    for (int t = 0; t < MAXK; ++t) {
      if (k0 + t * 32 >= nk) break;
      for (int src = 0; src < 32; ++src) {
        const int kj = k0 + t * 32 + src;
        const float pj = __shfl_sync(0xffffffffu, sc[t], src);
        if (kj < nk) {
          const __nv_bfloat162 vv =
              *reinterpret_cast<const __nv_bfloat162*>(p.v + b * p.bs + (long long)kj * p.ld + h * 64 + 2 * lane);
          const float2 f = __bfloat1622float2(vv);
          o0 += pj * f.x;
          o1 += pj * f.y;
        }
      }
    }
  }
  const float inv = 1.0f / sum;
  __nv_bfloat16* op = p.out + b * p.out_bs + (long long)qi * p.ld_out + h * 64 + 2 * lane;
  *reinterpret_cast<__nv_bfloat162*>(op) = __floats2bfloat162_rn(o0 * inv, o1 * inv);
}

// ------------------------------------------------------------------------------------------
// The same attention on the tensor cores (warp-level mma.sync.m16n8k16, bf16 in / fp32 accumulate): a flash loop over 64-key
// chunks staged in shared memory, 64 queries per CTA (16 per warp), S and O in registers.  The text encoders run once per
// prompt on a few thousand query rows per head -- far too small for the tcgen05 / TMEM kernel above (its 256-row CTA tiles
// and 128-wide heads do not apply: head_dim is 64, T5 adds a relative-position bias to every score, CLIP is causal) -- but
// the warp-per-query kernel re-read every key row from L2 for every query: 285 us per T5 layer, 66 % of the text-encode time.
// Fragment layouts (PTX ISA, mma.m16n8k16 .row.col): g = lane / 4, t = lane % 4;
//   A (16 x 16): a0 = (g, 2t..2t+1), a1 = (g + 8, 2t..), a2 = (g, 2t + 8..), a3 = (g + 8, 2t + 8..)
//   B (16 x 8):  b0 = (k = 2t..2t+1, n = g), b1 = (k = 2t + 8.., n = g);   C (16 x 8): c0,c1 = (g, 2t..2t+1), c2,c3 = (g + 8, ..)
// ------------------------------------------------------------------------------------------
constexpr int AS_PAD = 72;  // padded row of the K / V chunks (bf16): 144-byte stride = conflict-free fragment loads
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(128) attn_small_mma_kernel(const AttnSmallParams p) {
  __shared__ __align__(16) __nv_bfloat16 Ks[64][AS_PAD];
  __shared__ __align__(16) __nv_bfloat16 Vs[64][AS_PAD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;  // this thread's two query rows
  const __nv_bfloat16* qb = p.q + b * p.bs + h * 64;
  uint32_t qa[4][4];  // Q fragments of the four 16-wide slices of head_dim (rows past the end: zeros)
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int c = ks * 16 + 2 * t;
    qa[ks][0] = r0 < p.seq ? *reinterpret_cast<const uint32_t*>(qb + (long long)r0 * p.ld + c) : 0u;
    qa[ks][1] = r1 < p.seq ? *reinterpret_cast<const uint32_t*>(qb + (long long)r1 * p.ld + c) : 0u;
    qa[ks][2] = r0 < p.seq ? *reinterpret_cast<const uint32_t*>(qb + (long long)r0 * p.ld + c + 8) : 0u;
    qa[ks][3] = r1 < p.seq ? *reinterpret_cast<const uint32_t*>(qb + (long long)r1 * p.ld + c + 8) : 0u;
  }
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int nk = p.causal ? min(p.seq, q0 + 64) : p.seq;  // keys any query of this CTA may see
  const float* bias0 = p.bias ? p.bias + ((long long)h * p.seq + min(r0, p.seq - 1)) * p.seq : nullptr;
  const float* bias1 = p.bias ? p.bias + ((long long)h * p.seq + min(r1, p.seq - 1)) * p.seq : nullptr;
  for (int k0 = 0; k0 < nk; k0 += 64) {
    __syncthreads();  // the previous chunk has been consumed
    for (int i = tid; i < 512; i += 128) {  // 64 rows x 8 pieces of 16 bytes, K and V
      const int row = i >> 3, c = (i & 7) * 8;
      uint4 kv = make_uint4(0, 0, 0, 0), vv = kv;
      if (k0 + row < p.seq) {
        const long long off = b * p.bs + (long long)(k0 + row) * p.ld + h * 64 + c;
        kv = *reinterpret_cast<const uint4*>(p.k + off);
        vv = *reinterpret_cast<const uint4*>(p.v + off);
      }
      *reinterpret_cast<uint4*>(&Ks[row][c]) = kv;
      *reinterpret_cast<uint4*>(&Vs[row][c]) = vv;
    }
    __syncthreads();
    // S = Q K^T for 16 queries x 64 keys
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][ks * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][ks * 16 + 2 * t + 8]);
        mma_bf16_16816(sc[nt], qa[ks], b0, b1);
      }
    }
    // scale, bias, masks, chunk maxima of the two rows
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int kj = k0 + nt * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = kj + e;
        float s0 = sc[nt][e] * p.scale, s1 = sc[nt][2 + e] * p.scale;
        if (key < p.seq) {
          if (bias0) { s0 += __ldg(bias0 + key); s1 += __ldg(bias1 + key); }
          if (p.causal && key > r0) s0 = -INFINITY;
          if (p.causal && key > r1) s1 = -INFINITY;
        } else {
          s0 = s1 = -INFINITY;
        }
        sc[nt][e] = s0; sc[nt][2 + e] = s1;
        mx0 = fmaxf(mx0, s0); mx1 = fmaxf(mx1, s1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
    const float a0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - n0), a1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - n1);
    m0 = n0; m1 = n1;
    const float e0 = (n0 == -INFINITY) ? 0.f : n0, e1 = (n1 == -INFINITY) ? 0.f : n1;  // a fully masked row so far: all p = 0
    float cs0 = 0.f, cs1 = 0.f;
    uint32_t pa[4][4];  // P as the A operand of P V: 16 queries x 64 keys = four 16-key slices
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p00 = __expf(sc[nt][0] - e0), p01 = __expf(sc[nt][1] - e0);
      const float p10 = __expf(sc[nt][2] - e1), p11 = __expf(sc[nt][3] - e1);
      cs0 += p00 + p01; cs1 += p10 + p11;
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16(p00, p01);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p10, p11);
    }
    l0 = l0 * a0 + cs0; l1 = l1 * a1 + cs1;  // (per-thread partial sums: reduced over the quad at the end)
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1; }
    // O += P V: B[k = key][n = d] from V[key][d] through ldmatrix.trans (two d-tiles per instruction)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dt = 0; dt < 8; dt += 2) {
        uint32_t v0, v1, v2, v3;
        const uint32_t addr = smem_u32(&Vs[kk * 16 + (lane & 15)][dt * 8 + (lane >> 4) * 8]);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
        mma_bf16_16816(o[dt], pa[kk], v0, v1);
        mma_bf16_16816(o[dt + 1], pa[kk], v2, v3);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = p.out + b * p.out_bs + h * 64 + 2 * t;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    if (r0 < p.seq) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * p.ld_out + dt * 8) = pack_bf16(o[dt][0] * i0, o[dt][1] * i0);
    if (r1 < p.seq) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * p.ld_out + dt * 8) = pack_bf16(o[dt][2] * i1, o[dt][3] * i1);
  }
}

}  // namespace fx

using namespace fx;

template <bool P_TMEM>
static int launch_attn(const fx_attn_args* a, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const AttnParams& p, cudaStream_t st) {
  using Cfg = AttnCfg<P_TMEM>;
  auto kern = attn_kernel<P_TMEM>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES); });
  if (attr_err != cudaSuccess) return fail(FX_ERR_CUDA, "attention smem attribute: %s", cudaGetErrorString(attr_err));
  dim3 grid((a->seq + 255) / 256, a->heads, a->batch);
  kern<<<grid, ATT_THREADS, Cfg::SMEM_BYTES, st>>>(tq, tk, tv, p);
  return launched("attn_kernel");
}

// FX_ATTN_PERSISTENT=0/1: which kernel variant 0 (the default) runs -- the persistent work loop or one CTA per item
static bool persistent_default() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FX_ATTN_PERSISTENT");
    v = e ? atoi(e) : 1;  // measured: 1.612 -> 1.565 ms isolated, 419 -> 401 ms per bench step (profiles/README.md)
  }
  return v != 0;
}

extern "C" int fx_attention(const fx_attn_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->q && a->k && a->v && (a->out || a->q_out), "fx_attention: null pointer");
  FX_REQUIRE(a->batch > 0 && a->heads > 0 && a->seq > 0, "fx_attention: empty problem");
  FX_REQUIRE(a->q_out || (a->ld_out % 8 == 0 && a->out_bs % 8 == 0 && aligned16(a->out)), "fx_attention: out must be 16-byte aligned rows");
  FX_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v), "fx_attention: q/k/v must be 16-byte aligned");
  AttnParams p{};
  p.seq = a->seq; p.heads = a->heads; p.kv_tiles = (a->seq + 127) / 128;
  p.sequence = (a->variant == 4) ? 1 : 0;  // turn taking lost its edge once the issuer was fixed: 3314 vs 3230 clocks per key tile
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = (__nv_bfloat16*)a->out; p.ld_out = a->ld_out; p.out_bs = a->out_bs;
  p.out4[0] = ChunkQ{nullptr, nullptr, nullptr, 0, 0};
  p.out4[1] = p.out4[0];
  p.out4_split = 0;
  if (a->q_out) {  // NVFP4 chunks instead of bf16 (persistent kernel only)
    FX_REQUIRE(a->variant == 0 || a->variant == 7, "fx_attention: q_out runs on the persistent kernel only (variant 0 / 7)");
    FX_REQUIRE(a->sf_out && a->e_out && a->out_kc % 128 == 0 && a->out_col0 % 64 == 0 && a->out_col0 + a->heads * 128 <= a->out_kc &&
                   aligned16(a->q_out) && aligned16(a->sf_out),
               "fx_attention: q_out needs sf_out, e_out, out_kc %% 128 == 0 and the heads inside out_kc");
    FX_REQUIRE(a->out_split >= 0 && a->out_split <= a->seq && a->out_split % 128 == 0 && (a->seq - a->out_split) % 128 == 0,
               "fx_attention: q_out needs multiples of 128 rows on both sides of out_split");
    FX_REQUIRE(a->out_split == 0 || (a->q_out2 && a->sf_out2 && a->e_out2 && a->out_kc2 % 128 == 0 && a->heads * 128 <= a->out_kc2),
               "fx_attention: out_split > 0 needs the second operand");
    p.out4[0] = ChunkQ{(uint8_t*)a->q_out, (uint8_t*)a->sf_out, (int8_t*)a->e_out, a->out_kc, a->out_col0};
    p.out4[1] = ChunkQ{(uint8_t*)a->q_out2, (uint8_t*)a->sf_out2, (int8_t*)a->e_out2, a->out_kc2, 0};
    p.out4_split = a->out_split;
  }
  CUtensorMap tq, tk, tv;
  const uint64_t esz = a->fp8 ? 1 : 2;  // fp8: e4m3 q / k / v, one 128-byte row per token and head
  const uint64_t dims[3] = {128, (uint64_t)a->seq, (uint64_t)a->batch * a->heads};
  const uint64_t strides[2] = {128 * esz, (uint64_t)a->seq * 128 * esz};
  const uint32_t box[3] = {a->fp8 ? 128u : 64u, 128, 1};
  int rc;
  if ((rc = make_tmap_bf16(&tq, a->q, 3, dims, strides, box, a->fp8 != 0))) return rc;
  if ((rc = make_tmap_bf16(&tk, a->k, 3, dims, strides, box, a->fp8 != 0))) return rc;
  if ((rc = make_tmap_bf16(&tv, a->v, 3, dims, strides, box, a->fp8 != 0))) return rc;
  FX_REQUIRE(!a->fp8 || a->variant == 0 || a->variant == 7, "fx_attention: fp8 operands run on the persistent kernel only (variant 0 / 7)");
  if (a->variant == 7 || a->fp8 || a->q_out || (a->variant == 0 && persistent_default())) {
    // persistent work loop (attn_pkernel): one CTA per SM over the (q-block, head, batch) items
    static std::once_flag oncep;
    static cudaError_t attrp_err = cudaSuccess;
    std::call_once(oncep, [&] {
      attrp_err = cudaFuncSetAttribute(attn_pkernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnPCfg<false>::SMEM_BYTES);
      if (attrp_err == cudaSuccess)
        attrp_err = cudaFuncSetAttribute(attn_pkernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnPCfg<true>::SMEM_BYTES);
    });
    if (attrp_err != cudaSuccess) return fail(FX_ERR_CUDA, "attention smem attribute: %s", cudaGetErrorString(attrp_err));
    const int q_blocks = (a->seq + 255) / 256;
    const long long items = (long long)q_blocks * a->heads * a->batch;
    FX_REQUIRE(items < (1ll << 31), "fx_attention: too many work items");
    const int grid = (int)(items < num_sms() ? items : num_sms());
    if (a->fp8)
      attn_pkernel<true><<<grid, ATT_THREADS, AttnPCfg<true>::SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, p, q_blocks, (int)items);
    else
      attn_pkernel<false><<<grid, ATT_THREADS, AttnPCfg<false>::SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, p, q_blocks, (int)items);
    return launched("attn_pkernel");
  }
  if (a->variant == 5 || a->variant == 6) {
    // decoupled schedule (attn3_kernel): P through shared memory, Q K(j+1)^T issued as soon as S(j) has been read
    p.sequence = a->variant == 5 ? 1 : 0;
    static std::once_flag once3;
    static cudaError_t attr3_err = cudaSuccess;
    std::call_once(once3, [&] { attr3_err = cudaFuncSetAttribute(attn3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Cfg::SMEM_BYTES); });
    if (attr3_err != cudaSuccess) return fail(FX_ERR_CUDA, "attention smem attribute: %s", cudaGetErrorString(attr3_err));
    dim3 grid((a->seq + 255) / 256, a->heads, a->batch);
    attn3_kernel<<<grid, ATT_THREADS, Attn3Cfg::SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, p);
    return launched("attn3_kernel");
  }
  if (a->variant == 1) return launch_attn<false>(a, tq, tk, tv, p, (cudaStream_t)stream);
  return launch_attn<true>(a, tq, tk, tv, p, (cudaStream_t)stream);
}

#ifdef FX_ATTN_PROBE
extern "C" int fx_dbg_attn_probe(void* buf) {  // profiling builds only: device buffer of 3*64*8 int64
  long long* b = (long long*)buf;
  FX_CUDA(cudaMemcpyToSymbol(fx::g_attn_probe, &b, sizeof(b)));
  return FX_OK;
}
#endif

extern "C" int fx_attention_small(const fx_attn_small_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->q && a->k && a->v && a->out, "fx_attention_small: null pointer");
  FX_REQUIRE(a->seq > 0, "fx_attention_small: empty sequence");
  FX_REQUIRE(a->ld % 8 == 0 && a->bs % 8 == 0 && a->ld_out % 2 == 0, "fx_attention_small: unaligned strides");
  AttnSmallParams p{(const __nv_bfloat16*)a->q, (const __nv_bfloat16*)a->k, (const __nv_bfloat16*)a->v, a->ld, a->bs,
                    a->bias, (__nv_bfloat16*)a->out, a->ld_out, a->out_bs, a->scale, a->batch, a->heads, a->seq, a->causal};
  static int use_mma = -1;  // FX_ATTN_SMALL_MMA=0: the warp-per-query CUDA-core kernel (A/B)
  if (use_mma < 0) {
    const char* e = getenv("FX_ATTN_SMALL_MMA");
    use_mma = e ? atoi(e) : 1;
  }
  if (use_mma && a->batch <= 65535 && a->heads <= 65535 && aligned16(a->q) && aligned16(a->k) && aligned16(a->v) &&
      (reinterpret_cast<uintptr_t>(a->out) & 3) == 0) {
    dim3 grid((unsigned)((a->seq + 63) / 64), (unsigned)a->heads, (unsigned)a->batch);
    attn_small_mma_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    return launched("attn_small_mma_kernel");
  }
  const long long warps = (long long)a->batch * a->heads * a->seq;
  const int blocks = int((warps + 7) / 8);
  attn_small_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return launched("attn_small_kernel");
}
