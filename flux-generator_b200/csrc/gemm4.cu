// NVFP4 (W4A4) GEMMs for every block Linear of the MMDiT under `--quantize` (qkv, proj, mlp.0, mlp.2, linear1, linear2), the
// activation quantisers and the finalise pass of producer-emitted operands.
//   out = epilogue( (A4 . W4^T) * a_scale[row] * w_scale[col] ),   A4, W4: e2m1 (two per byte, K contiguous) with one
//   UE4M3 scale per 16 elements of K -- tcgen05.mma.kind::mxf4nvf4.block_scale.block16: the tensor core applies the
//   block scales itself, from TMEM, at four times the bf16 MAC rate.
// The honest Blackwell analogue of the reference's 4-bit `nn.quantize(group_size=64)` of the Linear layers
// (txt2image.py:28-29,79-82): there 4-bit weights are dequantised into a bf16 matmul, here both operands stay 4-bit.
//
// Layouts (validated on hardware with tests/gpu_bs_probe.py, profiles/r02_blockscale_probe.txt):
//   data     [rows][K / 2] bytes, element 2i in the low nibble; a 128-byte row segment = 256 elements = 4 MMAs (K = 64)
//   scales   512-byte atoms of 128 rows x 4 scales: byte (r % 32) * 16 + (r / 32) * 4 + s.  One tcgen05.cp.32x128b.warpx4
//            drops an atom into 4 TMEM columns (lane r % 32 of every sub-partition, column r / 32, byte s), which is
//            where the MMA reads the four scales of its 64 elements of K for row r.
//            A: atoms [row block of 128][K / 64]; W: atoms [column tile][K / 64][atoms per tile] (192-row tiles: rows 0-127 and
//            128-191; 128-row tiles for the QKV epilogue)
//   a_scale  fp32 per row, w_scale fp32 per output channel: the second quantisation level (see fx_quantize_rows_fp4)
// Kernel: persistent CTA pairs (cta_group::2: the W tile is split over the pair) -- single CTAs when a batch element is not a
// whole number of 256-row tiles -- with 16 warps: 12 epilogue warps, TMA producer, MMA issuer (both in uniform control flow).
//   EPI_GENERIC: 256 x 192 tiles (two accumulators AND two scale-factor slots fit the 512 TMEM columns: 2 x 192 + 2 x 48);
//                bias / act / gate / residual; bf16 output through TMA stores, fp32 output, or -- q_out -- the epilogue emits
//                the NEXT GEMM's NVFP4 operand (fp4.cuh) and the result never exists in bf16.
//   EPI_QKV:     256 x 128 tiles = one head, three accumulators (3 x 128 + 2 x 32 columns); QK-RMSNorm + RoPE + scatter.
// Ring: A + W + scale atoms by TMA; per stage 4 (SFA) + 4 or 8 (SFB) tcgen05.cp and 4 MMAs of K = 64.
#include <cuda_fp4.h>
#include <cuda_fp8.h>
#include <stdlib.h>

#include <mutex>

#include "api_common.cuh"
#include "fp4.cuh"
#include "gemm.cuh"
#include "tmap.cuh"

namespace fx {

constexpr int G4_BN = 192;                      // generic epilogue tiles; the QKV epilogue uses 128-column tiles (= one head)
constexpr int G4_A_BYTES = 128 * 128;           // 128 rows x 256 e2m1
constexpr int G4_SFA_BYTES = 4 * 512;           // 4 K-groups (of 64 elements) x one 128-row atom
// 16 warps: 12 epilogue warps (3 column groups x 4 TMEM lane quarters) + producer, MMA issuer and two idle warps.  At 4-bit MAC
// rates a K = 3072 tile lasts ~4600 clocks and the epilogue (dequantise, GELU / RMSNorm + RoPE, pack, store) sets the pace: with
// 8 epilogue warps (two per scheduler) it ran at 30-45 % issue utilisation, latency-bound; three per scheduler overlap better
// and each owns a third of the tile's columns (generic) or every third tile (QKV: one head per tile, THREE accumulators).
constexpr int G4_THREADS = 512, G4_EPI_WARPS = 12, G4_CTRL0 = 12, G4_WARP_TMA = 12, G4_WARP_MMA = 13, G4_GROUPS = 3;
template <int NCTA, int BN>
struct G4Cfg {
  // NCTA = 2: a CTA pair computes a 256 x BN tile with cta_group::2 MMAs; each CTA stages its 128 A rows, HALF of the W
  // tile and the scale atoms of its own A rows and of ALL BN W rows (tcgen05.cp.cta_group::2 copies, in each CTA, from
  // that CTA's shared memory into that CTA's TMEM: tests/gpu_bs_probe.py pair)
  static constexpr int NSFB = (BN + 127) / 128;             // W scale atoms per K-group: the WHOLE tile's columns, in every CTA
  static constexpr int SFB_BYTES = 4 * NSFB * 512;
  static constexpr int B_BYTES = (BN / NCTA) * 128;
  static constexpr int STAGE_BYTES = G4_A_BYTES + B_BYTES + G4_SFA_BYTES + SFB_BYTES;
  // behind the ring: barriers (1 KB slot), store staging (12 x 2 KB, 512-byte aligned: TMA SWIZZLE_64B), epilogue vectors
  static constexpr int TAIL = 1024 + G4_EPI_WARPS * (2048 + 384 * 4) + 1024;
  static constexpr int STAGES = (232448 - TAIL) / STAGE_BYTES;        // 5 / 3 (BN 192: pair / single), 6 / 5 (BN 128)
  static constexpr int STORE_OFF = STAGES * STAGE_BYTES + 1024;
  static constexpr int EPI_OFF = STORE_OFF + G4_EPI_WARPS * 2048;
  static constexpr int SMEM = EPI_OFF + G4_EPI_WARPS * 384 * 4 + 1024;
  static constexpr int NACC = BN == 128 ? 3 : 2;            // accumulators in TMEM: 3 x 128 (QKV) or 2 x 192 columns
  static constexpr int TMEM_SF = NACC * BN;                 // scale-factor slots start behind the accumulators
  static constexpr int SF_SLOT = 16 + 16 * NSFB;            // 16 columns of A scales + 16 per W atom row, per stage
  static_assert(TMEM_SF + 2 * SF_SLOT <= 512 && STAGES >= 3, "TMEM / shared memory budget");
};

__device__ __forceinline__ uint64_t make_smem_desc_sf(uint32_t smem_addr) {  // no swizzle, SBO = 128 B (8 rows x 16 B), version 1
  return uint64_t((smem_addr & 0x3FFFF) >> 4) | (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46);
}
template <int NCTA>
__device__ __forceinline__ void tc_cp_sf(uint32_t taddr, uint64_t sdesc) {
  if (NCTA == 2) asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
  else asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
template <int NCTA>
__device__ __forceinline__ void umma_nvf4(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc, uint32_t sfa, uint32_t sfb) {
  if (NCTA == 2)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4nvf4.block_scale.block16 [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d),
        "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4nvf4.block_scale.block16 [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d),
        "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb)
        : "memory");
}
// cute::UMMA::InstrDescriptorBlockScaled: a/b format E2M1 (1) at [7,10) / [10,13), N >> 3 at [17,23), scale format UE4M3 (0)
// at [23], M >> 4 at [24,29), scale-factor ids 0
__host__ __device__ constexpr uint32_t make_idesc_nvf4(int M, int N) {
  return (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

struct Gemm4Params {
  GemmParams g;              // shapes, raster, generic epilogue (a_scale / w_scale = the second-level scales)
  int k_groups;              // K / 64
  int tma_out;               // generic epilogue, bf16 output: chunks leave through TMA stores (tmap_out)
  int sfb_mcast;             // CTA pairs: the W scale atoms (needed whole by BOTH CTAs) are fetched once and multicast
  ChunkQ out4;               // q != nullptr: the generic epilogue writes the NEXT GEMM's NVFP4 operand instead of `out`
};

// tmap_sfa / tmap_sfb: the scale-atom buffers viewed as [bytes / 128][128] byte matrices (no swizzle): one stage's atoms of a
// row block / column tile are 16 / 32 consecutive rows, fetched by TMA like the operands (and, for a pair, credited to the
// leader's barrier like them)
template <int NCTA, int BN, int EPI>
__global__ void __launch_bounds__(G4_THREADS, 1)
gemm_nvfp4_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_sfa, const __grid_constant__ CUtensorMap tmap_sfb,
                  const __grid_constant__ CUtensorMap tmap_out, const Gemm4Params q) {
  const GemmParams& p = q.g;
  using Cfg = G4Cfg<NCTA, BN>;
  constexpr int STAGES = Cfg::STAGES, NSFB = Cfg::NSFB, NACC = Cfg::NACC;
  static_assert(EPI == EPI_GENERIC || BN == 128, "the QKV epilogue works on one head (128 columns) per tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 3);
  // warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the issuer's descriptors in uniform registers
  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const int first_tile = blockIdx.x / NCTA, tile_stride = gridDim.x / NCTA;

  if (warp == G4_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_sfa);
    tma_prefetch_desc(&tmap_sfb);
    if (q.tma_out) tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], (EPI == EPI_QKV ? 4 : G4_EPI_WARPS) * NCTA);
    }
    fence_barrier_init();
  }
  if (warp == G4_WARP_MMA) {
    if (NCTA == 2) {
      tmem_alloc2(tmem_slot, 512);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= G4_CTRL0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == G4_WARP_TMA) {
      // ================= TMA producer: A rows, this CTA's share of the W tile, the scale atoms (already in atom order)
      {
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = first_tile; tile < p.num_tiles; tile += tile_stride) {
          int tm, tn;
          gemm_tile_coords(p, tile, tm, tn);
          const int rb = tm * NCTA + int(cta_rank);  // 128-row block of the flattened activation rows
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + G4_A_BYTES;
            uint8_t* ssfa = sb + Cfg::B_BYTES;
            uint8_t* ssfb = ssfa + G4_SFA_BYTES;
            const int sfa_row = (rb * q.k_groups + kb * 4) * 4;          // 4 rows of 128 B per atom
            const int sfb_row = (tn * q.k_groups + kb * 4) * 4 * NSFB;   // NSFB atoms per K-group
            if (!elect_one()) {
            } else if (NCTA == 2) {
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
              tma2_load_2d(sa, &tmap_a, &full_bar[stage], kb * 128, rb * GEMM_BM);
              tma2_load_2d(sb, &tmap_w, &full_bar[stage], kb * 128, tn * BN + int(cta_rank) * (BN / 2));
              tma2_load_2d(ssfa, &tmap_sfa, &full_bar[stage], 0, sfa_row);
              if (!q.sfb_mcast) tma2_load_2d(ssfb, &tmap_sfb, &full_bar[stage], 0, sfb_row);
              else if (cta_rank == 0) tma2_load_2d_mcast(ssfb, &tmap_sfb, &full_bar[stage], 0, sfb_row, uint16_t(3), kL2EvictNormal);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * 128, rb * GEMM_BM);
              tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * 128, tn * BN);
              tma_load_2d(ssfa, &tmap_sfa, &full_bar[stage], 0, sfa_row);
              tma_load_2d(ssfb, &tmap_sfb, &full_bar[stage], 0, sfb_row);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == G4_WARP_MMA) {
      // ================= MMA issuer: per stage 12 scale-atom copies into the stage's TMEM slot, then 4 MMAs (K = 64 each)
      if (cta_rank == 0) {
        // the whole warp runs the loop (uniform control flow); one elected lane issues each tcgen05 instruction
        constexpr uint32_t idesc = make_idesc_nvf4(GEMM_BM * NCTA, BN);
        const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t smem0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
        int stage = 0;
        uint32_t phase = 0, n = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = first_tile; tile < p.num_tiles; tile += tile_stride) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tbase + acc * BN;
          for (int kb = 0; kb < p.k_blocks; ++kb, ++n) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem0 + stage * Cfg::STAGE_BYTES;
            const uint32_t sb = sa + G4_A_BYTES;
            const uint32_t ssfa = sb + Cfg::B_BYTES, ssfb = ssfa + G4_SFA_BYTES;
            // tcgen05.cp and tcgen05.mma execute in issue order: slot (n & 1) was last read by the MMAs of stage n - 2
            const uint32_t t_sfa = tbase + Cfg::TMEM_SF + (n & 1) * Cfg::SF_SLOT, t_sfb = t_sfa + 16;
            if (elect_one()) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                tc_cp_sf<NCTA>(t_sfa + g * 4, make_smem_desc_sf(ssfa + g * 512));
#pragma unroll
                for (int h = 0; h < NSFB; ++h) tc_cp_sf<NCTA>(t_sfb + (g * NSFB + h) * 4, make_smem_desc_sf(ssfb + (g * NSFB + h) * 512));
              }
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_nvf4<NCTA>(d_tmem, make_smem_desc_sw128(sa + k * 32, 16, 1024), make_smem_desc_sw128(sb + k * 32, 16, 1024), idesc,
                                (kb | k) != 0 ? 1u : 0u, t_sfa + k * 4, t_sfb + k * 4 * NSFB);
              if (NCTA == 2) tc_commit2(&empty_bar[stage]);
              else tc_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) {
            if (NCTA == 2) tc_commit2(&tfull_bar[acc]);
            else tc_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue warps (the generic epilogue of gemm_kernel in its dequantising form)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");   // 12 x 32 x 152 + 4 x 32 x 40 registers <= 64 K
    const int quarter = warp & 3;
    const int group = warp >> 2;   // column group (generic) / accumulator (QKV) this warp serves
    const int r = quarter * 32 + lane;
    float* sb = reinterpret_cast<float*>(smem + Cfg::EPI_OFF) + warp * 384;
    float* sg = sb + 128;
    float* sw = sb + 256;
    uint8_t* wst = smem + Cfg::STORE_OFF + warp * 2048;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr int WN = BN / G4_GROUPS, CH = (WN + 31) / 32;   // generic: 64 columns = 2 chunks per warp (BN = 192)
    // The tile's main loop is short at 4-bit MAC rates (K = 3072: ~4600 clocks), so the epilogue must not expose latencies:
    //  * the tile-uniform vectors (bias, column scales, gate) and the row scale of the NEXT tile are fetched into registers
    //    before this tile's chunk loop and written to the staging slots after it (pre_*: one global round trip per tile, hidden);
    //  * inside the chunk loop the tcgen05.ld of chunk c + 1 is in flight while chunk c is processed.
    float pre_b[CH], pre_w[CH], pre_g[CH], pre_rs = 1.f;
    int nx_tn = 0, nx_b = 0;       // coordinates of the prefetched tile (computed once, by the prefetch)
    long long nx_row = 0;
    auto prefetch = [&](int tile) {
      int tm;
      gemm_tile_coords(p, tile, tm, nx_tn);
      nx_b = tm / p.tiles_m_per_batch;
      nx_row = (long long)(tm - nx_b * p.tiles_m_per_batch) * (GEMM_BM * NCTA) + cta_rank * GEMM_BM + r;
      const int nw0 = nx_tn * BN + group * WN;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int n = nw0 + i * 32 + lane;
        const bool in = n < p.N;
        pre_b[i] = (p.bias != nullptr && in) ? __bfloat162float(__ldg(p.bias + n)) : 0.f;
        pre_w[i] = in ? __ldg(p.w_scale + n) : 0.f;
        pre_g[i] = (p.gate != nullptr && in) ? __bfloat162float(__ldg(p.gate + (long long)nx_b * p.gate_bs + n)) : 1.f;
      }
      pre_rs = nx_row < p.rows ? __ldg(p.a_scale + (long long)nx_b * p.a_scale_bs + nx_row) : 1.f;
    };
    // QKV: a tile is one head; the two warp groups take alternate accumulators, i.e. every other tile of the CTA's queue
    QkvPre qpre;
    auto qkv_coords = [&](int tile, QkvNext& n) {
      int tm, tn;
      gemm_tile_coords(p, tile, tm, tn);
      n.b = tm / p.tiles_m_per_batch;
      n.row = (long long)(tm - n.b * p.tiles_m_per_batch) * (GEMM_BM * NCTA) + cta_rank * GEMM_BM + r;
      n.valid = n.row < p.rows;
      n.g0 = tn * BN;
    };
    if (EPI == EPI_GENERIC && first_tile < p.num_tiles) prefetch(first_tile);
    if (EPI == EPI_QKV && first_tile + group * tile_stride < p.num_tiles) {
      QkvNext n;
      qkv_coords(first_tile + group * tile_stride, n);
      epi_qkv_prefetch<true>(p, n.b, n.row, n.valid, lane, n.g0, qpre);
    }
    for (int tile = first_tile; tile < p.num_tiles; tile += tile_stride) {
      if (EPI == EPI_QKV && group != acc) {
        if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      int tn, b;
      long long row;
      if (EPI == EPI_GENERIC && p.dbg_skip_w != 1) {  // this tile's coordinates came with its prefetch
        tn = nx_tn; b = nx_b; row = nx_row;
      } else {
        int tm;
        gemm_tile_coords(p, tile, tm, tn);
        b = tm / p.tiles_m_per_batch;
        row = (long long)(tm - b * p.tiles_m_per_batch) * (GEMM_BM * NCTA) + cta_rank * GEMM_BM + r;
      }
      const bool valid = row < p.rows;
      const uint32_t taddr = tmem_base + acc * BN + (uint32_t(quarter * 32) << 16);
      // FX_GEMM4_DBG_NOEPI=2 (measurement only): the whole epilogue except its global stores (every row masked)
      const uint32_t vmask = p.dbg_skip_w == 2 ? 0u : __ballot_sync(0xffffffffu, valid);
      if (p.dbg_skip_w == 1) {  // FX_GEMM4_DBG_NOEPI=1 (measurement only): the main loop without the epilogue's work
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
      } else if (EPI == EPI_QKV) {
        QkvNext n;
        n.has = tile + NACC * tile_stride < p.num_tiles;
        if (n.has) qkv_coords(tile + NACC * tile_stride, n);
        epi_qkv_group<true>(p, b, row, valid, vmask, lane, taddr, tn * BN, sb, sg, sw, wst, &tfull_bar[acc], acc_phase, qpre, n);
      } else {
        const int nw0 = tn * BN + group * WN;
        const long long out_off = (long long)b * p.out_bs + row * p.ldo;
        const long long res_off = (long long)b * p.resid_bs + row * p.ldr;
        const bool vec_ok = ((p.ldo | p.ldr) & 7) == 0;
        __syncwarp();   // the previous tile's chunk loop is done with the staging slots
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          sb[i * 32 + lane] = pre_b[i];
          sw[i * 32 + lane] = pre_w[i];
          sg[i * 32 + lane] = pre_g[i];
        }
        const float rs = pre_rs;
        const bool rr_ok = p.resid != nullptr && valid && vec_ok && (nw0 + WN <= p.N);
        uint4 rcur[4], rnxt[4];
        if (rr_ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i) rcur[i] = *reinterpret_cast<const uint4*>(p.resid + res_off + nw0 + i * 8);
        }
        if (tile + tile_stride < p.num_tiles) prefetch(tile + tile_stride);
        __syncwarp();
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_x32(taddr + group * WN, v);
#pragma unroll 1
        for (int c = 0; c < CH; ++c) {
          const int n0 = nw0 + c * 32;
          if (n0 >= p.N) {
            tmem_ld_wait_x32(v);
            break;
          }
          if (rr_ok && c + 1 < CH) {
#pragma unroll
            for (int i = 0; i < 4; ++i) rnxt[i] = *reinterpret_cast<const uint4*>(p.resid + res_off + n0 + 32 + i * 8);
          }
          float f[32];
          tmem_ld_wait_x32(v);
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (c + 1 < CH) tmem_ld_x32(taddr + group * WN + (c + 1) * 32, v);   // in flight while this chunk is processed
          if (q.out4.q != nullptr) {  // bias + activation, then straight to the next GEMM's NVFP4 operand (no bf16 round trip)
            epi_generic_chunk<true>(p, f, sb + c * 32, sg + c * 32, rcur, false, 0, 0, n0, true, sw + c * 32, rs, nullptr, lane, vmask,
                                    valid, -1, -1, nullptr, 0, 0, true);
            fp4_chunk_quantise(f, (long long)b * p.rows + row, q.out4.col0 + n0, q.out4, valid && vmask != 0u);
          } else
          epi_generic_chunk<true>(p, f, sb + c * 32, sg + c * 32, rcur, rr_ok, out_off, res_off, n0, vec_ok && (n0 + 32 <= p.N),
                                  sw + c * 32, rs, wst, lane, vmask, valid, -1, -1, (q.tma_out && vmask != 0u) ? &tmap_out : nullptr,
                                  int(row) - lane, b);
#pragma unroll
          for (int i = 0; i < 4; ++i) rcur[i] = rnxt[i];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_leader(&tempty_bar[acc]);
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
    }
    if (q.tma_out && lane == 0) tma_store_wait_all();   // the staging buffer must outlive the bulk stores that read it
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync();
  else __syncthreads();
  if (warp == G4_WARP_MMA) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ NVFP4 row quantiser: one 256-thread block per row
// x bf16 [rows][K] (K % 64 == 0, K <= 16384) -> e2m1 bytes [rows][K / 2], UE4M3 block scales in atom order, fp32 row scale.
// Two levels, every step a single IEEE fp32 operation so that the CPU oracle reproduces bytes and scales exactly:
//   g  = absmax(row) * (1 / 2688)                 (2688 = 6 * 448: the largest block scale then encodes as 448)
//   sf = e4m3_rn(absmax(block of 16) * (1 / 6) * rcp(g)),   d = float(sf) * g
//   q  = e2m1_rn_sat(x * rcp(d))   (0 when d == 0);    x ~= q * float(sf) * g      (rcp = correctly rounded 1 / x)
struct Quant4Params {
  const __nv_bfloat16* x; long long ldx, x_bs;
  uint8_t* q;        // [batch * rows][K / 2]
  uint8_t* sf;       // atoms [ceil(batch * rows / 128)][K / 64][512]
  float* scale;      // [batch * rows]
  int batch, rows, K;
};
template <int ITERS>  // ITERS = ceil(K / 2048)
__global__ void __launch_bounds__(256, 4) quantize_rows_fp4_kernel(const Quant4Params p) {
  // The row stays in registers as PACKED bf16 (4 registers per 8 elements; unpacked fp32 copies cost 90 registers at
  // K = 15360 -> two resident blocks per SM and 2.3 TB/s, against 5.7 TB/s for the K = 12288 instance); the block maxima are
  // taken on packed pairs (|x| and max are exact in bf16).
  __shared__ float s_max[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long m = blockIdx.x;  // flattened row
  const int b = int(m / p.rows);
  const long long r = m - (long long)b * p.rows;
  const __nv_bfloat16* xr = p.x + b * p.x_bs + r * p.ldx;
  uint4 raw[ITERS];
  float bmax[ITERS];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = i * 2048 + tid * 8;
    raw[i] = (c < p.K) ? *reinterpret_cast<const uint4*>(xr + c) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const uint32_t w4[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
    __nv_bfloat162 a2 = __habs2(*reinterpret_cast<const __nv_bfloat162*>(&w4[0]));
#pragma unroll
    for (int j = 1; j < 4; ++j) a2 = __hmax2(a2, __habs2(*reinterpret_cast<const __nv_bfloat162*>(&w4[j])));
    float mx = fmaxf(__low2float(a2), __high2float(a2));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));  // a block of 16 = two neighbouring threads
    bmax[i] = mx;
    amax = fmaxf(amax, mx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) s_max[warp] = amax;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) amax = fmaxf(amax, s_max[w]);
  const float g = amax > 0.f ? __fmul_rn(amax, 1.0f / 2688.0f) : 1.0f;
  const float rg = __frcp_rn(g);
  if (tid == 0) p.scale[m] = g;
  const long long rb = m >> 7;
  const int rr = int(m & 127);
  uint8_t* sf_row = p.sf + rb * (long long)(p.K / 64) * 512 + (rr & 31) * 16 + (rr >> 5) * 4;
  uint8_t* qr = p.q + m * (long long)(p.K / 2);
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = i * 2048 + tid * 8;
    const bool in = c < p.K;  // (no early exit: the warp shuffles below need every lane)
    const float u = __fmul_rn(__fmul_rn(bmax[i], 1.0f / 6.0f), rg);
    const __nv_fp8_storage_t sf8 = __nv_cvt_float_to_fp8(u, __NV_SATFINITE, __NV_E4M3);
    const float d = __fmul_rn(__half2float(__half(__nv_cvt_fp8_to_halfraw(sf8, __NV_E4M3))), g);
    const int j = c >> 4;  // scale index in the row; the four scales of a K-group sit in lanes 8m, 8m+2, 8m+4, 8m+6
    uint32_t sfw = uint32_t(sf8);
    sfw |= __shfl_down_sync(0xffffffffu, sfw, 2) << 8;
    sfw |= __shfl_down_sync(0xffffffffu, sfw, 4) << 16;   // (lanes 8m and 8m+4 each hold two bytes by now)
    if (in && (tid & 7) == 0) *reinterpret_cast<uint32_t*>(sf_row + (j >> 2) * 512) = sfw;
    const float rd = d > 0.f ? __frcp_rn(d) : 0.f;
    const uint32_t w4[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
    uint32_t w = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 v = unpack_bf16(w4[e]);
      w |= uint32_t(__nv_cvt_float2_to_fp4x2(make_float2(__fmul_rn(v.x, rd), __fmul_rn(v.y, rd)), __NV_E2M1, cudaRoundNearest)) << (8 * e);
    }
    if (in) *reinterpret_cast<uint32_t*>(qr + (c >> 1)) = w;
  }
}

// ------------------------------------------------------------------ chunked NVFP4 producer for a bf16 tensor + finalise
// x bf16 [batch][rows][C] (C % 32 == 0) -> columns [col0, col0 + C) of the operand described by ChunkQ: one thread per chunk.
struct QuantChunksParams {
  const __nv_bfloat16* x; long long ldx, x_bs;
  int batch, rows, C;
  ChunkQ o;
};
__global__ void __launch_bounds__(256) quantize_chunks_fp4_kernel(const QuantChunksParams p) {
  const int cpr = p.C / 32;  // chunks per row
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long m = idx / cpr;
  if (m >= (long long)p.batch * p.rows) return;
  const int c = int(idx - m * cpr);
  const int b = int(m / p.rows);
  const long long r = m - (long long)b * p.rows;
  const uint4* src = reinterpret_cast<const uint4*>(p.x + b * p.x_bs + r * p.ldx + c * 32);
  float f[32];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = __ldg(src + i);
    const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
    f[8 * i] = a0.x; f[8 * i + 1] = a0.y; f[8 * i + 2] = a1.x; f[8 * i + 3] = a1.y;
    f[8 * i + 4] = a2.x; f[8 * i + 5] = a2.y; f[8 * i + 6] = a3.x; f[8 * i + 7] = a3.y;
  }
  fp4_chunk_quantise(f, m, p.o.col0 + c * 32, p.o, true);
}

// Finalise: row exponent = max over the row's chunk exponents, scale[row] = 2^that, every block scale shifted from its chunk's
// exponent to the row's (e4m3_rn of an exact product: an exponent shift unless the result is subnormal).  One 256-thread block
// per 128-row block; lane l of every warp owns rows l, l + 32, l + 64, l + 96 (one 16-byte piece of every scale atom, one
// 8-byte piece of every exponent group); the eight warps split the K-groups.
__global__ void __launch_bounds__(256) fp4_finalize_kernel(uint8_t* sf, const int8_t* e, float* scale, int kc) {
  __shared__ int s_max[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rb = blockIdx.x;
  const int kg = kc / 64;
  const uint2* eg = reinterpret_cast<const uint2*>(e + rb * (long long)kg * 256) + lane;
  int mx[4] = {-128, -128, -128, -128};
#pragma unroll 4
  for (int k = warp; k < kg; k += 8) {
    const uint2 w = eg[k * 32];
    const uint32_t ws[2] = {w.x, w.y};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t h = ws[q >> 1] >> (16 * (q & 1));
      mx[q] = max(mx[q], max(int(int8_t(h)), int(int8_t(h >> 8))));
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) s_max[warp][q * 32 + lane] = mx[q];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int m = s_max[0][q * 32 + lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = max(m, s_max[w][q * 32 + lane]);
    mx[q] = m;
    if (warp == 0) scale[rb * 128 + q * 32 + lane] = __uint_as_float(uint32_t(127 + m) << 23);
  }
  uint4* atoms = reinterpret_cast<uint4*>(sf + rb * (long long)kg * 512) + lane;
#pragma unroll 2
  for (int k = warp; k < kg; k += 8) {
    const uint2 ew = eg[k * 32];
    const uint32_t es[2] = {ew.x, ew.y};
    const uint4 w = atoms[k * 32];
    uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t ee = es[q >> 1] >> (16 * (q & 1));
      uint32_t out = 0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {   // the two chunks (= two blocks of 16 each) of this K-group
        int sh = mx[q] - int(int8_t(ee >> (8 * c)));
        sh = sh > 60 ? 60 : sh;
        const float f2 = __uint_as_float(uint32_t(127 - sh) << 23);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int s4 = 2 * c + t;
          const float v = __half2float(__half(__nv_cvt_fp8_to_halfraw(__nv_fp8_storage_t((ws[q] >> (8 * s4)) & 255u), __NV_E4M3)));
          out |= uint32_t(__nv_cvt_float_to_fp8(__fmul_rn(v, f2), __NV_SATFINITE, __NV_E4M3)) << (8 * s4);
        }
      }
      ws[q] = out;
    }
    atoms[k * 32] = make_uint4(ws[0], ws[1], ws[2], ws[3]);
  }
}

}  // namespace fx

using namespace fx;

extern "C" int fx_quantize_rows_fp4(const fx_quant4_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->x && a->q && a->sf && a->scale, "fx_quantize_rows_fp4: null pointer");
  FX_REQUIRE(a->K > 0 && a->K % 64 == 0 && a->K <= 16384 && a->ldx % 8 == 0 && a->x_bs % 8 == 0,
             "fx_quantize_rows_fp4: K (%d) must be a multiple of 64 up to 16384, strides multiples of 8 elements", a->K);
  FX_REQUIRE(aligned16(a->x) && (reinterpret_cast<uintptr_t>(a->q) & 3) == 0, "fx_quantize_rows_fp4: unaligned pointers");
  if (a->batch <= 0 || a->rows <= 0) return FX_OK;
  const long long rows = (long long)a->batch * a->rows;
  FX_REQUIRE(rows < (1ll << 31), "fx_quantize_rows_fp4: too many rows");
  Quant4Params p{(const __nv_bfloat16*)a->x, a->ldx, a->x_bs, (uint8_t*)a->q, (uint8_t*)a->sf, a->scale, a->batch, a->rows, a->K};
  cudaStream_t st = (cudaStream_t)stream;
  switch ((a->K + 2047) / 2048) {
#define FX_Q4(I) case I: quantize_rows_fp4_kernel<I><<<(unsigned)rows, 256, 0, st>>>(p); break;
    FX_Q4(1) FX_Q4(2) FX_Q4(3) FX_Q4(4) FX_Q4(5) FX_Q4(6) FX_Q4(7) FX_Q4(8)
#undef FX_Q4
  }
  return launched("quantize_rows_fp4_kernel");
}

// scale-atom buffer as a [bytes / 128][128] byte matrix, no swizzle (the atoms are consumed by tcgen05.cp, not by the MMA)
static int make_tmap_sf(CUtensorMap* out, const void* base, uint64_t bytes, uint32_t box_rows) {
  tmap_encode_fn fn = get_tmap_encode();
  if (!fn) return fail(FX_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gd[2] = {128, bytes / 128};
  cuuint64_t gs[1] = {128};
  cuuint32_t bx[2] = {128, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FX_ERR_CUDA, "cuTensorMapEncodeTiled (scale atoms) failed (%d)", (int)r);
  return FX_OK;
}

template <int NCTA, int BN, int EPI>
static int launch_fp4(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tsa, const CUtensorMap& tsb, const CUtensorMap& to,
                      const Gemm4Params& q, cudaStream_t st) {
  using Cfg = G4Cfg<NCTA, BN>;
  auto kern = gemm_nvfp4_kernel<NCTA, BN, EPI>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM); });
  if (attr_err != cudaSuccess) return fail(FX_ERR_CUDA, "gemm_fp4 smem attribute: %s", cudaGetErrorString(attr_err));
  const int units = num_sms() / NCTA;
  const int grid = (q.g.num_tiles < units ? q.g.num_tiles : units) * NCTA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(G4_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tw, tsa, tsb, to, q);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(FX_ERR_CUDA, "gemm_nvfp4_kernel launch: %s", cudaGetErrorString(e));
  }
  return launched("gemm_nvfp4_kernel");
}

// CTA pairs (256-row tiles) when every batch element is a whole number of 256-row tiles; FX_GEMM4_NCTA=1 forces single CTAs
static int fp4_ncta(int rows) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("FX_GEMM4_NCTA");
    forced = e ? atoi(e) : 0;
  }
  return (forced == 1 || rows % 256 != 0) ? 1 : 2;
}

// shapes, tiling and the four tensor maps shared by the two entry points (bn = column tile: 192 generic, 128 QKV)
static int fp4_setup(Gemm4Params& q, const void* A, const void* sfa, const void* W, const void* sfw, int batch, int rows, int N, int K,
                     int bn, int ncta, CUtensorMap* ta, CUtensorMap* tw, CUtensorMap* tsa, CUtensorMap* tsb) {
  GemmParams& p = q.g;
  p.batch = batch; p.rows = rows; p.N = N; p.K = K;
  p.k_blocks = K / 256;
  q.k_groups = K / 64;
  p.a_scale_bs = rows;
  p.tiles_m_per_batch = (rows + GEMM_BM * ncta - 1) / (GEMM_BM * ncta);
  p.tiles_m = p.tiles_m_per_batch * batch;
  p.tiles_n = (N + bn - 1) / bn;
  p.num_tiles = p.tiles_m * p.tiles_n;
  p.group_m = p.tiles_m < 8 ? p.tiles_m : 8;
  p.group_n = 0;
  p.stream_out = 1;
  {
    const char* e = getenv("FX_GEMM4_DBG_NOEPI");
    p.dbg_skip_w = e ? atoi(e) : 0;
    // FX_GEMM4_SFB_MCAST=1: the pair's leader fetches the W scale atoms once and multicasts them (6 % fewer bytes leave the L2).
    // Measured: no difference (+-0.3 % on linear2 / fc2 / qkv / fc1, profiles/r02_nvfp4_sfb_multicast.txt) -- off by default.
    static int mcast = -1;
    if (mcast < 0) {
      const char* m = getenv("FX_GEMM4_SFB_MCAST");
      mcast = m ? atoi(m) : 0;
    }
    q.sfb_mcast = mcast;
  }
  const uint64_t rows_total = (uint64_t)batch * rows;
  {  // A: flattened rows [batch * rows][K / 2] bytes (the quantiser writes a compact operand)
    const uint64_t dims[2] = {(uint64_t)K / 2, rows_total};
    const uint64_t strides[1] = {(uint64_t)K / 2};
    const uint32_t box[2] = {128, GEMM_BM};
    int rc = make_tmap_bf16(ta, A, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K / 2, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)K / 2};
    const uint32_t box[2] = {128, (uint32_t)(bn / ncta)};
    int rc = make_tmap_bf16(tw, W, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  const int nsfb = (bn + 127) / 128;
  int rc = make_tmap_sf(tsa, sfa, ((rows_total + 127) / 128) * (uint64_t)q.k_groups * 512, 16);
  if (rc) return rc;
  return make_tmap_sf(tsb, sfw, (uint64_t)p.tiles_n * q.k_groups * nsfb * 512, 16 * nsfb);
}

extern "C" int fx_gemm_fp4(const fx_gemm4_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->A && a->W && a->sfa && a->sfw && a->a_scale && a->w_scale && (a->out || a->q_out), "fx_gemm_fp4: null pointer");
  FX_REQUIRE(a->batch > 0 && a->rows > 0 && a->N > 0 && a->K > 0, "fx_gemm_fp4: empty problem");
  FX_REQUIRE(a->K % 256 == 0, "fx_gemm_fp4: K (%d) must be a multiple of 256", a->K);
  FX_REQUIRE(a->rows % 128 == 0 || a->batch == 1, "fx_gemm_fp4: rows per batch element (%d) must be a multiple of 128", a->rows);
  FX_REQUIRE(aligned16(a->A) && aligned16(a->W) && aligned16(a->sfa) && aligned16(a->sfw), "fx_gemm_fp4: operands must be 16-byte aligned");
  const int ncta = fp4_ncta(a->rows);
  Gemm4Params q{};
  GemmParams& p = q.g;
  CUtensorMap ta, tw, tsa, tsb;
  int rc = fp4_setup(q, a->A, a->sfa, a->W, a->sfw, a->batch, a->rows, a->N, a->K, G4_BN, ncta, &ta, &tw, &tsa, &tsb);
  if (rc) return rc;
  p.a_scale = a->a_scale; p.w_scale = a->w_scale;
  p.bias = (const __nv_bfloat16*)a->bias;
  p.out = a->out; p.ldo = a->ldo; p.out_bs = a->out_bs; p.out_f32 = a->out_f32; p.act = a->act;
  p.gate = (const __nv_bfloat16*)a->gate; p.gate_bs = a->gate_bs;
  p.resid = (const __nv_bfloat16*)a->resid; p.ldr = a->ldr; p.resid_bs = a->resid_bs;
  if (a->q_out) {  // emit the next GEMM's NVFP4 operand instead of `out`
    FX_REQUIRE(a->sf_out && a->e_out && !a->gate && !a->resid, "fx_gemm_fp4: q_out needs sf_out and e_out and takes no gate / residual");
    FX_REQUIRE(a->N % 64 == 0 && a->out_kc % 128 == 0 && a->out_col0 % 64 == 0 && a->out_col0 + a->N <= a->out_kc &&
                   (a->rows % 128 == 0) && aligned16(a->q_out) && aligned16(a->sf_out),
               "fx_gemm_fp4: q_out needs N %% 64 == 0, rows %% 128 == 0, out_kc %% 128 == 0, out_col0 %% 64 == 0 inside out_kc");
    q.out4 = ChunkQ{(uint8_t*)a->q_out, (uint8_t*)a->sf_out, (int8_t*)a->e_out, a->out_kc, a->out_col0};
  }
  // bf16 outputs leave through TMA stores when the output view is TMA-addressable (16-byte aligned base and strides)
  CUtensorMap to = tsa;
  static int tma_out = -1;
  if (tma_out < 0) {
    const char* e = getenv("FX_GEMM4_TMA_OUT");
    tma_out = e ? atoi(e) : 1;
  }
  q.tma_out = 0;
  if (tma_out && !a->q_out && !a->out_f32 && aligned16(a->out) && a->ldo % 8 == 0 && a->out_bs % 8 == 0 && a->N % 8 == 0) {
    const uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)a->rows, (uint64_t)a->batch};
    const uint64_t strides[2] = {(uint64_t)a->ldo * 2, (uint64_t)(a->batch > 1 ? a->out_bs : (long long)a->rows * a->ldo) * 2};
    const uint32_t box[3] = {32, 32, 1};
    rc = make_tmap_bf16(&to, a->out, 3, dims, strides, box, false, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    q.tma_out = 1;
  }
  if (ncta == 2) return launch_fp4<2, G4_BN, EPI_GENERIC>(ta, tw, tsa, tsb, to, q, (cudaStream_t)stream);
  return launch_fp4<1, G4_BN, EPI_GENERIC>(ta, tw, tsa, tsb, to, q, (cudaStream_t)stream);
}

extern "C" int fx_gemm_fp4_qkv(const fx_gemm4_qkv_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->A && a->W && a->sfa && a->sfw && a->a_scale && a->w_scale && a->q && a->k && a->v && a->pe && a->q_scale && a->k_scale,
             "fx_gemm_fp4_qkv: null pointer");
  FX_REQUIRE(a->batch > 0 && a->rows > 0 && a->K > 0 && a->heads > 0, "fx_gemm_fp4_qkv: empty problem");
  FX_REQUIRE(a->K % 256 == 0, "fx_gemm_fp4_qkv: K (%d) must be a multiple of 256", a->K);
  FX_REQUIRE(a->rows % 128 == 0 || a->batch == 1, "fx_gemm_fp4_qkv: rows per batch element (%d) must be a multiple of 128", a->rows);
  FX_REQUIRE(a->seq_off >= 0 && a->seq_off + a->rows <= a->seq_total, "fx_gemm_fp4_qkv: rows exceed seq_total");
  FX_REQUIRE(!a->pe_blocked || (a->seq_off % 32 == 0), "fx_gemm_fp4_qkv: the blocked pe layout needs seq_off %% 32 == 0");
  FX_REQUIRE(aligned16(a->A) && aligned16(a->W) && aligned16(a->sfa) && aligned16(a->sfw) && aligned16(a->pe),
             "fx_gemm_fp4_qkv: operands must be 16-byte aligned");
  FX_REQUIRE(!a->qkv_fp8 || (aligned16(a->q) && aligned16(a->k) && aligned16(a->v)), "fx_gemm_fp4_qkv: q/k/v must be 16-byte aligned");
  const int ncta = fp4_ncta(a->rows);
  Gemm4Params q{};
  GemmParams& p = q.g;
  CUtensorMap ta, tw, tsa, tsb;
  int rc = fp4_setup(q, a->A, a->sfa, a->W, a->sfw, a->batch, a->rows, 3 * a->heads * 128, a->K, 128, ncta, &ta, &tw, &tsa, &tsb);
  if (rc) return rc;
  p.a_scale = a->a_scale; p.w_scale = a->w_scale;
  p.bias = (const __nv_bfloat16*)a->bias;
  p.heads = a->heads; p.seq_total = a->seq_total; p.seq_off = a->seq_off; p.rms_eps = a->rms_eps;
  p.qnorm_w = (const __nv_bfloat16*)a->q_scale; p.knorm_w = (const __nv_bfloat16*)a->k_scale;
  p.pe = (const uint32_t*)a->pe;
  p.pe_blocked = a->pe_blocked;
  p.q = (__nv_bfloat16*)a->q; p.k = (__nv_bfloat16*)a->k; p.v = (__nv_bfloat16*)a->v;
  p.qkv_f8 = a->qkv_fp8 ? 1 : 0;
  if (ncta == 2) return launch_fp4<2, 128, EPI_QKV>(ta, tw, tsa, tsb, tsa, q, (cudaStream_t)stream);
  return launch_fp4<1, 128, EPI_QKV>(ta, tw, tsa, tsb, tsa, q, (cudaStream_t)stream);
}

extern "C" int fx_quantize_chunks_fp4(const fx_quant4c_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->x && a->q && a->sf && a->e, "fx_quantize_chunks_fp4: null pointer");
  FX_REQUIRE(a->C > 0 && a->C % 64 == 0 && a->kc % 128 == 0 && a->col0 % 64 == 0 && a->col0 + a->C <= a->kc && a->ldx % 8 == 0 &&
                 a->x_bs % 8 == 0 && aligned16(a->x) && aligned16(a->q) && aligned16(a->sf),
             "fx_quantize_chunks_fp4: C %% 64, kc %% 128, col0 %% 64, 16-byte aligned rows required");
  if (a->batch <= 0 || a->rows <= 0) return FX_OK;
  QuantChunksParams p{(const __nv_bfloat16*)a->x, a->ldx, a->x_bs, a->batch, a->rows, a->C,
                      ChunkQ{(uint8_t*)a->q, (uint8_t*)a->sf, (int8_t*)a->e, a->kc, a->col0}};
  const long long chunks = (long long)a->batch * a->rows * (a->C / 32);
  quantize_chunks_fp4_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  return launched("quantize_chunks_fp4_kernel");
}

extern "C" int fx_fp4_finalize(void* sf, const void* e, float* scale, int64_t rows, int32_t kc, fx_stream stream) {
  FX_REQUIRE(sf && e && scale && rows > 0 && rows % 128 == 0 && kc > 0 && kc % 128 == 0 && aligned16(sf) &&
                 (reinterpret_cast<uintptr_t>(e) & 7) == 0,
             "fx_fp4_finalize: rows and kc must be multiples of 128, buffers aligned");
  fp4_finalize_kernel<<<(unsigned)(rows / 128), 256, 0, (cudaStream_t)stream>>>((uint8_t*)sf, (const int8_t*)e, scale, kc);
  return launched("fp4_finalize_kernel");
}
