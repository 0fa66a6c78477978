// Host launchers for the tcgen05 GEMM family (fx_gemm, fx_gemm_qkv, fx_conv3x3).
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <mutex>

#include "api_common.cuh"
#include "gemm.cuh"
#include "tmap.cuh"

namespace fx {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static unsigned long long l2_policy(int code) { return code == 1 ? kL2EvictFirst : (code == 2 ? kL2EvictLast : kL2EvictNormal); }

// L2 management (the GEMMs stream 1-13 GB per launch through a 126 MB L2; measured with tests/gpu_l2_probe.py under
// ncu + tests/gpu_microbench.py, profiles/r01_l2_experiments.txt):
//  * eviction-priority hints on the tile loads: EVICT_FIRST on either operand multiplies the DRAM reads (linear1 3.6 ->
//    10.7 GB, +18 % sustained time): the CTAs of a wave re-use each other's tiles at normal priority only.  EVICT_LAST
//    changes nothing.  Default: normal / normal (FX_GEMM_HINT_A / _W: 0 normal, 1 first, 2 last).
//  * outputs the kernel never re-reads are stored evict-first (st.global.cs; FX_GEMM_STCS=0 disables): -7 % DRAM reads.
//  * wave lockstep (gemm.cuh), opt-in: FX_GEMM_LOCKSTEP = k-blocks per group (0 off, default), FX_GEMM_LS_SLACK = groups
//    of run-ahead.  For the long-K GEMMs it cuts linear2's DRAM reads from 12.9 to 4.4 GB and its isolated sustained time
//    from 2.87 to 2.74 ms, but inside the pipeline (clocks set by the whole mix) the coupling costs more than the DRAM
//    energy it saves: bench 4.017 -> 3.980 images/s.  Measured and left off.
static void fill_l2_policy(GemmParams& p, int esz) {
  static int ha = env_int("FX_GEMM_HINT_A", 0), hw = env_int("FX_GEMM_HINT_W", 0), cs = env_int("FX_GEMM_STCS", 1);
  p.hint_a = l2_policy(ha);
  p.hint_w = l2_policy(hw);
  p.stream_out = cs;
  // (a group must outlast the sync warp's atomic + poll round trip, ~2 us: 16 k-blocks = 6 us of bf16 MMAs)
  static int ls = env_int("FX_GEMM_LOCKSTEP", -1), slack = env_int("FX_GEMM_LS_SLACK", 2);
  p.ls_group = ls > 0 ? ls : 0;
  (void)esz;
  p.ls_slack = slack < 1 ? 1 : slack;
  p.dbg_skip_w = env_int("FX_GEMM_DBG_SKIP_W", 0);
}

static void fill_tiling(GemmParams& p, int tiles_m_per_batch, int bn) {
  p.tiles_m_per_batch = tiles_m_per_batch;
  p.tiles_m = tiles_m_per_batch * p.batch;
  p.tiles_n = (p.N + bn - 1) / bn;
  p.num_tiles = p.tiles_m * p.tiles_n;
  // raster: GROUP_M m-tiles x all n-tiles per group; a wave of ~148 CTAs then touches
  // group_m A-tiles and ~148/group_m W-slabs -> both operands are re-read from L2, not HBM.
  // Narrow outputs (N = 3072: 12 column tiles) run all their column tiles of ~6 row tiles in one wave, so every
  // A k-slice is fetched once per wave; wide outputs keep 16 row tiles per group (measured, sustained regime:
  // linear2 1136 -> 1159 TFLOP/s with small groups, linear1 / fc1 best at 16).  FX_GEMM_GROUP_M overrides.
  // (the raster knobs are read on every call -- a getenv is nanoseconds -- so one tuning process can sweep them)
  const int forced = env_int("FX_GEMM_GROUP_M", 0);
  int gm = forced > 0 ? forced : (p.tiles_n <= 16 ? 6 : 16);
  p.group_m = p.tiles_m < gm ? p.tiles_m : gm;
  // column bands for the long-K narrow-N members (see gemm_tile_coords): FX_GEMM_GROUP_N n-tiles per band when
  // K >= FX_GEMM_GROUP_N_MINK; FX_GEMM_GROUP_M_BAND overrides group_m for banded launches
  const int gn = env_int("FX_GEMM_GROUP_N", 0), gn_mink = env_int("FX_GEMM_GROUP_N_MINK", 8192),
            gm_band = env_int("FX_GEMM_GROUP_M_BAND", 0);
  p.group_n = 0;
  if (gn > 0 && p.K >= gn_mink && p.tiles_n > gn) {
    p.group_n = gn;
    if (gm_band > 0) p.group_m = p.tiles_m < gm_band ? p.tiles_m : gm_band;
  }
}

template <int BN, int EPI, bool CONV, int NCTA = 1, bool F8 = false, int CL = 1>
static int launch(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN, NCTA>;
  auto kern = gemm_kernel<BN, EPI, CONV, NCTA, F8, CL>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return fail(FX_ERR_CUDA, "gemm smem attribute: %s", cudaGetErrorString(attr_err));
  // persistent: one CTA (or CTA pair) per SM (pair); clusters of two pairs: only 33 clusters of 4 CTAs are co-resident on
  // the 148 SMs (cudaOccupancyMaxActiveClusters; the GPCs hold 16 / 18 / 20 SMs)
  const int units = CL == 2 ? 33 : num_sms() / NCTA;
  const int grid = (p.num_tiles < units ? p.num_tiles : units) * NCTA * CL;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA * CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tw, p);
  if (e != cudaSuccess) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaGetLastError();
    return fail(FX_ERR_CUDA, "gemm_kernel launch: %s", cudaGetErrorString(e));
  }
  return launched("gemm_kernel");
}

// CTA pairs (256 x 256 tiles, cta_group::2) for the wide dense GEMMs; FX_GEMM_NCTA=1 forces single-CTA tiles.
static int want_ncta(int bn) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("FX_GEMM_NCTA");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 1 || forced == 2) return bn >= 128 ? forced : 1;
  return bn >= 128 ? 2 : 1;
}

template <int EPI, bool CONV>
static int launch_bn(int bn, int ncta, const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, cudaStream_t st) {
  if (bn == 256 && ncta == 2) return launch<256, EPI, CONV, 2>(ta, tw, p, st);
  if (bn == 128 && ncta == 2) return launch<128, EPI, CONV, 2>(ta, tw, p, st);
  if (bn == 256) return launch<256, EPI, CONV>(ta, tw, p, st);
  if (bn == 128) return launch<128, EPI, CONV>(ta, tw, p, st);
  return launch<64, EPI, CONV>(ta, tw, p, st);
}

static int pick_bn(int N) { return N > 128 ? 256 : (N > 64 ? 128 : 64); }

// Clusters of two CTA pairs sharing (multicasting) their W tile (gemm.cuh, CL = 2): FX_GEMM_CL=2 enables them for the wide
// CTA-pair GEMMs with at least `FX_GEMM_CL_MIN_TILES` row tiles.  Returns 2 and rewrites the raster to super row tiles.
static int want_cluster(GemmParams& p, int bn, int ncta) {
  static int cl = env_int("FX_GEMM_CL", 1), min_tiles = env_int("FX_GEMM_CL_MIN_TILES", 16);
  if (cl != 2 || ncta != 2 || bn != 256 || p.tiles_m < min_tiles) return 1;
  p.tiles_m_real = p.tiles_m;
  p.tiles_m = (p.tiles_m + 1) / 2;
  p.num_tiles = p.tiles_m * p.tiles_n;
  p.group_m = (p.group_m + 1) / 2;
  if (p.group_m > p.tiles_m) p.group_m = p.tiles_m;
  p.ls_group = 0;
  return 2;
}

// FP8 (e4m3) operands: only the wide CTA-pair tiles are instantiated (the quantised path covers the big
// Linear layers of the MMDiT blocks, all N >= 3072)
template <int EPI>
static int launch_f8(int ncta, const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, cudaStream_t st, int cl = 1) {
  if (ncta == 2 && cl == 2) return launch<256, EPI, false, 2, true, 2>(ta, tw, p, st);
  if (ncta == 2) return launch<256, EPI, false, 2, true>(ta, tw, p, st);
  return launch<256, EPI, false, 1, true>(ta, tw, p, st);
}

// A [batch][rows][K] and W [N][K] tensor maps; esz = bytes per element (2 bf16, 1 e4m3); the box is always
// 128 bytes of K by (128 | bn / ncta) rows
static int make_operand_maps(CUtensorMap* ta, CUtensorMap* tw, const void* A, int64_t lda, int64_t a_bs, const void* W,
                             int64_t ldw, int batch, int rows, int N, int K, int bn, int ncta, int esz, int cl = 1) {
  const bool u8 = esz == 1;
  const uint32_t bk = 128 / esz;
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)rows, (uint64_t)batch};
    const uint64_t strides[2] = {(uint64_t)lda * esz, (uint64_t)(batch > 1 ? a_bs : lda * (int64_t)rows) * esz};
    const uint32_t box[3] = {bk, GEMM_BM, 1};
    int rc = make_tmap_bf16(ta, A, 3, dims, strides, box, u8);
    if (rc) return rc;
  }
  const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
  const uint64_t strides[1] = {(uint64_t)ldw * esz};
  const uint32_t box[2] = {bk, (uint32_t)(bn / ncta / cl)};  // cl = 2: each CTA fetches (and multicasts) a quarter of the tile
  return make_tmap_bf16(tw, W, 2, dims, strides, box, u8);
}

}  // namespace fx

using namespace fx;

extern "C" int fx_gemm(const fx_gemm_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->A && a->W && a->out, "fx_gemm: null pointer");
  FX_REQUIRE(a->batch > 0 && a->rows > 0 && a->N > 0 && a->K > 0, "fx_gemm: empty problem (batch %d rows %d N %d K %d)",
             a->batch, a->rows, a->N, a->K);
  const int esz = a->fp8 ? 1 : 2;
  FX_REQUIRE(a->K % (16 / esz) == 0 && a->lda % (16 / esz) == 0 && a->ldw % (16 / esz) == 0 && a->a_bs % (16 / esz) == 0,
             "fx_gemm: K, lda, ldw, a_bs must be multiples of 16 bytes (TMA strides)");
  FX_REQUIRE(aligned16(a->A) && aligned16(a->W), "fx_gemm: A and W must be 16-byte aligned");
  FX_REQUIRE(!a->fp8 || (a->a_scale && a->w_scale && a->N > 128), "fx_gemm: fp8 needs a_scale, w_scale and N > 128");
  GemmParams p{};
  p.batch = a->batch; p.rows = a->rows; p.N = a->N; p.K = a->K;
  p.k_blocks = (a->K * esz + 127) / 128;
  p.a_scale = a->a_scale; p.a_scale_bs = a->a_scale_bs; p.w_scale = a->w_scale;
  p.bias = (const __nv_bfloat16*)a->bias;
  p.out = a->out; p.ldo = a->ldo; p.out_bs = a->out_bs; p.out_f32 = a->out_f32; p.act = a->act;
  p.gate = (const __nv_bfloat16*)a->gate; p.gate_bs = a->gate_bs;
  p.resid = (const __nv_bfloat16*)a->resid; p.ldr = a->ldr; p.resid_bs = a->resid_bs;
  const int bn = pick_bn(a->N);
  const int ncta = want_ncta(bn);
  fill_tiling(p, (a->rows + GEMM_BM * ncta - 1) / (GEMM_BM * ncta), bn);
  fill_l2_policy(p, esz);
  const int cl = want_cluster(p, bn, ncta);
  CUtensorMap ta, tw;
  int rc = make_operand_maps(&ta, &tw, a->A, a->lda, a->a_bs, a->W, a->ldw, a->batch, a->rows, a->N, a->K, bn, ncta, esz, cl);
  if (rc) return rc;
  if (a->fp8) return launch_f8<EPI_GENERIC>(ncta, ta, tw, p, (cudaStream_t)stream, cl);
  if (cl == 2) return launch<256, EPI_GENERIC, false, 2, false, 2>(ta, tw, p, (cudaStream_t)stream);
  return launch_bn<EPI_GENERIC, false>(bn, ncta, ta, tw, p, (cudaStream_t)stream);
}

extern "C" int fx_gemm_qkv(const fx_qkv_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->A && a->W && a->q && a->k && a->v && a->pe && a->q_scale && a->k_scale, "fx_gemm_qkv: null pointer");
  FX_REQUIRE(a->batch > 0 && a->rows > 0 && a->K > 0 && a->heads > 0, "fx_gemm_qkv: empty problem");
  const int D3 = 3 * a->heads * 128;
  FX_REQUIRE(a->N >= D3 && a->N % 128 == 0, "fx_gemm_qkv: N (%d) must be >= 3*heads*128 and a multiple of 128", a->N);
  FX_REQUIRE(a->N == D3 || a->mlp_out, "fx_gemm_qkv: mlp_out required when N > 3*heads*128");
  FX_REQUIRE(a->seq_off >= 0 && a->seq_off + a->rows <= a->seq_total, "fx_gemm_qkv: rows exceed seq_total");
  const int esz = a->fp8 ? 1 : 2;
  FX_REQUIRE(a->K % (16 / esz) == 0 && a->lda % (16 / esz) == 0 && a->ldw % (16 / esz) == 0 && a->a_bs % (16 / esz) == 0 &&
                 a->ld_mlp % 8 == 0 && a->mlp_bs % 8 == 0,
             "fx_gemm_qkv: strides must be multiples of 16 bytes");
  FX_REQUIRE(!a->fp8 || (a->a_scale && a->w_scale), "fx_gemm_qkv: fp8 needs a_scale and w_scale");
  FX_REQUIRE(!a->pe_blocked || (a->seq_off % 32 == 0), "fx_gemm_qkv: the blocked pe layout needs seq_off %% 32 == 0");
  GemmParams p{};
  p.batch = a->batch; p.rows = a->rows; p.N = a->N; p.K = a->K;
  p.k_blocks = (a->K * esz + 127) / 128;
  p.a_scale = a->a_scale; p.a_scale_bs = a->a_scale_bs; p.w_scale = a->w_scale;
  p.bias = (const __nv_bfloat16*)a->bias;
  p.out = a->mlp_out; p.ldo = a->ld_mlp; p.out_bs = a->mlp_bs; p.out_f32 = 0; p.act = FX_ACT_GELU_TANH;
  p.heads = a->heads; p.seq_total = a->seq_total; p.seq_off = a->seq_off; p.rms_eps = a->rms_eps;
  p.qnorm_w = (const __nv_bfloat16*)a->q_scale; p.knorm_w = (const __nv_bfloat16*)a->k_scale;
  p.pe = (const uint32_t*)a->pe;
  p.pe_blocked = a->pe_blocked;
  p.q = (__nv_bfloat16*)a->q; p.k = (__nv_bfloat16*)a->k; p.v = (__nv_bfloat16*)a->v;
  p.qkv_f8 = a->qkv_fp8 ? 1 : 0;
  FX_REQUIRE(!p.qkv_f8 || (aligned16(a->q) && aligned16(a->k) && aligned16(a->v)), "fx_gemm_qkv: q/k/v must be 16-byte aligned");
  const int ncta = want_ncta(256);
  fill_tiling(p, (a->rows + GEMM_BM * ncta - 1) / (GEMM_BM * ncta), 256);
  fill_l2_policy(p, esz);
  const int cl = want_cluster(p, 256, ncta);
  CUtensorMap ta, tw;
  int rc = make_operand_maps(&ta, &tw, a->A, a->lda, a->a_bs, a->W, a->ldw, a->batch, a->rows, a->N, a->K, 256, ncta, esz, cl);
  if (rc) return rc;
  if (a->fp8) return launch_f8<EPI_QKV>(ncta, ta, tw, p, (cudaStream_t)stream, cl);
  if (cl == 2) return launch<256, EPI_QKV, false, 2, false, 2>(ta, tw, p, (cudaStream_t)stream);
  if (ncta == 2) return launch<256, EPI_QKV, false, 2>(ta, tw, p, (cudaStream_t)stream);
  return launch<256, EPI_QKV, false>(ta, tw, p, (cudaStream_t)stream);
}

extern "C" int fx_conv3x3(const fx_conv3x3_args* a, fx_stream stream) {
  FX_REQUIRE(a && a->x && a->W && a->out, "fx_conv3x3: null pointer");
  FX_REQUIRE(a->batch > 0 && a->H > 0 && a->Wd > 0 && a->Cout > 0, "fx_conv3x3: empty problem");
  FX_REQUIRE(a->Cin % 64 == 0 && a->Cin > 0, "fx_conv3x3: Cin (%d) must be a multiple of 64", a->Cin);
  const int up = a->upsample2x ? 1 : 0;
  const int taps = up ? 4 : 9;
  FX_REQUIRE(!up || (a->Cout % 128 == 0 && !a->resid && !a->out_f32 && a->batch < (1 << 28)),
             "fx_conv3x3: upsample2x needs Cout %% 128 == 0, a bf16 output and no residual");
  GemmParams p{};
  // upsample2x: four parity convolutions per image ride in the batch index of the tile raster (b' = 4 * image + parity)
  p.batch = up ? 4 * a->batch : a->batch;
  p.rows = a->H * a->Wd; p.N = a->Cout; p.K = taps * a->Cin;
  p.cin_blocks = a->Cin / GEMM_BK;
  p.k_blocks = taps * p.cin_blocks;
  p.conv_H = a->H; p.conv_W = a->Wd;
  p.conv_up = up;
  const int bn = pick_bn(a->Cout);
  const int ncta = want_ncta(bn);  // CTA pair: two horizontally adjacent 8x16-pixel patches form one 256-row tile
  FX_REQUIRE(!up || a->Cout % bn == 0, "fx_conv3x3: upsample2x needs Cout to be a multiple of the column tile (%d)", bn);
  p.conv_tiles_x = (a->Wd + 16 * ncta - 1) / (16 * ncta);
  p.conv_tiles_y = (a->H + 7) / 8;
  p.bias = (const __nv_bfloat16*)a->bias;
  p.out = a->out; p.out_f32 = a->out_f32;
  if (up) {  // output pixel (2y + py, 2x + px): pixels of one parity are 2 Cout apart, their lines 2 * (2W) * Cout
    p.ldo = 2ll * a->Cout;
    p.conv_line = 4ll * a->Wd * a->Cout;
    p.out_bs = 4ll * a->H * a->Wd * a->Cout;
  } else {
    p.ldo = a->Cout;
    p.conv_line = (long long)a->Wd * a->Cout;
    p.out_bs = (long long)a->H * a->Wd * a->Cout;
  }
  p.resid = (const __nv_bfloat16*)a->resid; p.ldr = a->Cout; p.resid_bs = (long long)a->H * a->Wd * a->Cout;
  fill_tiling(p, p.conv_tiles_x * p.conv_tiles_y, bn);
  fill_l2_policy(p, 2);
  if (a->gn_partials) {
    FX_REQUIRE(a->Cout % 128 == 0 && !a->out_f32, "fx_conv3x3: gn_partials needs Cout %% 128 == 0 and a bf16 output");
    p.gn_partials = a->gn_partials;
    p.gn_gs = a->Cout / 32;
    p.gn_nblk = p.conv_tiles_x * p.conv_tiles_y * ncta * 4 * (up ? 4 : 1);
  }
  CUtensorMap ta, tw;
  {
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->Wd, (uint64_t)a->H, (uint64_t)a->batch};
    const uint64_t strides[3] = {(uint64_t)a->Cin * 2, (uint64_t)a->Wd * a->Cin * 2, (uint64_t)a->H * a->Wd * a->Cin * 2};
    const uint32_t box[4] = {GEMM_BK, 16, 8, 1};
    int rc = make_tmap_bf16(&ta, a->x, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)a->Cout * (up ? 4 : 1)};
    const uint64_t strides[1] = {(uint64_t)p.K * 2};
    const uint32_t box[2] = {GEMM_BK, (uint32_t)(bn / ncta)};
    int rc = make_tmap_bf16(&tw, a->W, 2, dims, strides, box);
    if (rc) return rc;
  }
  return launch_bn<EPI_GENERIC, true>(bn, ncta, ta, tw, p, (cudaStream_t)stream);
}

extern "C" int64_t fx_conv3x3_gn_blocks(int32_t H, int32_t Wd, int32_t Cout, int32_t upsample2x) {
  const int ncta = want_ncta(pick_bn(Cout));
  return (int64_t)((Wd + 16 * ncta - 1) / (16 * ncta)) * ((H + 7) / 8) * ncta * 4 * (upsample2x ? 4 : 1);
}

// ------------------------------------------------------------------------------------------
// misc API
// ------------------------------------------------------------------------------------------
extern "C" int fx_version(void) { return 100; }
extern "C" const char* fx_last_error(void) { return g_err; }
extern "C" uint64_t fx_launch_count(void) { return g_launches.load(); }
extern "C" int fx_check_device(int device) {
  int major = 0, minor = 0;
  FX_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  FX_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) return fail(FX_ERR_ARCH, "device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
  return FX_OK;
}
