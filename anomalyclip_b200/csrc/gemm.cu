// Host launcher for the tcgen05 GEMM: builds the TMA tensor maps and enqueues the kernel.
#include "gemm.cuh"

#include <cstdarg>
#include <mutex>

#include "common.h"

namespace aclip {

// ------------------------------------------------------------------ error / device plumbing
std::string& last_error() {
  thread_local std::string msg;
  return msg;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

int sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      return 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

std::atomic<long long> g_launches{0};

// Programmatic dependent launch is OPT-IN (ACLIP_PDL=1; ACLIP_NO_PDL=1 still forces it off): with it
// about one encoder run in 1 500 - 4 000 returns the first frames of a micro-batch slightly off
// (scripts/stress_determinism.py); without it 0 of 2 800.  It is worth 1.7 % of the step.
bool pdl_enabled() {
  static const bool on = [] {
    const char* on_ = getenv("ACLIP_PDL");
    const char* off = getenv("ACLIP_NO_PDL");
    return on_ != nullptr && on_[0] == '1' && !(off != nullptr && off[0] == '1');
  }();
  return on;
}

bool pdl_mx_enabled(int kind) {
  static const int mask = [] {
    const char* v = getenv("ACLIP_MX_PDL");
    if (v == nullptr) return 0;
    return v[0] == '1' ? 3 : v[0] == 'g' ? 1 : v[0] == 'l' ? 2 : 0;
  }();
  return (mask >> kind) & 1;
}

// ------------------------------------------------------------------ tensor maps
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 tensor map with the 128-byte swizzle; dims/strides innermost first, strides in bytes for
// dims 1..rank-1.
int make_tmap(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
              const cuuint64_t* strides_bytes, const cuuint32_t* box,
              CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
              CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(ACLIP_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  cuuint32_t elem_strides[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dtype, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(base), dims, strides_bytes, box, elem_strides,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(ACLIP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return ACLIP_OK;
}

template <int BLOCK_N, int PASSES, int EPI = 0, int BLOCK_M = 128,
          int KATOMS = default_katoms(BLOCK_N, PASSES, BLOCK_M)>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                  int max_ctas, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, PASSES, BLOCK_M, KATOMS>;
  auto kernel = gemm_tcgen05_kernel<BLOCK_N, PASSES, EPI, BLOCK_M, KATOMS>;
  static PerDeviceOnce once;
  int once_dev;
  if (once.need(once_dev)) {
    ACLIP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    once.mark(once_dev);
  }
  const int m_tiles = (p.M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  int ctas = max_ctas > 0 ? max_ctas : sm_count();
  if (ctas > m_tiles * n_tiles) ctas = m_tiles * n_tiles;
  timing_begin(KIND_GEMM, stream);
  ACLIP_CUDA_OK(launch_pdl(kernel, dim3(ctas), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, p));
  {
    const double planes = (PASSES == 1 || PASSES == 4) ? 1.0 : 2.0;
    const double out_b = (p.out_f32 ? 4.0 : 0.0) + (p.out_split ? (p.out_enc == 2 ? 2.0 : 4.0) : 0.0) +
                         (p.residual ? 4.0 : 0.0);
    // algorithmic: 2MNK flops (one product per term, whatever the number of passes); bytes = each
    // operand once + outputs (+ residual) once; the conv mode reads each activation once, not 9x
    const double a_elems = p.a_mode == 1 ? (double)p.M * (p.K / 9) : (double)p.M * p.K;
    timing_end(KIND_GEMM, stream, 2.0 * p.M * (double)p.N * p.K,
               planes * 2.0 * (a_elems + (double)p.N * p.K) + out_b * (double)p.M * p.N);
  }
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

template <int PASSES, int EPI = 0>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA8,
                       const CUtensorMap& tmB8, const GemmParams& p, int max_ctas,
                       cudaStream_t stream) {
  using Cfg = Gemm2Cfg<PASSES>;
  auto kernel = gemm2_tcgen05_kernel<PASSES, EPI>;
  static PerDeviceOnce once;
  int once_dev;
  if (once.need(once_dev)) {
    ACLIP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    once.mark(once_dev);
  }
  const int m_tiles = (p.M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int n_tiles = (p.N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N;
  int clusters = (max_ctas > 0 ? max_ctas : sm_count()) / 2;
  if (clusters > m_tiles * n_tiles) clusters = m_tiles * n_tiles;
  if (clusters < 1) clusters = 1;
  timing_begin(KIND_GEMM, stream);
  ACLIP_CUDA_OK(launch_pdl(kernel, dim3(2 * clusters), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB,
                           tmA8, tmB8, p));
  {
    const double planes = (PASSES == 1 || PASSES == 4) ? 1.0 : PASSES == 6 ? 1.5 : 2.0;
    const double out_b = (p.out_f32 ? 4.0 : 0.0) + (p.out_split ? (p.out_enc == 2 ? 2.0 : 4.0) : 0.0) +
                         (p.residual ? 4.0 : 0.0);
    const double a_elems = p.a_mode == 1 ? (double)p.M * (p.K / 9) : (double)p.M * p.K;
    timing_end(KIND_GEMM, stream, 2.0 * p.M * (double)p.N * p.K,
               planes * 2.0 * (a_elems + (double)p.N * p.K) + out_b * (double)p.M * p.N);
  }
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

int gemm(const AclipGemmArgs& g, cudaStream_t stream) {
  ACLIP_REQUIRE(g.a != nullptr && g.w != nullptr, "gemm: null operand");
  ACLIP_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  if (g.passes == 7) return gemm_mx(g, stream);   // f16mx operands (gemm_mx.cu)
  ACLIP_REQUIRE((g.passes >= 1 && g.passes <= 4) || g.passes == 6,
                "gemm: passes must be 1, 2, 3, 4, 6 or 7 (got %d)", g.passes);
  ACLIP_REQUIRE(g.out_enc >= 0 && g.out_enc <= 2,
                "gemm: out_enc must be 0 (bf16 hi/lo), 1 (f16f8) or 2 (fp16 plane)");
  if (g.passes == 4) {
    // fp16 operands, one pass: A is an fp16 matrix [M][lda] (or NHWC grid), W the fp16 plane of a
    // weight packed with aclip_encode_f16f8
    ACLIP_REQUIRE(g.out_scale > 0.0f, "gemm: passes=4 needs out_scale = 2^-(e_act + e_weight)");
  }
  const bool f16f8 = g.passes == 2 || g.passes == 6;   // 6: without the weight-residual cross term
  if (f16f8) {
    // f16f8 operands (split.cuh): CTA-pair kernel only
    ACLIP_REQUIRE(g.a_mode == 0 || (g.a_mode == 1 && g.passes == 2), "gemm: unsupported a_mode %d", g.a_mode);
    ACLIP_REQUIRE(g.N % 256 == 0 && g.kernel != 1,
                  "gemm: passes=2 (f16f8 operands) runs on the CTA-pair kernel: N %% 256 == 0 (N=%d)", g.N);
    ACLIP_REQUIRE(g.lda % 16 == 0 && g.ldw % 16 == 0 && g.a_plane_stride % 16 == 0 &&
                      g.w_plane_stride % 16 == 0,
                  "gemm: passes=2 needs pitches and plane strides that are multiples of 16");
    ACLIP_REQUIRE(g.out_scale > 0.0f, "gemm: passes=2 needs out_scale = 2^-(e_act + e_weight)");
  }
  ACLIP_REQUIRE(g.N % 32 == 0, "gemm: N=%d must be a multiple of 32", g.N);
  ACLIP_REQUIRE(g.ldw % 8 == 0 && g.K % 8 == 0, "gemm: K=%d / ldw=%d must be multiples of 8", g.K,
                g.ldw);
  ACLIP_REQUIRE((reinterpret_cast<uintptr_t>(g.a) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(g.w) & 15) == 0,
                "gemm: operands must be 16-byte aligned");
  ACLIP_REQUIRE(g.out_f32 != nullptr || g.out_split != nullptr, "gemm: no output");
  ACLIP_REQUIRE(g.out_f32 == nullptr || (g.ldc % 4 == 0 && g.ldc >= g.N),
                "gemm: ldc=%d invalid for N=%d", g.ldc, g.N);
  {
    const int lds = g.ld_split > 0 ? g.ld_split : g.ldc;
    ACLIP_REQUIRE(g.out_split == nullptr ||
                      (lds % 8 == 0 && lds >= g.N && g.split_plane_stride % 8 == 0 &&
                       (reinterpret_cast<uintptr_t>(g.out_split) & 15) == 0),
                  "gemm: split output pitch %d / alignment invalid", lds);
  }
  ACLIP_REQUIRE(g.residual == nullptr || (g.ldr % 4 == 0 && g.ldr >= g.N), "gemm: ldr=%d invalid",
                g.ldr);
  ACLIP_REQUIRE(g.act >= 0 && g.act <= 2, "gemm: unknown activation %d", g.act);
  const int planes = (g.passes == 1 || g.passes == 4) ? 1 : 2;
  const CUtensorMapDataType op_dtype =
      g.passes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  // 128-wide tiles when they waste fewer padded columns than 256-wide ones (e.g. N = 128, 384)
  // CTA-pair kernel (256 x 256 tiles over two SMs) whenever N tiles evenly and there is enough
  // work to fill the machine; kernel = 1 / 2 forces the single-CTA / pair kernel (tests).
  ACLIP_REQUIRE(g.kernel == 0 || g.kernel == 1 || g.kernel == 2,
                "gemm: kernel must be 0 (auto), 1 (single CTA) or 2 (CTA pair)");
  ACLIP_REQUIRE(g.kernel != 2 || g.N % 256 == 0, "gemm: the CTA-pair kernel needs N %% 256 == 0");
  ACLIP_REQUIRE(g.tile >= 0 && g.tile <= 2, "gemm: tile must be 0 (auto), 1 (128 rows) or 2 (64 x 32)");
  // (a K-heavy, narrow GEMM -- conv2 of 8..16 sub-videos: N = 256 -- has too few 256 x 256 tiles to
  // occupy the machine: it stays on single-CTA tiles until a quarter of the SMs would be paired)
  const long long pair_tiles = ((g.M + 255) / 256) * (long long)((g.N + 255) / 256);
  const bool pair = g.kernel == 2 || f16f8 ||
                    (g.kernel == 0 && g.N % 256 == 0 && g.M >= 4096 && pair_tiles >= sm_count() / 4);
  // Single-CTA kernel: the widest tile (256, 128 or 64 columns) that still yields at least half a
  // wave of tiles; small problems (the temporal path at a few sub-videos) get narrow tiles so that
  // more SMs share the K loop.  128 is also preferred when it wastes fewer padded columns.
  // Below a quarter of a wave of 128 x 64 tiles (one or two sub-videos of the temporal stage) the
  // tile shrinks to 64 x 32 with four K atoms per barrier round: these GEMMs are bound by the
  // number of L2 round trips of their K loop, not by bytes or flops.  Same accumulation order per
  // element: results do not depend on the tile choice.  tile = 1 / 2 forces 128-row / 64-row tiles
  // (tests).
  int block_n = 128;  // rows of W per TMA box (the pair kernel loads 128 per CTA)
  int block_m = 128;  // rows of A per TMA box
  if (!pair) {
    const int m_tiles = (g.M + 127) / 128;
    const int want = sm_count() / 2;
    auto tiles = [&](int bn) { return m_tiles * ((g.N + bn - 1) / bn); };
    const bool narrow = ((g.N + 127) / 128) * 128 < ((g.N + 255) / 256) * 256;
    if (!narrow && tiles(256) >= want) block_n = 256;
    else if (tiles(128) >= want) block_n = 128;
    else block_n = 64;
    const bool small_ok = (g.passes == 3 || g.passes == 4) && g.gather == nullptr &&
                          (g.a_mode == 0 || (g.conv_w > 0 && 64 % g.conv_w == 0 &&
                                             (g.conv_h * g.conv_w) % 64 == 0));
    const bool small_auto = block_n == 64 && tiles(64) < sm_count() / 4;
    if (small_ok && g.tile != 1 && (g.tile == 2 || small_auto)) { block_m = 64; block_n = 32; }
  }

  GemmParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.num_kb = (g.K + 63) / 64;
  p.a_mode = g.a_mode;
  p.bias = g.bias;
  p.residual = g.residual;
  p.res_mod = g.res_mod;
  p.ldr = g.ldr;
  p.act = g.act;
  p.out_f32 = g.out_f32;
  p.out_split = static_cast<__nv_bfloat16*>(g.out_split);
  p.split_plane_stride = g.split_plane_stride;
  p.ldc = g.ldc;
  p.ld_split = g.ld_split > 0 ? g.ld_split : g.ldc;
  p.row_group = g.row_group > 0 ? g.row_group : 1;
  p.row_group_stride = g.row_group > 0 ? g.row_group_stride : 0;  // 0 = no remap
  ACLIP_REQUIRE(g.row_group <= 0 || g.row_group_stride > 0, "gemm: row_group_stride must be > 0 with a row remap");
  p.row_offset = g.row_offset;
  p.out_scale = g.out_scale > 0.0f ? g.out_scale : 1.0f;
  p.out_enc = g.out_enc;
  p.sat = (g.out_split != nullptr && g.out_enc != 0) ? saturation_counter() : nullptr;
  if (g.gather != nullptr) {
    // fused all-gather of the fp32 output rows (frame-sharded image encoder, SURVEY 8e option 1)
    const AclipPeerGather& pg = *g.gather;
    ACLIP_REQUIRE(g.out_f32 != nullptr && pg.world >= 1 && pg.world <= 8 && pg.rank >= 0 &&
                      pg.rank < pg.world && pg.width == g.ldc && pg.epoch > 0 && pg.counter != nullptr &&
                      g.gather_row0 >= 0 && g.gather_row0 + g.M <= pg.rows_per_rank,
                  "gemm: bad peer-gather descriptor (needs an fp32 output of pitch == width, rows inside "
                  "this rank's block)");
    p.peer_world = pg.world; p.peer_rank = pg.rank; p.peer_signal = g.gather_signal;
    p.peer_epoch = pg.epoch; p.peer_counter = pg.counter;
    for (int r = 0; r < pg.world; ++r) {
      ACLIP_REQUIRE(pg.rows[r] != nullptr && pg.flags[r] != nullptr, "gemm: null peer pointer %d", r);
      p.peer_out[r] = pg.rows[r] + (pg.rank * pg.rows_per_rank + g.gather_row0) * pg.width;
      p.peer_flags[r] = pg.flags[r];
    }
  }
  {
    // profiling experiments (results wrong by construction) only with the environment switch;
    // read once per process
    static const int debug_mask = [] {
      const char* allow = getenv("ACLIP_PROFILING_EXPERIMENTS");
      const char* dbg = getenv("ACLIP_GEMM_DEBUG");
      return (allow != nullptr && allow[0] == '1' && dbg != nullptr) ? atoi(dbg) : 0;
    }();
    p.debug = debug_mask;
  }
  ACLIP_REQUIRE(g.out_enc == 0 || g.out_split == nullptr ||
                    (p.ld_split % 16 == 0 && g.split_plane_stride % 16 == 0),
                "gemm: an f16f8 output needs a pitch and plane stride that are multiples of 16");

  CUtensorMap tmA, tmB, tmA8, tmB8;
  if (f16f8) {
    // fp16 plane: [1][rows][ld] (128-byte swizzle rows of 64 values); e4m3 planes L, C: one
    // [2][rows][ld] byte tensor starting 2 * plane_stride bytes in (64-byte swizzle rows)
    for (int op = (g.a_mode == 1 ? 1 : 0); op < 2; ++op) {
      const void* base = op == 0 ? g.a : g.w;
      const cuuint64_t rows = op == 0 ? g.M : g.N, ld = op == 0 ? g.lda : g.ldw;
      const cuuint64_t ps = op == 0 ? g.a_plane_stride : g.w_plane_stride;
      ACLIP_REQUIRE(ps >= rows * ld, "gemm: plane stride smaller than the operand");
      cuuint64_t dims_h[3] = {(cuuint64_t)g.K, rows, 1};
      cuuint64_t str_h[2] = {ld * 2, ld * 2 * rows};
      cuuint32_t box_h[3] = {64, 128u, 1};
      ACLIP_TRY(make_tmap(op == 0 ? &tmA : &tmB, base, 3, dims_h, str_h, box_h,
                          CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_128B));
      cuuint64_t dims_8[3] = {(cuuint64_t)g.K, rows, 2};
      cuuint64_t str_8[2] = {ld, ps};
      cuuint32_t box_8[3] = {64, 128u, g.passes == 6 ? 1u : 2u};   // 6: one plane per box (L of A, C of W)
      ACLIP_TRY(make_tmap(op == 0 ? &tmA8 : &tmB8, static_cast<const uint8_t*>(base) + 2 * ps, 3,
                          dims_8, str_8, box_8, CU_TENSOR_MAP_DATA_TYPE_UINT8,
                          CU_TENSOR_MAP_SWIZZLE_64B));
    }
    if (g.a_mode == 1) {
      // conv3x3 over an f16f8 NHWC grid: 5-D maps {C, W, H, S, plane} for the fp16 plane and for
      // the two e4m3 planes; a box is 64 channels x one 128-row block of the grid (x both planes)
      ACLIP_REQUIRE(g.conv_c % 64 == 0, "conv3x3: C=%d must be a multiple of 64", g.conv_c);
      ACLIP_REQUIRE(g.conv_w > 0 && 128 % g.conv_w == 0 && (g.conv_h * g.conv_w) % 128 == 0,
                    "conv3x3: grid %dx%d unsupported (need W | 128 and 128 | H*W)", g.conv_h, g.conv_w);
      ACLIP_REQUIRE(g.M == g.conv_s * g.conv_h * g.conv_w && g.K == 9 * g.conv_c,
                    "conv3x3: M/K inconsistent with the grid");
      p.conv_cin_kb = g.conv_c / 64;
      p.conv_w = g.conv_w;
      p.conv_h = g.conv_h;
      const cuuint64_t C = g.conv_c, W = g.conv_w, H = g.conv_h, S = g.conv_s;
      const cuuint64_t ps = g.a_plane_stride;
      ACLIP_REQUIRE(ps >= S * H * W * C, "gemm: plane stride smaller than the operand");
      cuuint64_t dims_h[5] = {C, W, H, S, 1};
      cuuint64_t str_h[4] = {C * 2, W * C * 2, H * W * C * 2, S * H * W * C * 2};
      cuuint32_t box_h[5] = {64, (cuuint32_t)g.conv_w, (cuuint32_t)(128 / g.conv_w), 1, 1};
      ACLIP_TRY(make_tmap(&tmA, g.a, 5, dims_h, str_h, box_h, CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                          CU_TENSOR_MAP_SWIZZLE_128B));
      cuuint64_t dims_8[5] = {C, W, H, S, 2};
      cuuint64_t str_8[4] = {C, W * C, H * W * C, ps};
      cuuint32_t box_8[5] = {64, (cuuint32_t)g.conv_w, (cuuint32_t)(128 / g.conv_w), 1, 2};
      ACLIP_TRY(make_tmap(&tmA8, static_cast<const uint8_t*>(g.a) + 2 * ps, 5, dims_8, str_8, box_8,
                          CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_SWIZZLE_64B));
    }
    const int epi = encoded_epilogue_kind(p);
    if (g.passes == 6)
      return epi == 3 ? launch_pair<6, 3>(tmA, tmB, tmA8, tmB8, p, g.max_ctas, stream)
                      : launch_pair<6, 0>(tmA, tmB, tmA8, tmB8, p, g.max_ctas, stream);
    return epi == 1   ? launch_pair<2, 1>(tmA, tmB, tmA8, tmB8, p, g.max_ctas, stream)
           : epi == 2 ? launch_pair<2, 2>(tmA, tmB, tmA8, tmB8, p, g.max_ctas, stream)
           : epi == 3 ? launch_pair<2, 3>(tmA, tmB, tmA8, tmB8, p, g.max_ctas, stream)
                      : launch_pair<2, 0>(tmA, tmB, tmA8, tmB8, p, g.max_ctas, stream);
  }
  if (g.a_mode == 0) {
    ACLIP_REQUIRE(g.lda % 8 == 0 && g.lda >= g.K, "gemm: lda=%d invalid for K=%d", g.lda, g.K);
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)g.M, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)g.lda * 2, (cuuint64_t)g.a_plane_stride * 2};
    if (planes == 1) strides[1] = (cuuint64_t)g.lda * 2 * (cuuint64_t)g.M;
    cuuint32_t box[3] = {64, (cuuint32_t)block_m, (cuuint32_t)planes};
    ACLIP_REQUIRE(planes == 1 || (g.a_plane_stride % 8 == 0 && g.a_plane_stride > 0),
                  "gemm: a_plane_stride must be a positive multiple of 8");
    ACLIP_TRY(make_tmap(&tmA, g.a, 3, dims, strides, box, op_dtype));
  } else if (g.a_mode == 1) {
    ACLIP_REQUIRE(g.conv_c % 64 == 0, "conv3x3: C=%d must be a multiple of 64", g.conv_c);
    ACLIP_REQUIRE(g.conv_w > 0 && 128 % g.conv_w == 0 && (g.conv_h * g.conv_w) % 128 == 0,
                  "conv3x3: grid %dx%d unsupported (need W | 128 and 128 | H*W)", g.conv_h,
                  g.conv_w);
    ACLIP_REQUIRE(g.M == g.conv_s * g.conv_h * g.conv_w && g.K == 9 * g.conv_c,
                  "conv3x3: M/K inconsistent with the grid");
    p.conv_cin_kb = g.conv_c / 64;
    p.conv_w = g.conv_w;
    p.conv_h = g.conv_h;
    const cuuint64_t C = g.conv_c, W = g.conv_w, H = g.conv_h, S = g.conv_s;
    cuuint64_t dims[5] = {C, W, H, S, (cuuint64_t)planes};
    cuuint64_t strides[4] = {C * 2, W * C * 2, H * W * C * 2,
                             planes == 1 ? S * H * W * C * 2 : (cuuint64_t)g.a_plane_stride * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)g.conv_w, (cuuint32_t)(block_m / g.conv_w), 1,
                         (cuuint32_t)planes};
    ACLIP_TRY(make_tmap(&tmA, g.a, 5, dims, strides, box, op_dtype));
  } else {
    return fail(ACLIP_ERR_INVALID, "gemm: unknown a_mode %d", g.a_mode);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)g.N, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)g.ldw * 2, planes == 1 ? (cuuint64_t)g.ldw * 2 * g.N
                                                                : (cuuint64_t)g.w_plane_stride * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)block_n, (cuuint32_t)planes};
    ACLIP_REQUIRE(planes == 1 || (g.w_plane_stride % 8 == 0 && g.w_plane_stride > 0),
                  "gemm: w_plane_stride must be a positive multiple of 8");
    ACLIP_TRY(make_tmap(&tmB, g.w, 3, dims, strides, box, op_dtype));
  }

  int epi = encoded_epilogue_kind(p);
  if (g.passes != 4 && epi != 3) epi = 0;   // encoded outputs pair with fp16-based operands only
  if (pair) {
    if (g.passes == 4)
      return epi == 2   ? launch_pair<4, 2>(tmA, tmB, tmA, tmB, p, g.max_ctas, stream)
             : epi == 1 ? launch_pair<4, 1>(tmA, tmB, tmA, tmB, p, g.max_ctas, stream)
             : epi == 3 ? launch_pair<4, 3>(tmA, tmB, tmA, tmB, p, g.max_ctas, stream)
                        : launch_pair<4, 0>(tmA, tmB, tmA, tmB, p, g.max_ctas, stream);
    if (g.passes == 3)
      return epi == 3 ? launch_pair<3, 3>(tmA, tmB, tmA, tmB, p, g.max_ctas, stream)
                      : launch_pair<3, 0>(tmA, tmB, tmA, tmB, p, g.max_ctas, stream);
    return launch_pair<1>(tmA, tmB, tmA, tmB, p, g.max_ctas, stream);
  }
#define ACLIP_LAUNCH_SINGLE(BN)                                                            \
  return g.passes == 3   ? (epi == 3 ? launch<BN, 3, 3>(tmA, tmB, p, g.max_ctas, stream)   \
                                     : launch<BN, 3, 0>(tmA, tmB, p, g.max_ctas, stream))  \
         : g.passes == 4 ? (epi == 2   ? launch<BN, 4, 2>(tmA, tmB, p, g.max_ctas, stream) \
                            : epi == 3 ? launch<BN, 4, 3>(tmA, tmB, p, g.max_ctas, stream) \
                                       : launch<BN, 4, 0>(tmA, tmB, p, g.max_ctas, stream)) \
                         : launch<BN, 1>(tmA, tmB, p, g.max_ctas, stream)
  if (block_m == 64) {
    // 64 x 32 tiles: the generic epilogue handles every output kind
    return g.passes == 3 ? launch<32, 3, 0, 64>(tmA, tmB, p, g.max_ctas, stream)
                         : launch<32, 4, 0, 64>(tmA, tmB, p, g.max_ctas, stream);
  }
  if (block_n == 256) { ACLIP_LAUNCH_SINGLE(256); }
  if (block_n == 128) { ACLIP_LAUNCH_SINGLE(128); }
  ACLIP_LAUNCH_SINGLE(64);
#undef ACLIP_LAUNCH_SINGLE
}

}  // namespace aclip

extern "C" int aclip_gemm(const AclipGemmArgs* args, void* stream) {
  if (args == nullptr) return aclip::fail(ACLIP_ERR_INVALID, "aclip_gemm: args is NULL");
  return aclip::gemm(*args, aclip::as_stream(stream));
}
