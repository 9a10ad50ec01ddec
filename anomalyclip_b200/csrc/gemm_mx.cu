// f16mx operands (mx.cuh): the packer used for weights and tests, and the CTA-pair GEMM that
// multiplies them (gemm_mx.cuh).
#include <cuda.h>

#include "common.h"
#include "gemm_mx.cuh"

namespace aclip {

namespace {

// one thread per 32-value block of a row
__global__ void __launch_bounds__(256)
encode_f16mx_kernel(const float* __restrict__ in, long long rows, int cols, int ld_in, MxOut out,
                    float s_main, bool vec_ok, unsigned int* __restrict__ sat) {
  const int blocks_per_row = out.ld >> 5;
  const long long total = rows * blocks_per_row;
  float amax = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / blocks_per_row;
    const int c = static_cast<int>(i - r * blocks_per_row) << 5;
    float v[32];
    const float* src = in + r * ld_in + c;
    if (vec_ok && c + 32 <= cols) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = reinterpret_cast<const float4*>(src)[j];
        v[4 * j] = a.x * s_main; v[4 * j + 1] = a.y * s_main;
        v[4 * j + 2] = a.z * s_main; v[4 * j + 3] = a.w * s_main;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (c + j < cols) ? src[j] * s_main : 0.0f;
    }
    uint32_t h[16], l4[4], c4[4], sf_l, sf_c;
    amax = fmaxf(amax, mx_pack32(v, h, l4, c4, sf_l, sf_c));
    mx_store32(out, r, c, h, l4, c4, sf_l, sf_c);
  }
  if (sat != nullptr && !(amax <= 65504.0f)) atomicAdd(sat, 1u);   // also counts NaN
}

int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 8 * (2048 / threads);
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

}  // namespace

int encode_f16mx(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out, int e_main,
                 cudaStream_t stream) {
  ACLIP_REQUIRE(in != nullptr && out != nullptr, "encode_f16mx: null pointer");
  ACLIP_REQUIRE(rows >= 0 && cols > 0 && ld_in >= cols, "encode_f16mx: bad shape");
  ACLIP_REQUIRE(ld_out % 64 == 0 && ld_out >= cols, "encode_f16mx: ld_out=%d must be a multiple of 64 >= cols",
                ld_out);
  ACLIP_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "encode_f16mx: output must be 16-byte aligned");
  ACLIP_REQUIRE(e_main >= -30 && e_main <= 30, "encode_f16mx: exponent out of range");
  if (rows == 0) return ACLIP_OK;
  const bool vec_ok = (ld_in % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  MxOut o{static_cast<uint8_t*>(out), rows * ld_out, ld_out, static_cast<int>((rows + 127) / 128)};
  const long long total = rows * (ld_out >> 5);
  timing_begin(KIND_SPLIT, stream);
  encode_f16mx_kernel<<<grid_for(total, 256), 256, 0, stream>>>(in, rows, cols, ld_in, o, exp2f((float)e_main),
                                                                vec_ok, saturation_counter());
  timing_end(KIND_SPLIT, stream, 0.0, (double)rows * (4.0 * cols + 3.1 * ld_out));
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

int make_tmap(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
              const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapDataType dtype,
              CUtensorMapSwizzle swizzle);   // gemm.cu

namespace {

template <int EPI>
int launch_mx(const CUtensorMap (&tm)[6], const GemmParams& p, const MxOut& mo, int max_ctas, cudaStream_t stream) {
  using Cfg = GemmMxCfg;
  auto kernel = gemm2mx_tcgen05_kernel<EPI>;
  static PerDeviceOnce once;
  int once_dev;
  if (once.need(once_dev)) {
    ACLIP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    once.mark(once_dev);
  }
  const int tiles = ((p.M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M) * (p.N / Cfg::BLOCK_N);
  int clusters = (max_ctas > 0 ? max_ctas : sm_count()) / 2;
  if (clusters > tiles) clusters = tiles;
  if (clusters < 1) clusters = 1;
  timing_begin(KIND_GEMM, stream);
  ACLIP_CUDA_OK(launch_serial(0, kernel, dim3(2 * clusters), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, tm[0], tm[1],
                           tm[2], tm[3], tm[4], tm[5], p, mo));
  {
    const double out_b = (p.out_f32 ? 4.0 : 0.0) + (EPI == 4 ? 3.06 : p.out_split ? 4.0 : 0.0) + (p.residual ? 4.0 : 0.0);
    timing_end(KIND_GEMM, stream, 2.0 * p.M * (double)p.N * p.K,
               3.06 * ((double)p.M * p.K + (double)p.N * p.K) + out_b * (double)p.M * p.N);
  }
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

}  // namespace

// passes == 7: A = f16mx tensor [M][lda] at g.a (a_plane_stride = M * lda), W = f16mx tensor [N][ldw]
// at g.w (w_plane_stride = N * ldw); out_enc == 3: out_split is an f16mx tensor [*][ld_split] with
// split_plane_stride = rows * ld_split.
int gemm_mx(const AclipGemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmMxCfg;
  ACLIP_REQUIRE(g.a != nullptr && g.w != nullptr, "gemm(mx): null operand");
  ACLIP_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.a_mode == 0, "gemm(mx): linear A, non-empty problem");
  ACLIP_REQUIRE(g.N % Cfg::BLOCK_N == 0 && g.K % 64 == 0 && g.lda % 64 == 0 && g.ldw % 64 == 0 &&
                    g.lda >= g.K && g.ldw >= g.K,
                "gemm(mx): N %% 192, K %% 64, pitches %% 64 (N=%d K=%d lda=%d ldw=%d)", g.N, g.K, g.lda, g.ldw);
  ACLIP_REQUIRE(g.a_plane_stride == (long long)g.M * g.lda && g.w_plane_stride == (long long)g.N * g.ldw,
                "gemm(mx): plane strides must be rows * pitch");
  ACLIP_REQUIRE((reinterpret_cast<uintptr_t>(g.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w) & 15) == 0,
                "gemm(mx): operands must be 16-byte aligned");
  ACLIP_REQUIRE(g.out_scale > 0.0f, "gemm(mx): needs out_scale = 2^-(e_act + e_weight)");
  ACLIP_REQUIRE(g.out_f32 != nullptr || g.out_split != nullptr, "gemm(mx): no output");
  ACLIP_REQUIRE(g.out_f32 == nullptr || (g.ldc % 4 == 0 && g.ldc >= g.N), "gemm(mx): ldc=%d invalid", g.ldc);
  ACLIP_REQUIRE(g.residual == nullptr || (g.ldr % 4 == 0 && g.ldr >= g.N), "gemm(mx): ldr=%d invalid", g.ldr);
  ACLIP_REQUIRE(g.act >= 0 && g.act <= 2, "gemm(mx): unknown activation %d", g.act);
  ACLIP_REQUIRE(g.gather == nullptr && g.row_group <= 0, "gemm(mx): no row map / peer gather");

  GemmParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.num_kb = g.K / 64;
  p.bias = g.bias;
  p.residual = g.residual; p.res_mod = g.res_mod; p.ldr = g.ldr;
  p.act = g.act;
  p.out_f32 = g.out_f32; p.ldc = g.ldc;
  p.row_group = 1; p.row_group_stride = 0; p.row_offset = g.row_offset;
  p.out_scale = g.out_scale;
  p.sat = saturation_counter();
  {
    static const int debug_mask = [] {
      const char* allow = getenv("ACLIP_PROFILING_EXPERIMENTS");
      const char* dbg = getenv("ACLIP_GEMM_DEBUG");
      return (allow != nullptr && allow[0] == '1' && dbg != nullptr) ? atoi(dbg) : 0;
    }();
    p.debug = debug_mask;
  }
  MxOut mo{};
  int epi = 0;
  if (g.out_split != nullptr && g.out_enc == 3) {
    ACLIP_REQUIRE(g.out_f32 == nullptr && g.residual == nullptr, "gemm(mx): an f16mx output excludes fp32 output / residual");
    ACLIP_REQUIRE(g.ld_split % 64 == 0 && g.ld_split >= g.N && g.split_plane_stride > 0 &&
                      g.split_plane_stride % g.ld_split == 0 && (reinterpret_cast<uintptr_t>(g.out_split) & 15) == 0,
                  "gemm(mx): f16mx output needs pitch %% 64 and split_plane_stride = rows * pitch");
    const long long rows = g.split_plane_stride / g.ld_split;
    ACLIP_REQUIRE(rows >= g.M + g.row_offset, "gemm(mx): f16mx output has fewer rows than the result");
    mo = MxOut{static_cast<uint8_t*>(g.out_split), g.split_plane_stride, g.ld_split, static_cast<int>((rows + 127) / 128)};
    epi = 4;
  } else {
    ACLIP_REQUIRE(g.out_split == nullptr || g.out_enc == 0, "gemm(mx): split outputs are bf16 hi/lo (0) or f16mx (3)");
    p.out_split = static_cast<__nv_bfloat16*>(g.out_split);
    p.split_plane_stride = g.split_plane_stride;
    p.ld_split = g.ld_split > 0 ? g.ld_split : g.ldc;
    if (encoded_epilogue_kind(p) == 3) epi = 3;
  }

  CUtensorMap tm[6];
  for (int op = 0; op < 2; ++op) {
    const uint8_t* base = static_cast<const uint8_t*>(op == 0 ? g.a : g.w);
    const cuuint64_t rows = op == 0 ? g.M : g.N, ld = op == 0 ? g.lda : g.ldw, P = rows * ld;
    const cuuint32_t box_rows = op == 0 ? Cfg::CTA_M : Cfg::CTA_N;
    cuuint64_t dims_h[3] = {(cuuint64_t)g.K, rows, 1};
    cuuint64_t str_h[2] = {ld * 2, P * 2};
    cuuint32_t box_h[3] = {64, box_rows, 1};
    ACLIP_TRY(make_tmap(&tm[3 * op], base, 3, dims_h, str_h, box_h, CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                        CU_TENSOR_MAP_SWIZZLE_128B));
    cuuint64_t dims_q[3] = {(cuuint64_t)g.K / 2, rows, 2};
    cuuint64_t str_q[2] = {ld / 2, P / 2};
    cuuint32_t box_q[3] = {32, box_rows, 2};
    ACLIP_TRY(make_tmap(&tm[3 * op + 1], base + 2 * P, 3, dims_q, str_q, box_q, CU_TENSOR_MAP_DATA_TYPE_UINT8,
                        CU_TENSOR_MAP_SWIZZLE_32B));
    const cuuint64_t blocks = (rows + 127) / 128, atoms = ld / 64;
    cuuint64_t dims_s[3] = {128, blocks, atoms};
    cuuint64_t str_s[2] = {512, blocks * 512};
    cuuint32_t box_s[3] = {128, op == 0 ? 1u : 2u, 1};
    ACLIP_TRY(make_tmap(&tm[3 * op + 2], base + 3 * P, 3, dims_s, str_s, box_s, CU_TENSOR_MAP_DATA_TYPE_UINT32,
                        CU_TENSOR_MAP_SWIZZLE_NONE));
  }
  if (epi == 4) return launch_mx<4>(tm, p, mo, g.max_ctas, stream);
  if (epi == 3) return launch_mx<3>(tm, p, mo, g.max_ctas, stream);
  return launch_mx<0>(tm, p, mo, g.max_ctas, stream);
}

}  // namespace aclip

extern "C" long long aclip_f16mx_bytes(long long rows, int ld) {
  if (rows < 0 || ld <= 0 || ld % 64 != 0) return -1;
  return 3 * rows * ld + (long long)(ld / 64) * ((rows + 127) / 128) * 512;
}

extern "C" int aclip_encode_f16mx(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out,
                                  int e_main, void* stream) {
  return aclip::encode_f16mx(in, rows, cols, ld_in, out, ld_out, e_main, aclip::as_stream(stream));
}
