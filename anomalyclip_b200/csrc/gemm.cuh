// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T )
//
// The reference path is fp32 end to end (/root/reference/src/models/components/anomaly_clip.py:66-67),
// so every fp32 operand value travels in a narrow encoding whose products reproduce the fp32 product
// (split.cuh), accumulated in one fp32 TMEM accumulator:
//   PASSES == 3  "split-bf16": hi = bf16(x), lo = bf16(x - hi); hi*hi + lo*hi + hi*lo, three
//                kind::f16 MMAs per K step, ~2^-16 relative error per product
//   PASSES == 2  "f16f8": fp16 main plane + e4m3 residual and coarse planes; one kind::f16 MMA
//                (K = 16) per K step plus two kind::f8f6f4 MMAs (K = 32, twice the rate) for the
//                cross terms: two bf16-pass equivalents, ~2^-16 as well (CTA-pair kernels only)
//   PASSES == 6  f16f8 WITHOUT the weight-residual term: x_H w_H + x_L w_C only (A travels as H + L,
//                W as H + C: 3 bytes per element each, 1.5 pass-equivalents); the weights are then
//                effectively fp16, ~1.9e-4 on the ViT features when used for ONE MLP GEMM (CTA pairs)
//   PASSES == 1  plain bf16 (hi plane only)
//   PASSES == 4  "f16": the fp16 main plane only, one kind::f16 MMA per K step (both kernels);
//                ~2^-12 relative per product, selected by the host after calibration (split.cuh)
//
// Two kernels share the producer / issuer / epilogue code:
//   gemm_tcgen05_kernel   one CTA per 128 x {64,128,256} tile, or per 64 x 32    (192 threads)
//   gemm2_tcgen05_kernel  a CTA pair (cta_group::2) per 256 x 256 tile           (320 threads per CTA)
// (A four-CTA variant sharing the W tile by TMA multicast was measured in round 1: ~8 % faster per
// SM but only 33 clusters of four are co-resident on 148 SMs, a net loss; removed.)
// Roles:
//   warp 0  lane 0 : TMA producer   (A and W tiles, swizzled boxes, mbarrier complete_tx)
//   warp 1  lane 0 : MMA issuer     (tcgen05.mma)
//   other warps    : epilogue       (tcgen05.ld -> scale + bias / activation -> smem transpose ->
//                                    residual -> row-contiguous fp32 / encoded stores)
// Pipelines: smem ring full/empty (TMA <-> MMA) and a double-buffered TMEM accumulator
// full/empty (MMA <-> epilogue), so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// A can be addressed in two ways (GemmParams::a_mode):
//   0  linear  : A is [plane][M][K] row-major                       (3-D tensor map)
//   1  conv3x3 : A is an NHWC grid [plane][S][H][W][C]; the K loop runs over 9 taps x C and
//                each tap is a TMA box shifted by (dy,dx) with hardware zero fill at the
//                borders, i.e. an implicit-GEMM 3x3 "same" convolution  (5-D tensor map)
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"
#include "split.cuh"

namespace aclip {

enum : int { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_LEAKYRELU = 2 };

struct GemmParams {
  int M, N, K;
  int num_kb;  // ceil(K / 64)
  int a_mode;  // 0 linear, 1 conv3x3
  int conv_cin_kb;  // Cin / 64             (conv mode)
  int conv_w;       // grid width  W        (conv mode; 128 % W == 0)
  int conv_h;       // grid height H        (conv mode; (H*W) % 128 == 0)
  // epilogue
  const float* bias;      // [N] or nullptr
  const float* residual;  // fp32 [*, ldr] or nullptr
  int res_mod;            // >0: residual row = m % res_mod, else residual row = output row
  int ldr;
  int act;
  float* out_f32;            // fp32 [*, ldc] or nullptr
  __nv_bfloat16* out_split;  // encoded output (out_enc) or nullptr: bf16 [2][rows][ld_split] planes
                             // hi, lo, or f16f8 planes H | L | C (split.cuh)
  long long split_plane_stride;  // plane stride in elements
  int ldc;
  int ld_split;  // pitch of out_split (elements)
  // output row remap: out_row = (m / row_group) * row_group_stride + (m % row_group) + row_offset
  int row_group, row_group_stride, row_offset;
  float out_scale;  // multiplies the accumulator before bias (f16f8 operands: 2^-(ex + ew)), else 1
  int out_enc;      // encoding of out_split: 0 = bf16 hi/lo planes, 1 = f16f8 activation planes,
                    // 2 = fp16 plane only
  unsigned int* sat;  // fp16 saturation counter of this device (split.cuh) or nullptr
  // Fused all-gather of the fp32 output over NVLink peer memory (peer_world > 0): every row this
  // GEMM stores to out_f32 is also stored to peer_out[r] (same pitch, same row remap) for every
  // rank r of the group -- peer-mapped buffers, the local rank's own is one of them -- and, when
  // peer_signal is set, the last CTA publishes peer_flags[r][peer_rank] = peer_epoch on every rank
  // (system-scope release) once all CTAs have fenced their stores.  aclip_peer_wait is the consumer.
  int peer_world, peer_rank, peer_signal;
  unsigned int peer_epoch;
  unsigned int* peer_counter;
  float* peer_out[8];
  unsigned int* peer_flags[8];
  int debug;        // profiling experiments only (ACLIP_PROFILING_EXPERIMENTS=1): 1 = issue no MMAs
                    // (operand feed + epilogue only; results are wrong by construction)
};

#ifndef ACLIP_KATOMS_NARROW
#define ACLIP_KATOMS_NARROW 2
#endif
#ifndef ACLIP_KATOMS_PAIR
#define ACLIP_KATOMS_PAIR 1
#endif

// KATOMS: 64-wide K atoms (one 128-byte swizzle row each) per barrier round of the operand ring.
// A round (TMA -> mbarrier -> MMA -> commit -> producer) costs an SM ~230 ns whatever it carries:
// measured the same for 12 KB and 24 KB, one TMA operation and two, a ring 3 or 12 deep, with MMAs
// or none, 1 or 128 CTAs on the machine (profiles/r2_small_gemm_round_trip_experiments.txt).
// Tiles whose MMAs per atom are shorter than that (one product per atom on <= 128 columns) carry
// two atoms per round, the 64 x 32 tile four.
constexpr int default_katoms(int block_n, int passes, int block_m) {
  return block_m == 64 ? 4 : ((passes == 1 || passes == 4) && block_n <= 128) ? ACLIP_KATOMS_NARROW : 1;
}

template <int BLOCK_N_, int PASSES_, int BLOCK_M_ = 128,
          int KATOMS_ = default_katoms(BLOCK_N_, PASSES_, BLOCK_M_)>
struct GemmCfg {
  static constexpr int BLOCK_M = BLOCK_M_;   // 128, or 64 (accumulator rows in lanes 0..15 of each quarter)
  static constexpr int BLOCK_N = BLOCK_N_;
  static constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
  static constexpr int UMMA_K = 16;
  static constexpr int PASSES = PASSES_;
  static constexpr int PLANES = (PASSES_ == 1 || PASSES_ == 4) ? 1 : 2;
  static constexpr int A_PLANE_BYTES = BLOCK_M * 128;
  static constexpr int B_PLANE_BYTES = BLOCK_N * 128;
  static constexpr int KATOMS = KATOMS_;
  static constexpr int A_ATOM_BYTES = PLANES * A_PLANE_BYTES;   // one TMA box of A
  static constexpr int B_ATOM_BYTES = PLANES * B_PLANE_BYTES;   // one TMA box of W
  static constexpr int ATOM_BYTES = A_ATOM_BYTES + B_ATOM_BYTES;
  static constexpr int STAGE_BYTES = KATOMS * ATOM_BYTES;       // [A atoms][W atoms]
  static constexpr int SMEM_BUDGET = 200 * 1024;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // double-buffered fp32 accumulator
  static constexpr int BAR_BYTES = 256;
  static constexpr int EPI_WARPS = 4;
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + BAR_BYTES + EPI_WARPS * 32 * 128 + 1024;  // +1024 align slack
  static constexpr int THREADS = 192;
  static_assert(STAGES >= 2, "need at least a double buffer");
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512,
                "TMEM columns must be a power of two");
  static_assert(BLOCK_M_ == 128 || BLOCK_M_ == 64, "UMMA M of one CTA is 64 or 128");
  static_assert(2 * 8 * STAGES + 4 * 8 + 4 <= BAR_BYTES, "barrier region too small");
};

// x * sigmoid(1.702 x) = x / (1 + 2^(-1.702 log2(e) x))   (reference: clip/model.py:183-185)
// one multiply, ex2.approx, one add, rcp.approx, one multiply.
__device__ __forceinline__ float quick_gelu(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -2.4554669595930157f));
  return __fdividef(x, 1.0f + e);
}

// Epilogue of one 32-row x 32-column chunk per warp.  The accumulator comes out of TMEM with one
// row per lane; scale, bias and activation are applied in that layout, then the chunk is transposed
// through a 4 KB per-warp staging buffer (16-byte pieces XOR-swizzled by the row) so that every
// global access is row-contiguous: a load/store instruction touches 4 rows x 128 B instead of 32
// rows x 16 B.  In the transposed layout lane = (row & 3 within a group of 4 rows, 16-byte piece
// 0..7).  The bias and residual loads are issued before the wait on the TMEM load so that their
// latency overlaps it and the arithmetic; every warp-uniform option is tested once per chunk, not
// once per row.
constexpr int EPI_STAGE_BYTES = 32 * 128;

// SPLIT: 0 = no split output, 1 = bf16 hi/lo planes, 2 = f16f8 activation planes, 3 = fp16 plane
template <bool F32, int SPLIT>
__device__ __forceinline__ void epilogue_store(const GemmParams& p, const int (&orow)[8],
                                               const float4 (&val)[8], int col, int lane,
                                               const uint8_t* stage) {
  const int piece = lane & 7;
  float amax = 0.f;  // max |value| stored in an fp16-based encoding (saturation guard)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    if (orow[i] >= 0) {
      const float4 a = *reinterpret_cast<const float4*>(stage + row * 128 + ((piece ^ (row & 7)) << 4));
      const float4 v4 = make_float4(a.x + val[i].x, a.y + val[i].y, a.z + val[i].z, a.w + val[i].w);
      if (F32) {
        *reinterpret_cast<float4*>(p.out_f32 + static_cast<long long>(orow[i]) * p.ldc + col) = v4;
        if (p.peer_world > 0) {
#pragma unroll 1
          for (int r = 0; r < p.peer_world; ++r)
            *reinterpret_cast<float4*>(p.peer_out[r] + static_cast<long long>(orow[i]) * p.ldc + col) = v4;
        }
      }
      if (SPLIT == 1) {
        uint32_t h0, l0, h1, l1;
        split_pack2(v4.x, v4.y, h0, l0);
        split_pack2(v4.z, v4.w, h1, l1);
        __nv_bfloat16* dst = p.out_split + static_cast<long long>(orow[i]) * p.ld_split + col;
        *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(dst + p.split_plane_stride) = make_uint2(l0, l1);
      } else if (SPLIT == 2) {
        f16f8_store4_act(p.out_split, p.split_plane_stride,
                         static_cast<long long>(orow[i]) * p.ld_split + col, v4.x, v4.y, v4.z, v4.w);
        amax = sat_track(amax, v4.x, v4.y, v4.z, v4.w);
      } else if (SPLIT == 3) {
        f16_store4_act(p.out_split, static_cast<long long>(orow[i]) * p.ld_split + col, v4.x, v4.y,
                       v4.z, v4.w);
        amax = sat_track(amax, v4.x, v4.y, v4.z, v4.w);
      }
    }
  }
  if (SPLIT >= 2) sat_report(p.sat, amax);
}

// m_end: first row this warp does NOT own (p.M, or m_base + 16 with 64-row tiles, whose accumulator
// keeps 16 rows in the first 16 lanes of every TMEM lane quarter).
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, uint32_t taddr, int n,
                                               int m_base, int lane, uint8_t* stage, int m_end) {
  uint32_t raw[32];
  ptx::tmem_ld_32x32(taddr, raw);
  // ---- everything that does not depend on the accumulator, while the TMEM load is in flight
  float4 b4[8];
  const bool has_bias = p.bias != nullptr;
  if (has_bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) b4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int col = n + (lane & 7) * 4;
  int orow[8];      // output row of the 8 rows this lane stores (-1: beyond M)
  if (p.row_group_stride != 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m_base + i * 4 + (lane >> 3);
      orow[i] = m < m_end ? (m / p.row_group) * p.row_group_stride + (m % p.row_group) + p.row_offset : -1;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m_base + i * 4 + (lane >> 3);
      orow[i] = m < m_end ? m + p.row_offset : -1;
    }
  }
  float4 val[8];    // residual (out_f32 may alias it: each lane reads exactly what it later writes)
#pragma unroll
  for (int i = 0; i < 8; ++i) val[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.residual != nullptr && m_base < p.M) {
    if (p.res_mod > 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (orow[i] >= 0) {
          const int m = m_base + i * 4 + (lane >> 3);
          val[i] = *reinterpret_cast<const float4*>(p.residual + static_cast<long long>(m % p.res_mod) * p.ldr + col);
        }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (orow[i] >= 0)
          val[i] = *reinterpret_cast<const float4*>(p.residual + static_cast<long long>(orow[i]) * p.ldr + col);
    }
  }
  ptx::tmem_ld_wait();
  if (m_base >= p.M) return;  // warp-uniform: the whole block is padding
  // ---- accumulator * scale + bias, activation (one row per lane)
  float v[32];
  const float sc = p.out_scale;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[4 * j + 0] = fmaf(__uint_as_float(raw[4 * j + 0]), sc, b4[j].x);
    v[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]), sc, b4[j].y);
    v[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]), sc, b4[j].z);
    v[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]), sc, b4[j].w);
  }
  if (p.act == ACT_QUICKGELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
  } else if (p.act == ACT_LEAKYRELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : 0.01f * v[j];
  }
  __syncwarp();  // the previous chunk's readers are done with the staging buffer
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  // ---- add the residual to the staged values and store row-contiguously
  const int kind = p.out_split == nullptr ? 0 : 1 + p.out_enc;
  if (p.out_f32 != nullptr) {
    if (kind == 0) epilogue_store<true, 0>(p, orow, val, col, lane, stage);
    else if (kind == 1) epilogue_store<true, 1>(p, orow, val, col, lane, stage);
    else if (kind == 2) epilogue_store<true, 2>(p, orow, val, col, lane, stage);
    else epilogue_store<true, 3>(p, orow, val, col, lane, stage);
  } else {
    if (kind == 1) epilogue_store<false, 1>(p, orow, val, col, lane, stage);
    else if (kind == 2) epilogue_store<false, 2>(p, orow, val, col, lane, stage);
    else if (kind == 3) epilogue_store<false, 3>(p, orow, val, col, lane, stage);
  }
}

// ------------------------------------------------------------------------------------------
// Encoded-output epilogue (the GEMMs whose only output feeds another GEMM or the attention:
// in_proj, c_fc, conv1): no fp32 output, no residual, no row remap.  Differences to the generic
// path above, all of them fewer instructions per element (these K = 768 GEMMs are epilogue-bound:
// 6 k / 12 k tensor cycles per 256 x 256 tile against ~30 thread instructions per element before):
//   * the values are ENCODED in the accumulator layout (one row per lane) and staged already
//     narrow: 64 B (fp16) resp. 64 + 32 + 32 B (f16f8 planes H | L | C) per row instead of 128 B of
//     fp32, so the transposed read-back and the global stores move 16 bytes per lane and
//     instruction (4 resp. 8 of each per chunk instead of 8 + 8..24 narrower stores);
//   * the TMEM load of chunk c + 1 is issued before chunk c is processed;
//   * no residual registers, no per-row remap arithmetic.
// Staging layouts (per warp, 4 KB): H rows of 64 B with the 16-byte piece XOR-ed by (row >> 1) & 3,
// L and C rows of 32 B with the piece XOR-ed by (row >> 2) & 1: conflict-free for the row-per-lane
// writes and for the 8-rows-per-instruction reads.
template <int ENC>   // 1 = f16f8 planes, 2 = fp16 plane
__device__ __forceinline__ void epilogue_encode_store(const GemmParams& p, const uint32_t (&raw)[32],
                                                      int n, int m_base, int lane, uint8_t* stage,
                                                      float& amax) {
  // ---- scale + bias + activation, one row per lane
  float v[32];
  const float sc = p.out_scale;
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
      v[4 * j + 0] = fmaf(__uint_as_float(raw[4 * j + 0]), sc, b.x);
      v[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]), sc, b.y);
      v[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]), sc, b.z);
      v[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]), sc, b.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * sc;
  }
  if (p.act == ACT_QUICKGELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
  } else if (p.act == ACT_LEAKYRELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : 0.01f * v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) amax = sat_track(amax, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  // ---- encode and stage (row = lane)
  __syncwarp();  // the previous chunk's readers are done with the staging buffer
  uint8_t* srow_h = stage + lane * 64;
  const int swz_h = (lane >> 1) & 3;
  if (ENC == 2) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // 8 values = 16 bytes of fp16 per piece
      uint2 a, b;
      f16_pack4(v[8 * q + 0], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], a);
      f16_pack4(v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7], b);
      *reinterpret_cast<uint4*>(srow_h + ((q ^ swz_h) << 4)) = make_uint4(a.x, a.y, b.x, b.y);
    }
  } else {
    uint8_t* srow_l = stage + 2048 + lane * 32;
    uint8_t* srow_c = stage + 3072 + lane * 32;
    const int swz_8 = (lane >> 2) & 1;
    uint32_t lw[8], cw[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint2 a, b;
      f16f8_pack4(v[8 * q + 0], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], kActScaleMain, kActScaleRes,
                  kActScaleCoarse, a, lw[2 * q], cw[2 * q]);
      f16f8_pack4(v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7], kActScaleMain, kActScaleRes,
                  kActScaleCoarse, b, lw[2 * q + 1], cw[2 * q + 1]);
      *reinterpret_cast<uint4*>(srow_h + ((q ^ swz_h) << 4)) = make_uint4(a.x, a.y, b.x, b.y);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {   // 16 values = 16 bytes of e4m3 per piece
      *reinterpret_cast<uint4*>(srow_l + ((q ^ swz_8) << 4)) =
          make_uint4(lw[4 * q], lw[4 * q + 1], lw[4 * q + 2], lw[4 * q + 3]);
      *reinterpret_cast<uint4*>(srow_c + ((q ^ swz_8) << 4)) =
          make_uint4(cw[4 * q], cw[4 * q + 1], cw[4 * q + 2], cw[4 * q + 3]);
    }
  }
  __syncwarp();
  // ---- row-contiguous stores: H 8 rows x 64 B per instruction, L / C 16 rows x 32 B
  uint8_t* out = reinterpret_cast<uint8_t*>(p.out_split);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + (lane >> 2), piece = lane & 3;
    const int m = m_base + row;
    if (m < p.M) {
      const uint4 h = *reinterpret_cast<const uint4*>(stage + row * 64 + ((piece ^ ((row >> 1) & 3)) << 4));
      const long long off = static_cast<long long>(m + p.row_offset) * p.ld_split + n + piece * 8;
      *reinterpret_cast<uint4*>(out + 2 * off) = h;
    }
  }
  if (ENC == 1) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = i * 16 + (lane >> 1), piece = lane & 1;
      const int m = m_base + row;
      if (m < p.M) {
        const int sw = (piece ^ ((row >> 2) & 1)) << 4;
        const uint4 l = *reinterpret_cast<const uint4*>(stage + 2048 + row * 32 + sw);
        const uint4 c = *reinterpret_cast<const uint4*>(stage + 3072 + row * 32 + sw);
        const long long off = static_cast<long long>(m + p.row_offset) * p.ld_split + n + piece * 16;
        *reinterpret_cast<uint4*>(out + 2 * p.split_plane_stride + off) = l;
        *reinterpret_cast<uint4*>(out + 3 * p.split_plane_stride + off) = c;
      }
    }
  }
}

// The warp's column chunks of one tile through the encoded-output epilogue, with the TMEM load of
// the next chunk in flight while the current one is processed.  chunks: 32-column chunks of this warp.
template <int ENC>
__device__ __forceinline__ void epilogue_tile_encoded(const GemmParams& p, uint32_t t_row, int n0,
                                                      int chunks, int m_base, int lane, uint8_t* stage) {
  if (m_base >= p.M) return;  // warp-uniform: the whole 32-row block is padding
  uint32_t raw[2][32];
  float amax = 0.f;
  ptx::tmem_ld_32x32(t_row, raw[0]);
#pragma unroll 1
  for (int c = 0; c < chunks; c += 2) {
    ptx::tmem_ld_wait();
    const bool more1 = c + 1 < chunks && n0 + (c + 1) * 32 < p.N;
    if (more1) ptx::tmem_ld_32x32(t_row + (c + 1) * 32, raw[1]);
    epilogue_encode_store<ENC>(p, raw[0], n0 + c * 32, m_base, lane, stage, amax);
    if (!more1) break;
    ptx::tmem_ld_wait();
    const bool more2 = c + 2 < chunks && n0 + (c + 2) * 32 < p.N;
    if (more2) ptx::tmem_ld_32x32(t_row + (c + 2) * 32, raw[0]);
    epilogue_encode_store<ENC>(p, raw[1], n0 + (c + 1) * 32, m_base, lane, stage, amax);
    if (!more2) break;
  }
  sat_report(p.sat, amax);
}

// 1 / 2 when the launch qualifies for the encoded-output epilogue with f16f8 / fp16 planes, 3 for
// the residual epilogue (fp32 output + fp32 residual, identity row map), else 0 (generic)
inline int encoded_epilogue_kind(const GemmParams& p) {
  if (p.row_group_stride != 0 || p.peer_world != 0) return 0;
  if (p.out_f32 == nullptr && p.out_split != nullptr && p.out_enc != 0 && p.residual == nullptr)
    return p.out_enc;
  if (p.out_f32 != nullptr && p.out_split == nullptr && p.residual != nullptr && p.res_mod == 0) return 3;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Residual epilogue (EPI == 3): out_f32 = act(acc * scale + bias) + residual, fp32 only, identity row
// map -- out_proj, c_proj, the axial to_out and conv2.  These launches read and write a full fp32
// residual tile per output tile and are bound by the bytes they keep in flight: the generic path
// issues a chunk's residual loads only when it reaches that chunk (32 KB per CTA outstanding).  Here
// the residual rows of chunk c + 1 are requested before chunk c is processed, which doubles the
// bytes in flight.  Same transposed, row-contiguous access pattern as the generic path.
__device__ __forceinline__ void epilogue_tile_residual(const GemmParams& p, uint32_t t_row, int n0,
                                                       int chunks, int m_base, int lane, uint8_t* stage) {
  if (m_base >= p.M) return;  // warp-uniform: the whole 32-row block is padding
  const int piece = lane & 7, rsub = lane >> 3;
  const int colp = piece * 4;
  long long roff[8];   // element offset of this lane's 8 rows (row i*4 + rsub), -1 = beyond M
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m_base + i * 4 + rsub;
    roff[i] = m < p.M ? static_cast<long long>(m + p.row_offset) : -1;
  }
  auto load_res = [&](float4 (&dst)[8], int n) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      dst[i] = roff[i] >= 0 ? *reinterpret_cast<const float4*>(p.residual + roff[i] * p.ldr + n + colp)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 cur[8], nxt[8];
  uint32_t raw[32];
  ptx::tmem_ld_32x32(t_row, raw);
  load_res(cur, n0);
  const float sc = p.out_scale;
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    const int n = n0 + c * 32;
    const bool more = c + 1 < chunks && n + 32 < p.N;
    if (more) load_res(nxt, n + 32);
    float4 b4[8];
    if (p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) b4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ptx::tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j + 0] = fmaf(__uint_as_float(raw[4 * j + 0]), sc, b4[j].x);
      v[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]), sc, b4[j].y);
      v[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]), sc, b4[j].z);
      v[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]), sc, b4[j].w);
    }
    if (more) ptx::tmem_ld_32x32(t_row + (c + 1) * 32, raw);   // raw is dead: next chunk's accumulator
    if (p.act == ACT_QUICKGELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
    } else if (p.act == ACT_LEAKYRELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : 0.01f * v[j];
    }
    __syncwarp();  // the previous chunk's readers are done with the staging buffer
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
          make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = i * 4 + rsub;
      if (roff[i] >= 0) {
        const float4 a = *reinterpret_cast<const float4*>(stage + row * 128 + ((piece ^ (row & 7)) << 4));
        *reinterpret_cast<float4*>(p.out_f32 + roff[i] * p.ldc + n + colp) =
            make_float4(a.x + cur[i].x, a.y + cur[i].y, a.z + cur[i].z, a.w + cur[i].w);
      }
    }
    if (!more) break;
#pragma unroll
    for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
  }
}

// End of a GEMM whose epilogue stored into peer memory: called by every thread after its last
// store; the block-level barrier that follows in the kernels orders the fences before the count.
__device__ __forceinline__ void peer_fence(const GemmParams& p) {
  if (p.peer_world > 0) __threadfence_system();
}
// One thread per CTA, after the barrier: the last CTA to arrive raises this rank's flag everywhere.
__device__ __forceinline__ void peer_publish(const GemmParams& p) {
  if (p.peer_world > 0 && p.peer_signal) {
    const unsigned int done = atomicAdd(p.peer_counter, 1u);
    if (done == gridDim.x - 1) {
      *p.peer_counter = 0u;  // ready for the next launch
      __threadfence_system();
      for (int r = 0; r < p.peer_world; ++r)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.peer_flags[r] + p.peer_rank),
                     "r"(p.peer_epoch)
                     : "memory");
    }
  }
}

// EPI: epilogue compiled into the instantiation -- 0 generic (epilogue_chunk), 1 / 2 encoded-output
// only with f16f8 / fp16 planes (epilogue_tile_encoded), 3 fp32 + residual (epilogue_tile_residual).  One path per instantiation keeps the code
// and the register allocation of each small (all paths in one kernel cost the generic one 23 %).
// BLOCK_M = 64 (with BLOCK_N = 32, KATOMS = 4): the tile of the SMALL problems (one or two
// sub-videos of the temporal stage): four times the CTAs of the 128 x 64 tiling, and four K atoms
// per barrier round (see GemmCfg).  Per element the accumulation order along K is the same, so
// results are bit-identical to the 128-row tiles.
template <int BLOCK_N, int PASSES, int EPI, int BLOCK_M = 128,
          int KATOMS = default_katoms(BLOCK_N, PASSES, BLOCK_M)>
__global__ void __launch_bounds__(192, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                    const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, PASSES, BLOCK_M, KATOMS>;
  static_assert(BLOCK_M == 128 || EPI == 0, "64-row tiles use the generic epilogue");
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET into the __shared__ array: the pointer keeps its address
  // space, so plain C++ accesses compile to LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 4);  // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();   // operands, residual and outputs belong to the predecessors until here

  const int m_tiles = (p.M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int total_tiles = m_tiles * n_tiles;
  const int rounds = (p.num_kb + KATOMS - 1) / KATOMS;   // barrier rounds per tile

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int n0 = (t % n_tiles) * BLOCK_N;
        const int mt = t / n_tiles;
        const int m0 = mt * Cfg::BLOCK_M;
        int img = 0, h0 = 0;
        if (p.a_mode == 1) {
          const int tiles_per_img = (p.conv_h * p.conv_w) / Cfg::BLOCK_M;
          img = mt / tiles_per_img;
          h0 = (mt % tiles_per_img) * (Cfg::BLOCK_M / p.conv_w);
        }
        // conv3x3: (tap, channel block) advance incrementally -- the divisions by the run-time
        // channel-block count cost this single thread more per K atom than the TMA issue itself
        int cb = 0, dy = -1, dx = -1;
        int kb = 0;
        for (int r = 0; r < rounds; ++r) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + KATOMS * Cfg::A_ATOM_BYTES;
          const int atoms = min(KATOMS, p.num_kb - kb);   // the last round may be short
          ptx::mbar_expect_tx(&full_bar[stage], atoms * Cfg::ATOM_BYTES);
#pragma unroll
          for (int j = 0; j < KATOMS; ++j) {
            if (j < atoms) {
              if (p.a_mode == 0) {
                ptx::tma_load_3d(sa + j * Cfg::A_ATOM_BYTES, &tmA, &full_bar[stage], kb * Cfg::BLOCK_K, m0, 0);
              } else {
                ptx::tma_load_5d(sa + j * Cfg::A_ATOM_BYTES, &tmA, &full_bar[stage], cb * Cfg::BLOCK_K, dx,
                                 h0 + dy, img, 0);
                if (++cb == p.conv_cin_kb) {
                  cb = 0;
                  if (++dx > 1) { dx = -1; ++dy; }
                }
              }
              ptx::tma_load_3d(sb + j * Cfg::B_ATOM_BYTES, &tmB, &full_bar[stage], kb * Cfg::BLOCK_K, n0, 0);
              ++kb;
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The warp stays converged; one elected lane issues (descriptors in uniform registers, offsets
    // are adds on the descriptor's low word, which counts 16-byte units).
    {
      const bool leader = ptx::elect_one();
      const bool issue = leader && (p.debug & 1) == 0;   // debug: profiling experiments only
      constexpr uint32_t idesc = PASSES == 4 ? ptx::make_idesc_fmt0_f32(Cfg::BLOCK_M, BLOCK_N)   // fp16
                                             : ptx::make_idesc_bf16_f32(Cfg::BLOCK_M, BLOCK_N);
      const uint64_t desc0 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem));
      uint32_t stage = 0, phase = 0;
      uint32_t acc = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        int kb = 0;
        for (int r = 0; r < rounds; ++r) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_st = desc0 + ((stage * Cfg::STAGE_BYTES) >> 4);
          const uint64_t b_st = a_st + ((KATOMS * Cfg::A_ATOM_BYTES) >> 4);
#pragma unroll
          for (int j = 0; j < KATOMS; ++j) {
            const uint64_t a_hi0 = a_st + j * (Cfg::A_ATOM_BYTES >> 4);
            const uint64_t b_hi0 = b_st + j * (Cfg::B_ATOM_BYTES >> 4);
            const bool atom = issue && kb + j < p.num_kb;   // the last round may be short
#pragma unroll
            for (int k = 0; k < Cfg::BLOCK_K / Cfg::UMMA_K; ++k) {
              const uint64_t a_hi = a_hi0 + 2 * k, b_hi = b_hi0 + 2 * k;   // 32 bytes along K per step
              const bool go = atom && (k == 0 || (p.debug & 2) == 0);
              ptx::mma_f16_ss_if(go, d_tmem, a_hi, b_hi, idesc, (kb | j | k) != 0 ? 1u : 0u);
              if (PASSES == 3) {
                ptx::mma_f16_ss_if(go, d_tmem, a_hi + (Cfg::A_PLANE_BYTES >> 4), b_hi, idesc, 1u);
                ptx::mma_f16_ss_if(go, d_tmem, a_hi, b_hi + (Cfg::B_PLANE_BYTES >> 4), idesc, 1u);
              }
            }
          }
          kb += KATOMS;
          ptx::mma_commit_if(leader, &empty_bar[stage]);  // smem slot is free once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit_if(leader, &tfull_bar[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    uint8_t* stage = smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + (warp - 2) * EPI_STAGE_BYTES;
    uint32_t acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int n0 = (t % n_tiles) * BLOCK_N;
      const int m_base = (t / n_tiles) * Cfg::BLOCK_M + quarter * (Cfg::BLOCK_M / 4);
      const int m_end = BLOCK_M == 128 ? p.M : min(p.M, m_base + 16);
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quarter * 32) << 16);
      if (EPI == 3) {
        epilogue_tile_residual(p, t_row, n0, BLOCK_N / 32, m_base, lane, stage);
      } else if (EPI != 0) {
        epilogue_tile_encoded<EPI>(p, t_row, n0, BLOCK_N / 32, m_base, lane, stage);
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          const int n = n0 + c * 32;
          if (n >= p.N) break;  // warp-uniform
          epilogue_chunk(p, t_row + c * 32, n, m_base, lane, stage, m_end);
        }
      }
      // hand the accumulator buffer back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  peer_fence(p);
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) peer_publish(p);
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs computes a 256 x 256 tile.
// Each CTA stages 128 rows of A (its half of M) and 128 rows of W (its half of N) per 64-wide
// K block, so a stage is 64 KB instead of 96 KB (3 stages instead of 2) and the L2 -> SMEM traffic
// per MMA drops by a third; the leader CTA's single MMA thread issues M=256 x N=256 x K=16
// instructions that read both CTAs' shared memory and write both CTAs' TMEM.
// Roles per CTA (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader only),
// warps 2..9 = epilogue (two warps per TMEM lane quarter, 128 columns each).
template <int PASSES_>
struct Gemm2Cfg {
  static constexpr int CTA_M = 128;     // rows of A per CTA
  static constexpr int CTA_N = 128;     // rows of W per CTA
  static constexpr int BLOCK_M = 256;   // per cluster
  static constexpr int BLOCK_N = 256;
  static constexpr int BLOCK_K = 64;
  static constexpr int UMMA_K = 16;
  static constexpr int PASSES = PASSES_;
  static constexpr int PLANES = (PASSES_ == 1 || PASSES_ == 4) ? 1 : 2;
  static constexpr int A_PLANE_BYTES = CTA_M * 128;
  static constexpr int B_PLANE_BYTES = CTA_N * 128;
  // bytes of one stage's A and W regions: main plane + second plane (bf16 lo, or the e4m3 L | C
  // pair) -- PASSES == 6 carries ONE e4m3 plane per operand (L of A, C of W), half a plane's bytes
  static constexpr int A_REGION = PASSES_ == 6 ? A_PLANE_BYTES + A_PLANE_BYTES / 2 : PLANES * A_PLANE_BYTES;
  static constexpr int B_REGION = PASSES_ == 6 ? B_PLANE_BYTES + B_PLANE_BYTES / 2 : PLANES * B_PLANE_BYTES;
  // K atoms per barrier round.  Even in the one-pass modes the four 131-clk MMAs of an atom outlast
  // a round (~450 clk, see GemmCfg): two atoms per round (3 stages instead of 6) measured 1.5-3 %
  // SLOWER on in_proj / c_fc / the temporal convs (profiles/r2_small_gemm_round_trip_experiments.txt),
  // so the pair kernel keeps one (-DACLIP_KATOMS_PAIR=2 rebuilds the variant).
  static constexpr int KATOMS = (PASSES_ == 1 || PASSES_ == 4) ? ACLIP_KATOMS_PAIR : 1;
  static constexpr int ATOM_BYTES = A_REGION + B_REGION;
  static constexpr int STAGE_BYTES = KATOMS * ATOM_BYTES;       // [A atoms][W atoms]
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int TMEM_COLS = 512;
  static constexpr int EPI_WARPS = 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + EPI_WARPS * 32 * 128 + 1024;
  static constexpr int THREADS = 64 + EPI_WARPS * 32;
};

template <int PASSES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmA8,
                     const __grid_constant__ CUtensorMap tmB8, const GemmParams p) {
  using Cfg = Gemm2Cfg<PASSES>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET into the __shared__ array: the pointer keeps its address
  // space, so plain C++ accesses compile to LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  ptx::pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (PASSES == 2 || PASSES == 6) {
      ptx::prefetch_tmap(&tmA8);
      ptx::prefetch_tmap(&tmB8);
    }
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);   // leader's producer (arrive.expect_tx for both CTAs)
      ptx::mbar_init(&empty_bar[s], 1);  // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);                        // multicast tcgen05.commit
      ptx::mbar_init(&tempty_bar[a], 2 * Cfg::EPI_WARPS);      // epilogue warps of both CTAs
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();  // barriers and TMEM of BOTH CTAs are ready before anyone touches them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();   // operands, residual and outputs belong to the predecessors until here

  const int m_tiles = (p.M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int n_tiles = (p.N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N;
  const int total_tiles = m_tiles * n_tiles;
  constexpr int KA = Cfg::KATOMS;
  const int rounds = (p.num_kb + KA - 1) / KA;   // barrier rounds per tile

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        const int n0 = (t % n_tiles) * Cfg::BLOCK_N + static_cast<int>(rank) * Cfg::CTA_N;
        const int mt = (t / n_tiles) * 2 + static_cast<int>(rank);  // 128-row tile index
        const int m0 = mt * Cfg::CTA_M;
        int img = 0, h0 = 0;
        if (p.a_mode == 1) {
          const int tiles_per_img = (p.conv_h * p.conv_w) / Cfg::CTA_M;
          img = mt / tiles_per_img;
          h0 = (mt % tiles_per_img) * (Cfg::CTA_M / p.conv_w);
        }
        // conv3x3: (tap, channel block) advance incrementally (no divisions in this single thread)
        int cb = 0, dy = -1, dx = -1;
        int kb = 0;
        for (int r = 0; r < rounds; ++r) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa0 = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb0 = sa0 + KA * Cfg::A_REGION;
          const int atoms = min(KA, p.num_kb - kb);   // the last round may be short
          if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * atoms * Cfg::ATOM_BYTES);
#pragma unroll
          for (int j = 0; j < KA; ++j) {
            if (j >= atoms) break;
            uint8_t* sa = sa0 + j * Cfg::A_REGION;
            uint8_t* sb = sb0 + j * Cfg::B_REGION;
            const int k0 = kb * Cfg::BLOCK_K;
            if (PASSES == 6) {
              // linear A only: fp16 planes + plane L of A (index 0) and plane C of W (index 1)
              ptx::tma_load_3d_pair(sa, &tmA, &full_bar[stage], k0, m0, 0);
              ptx::tma_load_3d_pair(sa + Cfg::A_PLANE_BYTES, &tmA8, &full_bar[stage], k0, m0, 0);
              ptx::tma_load_3d_pair(sb, &tmB, &full_bar[stage], k0, n0, 0);
              ptx::tma_load_3d_pair(sb + Cfg::B_PLANE_BYTES, &tmB8, &full_bar[stage], k0, n0, 1);
            } else {
              // PASSES == 2, f16f8 operands: fp16 plane (128 B rows, SWIZZLE_128B) + the two e4m3
              // planes in one box (64 B rows, SWIZZLE_64B); same bytes per atom as two bf16 planes
              if (p.a_mode == 0) {
                ptx::tma_load_3d_pair(sa, &tmA, &full_bar[stage], k0, m0, 0);
                if (PASSES == 2)
                  ptx::tma_load_3d_pair(sa + Cfg::A_PLANE_BYTES, &tmA8, &full_bar[stage], k0, m0, 0);
              } else {  // conv3x3: the same shifted, zero-filled boxes for every plane
                ptx::tma_load_5d_pair(sa, &tmA, &full_bar[stage], cb * Cfg::BLOCK_K, dx, h0 + dy, img, 0);
                if (PASSES == 2)
                  ptx::tma_load_5d_pair(sa + Cfg::A_PLANE_BYTES, &tmA8, &full_bar[stage], cb * Cfg::BLOCK_K, dx,
                                        h0 + dy, img, 0);
                if (++cb == p.conv_cin_kb) {
                  cb = 0;
                  if (++dx > 1) { dx = -1; ++dy; }
                }
              }
              ptx::tma_load_3d_pair(sb, &tmB, &full_bar[stage], k0, n0, 0);
              if (PASSES == 2)
                ptx::tma_load_3d_pair(sb + Cfg::B_PLANE_BYTES, &tmB8, &full_bar[stage], k0, n0, 0);
            }
            ++kb;
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    // The warp stays converged; one elected lane issues (descriptors in uniform registers).
    if (rank == 0) {
      const bool leader = ptx::elect_one();
      // kind::f16 with fp16 operands and kind::f8f6f4 with e4m3 operands both encode format 0
      constexpr uint32_t idesc = (PASSES == 2 || PASSES == 4 || PASSES == 6)
                                     ? ptx::make_idesc_fmt0_f32(Cfg::BLOCK_M, Cfg::BLOCK_N)
                                     : ptx::make_idesc_bf16_f32(Cfg::BLOCK_M, Cfg::BLOCK_N);
      const uint64_t desc0 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem));
      const uint64_t desc8 = ptx::make_kmajor_sw64_desc(ptx::smem_u32(smem));
      uint32_t stage = 0, phase = 0;
      uint32_t acc = 0, acc_phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::BLOCK_N;
        int kb = 0;
        for (int r = 0; r < rounds; ++r) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
#pragma unroll
          for (int j = 0; j < KA; ++j) {
            // descriptor low word counts 16-byte units: stage / atom / plane / K-step offsets are adds
            const uint32_t a_off = stage * Cfg::STAGE_BYTES + j * Cfg::A_REGION;
            const uint32_t b_off = stage * Cfg::STAGE_BYTES + KA * Cfg::A_REGION + j * Cfg::B_REGION;
            const uint64_t a_hi0 = desc0 + (a_off >> 4);
            const uint64_t b_hi0 = desc0 + (b_off >> 4);
            const bool go = leader && (KA == 1 || kb + j < p.num_kb);   // the last round may be short
            const uint32_t first = (kb | j) != 0 ? 1u : 0u;
            if (p.debug & 1) {
              // feed-rate experiment: consume the stage without issuing MMAs
            } else if (PASSES == 6) {
              // x_H w_H: four K=16 fp16 MMAs; x_L w_C: two K=32 e4m3 MMAs
#pragma unroll
              for (int k = 0; k < Cfg::BLOCK_K / Cfg::UMMA_K; ++k)
                ptx::mma_bf16_ss_pair_if(go, d_tmem, a_hi0 + 2 * k, b_hi0 + 2 * k, idesc, k != 0 ? 1u : first);
              const uint64_t a_l0 = desc8 + ((a_off + Cfg::A_PLANE_BYTES) >> 4);
              const uint64_t b_c0 = desc8 + ((b_off + Cfg::B_PLANE_BYTES) >> 4);
#pragma unroll
              for (int k = 0; k < Cfg::BLOCK_K / 32; ++k)
                ptx::mma_f8_ss_pair_if(go, d_tmem, a_l0 + 2 * k, b_c0 + 2 * k, idesc, 1u);
            } else if (PASSES == 2) {
              // x_H w_H: four K=16 fp16 MMAs; x_L w_C and x_C w_L: two K=32 e4m3 MMAs each (the
              // e4m3 tile sits behind the fp16 tile)
              if (!(p.debug & 4)) {
#pragma unroll
                for (int k = 0; k < Cfg::BLOCK_K / Cfg::UMMA_K; ++k)
                  ptx::mma_bf16_ss_pair_if(go, d_tmem, a_hi0 + 2 * k, b_hi0 + 2 * k, idesc, k != 0 ? 1u : first);
              }
              const uint64_t a_l0 = desc8 + ((a_off + Cfg::A_PLANE_BYTES) >> 4);
              const uint64_t b_l0 = desc8 + ((b_off + Cfg::B_PLANE_BYTES) >> 4);
              constexpr uint64_t kCoarse = (Cfg::A_PLANE_BYTES / 2) >> 4;
              if (!(p.debug & 2))
#pragma unroll
              for (int k = 0; k < Cfg::BLOCK_K / 32; ++k) {
                ptx::mma_f8_ss_pair_if(go, d_tmem, a_l0 + 2 * k, b_l0 + kCoarse + 2 * k, idesc, 1u);
                ptx::mma_f8_ss_pair_if(go, d_tmem, a_l0 + kCoarse + 2 * k, b_l0 + 2 * k, idesc, 1u);
              }
            } else {
#pragma unroll
              for (int k = 0; k < Cfg::BLOCK_K / Cfg::UMMA_K; ++k) {
                const uint64_t a_hi = a_hi0 + 2 * k, b_hi = b_hi0 + 2 * k;
                ptx::mma_bf16_ss_pair_if(go, d_tmem, a_hi, b_hi, idesc, k != 0 ? 1u : first);
                if (PASSES == 3) {
                  ptx::mma_bf16_ss_pair_if(go, d_tmem, a_hi + (Cfg::A_PLANE_BYTES >> 4), b_hi, idesc, 1u);
                  ptx::mma_bf16_ss_pair_if(go, d_tmem, a_hi, b_hi + (Cfg::B_PLANE_BYTES >> 4), idesc, 1u);
                }
              }
            }
          }
          kb += KA;
          ptx::mma_commit_pair_if(leader, &empty_bar[stage], 0x3);  // frees the slot in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit_pair_if(leader, &tfull_bar[acc], 0x3);  // accumulator complete -> both epilogues
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int col_half = (warp - 2) >> 2;    // which 128 of the 256 columns
    uint8_t* stage = smem + STAGES * Cfg::STAGE_BYTES + 256 + (warp - 2) * EPI_STAGE_BYTES;
    uint32_t acc = 0, acc_phase = 0;
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      const int n0 = (t % n_tiles) * Cfg::BLOCK_N + col_half * 128;
      const int m_base = (t / n_tiles) * Cfg::BLOCK_M + static_cast<int>(rank) * Cfg::CTA_M + quarter * 32;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + acc * Cfg::BLOCK_N + col_half * 128 +
                             (static_cast<uint32_t>(quarter * 32) << 16);
      if (EPI == 3) {
        epilogue_tile_residual(p, t_row, n0, 4, m_base, lane, stage);
      } else if (EPI != 0) {
        epilogue_tile_encoded<EPI>(p, t_row, n0, 4, m_base, lane, stage);
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int n = n0 + c * 32;
          if (n >= p.N) break;  // warp-uniform
          epilogue_chunk(p, t_row + c * 32, n, m_base, lane, stage, p.M);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_remote(&tempty_bar[acc], 0);  // leader's barrier
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  peer_fence(p);
  ptx::tc_fence_before();
  ptx::cluster_sync();  // nobody leaves while the pair may still touch its smem / TMEM / barriers
  if (threadIdx.x == 0) peer_publish(p);
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace aclip
