// CTA-pair GEMM on f16mx operands (mx.cuh):  D = x_H w_H + x_L4 w_C4 + x_C4 w_L4
//   four kind::f16 MMAs (K = 16) on the fp16 planes + two block-scaled kind::mxf4 MMAs (K = 64) per
//   64-wide K atom = 1.5 fp16-pass equivalents (f16f8: 2).
// Same roles and pipelines as gemm2_tcgen05_kernel (gemm.cuh); what differs:
//   * the tile is 256 x 192: the scale factors of the block-scaled MMAs live in TMEM, and two
//     256-column accumulators would fill all 512 columns.  2 x 192 accumulator columns + 4 columns of
//     A scales + 8 columns of W scales;
//   * a stage carries per CTA: fp16 A 128 x 64 (SWIZZLE_128B), fp16 W 96 x 64, the e2m1 planes L4 | C4
//     of both (rows of 32 B, SWIZZLE_32B), the 512-byte scale chunk of the CTA's 128 A rows and the
//     two scale chunks (256 rows, 128-row aligned) that cover the tile's 192 W rows -- 43.5 KB, 4 stages;
//   * per round the issuer copies the scale chunks shared memory -> TMEM (tcgen05.cp, both CTAs)
//     ahead of the MMAs; tcgen05 executes them in issue order, so one TMEM copy of the scales
//     suffices, and the stage is released by the same commit as before;
//   * a W tile that starts in the middle of a 128-row scale chunk (odd tiles: 192 j mod 128 = 64)
//     reads its scales two TMEM columns further.
#pragma once
#include "gemm.cuh"
#include "mx.cuh"

namespace aclip {

struct GemmMxCfg {
  static constexpr int CTA_M = 128, CTA_N = 96;
  static constexpr int BLOCK_M = 256, BLOCK_N = 192, BLOCK_K = 64, UMMA_K = 16;
  static constexpr int A_H = 0, A_H_BYTES = CTA_M * 128;
  static constexpr int B_H = A_H + A_H_BYTES, B_H_BYTES = CTA_N * 128;
  static constexpr int A_Q = B_H + B_H_BYTES, A_Q_PLANE = CTA_M * 32;       // L4 then C4
  static constexpr int B_Q = A_Q + 2 * A_Q_PLANE, B_Q_PLANE = CTA_N * 32;
  static constexpr int SFA = B_Q + 2 * B_Q_PLANE, SFA_BYTES = 512;
  static constexpr int SFB = SFA + SFA_BYTES, SFB_BYTES = 1024;
  static constexpr int TX_BYTES = SFB + SFB_BYTES;                          // bytes landing per CTA per round
  static constexpr int STAGE_BYTES = (TX_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
  static constexpr int TMEM_COLS = 512;
  static constexpr int TMEM_SFA = 2 * BLOCK_N, TMEM_SFB = 2 * BLOCK_N + 8;
  static constexpr int EPI_WARPS = 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + EPI_WARPS * 32 * 128 + 1024;
  static constexpr int THREADS = 64 + EPI_WARPS * 32;
  static_assert(A_Q % 1024 == 0 && B_H % 1024 == 0 && B_Q % 256 == 0 && SFA % 128 == 0, "tile alignment");
  static_assert(STAGES >= 3, "operand ring too shallow");
};

// Epilogue of a warp's 32 rows x (chunks x 32) columns whose output is an f16mx tensor (the A
// operand of the next GEMM).  A lane holds one row's 32 consecutive columns of a chunk, which is
// exactly one scale block of that operand, so packing needs no exchange between lanes.  The next
// chunk's accumulator is requested from TMEM before the current one is processed; the fp16 plane
// (64 B per row and chunk) goes through the warp's staging buffer so that a store instruction
// writes 8 rows x 64 contiguous bytes instead of 32 rows x 16; the two e2m1 planes (16 B per row and
// chunk each) and the scale bytes are stored by the owning lane.
__device__ __forceinline__ float epilogue_tile_mx(const GemmParams& p, const MxOut& out, uint32_t t_row, int n0,
                                                  int chunks, int m_base, int lane, uint8_t* stage) {
  if (m_base >= p.M) return 0.f;   // warp-uniform: the whole 32-row block is padding
  uint32_t raw[2][32];
  float amax = 0.f;
  const int m = m_base + lane;
  const bool row_ok = m < p.M;
  const float sc = p.out_scale;
  ptx::tmem_ld_32x32(t_row, raw[0]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (c >= chunks) break;
    const int n = n0 + c * 32;
    float4 b4[8];
    if (p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) b4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ptx::tmem_ld_wait();
    if (c + 1 < chunks) ptx::tmem_ld_32x32(t_row + (c + 1) * 32, raw[(c + 1) & 1]);
    const uint32_t (&r)[32] = raw[c & 1];
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), sc, b4[j].x);
      v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), sc, b4[j].y);
      v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), sc, b4[j].z);
      v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), sc, b4[j].w);
    }
    if (p.act == ACT_QUICKGELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]) * kActScaleMain;
    } else if (p.act == ACT_LEAKYRELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (v[j] > 0.0f ? v[j] : 0.01f * v[j]) * kActScaleMain;
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= kActScaleMain;
    }
    uint32_t h[16], l4[4], c4[4], sf_l, sf_c;
    const float mx = mx_pack32(v, h, l4, c4, sf_l, sf_c);
    if (row_ok) amax = fmaxf(amax, mx);
    // fp16 plane through the staging buffer: lane = row, 4 pieces of 16 B, XOR-swizzled by the row pair
    __syncwarp();   // the previous chunk's readers are done
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<uint4*>(stage + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
          make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
    __syncwarp();
    const int piece = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = i * 8 + (lane >> 2);
      const int mm = m_base + row;
      if (mm < p.M) {
        const uint4 q = *reinterpret_cast<const uint4*>(stage + row * 64 + ((piece ^ ((row >> 1) & 3)) << 4));
        *reinterpret_cast<uint4*>(out.base + 2 * (static_cast<long long>(mm + p.row_offset) * out.ld + n) + piece * 16) = q;
      }
    }
    if (row_ok) {
      const long long mo = m + p.row_offset;
      const long long e = mo * out.ld + n;
      *reinterpret_cast<uint4*>(out.base + 2 * out.plane + (e >> 1)) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
      *reinterpret_cast<uint4*>(out.base + 2 * out.plane + (out.plane >> 1) + (e >> 1)) =
          make_uint4(c4[0], c4[1], c4[2], c4[3]);
      const int kb = n >> 5;
      uint8_t* sf = out.base + 3 * out.plane + (static_cast<long long>(kb >> 1) * out.row_blocks + (mo >> 7)) * 512 +
                    (mo & 31) * 16 + ((mo >> 5) & 3) * 4 + (kb & 1);
      sf[0] = static_cast<uint8_t>(sf_l);
      sf[2] = static_cast<uint8_t>(sf_c);
    }
  }
  return amax;
}

// EPI: 0 generic (fp32 / split outputs through epilogue_chunk), 3 fp32 + residual, 4 f16mx output
template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
gemm2mx_tcgen05_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAq,
                       const __grid_constant__ CUtensorMap tmAs, const __grid_constant__ CUtensorMap tmBh,
                       const __grid_constant__ CUtensorMap tmBq, const __grid_constant__ CUtensorMap tmBs,
                       const GemmParams p, const MxOut mx_out) {
  using Cfg = GemmMxCfg;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
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
    ptx::prefetch_tmap(&tmAh); ptx::prefetch_tmap(&tmAq); ptx::prefetch_tmap(&tmAs);
    ptx::prefetch_tmap(&tmBh); ptx::prefetch_tmap(&tmBq); ptx::prefetch_tmap(&tmBs);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 2 * Cfg::EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();

  const int m_tiles = (p.M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int n_tiles = p.N / Cfg::BLOCK_N;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        const int n_tile0 = (t % n_tiles) * Cfg::BLOCK_N;
        const int n0 = n_tile0 + static_cast<int>(rank) * Cfg::CTA_N;
        const int mblk = (t / n_tiles) * 2 + static_cast<int>(rank);   // 128-row block of A
        const int m0 = mblk * Cfg::CTA_M;
        const int nblk = n_tile0 >> 7;                                  // first 128-row scale chunk of the W tile
        for (int kb = 0; kb < p.num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
          const bool skip_q = (p.debug & 16) != 0;   // profiling experiment: no e2m1 planes
          if (rank == 0)
            ptx::mbar_expect_tx(&full_bar[stage], 2 * (Cfg::TX_BYTES - (skip_q ? 2 * (Cfg::A_Q_PLANE + Cfg::B_Q_PLANE) : 0)));
          ptx::tma_load_3d_pair(st + Cfg::A_H, &tmAh, &full_bar[stage], kb * 64, m0, 0);
          ptx::tma_load_3d_pair(st + Cfg::B_H, &tmBh, &full_bar[stage], kb * 64, n0, 0);
          if (!skip_q) {
            ptx::tma_load_3d_pair(st + Cfg::A_Q, &tmAq, &full_bar[stage], kb * 32, m0, 0);
            ptx::tma_load_3d_pair(st + Cfg::B_Q, &tmBq, &full_bar[stage], kb * 32, n0, 0);
          }
          ptx::tma_load_3d_pair(st + Cfg::SFA, &tmAs, &full_bar[stage], 0, mblk, kb);
          ptx::tma_load_3d_pair(st + Cfg::SFB, &tmBs, &full_bar[stage], 0, nblk, kb);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const bool leader = ptx::elect_one();
      const bool issue = leader && (p.debug & 1) == 0;
      constexpr uint32_t idesc_h = ptx::make_idesc_fmt0_f32(Cfg::BLOCK_M, Cfg::BLOCK_N);
      constexpr uint32_t idesc_lc = ptx::make_idesc_mxf4(Cfg::BLOCK_M, Cfg::BLOCK_N, 0, 2);   // x_L4 w_C4
      constexpr uint32_t idesc_cl = ptx::make_idesc_mxf4(Cfg::BLOCK_M, Cfg::BLOCK_N, 2, 0);   // x_C4 w_L4
      const uint64_t desc128 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem));
      const uint64_t desc32 = ptx::make_kmajor_sw32_desc(ptx::smem_u32(smem));
      const uint64_t desc_sf = ptx::make_chunk16_desc(ptx::smem_u32(smem));
      const uint32_t tsfa = tmem_base + Cfg::TMEM_SFA;
      uint32_t stage = 0, phase = 0;
      uint32_t acc = 0, acc_phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::BLOCK_N;
        // odd tiles start 64 rows into their first scale chunk: two TMEM columns further
        const uint32_t tsfb = tmem_base + Cfg::TMEM_SFB + ((((t % n_tiles) * Cfg::BLOCK_N) & 127) >> 5);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t off = stage * Cfg::STAGE_BYTES;
          ptx::tmem_cp_32x128b_pair_if(leader, tsfa, desc_sf + ((off + Cfg::SFA) >> 4));
          ptx::tmem_cp_32x128b_pair_if(leader, tmem_base + Cfg::TMEM_SFB, desc_sf + ((off + Cfg::SFB) >> 4));
          ptx::tmem_cp_32x128b_pair_if(leader, tmem_base + Cfg::TMEM_SFB + 4, desc_sf + ((off + Cfg::SFB + 512) >> 4));
          const uint64_t a_h = desc128 + ((off + Cfg::A_H) >> 4), b_h = desc128 + ((off + Cfg::B_H) >> 4);
#pragma unroll
          for (int k = 0; k < Cfg::BLOCK_K / Cfg::UMMA_K; ++k)
            ptx::mma_bf16_ss_pair_if(issue && !(p.debug & 4), d_tmem, a_h + 2 * k, b_h + 2 * k, idesc_h,
                                     (kb | k) != 0 ? 1u : 0u);
          const uint64_t a_l = desc32 + ((off + Cfg::A_Q) >> 4), a_c = a_l + (Cfg::A_Q_PLANE >> 4);
          const uint64_t b_l = desc32 + ((off + Cfg::B_Q) >> 4), b_c = b_l + (Cfg::B_Q_PLANE >> 4);
          const bool cross = issue && !(p.debug & 2);
          ptx::mma_mxf4_ss_pair_if(cross, d_tmem, a_l, b_c, idesc_lc, tsfa, tsfb, (p.debug & 4) && kb == 0 ? 0u : 1u);
          ptx::mma_mxf4_ss_pair_if(cross && !(p.debug & 8), d_tmem, a_c, b_l, idesc_cl, tsfa, tsfb, 1u);
          ptx::mma_commit_pair_if(leader, &empty_bar[stage], 0x3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit_pair_if(leader, &tfull_bar[acc], 0x3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int quarter = warp & 3;
    const int col_half = (warp - 2) >> 2;    // which 96 of the 192 columns
    uint8_t* stage = smem + STAGES * Cfg::STAGE_BYTES + 256 + (warp - 2) * EPI_STAGE_BYTES;
    uint32_t acc = 0, acc_phase = 0;
    float amax = 0.f;
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      const int n0 = (t % n_tiles) * Cfg::BLOCK_N + col_half * 96;
      const int m_base = (t / n_tiles) * Cfg::BLOCK_M + static_cast<int>(rank) * Cfg::CTA_M + quarter * 32;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + acc * Cfg::BLOCK_N + col_half * 96 +
                             (static_cast<uint32_t>(quarter * 32) << 16);
      if (EPI == 3) {
        epilogue_tile_residual(p, t_row, n0, 3, m_base, lane, stage);
      } else if (EPI == 4) {
        amax = fmaxf(amax, epilogue_tile_mx(p, mx_out, t_row, n0, 3, m_base, lane, stage));
      } else {
#pragma unroll 1
        for (int c = 0; c < 3; ++c) epilogue_chunk(p, t_row + c * 32, n0 + c * 32, m_base, lane, stage, p.M);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_remote(&tempty_bar[acc], 0);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (EPI == 4 && p.sat != nullptr && !(amax <= 65504.0f)) atomicAdd(p.sat, 1u);
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace aclip
