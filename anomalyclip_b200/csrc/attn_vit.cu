// Multi-head self-attention of the ViT blocks: softmax(Q K^T / sqrt(64)) V, no mask
// (nn.MultiheadAttention(need_weights=False, attn_mask=None), clip/model.py:191,206-212).
//
// Attention is 4 % of a block's FLOPs, so it runs on the warp-level tensor-core
// path (mma.sync m16n8k16) with the same split-bf16 scheme as the GEMMs: Q, K, V and P are each
// carried as hi + lo bf16 and every product issues hi*hi + lo*hi + hi*lo into fp32 accumulators.
//
// One CTA = (query half, head, frame); 7 warps, each owning a 16-row query tile.  K and V
// (hi and lo planes, all L keys) live in shared memory with a 16-byte XOR swizzle so that
// ldmatrix is conflict free; scores never leave registers (online softmax over 64-key chunks).
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.h"
#include "split.cuh"

namespace aclip {

namespace {

constexpr int HD = 64;           // head dim
constexpr int ROW_BYTES = HD * 2;  // one K/V row of bf16
constexpr int ATT_WARPS = 7;
constexpr int ATT_THREADS = ATT_WARPS * 32;

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  split_pack2(a, b, hi, lo);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct AttnState {
  float o[8][4];     // 16 x 64 output accumulator (8 n-tiles of 8 dims)
  float m[2], l[2];  // running max (raw scores) / partial row sums for rows g and g+8
};

// One chunk of NG*16 keys starting at key0.
template <int NG>
__device__ __forceinline__ void attend_chunk(AttnState& st, const uint32_t (&qh)[4][4],
                                             const uint32_t (&ql)[4][4], uint32_t k_hi,
                                             uint32_t k_lo, uint32_t v_hi, uint32_t v_lo,
                                             int key0, int L, float sl2, int lane) {
  constexpr int NT = NG * 2;
  float s[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    const int row = key0 + j * 8 + (lane & 7);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int chunk = (lane >> 3) + 4 * half;
      const uint32_t off = row * ROW_BYTES + ((chunk ^ (row & 7)) << 4);
      uint32_t bh[4], bl[4];
      ldmatrix_x4(k_hi + off, bh);
      ldmatrix_x4(k_lo + off, bl);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ks = half * 2 + kk;
        mma16816(s[j], qh[ks], bh[2 * kk], bh[2 * kk + 1]);
        mma16816(s[j], ql[ks], bh[2 * kk], bh[2 * kk + 1]);
        mma16816(s[j], qh[ks], bl[2 * kk], bl[2 * kk + 1]);
      }
    }
  }
  // mask the padded keys, chunk row maxima
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int key = key0 + j * 8 + 2 * (lane & 3);
    if (key >= L) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
    if (key + 1 >= L) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
    mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
    mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
  }
  const float mn0 = fmaxf(st.m[0], quad_max(mx0));
  const float mn1 = fmaxf(st.m[1], quad_max(mx1));
  const float a0 = exp2f((st.m[0] - mn0) * sl2);
  const float a1 = exp2f((st.m[1] - mn1) * sl2);
  st.m[0] = mn0; st.m[1] = mn1;
  const float b0 = mn0 * sl2, b1 = mn1 * sl2;
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    s[j][0] = exp2f(s[j][0] * sl2 - b0);
    s[j][1] = exp2f(s[j][1] * sl2 - b0);
    s[j][2] = exp2f(s[j][2] * sl2 - b1);
    s[j][3] = exp2f(s[j][3] * sl2 - b1);
    sum0 += s[j][0] + s[j][1];
    sum1 += s[j][2] + s[j][3];
  }
  st.l[0] = st.l[0] * a0 + sum0;
  st.l[1] = st.l[1] * a1 + sum1;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    st.o[d][0] *= a0; st.o[d][1] *= a0; st.o[d][2] *= a1; st.o[d][3] *= a1;
  }
  // O += P V
#pragma unroll
  for (int jj = 0; jj < NG; ++jj) {
    uint32_t ph[4], pl[4];
    split_pair(s[2 * jj][0], s[2 * jj][1], ph[0], pl[0]);
    split_pair(s[2 * jj][2], s[2 * jj][3], ph[1], pl[1]);
    split_pair(s[2 * jj + 1][0], s[2 * jj + 1][1], ph[2], pl[2]);
    split_pair(s[2 * jj + 1][2], s[2 * jj + 1][3], ph[3], pl[3]);
    const int row = key0 + jj * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-dim n-tiles
      const int chunk = dp * 2 + (lane >> 4);
      const uint32_t off = row * ROW_BYTES + ((chunk ^ (row & 7)) << 4);
      uint32_t vh[4], vl[4];
      ldmatrix_x4_trans(v_hi + off, vh);
      ldmatrix_x4_trans(v_lo + off, vl);
      mma16816(st.o[2 * dp], ph, vh[0], vh[1]);
      mma16816(st.o[2 * dp], pl, vh[0], vh[1]);
      mma16816(st.o[2 * dp], ph, vl[0], vl[1]);
      mma16816(st.o[2 * dp + 1], ph, vh[2], vh[3]);
      mma16816(st.o[2 * dp + 1], pl, vh[2], vh[3]);
      mma16816(st.o[2 * dp + 1], ph, vl[2], vl[3]);
    }
  }
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
vit_attention_kernel(const __nv_bfloat16* __restrict__ qkv, long long in_plane_stride, int ld_in,
                     int L, int LP, int width, __nv_bfloat16* __restrict__ out,
                     long long out_plane_stride, int ld_out, float sl2) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int plane_bytes = LP * ROW_BYTES;
  const long long row0 = static_cast<long long>(b) * L;

  // ---- stage K and V (hi, lo) of this (frame, head) into shared memory
  const uint32_t smem_base = smem_addr(att_smem);
  const int total_chunks = 4 * LP * 8;
  for (int i = tid; i < total_chunks; i += ATT_THREADS) {
    const int arr = i / (LP * 8);  // 0 K.hi, 1 K.lo, 2 V.hi, 3 V.lo
    const int rem = i - arr * LP * 8;
    const int row = rem >> 3, ch = rem & 7;
    const uint32_t dst = smem_base + arr * plane_bytes + row * ROW_BYTES + ((ch ^ (row & 7)) << 4);
    if (row < L) {
      const __nv_bfloat16* src = qkv + (arr & 1) * in_plane_stride + (row0 + row) * ld_in +
                                 (arr < 2 ? width : 2 * width) + h * HD + ch * 8;
      cp_async16(dst, src);
    } else {
      *reinterpret_cast<uint4*>(att_smem + (dst - smem_base)) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // ---- this warp's query tile (A fragments straight from global memory)
  const int tiles = LP >> 4;
  const int tiles_per_half = (tiles + 1) >> 1;
  const int tile = half * tiles_per_half + warp;
  const bool active = warp < tiles_per_half && tile < tiles;
  const int g = lane >> 2, c = lane & 3;
  const int rA = tile * 16 + g, rB = rA + 8;
  uint32_t qh[4][4], ql[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = (e & 1) ? rB : rA;
      const int col = ks * 16 + ((e >> 1) ? 8 : 0) + 2 * c;
      uint32_t vh = 0u, vl = 0u;
      if (active && r < L) {
        const __nv_bfloat16* src = qkv + (row0 + r) * ld_in + h * HD + col;
        vh = *reinterpret_cast<const uint32_t*>(src);
        vl = *reinterpret_cast<const uint32_t*>(src + in_plane_stride);
      }
      qh[ks][e] = vh;
      ql[ks][e] = vl;
    }

  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (!active) return;

  AttnState st;
#pragma unroll
  for (int d = 0; d < 8; ++d) st.o[d][0] = st.o[d][1] = st.o[d][2] = st.o[d][3] = 0.f;
  st.m[0] = st.m[1] = -INFINITY;
  st.l[0] = st.l[1] = 0.f;

  const uint32_t k_hi = smem_base, k_lo = smem_base + plane_bytes;
  const uint32_t v_hi = smem_base + 2 * plane_bytes, v_lo = smem_base + 3 * plane_bytes;
  int key0 = 0;
  for (; key0 + 64 <= LP; key0 += 64)
    attend_chunk<4>(st, qh, ql, k_hi, k_lo, v_hi, v_lo, key0, L, sl2, lane);
  const int rest = (LP - key0) >> 4;  // 0..3 groups of 16 keys left
  if (rest == 1) attend_chunk<1>(st, qh, ql, k_hi, k_lo, v_hi, v_lo, key0, L, sl2, lane);
  else if (rest == 2) attend_chunk<2>(st, qh, ql, k_hi, k_lo, v_hi, v_lo, key0, L, sl2, lane);
  else if (rest == 3) attend_chunk<3>(st, qh, ql, k_hi, k_lo, v_hi, v_lo, key0, L, sl2, lane);

  const float inv0 = 1.0f / quad_sum(st.l[0]);
  const float inv1 = 1.0f / quad_sum(st.l[1]);
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    const int col = h * HD + d * 8 + 2 * c;
    uint32_t hi, lo;
    if (rA < L) {
      split_pair(st.o[d][0] * inv0, st.o[d][1] * inv0, hi, lo);
      __nv_bfloat16* dst = out + (row0 + rA) * ld_out + col;
      *reinterpret_cast<uint32_t*>(dst) = hi;
      *reinterpret_cast<uint32_t*>(dst + out_plane_stride) = lo;
    }
    if (rB < L) {
      split_pair(st.o[d][2] * inv1, st.o[d][3] * inv1, hi, lo);
      __nv_bfloat16* dst = out + (row0 + rB) * ld_out + col;
      *reinterpret_cast<uint32_t*>(dst) = hi;
      *reinterpret_cast<uint32_t*>(dst + out_plane_stride) = lo;
    }
  }
}

}  // namespace

int vit_attention_mma(const void* qkv_split, long long in_plane_stride, int ld_in, int B, int L,
                      int heads, void* out_split, long long out_plane_stride, int ld_out,
                      cudaStream_t stream) {
  ACLIP_REQUIRE(qkv_split != nullptr && out_split != nullptr, "vit_attention: null pointer");
  ACLIP_REQUIRE(B > 0 && heads > 0 && L > 0, "vit_attention: empty problem");
  const int LP = (L + 15) / 16 * 16;
  const int tiles = LP / 16;
  ACLIP_REQUIRE((tiles + 1) / 2 <= ATT_WARPS, "vit_attention: L=%d exceeds the %d-row limit", L,
                ATT_WARPS * 32);
  ACLIP_REQUIRE(ld_in % 8 == 0 && ld_in >= 3 * heads * HD && ld_out % 2 == 0 &&
                    ld_out >= heads * HD && in_plane_stride % 8 == 0 && out_plane_stride % 2 == 0,
                "vit_attention: bad pitches");
  ACLIP_REQUIRE((reinterpret_cast<uintptr_t>(qkv_split) & 15) == 0,
                "vit_attention: qkv must be 16-byte aligned");
  const int smem = 4 * LP * ROW_BYTES;
  static PerDeviceOnce once;
  int once_dev;
  if (once.need(once_dev)) {
    ACLIP_CUDA_OK(cudaFuncSetAttribute(vit_attention_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       4 * ATT_WARPS * 32 * ROW_BYTES));
    once.mark(once_dev);
  }
  const float sl2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid(2, heads, B);
  timing_begin(KIND_VIT_ATTENTION, stream);
  vit_attention_kernel<<<grid, ATT_THREADS, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(qkv_split), in_plane_stride, ld_in, L, LP, heads * HD,
      static_cast<__nv_bfloat16*>(out_split), out_plane_stride, ld_out, sl2);
  timing_end(KIND_VIT_ATTENTION, stream, 4.0 * B * heads * (double)L * L * HD,
             (double)B * L * heads * HD * (3 * 4.0 + 4.0));
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

}  // namespace aclip

namespace aclip {
// kernel: 0 = default (tcgen05), 1 = warp-level mma.sync kernel, 2 = tcgen05 kernel
int vit_attention(const void* qkv_split, long long in_plane_stride, int ld_in, int B, int L,
                  int heads, void* out_split, long long out_plane_stride, int ld_out, int kernel,
                  int out_enc, cudaStream_t stream) {
  ACLIP_REQUIRE(out_enc == 0 || kernel == 0 || kernel == 2,
                "vit_attention: only the tcgen05 kernel writes f16f8 output");
  if (kernel >= 16) {
    // profiling experiments (kernel = 16 + mask: skip softmax math / output stores / TMEM stores;
    // results are WRONG by construction) -- only with ACLIP_PROFILING_EXPERIMENTS=1 in the environment
    const char* allow = getenv("ACLIP_PROFILING_EXPERIMENTS");
    ACLIP_REQUIRE(allow != nullptr && allow[0] == '1',
                  "vit_attention: kernel must be 0, 1 or 2 (got %d)", kernel);
    return vit_attention_tc(qkv_split, in_plane_stride, ld_in, B, L, heads, out_split,
                            out_plane_stride, ld_out, stream, kernel - 16);
  }
  ACLIP_REQUIRE(kernel >= 0 && kernel <= 2, "vit_attention: kernel must be 0, 1 or 2");
  if (kernel == 1)
    return vit_attention_mma(qkv_split, in_plane_stride, ld_in, B, L, heads, out_split,
                             out_plane_stride, ld_out, stream);
  return vit_attention_tc(qkv_split, in_plane_stride, ld_in, B, L, heads, out_split,
                          out_plane_stride, ld_out, stream, 0, out_enc);
}
}  // namespace aclip

extern "C" int aclip_vit_attention(const void* qkv_split, long long in_plane_stride, int ld_in,
                                   int B, int L, int heads, void* out_split,
                                   long long out_plane_stride, int ld_out, int kernel,
                                   int out_enc, void* stream) {
  return aclip::vit_attention(qkv_split, in_plane_stride, ld_in, B, L, heads, out_split,
                              out_plane_stride, ld_out, kernel, out_enc, aclip::as_stream(stream));
}
