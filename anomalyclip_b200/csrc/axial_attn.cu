// Axial self-attention of the temporal transformer (restated axial_attention.SelfAttention,
// call site src/models/components/temporal_model.py:32-39,64): sequences of 32 segments (long
// range, axis n) or 16 frames (short range, axis l), 8 heads of E/8 dims, no mask.
//
// The sequences are tiny (L <= 32), so this is exact fp32 SIMT work: one CTA per sequence, one
// warp per head, lane i owns query row i; the head's K and V slices sit in shared memory and are
// read as warp-wide broadcasts.  Rows are addressed in sub-video order: grid cell (i, k) of
// sub-video S is row S*n*l + i*l + k, so an axis-n sequence is a stride-l walk and an axis-l
// sequence is l consecutive rows.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"
#include "split.cuh"

namespace aclip {

namespace {

constexpr int MAX_L = 32;

template <int EH>  // dims per head
__global__ void __launch_bounds__(256)
axial_attention_kernel(const float* __restrict__ qkv, int E, int heads, int L, long long unit,
                       int inner, int inner_mul, int stride, float scale,
                       __nv_bfloat16* __restrict__ out, long long plane_stride) {
  extern __shared__ float ax_smem[];
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= heads) return;
  const long long q = blockIdx.x;
  const long long base = (q / inner) * unit + (q % inner) * inner_mul;
  const int ld = 3 * E;
  float* ks = ax_smem + warp * (2 * MAX_L * EH);
  float* vs = ks + MAX_L * EH;

  // stage K and V of this head: L rows x EH floats each
  for (int i = lane; i < L * (EH / 4); i += 32) {
    const int j = i / (EH / 4), d4 = i - j * (EH / 4);
    const float* row = qkv + (base + static_cast<long long>(j) * stride) * ld + warp * EH + 4 * d4;
    reinterpret_cast<float4*>(ks)[i] = *reinterpret_cast<const float4*>(row + E);
    reinterpret_cast<float4*>(vs)[i] = *reinterpret_cast<const float4*>(row + 2 * E);
  }
  __syncwarp();
  if (lane >= L) return;

  const long long my_row = base + static_cast<long long>(lane) * stride;
  float qr[EH];
  {
    const float4* q4 = reinterpret_cast<const float4*>(qkv + my_row * ld + warp * EH);
#pragma unroll
    for (int d = 0; d < EH / 4; ++d) {
      const float4 t = q4[d];
      qr[4 * d] = t.x; qr[4 * d + 1] = t.y; qr[4 * d + 2] = t.z; qr[4 * d + 3] = t.w;
    }
  }
  float s[MAX_L];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < MAX_L; ++j) {
    if (j < L) {
      float acc = 0.f;
      const float4* k4 = reinterpret_cast<const float4*>(ks + j * EH);
#pragma unroll
      for (int d = 0; d < EH / 4; ++d) {
        const float4 t = k4[d];
        acc = fmaf(qr[4 * d], t.x, acc);
        acc = fmaf(qr[4 * d + 1], t.y, acc);
        acc = fmaf(qr[4 * d + 2], t.z, acc);
        acc = fmaf(qr[4 * d + 3], t.w, acc);
      }
      s[j] = acc * scale;
      mx = fmaxf(mx, s[j]);
    } else {
      s[j] = -INFINITY;
    }
  }
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < MAX_L; ++j) {
    s[j] = j < L ? expf(s[j] - mx) : 0.f;
    den += s[j];
  }
  const float inv = 1.0f / den;
  float o[EH];
#pragma unroll
  for (int d = 0; d < EH; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < MAX_L; ++j) {
    if (j < L) {
      const float p = s[j] * inv;
      const float4* v4 = reinterpret_cast<const float4*>(vs + j * EH);
#pragma unroll
      for (int d = 0; d < EH / 4; ++d) {
        const float4 t = v4[d];
        o[4 * d] = fmaf(p, t.x, o[4 * d]);
        o[4 * d + 1] = fmaf(p, t.y, o[4 * d + 1]);
        o[4 * d + 2] = fmaf(p, t.z, o[4 * d + 2]);
        o[4 * d + 3] = fmaf(p, t.w, o[4 * d + 3]);
      }
    }
  }
  __nv_bfloat16* dst = out + my_row * E + warp * EH;
#pragma unroll
  for (int d = 0; d < EH; d += 8) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      split_pack2(o[d + 2 * t], o[d + 2 * t + 1], hi[t], lo[t]);
    }
    *reinterpret_cast<uint4*>(dst + d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + d + plane_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}


// ---------------------------------------------------------------------------------------------
// Tensor-core version for the reference's grid (L = 32 segments or 16 frames, 16 or 32 dims per
// head): one warp per (sequence, head) again, but S = Q K^T and O = P V run as warp-level
// mma.sync m16n8k16 tiles on split-bf16 operands (hi*hi + lo*hi + hi*lo, the GEMMs' three-pass
// product), with every operand fragment read straight from global memory in its MMA layout
// (8-byte loads, 32-byte sectors fully used), the scores kept in the accumulator registers through
// the softmax and re-used as the A fragments of P V (the m16n8 C layout of two adjacent key tiles IS
// the m16k16 A layout).  About 700 instructions per warp instead of ~2 600 (the SIMT kernel issues
// one shared-memory broadcast load per four FMAs), no shared memory, all loads in flight at once:
// the stage is HBM-bound, not issue-bound.  (tcgen05's M = 128 tiles would be 4-8x padding here.)
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_16816(c, ah, bh);
  mma_16816(c, al, bh);
  mma_16816(c, ah, bl);
}

template <int L, int EH>
__global__ void __launch_bounds__(256)
axial_attention_mma_kernel(const float* __restrict__ qkv, int E, int heads, long long unit, int inner,
                           int inner_mul, int stride, float scale_log2,
                           __nv_bfloat16* __restrict__ out, long long plane_stride) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= heads) return;
  constexpr int MT = L / 16, NT = L / 8, KS = EH / 16, DT = EH / 8;
  const int g = lane >> 2, t = lane & 3;
  const long long q = blockIdx.x;
  const long long base = (q / inner) * unit + (q % inner) * inner_mul;
  const long long ld = 3ll * E;
  const float* Q = qkv + base * ld + warp * EH;   // row j of the sequence: + j * stride * ld
  const float* K = Q + E;
  const float* V = Q + 2 * E;
  const long long rs = static_cast<long long>(stride) * ld;

  // ---- Q fragments (A operand of S), hi / lo
  uint32_t qh[MT][KS][4], ql[MT][KS][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const float* r0 = Q + (mt * 16 + g) * rs + ks * 16 + 2 * t;
      const float* r1 = r0 + 8 * rs;
      const float2 a = *reinterpret_cast<const float2*>(r0), b = *reinterpret_cast<const float2*>(r1);
      const float2 c = *reinterpret_cast<const float2*>(r0 + 8), d = *reinterpret_cast<const float2*>(r1 + 8);
      split_pack2(a.x, a.y, qh[mt][ks][0], ql[mt][ks][0]);
      split_pack2(b.x, b.y, qh[mt][ks][1], ql[mt][ks][1]);
      split_pack2(c.x, c.y, qh[mt][ks][2], ql[mt][ks][2]);
      split_pack2(d.x, d.y, qh[mt][ks][3], ql[mt][ks][3]);
    }
  // ---- S = Q K^T (B fragment of key tile nt: key nt*8 + g, dims 2t, 2t+1 and +8)
  float s[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[mt][nt][i] = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const float* kr = K + (nt * 8 + g) * rs + ks * 16 + 2 * t;
      const float2 a = *reinterpret_cast<const float2*>(kr), b = *reinterpret_cast<const float2*>(kr + 8);
      uint32_t kh[2], kl[2];
      split_pack2(a.x, a.y, kh[0], kl[0]);
      split_pack2(b.x, b.y, kh[1], kl[1]);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma3(s[mt][nt], qh[mt][ks], ql[mt][ks], kh, kl);
    }
  // ---- softmax over the keys: a row's scores sit in the four lanes of a quad
  uint32_t ph[MT][L / 16][4], pl[MT][L / 16][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      m0 = fmaxf(m0, fmaxf(s[mt][nt][0], s[mt][nt][1]));
      m1 = fmaxf(m1, fmaxf(s[mt][nt][2], s[mt][nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[mt][nt][0] = exp2f((s[mt][nt][0] - m0) * scale_log2);
      s[mt][nt][1] = exp2f((s[mt][nt][1] - m0) * scale_log2);
      s[mt][nt][2] = exp2f((s[mt][nt][2] - m1) * scale_log2);
      s[mt][nt][3] = exp2f((s[mt][nt][3] - m1) * scale_log2);
      d0 += s[mt][nt][0] + s[mt][nt][1];
      d1 += s[mt][nt][2] + s[mt][nt][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    const float i0 = 1.0f / d0, i1 = 1.0f / d1;
#pragma unroll
    for (int kk = 0; kk < L / 16; ++kk) {   // key tiles 2kk, 2kk+1 -> one 16-key A fragment
      split_pack2(s[mt][2 * kk][0] * i0, s[mt][2 * kk][1] * i0, ph[mt][kk][0], pl[mt][kk][0]);
      split_pack2(s[mt][2 * kk][2] * i1, s[mt][2 * kk][3] * i1, ph[mt][kk][1], pl[mt][kk][1]);
      split_pack2(s[mt][2 * kk + 1][0] * i0, s[mt][2 * kk + 1][1] * i0, ph[mt][kk][2], pl[mt][kk][2]);
      split_pack2(s[mt][2 * kk + 1][2] * i1, s[mt][2 * kk + 1][3] * i1, ph[mt][kk][3], pl[mt][kk][3]);
    }
  }
  // ---- O = P V (B fragment of dim tile dt: dim dt*8 + g, keys 2t, 2t+1 and +8 of the 16-key step)
  float o[MT][DT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int dt = 0; dt < DT; ++dt)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[mt][dt][i] = 0.f;
#pragma unroll
  for (int dt = 0; dt < DT; ++dt)
#pragma unroll
    for (int kk = 0; kk < L / 16; ++kk) {
      const float* vr = V + (kk * 16 + 2 * t) * rs + dt * 8 + g;
      uint32_t vh[2], vl[2];
      split_pack2(vr[0], vr[rs], vh[0], vl[0]);
      split_pack2(vr[8 * rs], vr[9 * rs], vh[1], vl[1]);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma3(o[mt][dt], ph[mt][kk], pl[mt][kk], vh, vl);
    }
  // ---- split-bf16 rows out (row stride of the output: E elements, sequence stride `stride` rows)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const long long r0 = (base + static_cast<long long>(mt * 16 + g) * stride) * E + warp * EH + dt * 8 + 2 * t;
      const long long r1 = r0 + 8ll * stride * E;
      uint32_t h, l;
      split_pack2(o[mt][dt][0], o[mt][dt][1], h, l);
      *reinterpret_cast<uint32_t*>(out + r0) = h;
      *reinterpret_cast<uint32_t*>(out + r0 + plane_stride) = l;
      split_pack2(o[mt][dt][2], o[mt][dt][3], h, l);
      *reinterpret_cast<uint32_t*>(out + r1) = h;
      *reinterpret_cast<uint32_t*>(out + r1 + plane_stride) = l;
    }
}

}  // namespace

// axis 0: attend along the n segments (long range); axis 1: along the l frames of a segment.
int axial_attention(const float* qkv, long long sub_videos, int n, int l, int E, int heads,
                    int axis, void* out_split, long long plane_stride, cudaStream_t stream) {
  ACLIP_REQUIRE(qkv != nullptr && out_split != nullptr, "axial_attention: null pointer");
  ACLIP_REQUIRE(heads >= 1 && heads <= 8 && E % heads == 0, "axial_attention: heads=%d E=%d", heads, E);
  const int eh = E / heads;
  ACLIP_REQUIRE(eh == 16 || eh == 32, "axial_attention: E/heads=%d unsupported (16 or 32)", eh);
  ACLIP_REQUIRE(n >= 1 && n <= MAX_L && l >= 1 && l <= MAX_L, "axial_attention: grid %dx%d too large", n, l);
  ACLIP_REQUIRE(axis == 0 || axis == 1, "axial_attention: axis must be 0 or 1");
  if (sub_videos <= 0) return ACLIP_OK;
  const long long unit = static_cast<long long>(n) * l;
  const int L = axis == 0 ? n : l;
  const int inner = axis == 0 ? l : n;
  const int inner_mul = axis == 0 ? 1 : l;
  const int stride = axis == 0 ? l : 1;
  const long long seqs = sub_videos * inner;
  ACLIP_REQUIRE(seqs < (1ll << 31), "axial_attention: too many sequences");
  const float scale = 1.0f / sqrtf(static_cast<float>(eh));
  const int smem = heads * 2 * MAX_L * eh * static_cast<int>(sizeof(float));
  auto* o = static_cast<__nv_bfloat16*>(out_split);
  timing_begin(KIND_AXIAL_ATTENTION, stream);
  // the reference's grid (32 segments x 16 frames) runs on the tensor cores; ACLIP_AXIAL_SIMT=1
  // forces the fp32 SIMT kernel (A/B runs), which also serves every other sequence length
  static const bool force_simt = [] {
    const char* e = getenv("ACLIP_AXIAL_SIMT");
    return e != nullptr && e[0] == '1';
  }();
  if (!force_simt && (L == 32 || L == 16)) {
    const float scale_log2 = scale * 1.4426950408889634f;
    const dim3 grid(static_cast<unsigned>(seqs)), block(heads * 32);
#define ACLIP_AXIAL_MMA(LL, EE)                                                                    \
  ACLIP_CUDA_OK(launch_pdl(axial_attention_mma_kernel<LL, EE>, grid, block, 0, stream, qkv, E, heads, unit, \
                           inner, inner_mul, stride, scale_log2, o, plane_stride))
    if (L == 32 && eh == 32) ACLIP_AXIAL_MMA(32, 32);
    else if (L == 32) ACLIP_AXIAL_MMA(32, 16);
    else if (eh == 32) ACLIP_AXIAL_MMA(16, 32);
    else ACLIP_AXIAL_MMA(16, 16);
#undef ACLIP_AXIAL_MMA
  } else if (eh == 32) {
    static PerDeviceOnce once;
    int once_dev;
    if (once.need(once_dev)) {
      ACLIP_CUDA_OK(cudaFuncSetAttribute(axial_attention_kernel<32>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * MAX_L * 32 * 4));
      once.mark(once_dev);
    }
    ACLIP_CUDA_OK(launch_pdl(axial_attention_kernel<32>, dim3(static_cast<unsigned>(seqs)), dim3(heads * 32), smem,
                             stream, qkv, E, heads, L, unit, inner, inner_mul, stride, scale, o, plane_stride));
  } else {
    ACLIP_CUDA_OK(launch_pdl(axial_attention_kernel<16>, dim3(static_cast<unsigned>(seqs)), dim3(heads * 32), smem,
                             stream, qkv, E, heads, L, unit, inner, inner_mul, stride, scale, o, plane_stride));
  }
  timing_end(KIND_AXIAL_ATTENTION, stream, 4.0 * seqs * (double)L * L * E, (double)seqs * L * E * 16.0);
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

}  // namespace aclip

extern "C" int aclip_axial_attention(const float* qkv, long long sub_videos, int n, int l, int E,
                                     int heads, int axis, void* out_split, long long plane_stride,
                                     void* stream) {
  return aclip::axial_attention(qkv, sub_videos, n, l, E, heads, axis, out_split, plane_stride,
                                aclip::as_stream(stream));
}
