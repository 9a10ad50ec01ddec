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
  if (eh == 32) {
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
