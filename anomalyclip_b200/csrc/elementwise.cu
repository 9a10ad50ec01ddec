// HBM-bound helper kernels: fp32 -> split-bf16 conversion, LayerNorm variants, frame patchify.
// All of them stream rows with 128-bit accesses; grids are sized as a multiple of the SM count.
#include <cuda_bf16.h>

#include "common.h"

namespace aclip {

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(a);
  const __nv_bfloat16 h1 = __float2bfloat16_rn(b);
  const __nv_bfloat16 l0 = __float2bfloat16_rn(a - __bfloat162float(h0));
  const __nv_bfloat16 l1 = __float2bfloat16_rn(b - __bfloat162float(h1));
  hi = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) |
       (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
  lo = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) |
       (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
}

// ------------------------------------------------------------------------------ split
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ in, long long rows, int cols, int ld_in,
             __nv_bfloat16* __restrict__ out, int ld_out, long long plane_stride, bool vec_ok) {
  const int groups_per_row = ld_out >> 3;
  const long long total = rows * groups_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups_per_row;
    const int c = static_cast<int>(i - r * groups_per_row) << 3;
    float v[8];
    const float* src = in + r * ld_in + c;
    if (vec_ok && c + 8 <= cols) {
      const float4 a = *reinterpret_cast<const float4*>(src);
      const float4 b = *reinterpret_cast<const float4*>(src + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? src[j] : 0.0f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
    __nv_bfloat16* dst = out + r * ld_out + c;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + plane_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

static int grid_for(long long work_items, int threads, int max_waves = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

int split_f32(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out,
              long long plane_stride, cudaStream_t stream) {
  ACLIP_REQUIRE(in != nullptr && out != nullptr, "split: null pointer");
  ACLIP_REQUIRE(rows >= 0 && cols > 0 && ld_in >= cols, "split: bad shape");
  ACLIP_REQUIRE(ld_out % 8 == 0 && ld_out >= cols, "split: ld_out=%d must be a multiple of 8 >= cols",
                ld_out);
  ACLIP_REQUIRE(plane_stride % 8 == 0 && plane_stride >= rows * ld_out,
                "split: plane_stride too small or unaligned");
  ACLIP_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "split: output must be 16-byte aligned");
  if (rows == 0) return ACLIP_OK;
  const bool vec_ok = (ld_in % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  const long long total = rows * (ld_out >> 3);
  split_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      in, rows, cols, ld_in, static_cast<__nv_bfloat16*>(out), ld_out, plane_stride, vec_ok);
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

}  // namespace aclip

extern "C" int aclip_split_f32(const float* in, long long rows, int cols, int ld_in,
                               void* out_split, int ld_out, long long plane_stride, void* stream) {
  return aclip::split_f32(in, rows, cols, ld_in, out_split, ld_out, plane_stride,
                          aclip::as_stream(stream));
}
