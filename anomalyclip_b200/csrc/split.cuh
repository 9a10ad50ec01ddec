// fp32 pair -> packed bf16x2 hi plane and lo plane (lo = bf16(x - hi)); two elements per
// conversion instruction (cvt.rn.bf16x2.f32 runs on the FMA-class pipe, not the 16-lane XU pipe).
#pragma once
#include <cuda_bf16.h>

#include <cstdint>

namespace aclip {

__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);  // .x = a (low half), .y = b (high half)
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ha = __uint_as_float(hi << 16);
  const float hb = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

}  // namespace aclip
