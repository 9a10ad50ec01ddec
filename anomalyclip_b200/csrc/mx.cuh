// "f16mx" operand encoding (GEMM passes == 7): the f16f8 idea (split.cuh) with the two correction
// planes as MXFP4 instead of e4m3.  An fp32 value v travels as
//     H  = fp16(v * 2^e_main)                        main product        (kind::f16 MMA)
//     L4 = mxfp4(v * 2^e_main - H)                   residual of H       (kind::mxf4 MMA, 4x rate)
//     C4 = mxfp4(v * 2^e_main)                       coarse copy, multiplies the OTHER operand's L4
// mxfp4 = e2m1 elements (0, .5, 1, 1.5, 2, 3, 4, 6 and their negatives) with one power-of-two
// scale (UE8M0 byte = exponent + 127) per 32 consecutive values along K, chosen as the smallest
// power of two that brings the block maximum to <= 6.  The GEMM accumulates
//     x_H w_H + x_L4 w_C4 + x_C4 w_L4
// in units of 2^(e_x + e_w): the block scales carry absolute exponents, so no extra shifts.  The
// correction terms only have to be right to a few bits (they repair the 2^-12 rounding of the
// fp16 planes); on the ViT-B/16 features this costs 9.0e-5 against f16f8's 8.0e-5 when used for
// the MLP pair (scripts/numerics_passes.py) at 1.5 instead of 2 tensor-pass equivalents.
//
// Memory layout of a tensor [rows][ld] (ld % 64 == 0), P = rows * ld:
//     byte 0        H   fp16 [rows][ld]
//     byte 2P       L4  packed e2m1 [rows][ld / 2]   (element k of a row in nibble k & 1 of byte k / 2)
//     byte 2P+P/2   C4  packed e2m1 [rows][ld / 2]
//     byte 3P       SF  [ld / 64 atoms][ceil(rows / 128) row blocks][512 bytes]: the 32 x 16-byte
//                       chunk tcgen05.cp (32x128b.warpx4) copies into TMEM -- row m of the block at
//                       (m & 31) * 16 + (m >> 5) * 4, bytes 0..1 = L4 scales of the atom's two
//                       32-value blocks, bytes 2..3 = C4 scales
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp4.h>

#include <cstdint>

namespace aclip {

// UE8M0 byte of the smallest power of two s with amax <= 6 s (0 for amax == 0)
__device__ __forceinline__ uint32_t mx_scale_byte(float amax) {
  const uint32_t b = __float_as_uint(amax * (1.0f / 6.0f));
  const uint32_t e = (b >> 23) + ((b & 0x7fffffu) != 0u ? 1u : 0u);
  return e > 254u ? 254u : e;
}
// 1 / scale of that byte (2^(127 - byte)); finite for every byte <= 253
__device__ __forceinline__ float mx_inv_scale(uint32_t byte) {
  return __uint_as_float((254u - byte) << 23);
}
__device__ __forceinline__ uint32_t mx_e2m1x2(float lo, float hi) {   // lo -> low nibble
  return static_cast<uint32_t>(__nv_cvt_float2_to_fp4x2(make_float2(lo, hi), __NV_E2M1, cudaRoundNearest));
}
// eight consecutive values -> 8 nibbles (value 0 in the low nibble of the low byte)
__device__ __forceinline__ uint32_t mx_e2m1x8(const float* v, float inv) {
  return mx_e2m1x2(v[0] * inv, v[1] * inv) | (mx_e2m1x2(v[2] * inv, v[3] * inv) << 8) |
         (mx_e2m1x2(v[4] * inv, v[5] * inv) << 16) | (mx_e2m1x2(v[6] * inv, v[7] * inv) << 24);
}

// One 32-value block of a row, values already multiplied by 2^e_main: fp16 plane (16 packed
// pairs), the two e2m1 planes (4 x 8 nibbles each) and their scale bytes.  Returns max |v|.
__device__ __forceinline__ float mx_pack32(const float (&v)[32], uint32_t (&h)[16], uint32_t (&l4)[4],
                                           uint32_t (&c4)[4], uint32_t& sf_l, uint32_t& sf_c) {
  float r[32];
  float amax_v = 0.f, amax_r = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
    r[2 * j] = v[2 * j] - hf.x;
    r[2 * j + 1] = v[2 * j + 1] - hf.y;
    amax_v = fmaxf(amax_v, fmaxf(fabsf(v[2 * j]), fabsf(v[2 * j + 1])));
    amax_r = fmaxf(amax_r, fmaxf(fabsf(r[2 * j]), fabsf(r[2 * j + 1])));
  }
  sf_l = mx_scale_byte(amax_r);
  sf_c = mx_scale_byte(amax_v);
  const float il = mx_inv_scale(sf_l), ic = mx_inv_scale(sf_c);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    l4[j] = mx_e2m1x8(r + 8 * j, il);
    c4[j] = mx_e2m1x8(v + 8 * j, ic);
  }
  return amax_v;
}

// Destination of an f16mx tensor (device pointers into one buffer, see the layout above).
struct MxOut {
  uint8_t* base;          // byte 0 of the tensor
  long long plane;        // P = rows * ld
  int ld;                 // elements per row (multiple of 64)
  int row_blocks;         // ceil(rows / 128)
};

// store one packed 32-value block: row m, columns [col, col + 32)
__device__ __forceinline__ void mx_store32(const MxOut& o, long long m, int col, const uint32_t (&h)[16],
                                           const uint32_t (&l4)[4], const uint32_t (&c4)[4],
                                           uint32_t sf_l, uint32_t sf_c) {
  const long long e = m * o.ld + col;
  uint4* hp = reinterpret_cast<uint4*>(o.base + 2 * e);
#pragma unroll
  for (int j = 0; j < 4; ++j) hp[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
  *reinterpret_cast<uint4*>(o.base + 2 * o.plane + (e >> 1)) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
  *reinterpret_cast<uint4*>(o.base + 2 * o.plane + (o.plane >> 1) + (e >> 1)) =
      make_uint4(c4[0], c4[1], c4[2], c4[3]);
  const int kb = col >> 5;   // 32-value block index along K
  uint8_t* sf = o.base + 3 * o.plane +
                (static_cast<long long>(kb >> 1) * o.row_blocks + (m >> 7)) * 512 + (m & 31) * 16 +
                ((m >> 5) & 3) * 4 + (kb & 1);
  sf[0] = static_cast<uint8_t>(sf_l);
  sf[2] = static_cast<uint8_t>(sf_c);
}

}  // namespace aclip
