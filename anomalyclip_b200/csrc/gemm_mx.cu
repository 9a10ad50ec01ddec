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

}  // namespace aclip

extern "C" long long aclip_f16mx_bytes(long long rows, int ld) {
  if (rows < 0 || ld <= 0 || ld % 64 != 0) return -1;
  return 3 * rows * ld + (long long)(ld / 64) * ((rows + 127) / 128) * 512;
}

extern "C" int aclip_encode_f16mx(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out,
                                  int e_main, void* stream) {
  return aclip::encode_f16mx(in, rows, cols, ld_in, out, ld_out, e_main, aclip::as_stream(stream));
}
