// GPU-side frame ingest: Pillow-exact bicubic resize + centre crop of decoded uint8 frames.
//
// Replaces, for frames that are already decoded (H x W x 3 uint8), the per-frame CPU work of the
// reference's test-mode transform: GroupScale(224, BICUBIC) + GroupCenterCrop(224)
// (/root/reference/src/utils/augmentations.py:25-29; torchvision -> PIL.Image.resize).  The
// arithmetic is Pillow's 8-bit two-pass resample (libImaging/Resample.c): 22-bit fixed-point taps,
// int32 accumulation, rounding and clipping to uint8 after EACH pass -- so the output is bit-exact
// with the reference's.  Tap tables come from the host (data.resize_crop_plan) and cover only the
// 224 x 224 crop; the output is planar (3, 224, 224), which is what patchify reads.
// HBM-bound integer work: one thread per output pixel, three channels, <= ~9 taps.
#include "common.h"

namespace aclip {

namespace {

constexpr int kBits = 22;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[f][r][x][c] for source rows [row0, row0 + rows), crop columns x in [0, size)
__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ src, long long total, int H, int W, int row0, int rows,
                int size, const int* __restrict__ bounds, const int* __restrict__ coeffs, int ksize,
                uint8_t* __restrict__ tmp) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = static_cast<int>(i % size);
    const long long t = i / size;
    const int r = static_cast<int>(t % rows);
    const long long f = t / rows;
    const int x0 = bounds[2 * x], n = bounds[2 * x + 1];
    const uint8_t* p = src + ((f * H + row0 + r) * W + x0) * 3;
    const int* k = coeffs + x * ksize;
    int s0 = 1 << (kBits - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < n; ++j) {
      const int w = __ldg(k + j);
      s0 += p[3 * j + 0] * w;
      s1 += p[3 * j + 1] * w;
      s2 += p[3 * j + 2] * w;
    }
    uint8_t* o = tmp + i * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
  }
}

// out[f][c][y][x] from tmp[f][y0 + j][x][c]
__global__ void __launch_bounds__(256)
resize_v_kernel(const uint8_t* __restrict__ tmp, long long total, int rows, int size,
                const int* __restrict__ bounds, const int* __restrict__ coeffs, int ksize,
                uint8_t* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = static_cast<int>(i % size);
    const long long t = i / size;
    const int y = static_cast<int>(t % size);
    const long long f = t / size;
    const int y0 = bounds[2 * y], n = bounds[2 * y + 1];
    const uint8_t* p = tmp + ((f * rows + y0) * size + x) * 3;
    const int* k = coeffs + y * ksize;
    int s0 = 1 << (kBits - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < n; ++j) {
      const int w = __ldg(k + j);
      const uint8_t* q = p + static_cast<long long>(j) * size * 3;
      s0 += q[0] * w;
      s1 += q[1] * w;
      s2 += q[2] * w;
    }
    const long long plane = static_cast<long long>(size) * size;
    uint8_t* o = out + f * 3 * plane + static_cast<long long>(y) * size + x;
    o[0] = clip8(s0); o[plane] = clip8(s1); o[2 * plane] = clip8(s2);
  }
}

int grid_for(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  return static_cast<int>(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace

}  // namespace aclip

extern "C" int aclip_resize_crop_u8(const uint8_t* frames_hwc, int num_frames, int H, int W, int row0,
                                    int rows, int size, const int* hbounds, const int* hcoeffs,
                                    int hk, const int* vbounds, const int* vcoeffs, int vk,
                                    uint8_t* tmp, uint8_t* out_chw, void* stream_) {
  using namespace aclip;
  ACLIP_REQUIRE(frames_hwc && hbounds && hcoeffs && vbounds && vcoeffs && tmp && out_chw,
                "resize_crop: null pointer");
  ACLIP_REQUIRE(num_frames >= 0 && H > 0 && W > 0 && size > 0 && hk > 0 && vk > 0 && row0 >= 0 &&
                    rows > 0 && row0 + rows <= H,
                "resize_crop: bad geometry (H=%d W=%d row0=%d rows=%d size=%d)", H, W, row0, rows, size);
  if (num_frames == 0) return ACLIP_OK;
  cudaStream_t stream = as_stream(stream_);
  const long long t1 = static_cast<long long>(num_frames) * rows * size;
  const long long t2 = static_cast<long long>(num_frames) * size * size;
  timing_begin(KIND_RESIZE, stream);
  resize_h_kernel<<<grid_for(t1), 256, 0, stream>>>(frames_hwc, t1, H, W, row0, rows, size, hbounds,
                                                    hcoeffs, hk, tmp);
  resize_v_kernel<<<grid_for(t2), 256, 0, stream>>>(tmp, t2, rows, size, vbounds, vcoeffs, vk, out_chw);
  timing_end(KIND_RESIZE, stream, 0.0,
             static_cast<double>(num_frames) * (3.0 * rows * W + 6.0 * rows * size + 3.0 * size * size));
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return ACLIP_OK;
}
