// HBM-bound helper kernels: fp32 -> split-bf16 conversion, LayerNorm variants, frame patchify.
// All of them stream rows with 128-bit accesses; grids are sized as a multiple of the SM count.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"
#include "split.cuh"
#include "rowmap.cuh"

namespace aclip {

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  split_pack2(a, b, hi, lo);
}

// ------------------------------------------------------------------------------ saturation guard
// Per-device counter of threads that stored a value beyond the fp16 range of the activation
// encodings (split.cuh: sat_track / sat_report).  Kernels of other translation units receive its
// device address as a parameter.
__device__ unsigned int g_f16_saturations = 0;

unsigned int* saturation_counter() {
  static std::atomic<unsigned int*> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  unsigned int* ptr = cached[dev].load(std::memory_order_acquire);
  if (ptr == nullptr) {
    void* sym = nullptr;
    if (cudaGetSymbolAddress(&sym, g_f16_saturations) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    ptr = static_cast<unsigned int*>(sym);
    cached[dev].store(ptr, std::memory_order_release);
  }
  return ptr;
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
  timing_begin(KIND_SPLIT, stream);
  split_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      in, rows, cols, ld_in, static_cast<__nv_bfloat16*>(out), ld_out, plane_stride, vec_ok);
  timing_end(KIND_SPLIT, stream, 0.0, (double)rows * (4.0 * cols + 4.0 * ld_out));
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}


// ------------------------------------------------------------------------------ f16f8 encode
// fp32 [rows][cols] -> f16f8 planes (split.cuh): used to pack weights (per-tensor exponent) and, in
// tests, activations.  Columns cols..ld_out-1 are zero-filled.
__global__ void __launch_bounds__(256)
encode_f16f8_kernel(const float* __restrict__ in, long long rows, int cols, int ld_in,
                    uint8_t* __restrict__ out, int ld_out, long long plane_stride, float s_main,
                    float s_res, float s_coarse, bool vec_ok) {
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
    uint2 h0, h1;
    uint32_t l0, l1, c0, c1;
    f16f8_pack4(v[0], v[1], v[2], v[3], s_main, s_res, s_coarse, h0, l0, c0);
    f16f8_pack4(v[4], v[5], v[6], v[7], s_main, s_res, s_coarse, h1, l1, c1);
    const long long off = r * ld_out + c;
    *reinterpret_cast<uint4*>(out + 2 * off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
    *reinterpret_cast<uint2*>(out + 2 * plane_stride + off) = make_uint2(l0, l1);
    *reinterpret_cast<uint2*>(out + 3 * plane_stride + off) = make_uint2(c0, c1);
  }
}

int encode_f16f8(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out,
                 long long plane_stride, int e_main, int e_res, int e_coarse, cudaStream_t stream) {
  ACLIP_REQUIRE(in != nullptr && out != nullptr, "encode_f16f8: null pointer");
  ACLIP_REQUIRE(rows >= 0 && cols > 0 && ld_in >= cols, "encode_f16f8: bad shape");
  ACLIP_REQUIRE(ld_out % 16 == 0 && ld_out >= cols,
                "encode_f16f8: ld_out=%d must be a multiple of 16 >= cols", ld_out);
  ACLIP_REQUIRE(plane_stride % 16 == 0 && plane_stride >= rows * ld_out,
                "encode_f16f8: plane_stride too small or unaligned");
  ACLIP_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "encode_f16f8: output must be 16-byte aligned");
  ACLIP_REQUIRE(e_main >= -30 && e_main <= 30 && e_res >= 0 && e_res <= 12 && e_coarse >= -40 && e_coarse <= 40,
                "encode_f16f8: exponents out of range");
  if (rows == 0) return ACLIP_OK;
  const bool vec_ok = (ld_in % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  const long long total = rows * (ld_out >> 3);
  timing_begin(KIND_SPLIT, stream);
  encode_f16f8_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      in, rows, cols, ld_in, static_cast<uint8_t*>(out), ld_out, plane_stride, exp2f((float)e_main),
      exp2f((float)e_res), exp2f((float)e_coarse), vec_ok);
  timing_end(KIND_SPLIT, stream, 0.0, (double)rows * (4.0 * cols + 4.0 * ld_out));
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

// ------------------------------------------------------------------------------ patchify
// Frames (B,3,R,R) -> im2col rows for the patch-embedding GEMM (clip/model.py:246-252,267):
// row = b*G*G + gy*G + gx, column k = c*P*P + py*P + px (the order of conv1.weight.reshape(width,-1)).
// U8 input is normalised on the fly exactly like torchvision's ToTensor + Normalize
// (src/utils/augmentations.py:21-34): ((v / 255) - mean) / std in fp32.
struct Norm3 { float mean[3]; float std[3]; };

template <bool U8, int ENC>
__global__ void __launch_bounds__(256)
patchify_kernel(const void* __restrict__ frames, int B, int R, int P, Norm3 nrm,
                __nv_bfloat16* __restrict__ out, long long plane_stride,
                unsigned int* __restrict__ sat) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int G = R / P;
  const int K = 3 * P * P;
  const int groups_per_row = K >> 3;
  const long long total = static_cast<long long>(B) * G * G * groups_per_row;
  float amax = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / groups_per_row;
    const int k = static_cast<int>(i - row * groups_per_row) << 3;
    const int c = k / (P * P);
    const int py = (k - c * P * P) / P;
    const int px = k - c * P * P - py * P;
    const int b = static_cast<int>(row / (G * G));
    const int cell = static_cast<int>(row - static_cast<long long>(b) * G * G);
    const int gy = cell / G, gx = cell - gy * G;
    const long long src = ((static_cast<long long>(b) * 3 + c) * R + gy * P + py) * R + gx * P + px;
    float v[8];
    if (U8) {
      const uint2 raw = *reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(frames) + src);
      const uint32_t w[2] = {raw.x, raw.y};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pix = static_cast<float>((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
        v[j] = __fdiv_rn(__fsub_rn(__fdiv_rn(pix, 255.0f), nrm.mean[c]), nrm.std[c]);
      }
    } else {
      const float4* s4 = reinterpret_cast<const float4*>(static_cast<const float*>(frames) + src);
      const float4 a = s4[0], bb = s4[1];
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      v[4] = bb.x; v[5] = bb.y; v[6] = bb.z; v[7] = bb.w;
    }
    if (ENC != 0) amax = sat_track(sat_track(amax, v[0], v[1], v[2], v[3]), v[4], v[5], v[6], v[7]);
    if (ENC == 1) {
      f16f8_store4_act(out, plane_stride, row * K + k, v[0], v[1], v[2], v[3]);
      f16f8_store4_act(out, plane_stride, row * K + k + 4, v[4], v[5], v[6], v[7]);
      continue;
    }
    if (ENC == 2) {
      uint2 h0, h1;
      f16_pack4(v[0], v[1], v[2], v[3], h0);
      f16_pack4(v[4], v[5], v[6], v[7], h1);
      *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + 2 * (row * K + k)) =
          make_uint4(h0.x, h0.y, h1.x, h1.y);
      continue;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
    __nv_bfloat16* dst = out + row * K + k;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + plane_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (ENC != 0) sat_report(sat, amax);
}

int patchify(const void* frames, int is_u8, int B, int R, int P, const float* mean3,
             const float* std3, void* out_split, long long plane_stride, int out_enc,
             cudaStream_t stream) {
  ACLIP_REQUIRE(frames != nullptr && out_split != nullptr, "patchify: null pointer");
  ACLIP_REQUIRE(out_enc == 0 || out_enc == 2 ||
                    (out_enc == 1 && plane_stride % 16 == 0 && (3 * P * P) % 16 == 0),
                "patchify: out_enc=%d unsupported", out_enc);
  ACLIP_REQUIRE(B > 0 && P % 8 == 0 && R % P == 0, "patchify: B=%d R=%d P=%d unsupported", B, R, P);
  ACLIP_REQUIRE((reinterpret_cast<uintptr_t>(frames) & 15) == 0, "patchify: frames must be 16-byte aligned");
  Norm3 nrm{{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
  if (is_u8) {
    ACLIP_REQUIRE(mean3 != nullptr && std3 != nullptr, "patchify: u8 frames need mean/std");
    for (int c = 0; c < 3; ++c) { nrm.mean[c] = mean3[c]; nrm.std[c] = std3[c]; }
  }
  const int G = R / P;
  const long long total = static_cast<long long>(B) * G * G * (3 * P * P / 8);
  auto* o = static_cast<__nv_bfloat16*>(out_split);
  timing_begin(KIND_PATCHIFY, stream);
  const int grid = grid_for(total, 256);
  unsigned int* sat = out_enc != 0 ? saturation_counter() : nullptr;
#define ACLIP_PATCHIFY(U8, ENC) \
  ACLIP_CUDA_OK(launch_pdl(patchify_kernel<U8, ENC>, dim3(grid), dim3(256), 0, stream, frames, B, R, P, nrm, o, \
                           plane_stride, sat))
  if (is_u8) {
    if (out_enc == 2) ACLIP_PATCHIFY(true, 2);
    else if (out_enc == 1) ACLIP_PATCHIFY(true, 1);
    else ACLIP_PATCHIFY(true, 0);
  } else {
    if (out_enc == 2) ACLIP_PATCHIFY(false, 2);
    else if (out_enc == 1) ACLIP_PATCHIFY(false, 1);
    else ACLIP_PATCHIFY(false, 0);
  }
#undef ACLIP_PATCHIFY
  timing_end(KIND_PATCHIFY, stream, 0.0,
             (double)B * 3 * R * R * ((is_u8 ? 1.0 : 4.0) + (out_enc == 2 ? 2.0 : 4.0)));
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

// ------------------------------------------------------------------------------ CLS rows
// x[b*tokens + 0, :] = class_embedding + positional_embedding[0]   (clip/model.py:270-278)
__global__ void cls_rows_kernel(float* __restrict__ x, int B, int tokens, int width,
                                const float* __restrict__ cls, const float* __restrict__ pos) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * width) return;
  const int b = i / width, c = i - b * width;
  x[static_cast<long long>(b) * tokens * width + c] = cls[c] + pos[c];
}

int cls_rows(float* x, int B, int tokens, int width, const float* cls, const float* pos,
             cudaStream_t stream) {
  ACLIP_REQUIRE(x && cls && pos && B > 0, "cls_rows: bad arguments");
  timing_begin(KIND_CLS_ROWS, stream);
  ACLIP_CUDA_OK(launch_pdl(cls_rows_kernel, dim3((B * width + 255) / 256), dim3(256), 0, stream, x, B, tokens,
                           width, cls, pos));
  timing_end(KIND_CLS_ROWS, stream, 0.0, 4.0 * B * width);
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

// ------------------------------------------------------------------------------ centre + regroup
// Feature rows (caller order "(b n s l)") -> (x - ncentroid) as split-bf16 rows in sub-video
// order, columns [0, D) of a buffer whose pitch may leave room for the similarity columns.
// (selector_model.py:54 and anomaly_clip.py:143 both subtract the same centroid; done once here.)
__global__ void __launch_bounds__(256)
center_regroup_kernel(const float* __restrict__ feats, long long rows, int D,
                      const float* __restrict__ centroid, RowMap map,
                      __nv_bfloat16* __restrict__ out, int ld_out, long long plane_stride) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int groups_per_row = D >> 3;
  const long long total = rows * groups_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups_per_row;
    const int c = static_cast<int>(i - r * groups_per_row) << 3;
    const float* src = feats + map.caller_row(r) * D + c;
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 b = *reinterpret_cast<const float4*>(src + 4);
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(centroid + c));
    const float4 m1 = __ldg(reinterpret_cast<const float4*>(centroid + c + 4));
    uint32_t hi[4], lo[4];
    split2(a.x - m0.x, a.y - m0.y, hi[0], lo[0]);
    split2(a.z - m0.z, a.w - m0.w, hi[1], lo[1]);
    split2(b.x - m1.x, b.y - m1.y, hi[2], lo[2]);
    split2(b.z - m1.z, b.w - m1.w, hi[3], lo[3]);
    __nv_bfloat16* dst = out + r * ld_out + c;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + plane_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

int center_regroup(const float* feats, long long rows, int D, const float* centroid,
                   const RowMap& map, void* out_split, int ld_out, long long plane_stride,
                   cudaStream_t stream) {
  ACLIP_REQUIRE(feats && centroid && out_split, "center_regroup: null pointer");
  ACLIP_REQUIRE(D % 8 == 0 && ld_out % 8 == 0 && ld_out >= D, "center_regroup: D=%d ld=%d unsupported", D, ld_out);
  ACLIP_REQUIRE((reinterpret_cast<uintptr_t>(feats) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(centroid) & 15) == 0,
                "center_regroup: inputs must be 16-byte aligned");
  if (rows <= 0) return ACLIP_OK;
  timing_begin(KIND_CENTER_REGROUP, stream);
  ACLIP_CUDA_OK(launch_pdl(center_regroup_kernel, dim3(grid_for(rows * (D >> 3), 256)), dim3(256), 0, stream,
                           feats, rows, D, centroid, map, static_cast<__nv_bfloat16*>(out_split), ld_out,
                           plane_stride));
  timing_end(KIND_CENTER_REGROUP, stream, (double)rows * D, (double)rows * D * 8.0);
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

extern "C" int aclip_encode_f16f8(const float* in, long long rows, int cols, int ld_in, void* out,
                                  int ld_out, long long plane_stride, int e_main, int e_res,
                                  int e_coarse, void* stream) {
  return aclip::encode_f16f8(in, rows, cols, ld_in, out, ld_out, plane_stride, e_main, e_res,
                             e_coarse, aclip::as_stream(stream));
}

extern "C" int aclip_center_regroup(const float* feats, long long rows, int D,
                                    const float* centroid, int num_segments, int segment_size,
                                    int seg_length, void* out_split, int ld_out,
                                    long long plane_stride, void* stream) {
  if (num_segments < 1 || segment_size < 1 || seg_length < 1)
    return aclip::fail(ACLIP_ERR_INVALID, "center_regroup: n, s, l must be >= 1");
  const aclip::RowMap map{num_segments, segment_size, seg_length, 0};
  return aclip::center_regroup(feats, rows, D, centroid, map, out_split, ld_out, plane_stride,
                               aclip::as_stream(stream));
}

extern "C" int aclip_patchify(const void* frames, int frames_are_u8, int B, int R, int P,
                              const float* mean3_host, const float* std3_host, void* out_split,
                              long long plane_stride, int out_enc, void* stream) {
  return aclip::patchify(frames, frames_are_u8, B, R, P, mean3_host, std3_host, out_split,
                         plane_stride, out_enc, aclip::as_stream(stream));
}

extern "C" long long aclip_saturation_count(int reset) {
  unsigned int* ptr = aclip::saturation_counter();
  if (ptr == nullptr) return aclip::fail(ACLIP_ERR_CUDA, "saturation_count: no counter on this device");
  unsigned int v = 0;
  if (cudaMemcpy(&v, ptr, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess)
    return aclip::fail(ACLIP_ERR_CUDA, "saturation_count: cudaMemcpy failed");
  if (reset != 0 && v != 0) {
    const unsigned int zero = 0;
    if (cudaMemcpy(ptr, &zero, sizeof(zero), cudaMemcpyHostToDevice) != cudaSuccess)
      return aclip::fail(ACLIP_ERR_CUDA, "saturation_count: reset failed");
  }
  return static_cast<long long>(v);
}
