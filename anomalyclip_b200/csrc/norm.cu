// Row-wise normalisation kernels (HBM-bound; one warp per row, 128-bit accesses):
//   layernorm_kernel  nn.LayerNorm / axial-attention ChanLayerNorm -> fp32 and/or split-bf16 rows
//   head_kernel       mean of the reversible halves -> LayerNorm -> Linear(E,1) -> sigmoid,
//                     plus softmax(similarity) * score, written back in the caller's row order
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"
#include "rowmap.cuh"
#include "split.cuh"
#include "mx.cuh"

namespace aclip {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kMaxVec = 6;  // float4 per lane: D <= 768

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(a)) |
         (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

// mode 0: (x - mean) / sqrt(var + eps) * g + b      nn.LayerNorm (clip/model.py:174-180)
// mode 1: (x - mean) / (sqrt(var) + eps) * g + b    ChanLayerNorm of axial_attention (eps on std)
// ENC: encoding of out_split, 0 = bf16 hi/lo planes, 1 = f16f8 activation planes, 2 = fp16 plane
// (split.cuh)
template <int MODE, int ENC>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
layernorm_kernel(const float* __restrict__ x, long long rows, int D, long long ldx,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 float* __restrict__ out_f32, long long ld_f32,
                 __nv_bfloat16* __restrict__ out_split, long long ld_split,
                 long long plane_stride, unsigned int* __restrict__ sat) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int lane = threadIdx.x & 31;
  // Rows are walked from the LAST to the first: the GEMM that produced x wrote its highest rows
  // last, so they are the ones still resident in L2 (x is larger than L2 at the bench's micro-batch),
  // and the rows written last here are the first ones the consuming GEMM reads.
  const long long row = rows - 1 - (static_cast<long long>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5));
  if (row < 0) return;
  const int nvec = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[kMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      v[i] = xr[c];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
  const float var = warp_sum(q) / static_cast<float>(D);
  const float rstd = MODE == 0 ? rsqrtf(var + eps) : 1.0f / (sqrtf(var) + eps);
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 y;
      y.x = v[i].x * rstd * g.x + b.x;
      y.y = v[i].y * rstd * g.y + b.y;
      y.z = v[i].z * rstd * g.z + b.z;
      y.w = v[i].w * rstd * g.w + b.w;
      if (out_f32 != nullptr) reinterpret_cast<float4*>(out_f32 + row * ld_f32)[c] = y;
      if (out_split != nullptr) {
        if (ENC == 0) {
          uint32_t h01, l01, h23, l23;
          split_pack2(y.x, y.y, h01, l01);
          split_pack2(y.z, y.w, h23, l23);
          __nv_bfloat16* dst = out_split + row * ld_split + 4 * c;
          *reinterpret_cast<uint2*>(dst) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(dst + plane_stride) = make_uint2(l01, l23);
        } else if (ENC == 1) {
          f16f8_store4_act(out_split, plane_stride, row * ld_split + 4 * c, y.x, y.y, y.z, y.w);
          amax = sat_track(amax, y.x, y.y, y.z, y.w);
        } else {
          f16_store4_act(out_split, row * ld_split + 4 * c, y.x, y.y, y.z, y.w);
          amax = sat_track(amax, y.x, y.y, y.z, y.w);
        }
      }
    }
  }
  if (ENC != 0) sat_report(sat, amax);
}

// nn.LayerNorm straight into the f16mx encoding (mx.cuh): one warp per row, a lane owns 8
// consecutive columns of every 256-column slab, so FOUR lanes hold one 32-value scale block (two
// shuffle steps for its maxima) and the stores are 16 B (fp16 plane) and 4 B (each e2m1 plane) per
// lane.  D must be a multiple of 256 (<= 768).
__global__ void __launch_bounds__(kWarpsPerCta * 32)
layernorm_mx_kernel(const float* __restrict__ x, long long rows, int D, long long ldx,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                    uint8_t* __restrict__ out, long long ld_out, unsigned int* __restrict__ sat) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = rows - 1 - (static_cast<long long>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5));
  if (row < 0) return;
  constexpr int kSlabs = kMaxVec / 2;   // 256-column slabs
  const int slabs = D >> 8;
  const float* xr = x + row * ldx;
  float v[kSlabs][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kSlabs; ++i) {
    if (i < slabs) {
      const float4 a = *reinterpret_cast<const float4*>(xr + i * 256 + lane * 8);
      const float4 b = *reinterpret_cast<const float4*>(xr + i * 256 + lane * 8 + 4);
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
      v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
      s += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kSlabs; ++i) {
    if (i < slabs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[i][j] -= mean; q += v[i][j] * v[i][j]; }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
  const long long plane = rows * ld_out;
  const long long row_blocks = (rows + 127) >> 7;
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < kSlabs; ++i) {
    if (i < slabs) {
      const int col = i * 256 + lane * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float y[8], r[8];
      uint32_t h[4];
      float mv = 0.f, mr = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (v[i][j] * rstd * gg[j] + bb[j]) * kActScaleMain;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(y[2 * j + 1]), "f"(y[2 * j]));
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
        r[2 * j] = y[2 * j] - f.x;
        r[2 * j + 1] = y[2 * j + 1] - f.y;
        mv = fmaxf(mv, fmaxf(fabsf(y[2 * j]), fabsf(y[2 * j + 1])));
        mr = fmaxf(mr, fmaxf(fabsf(r[2 * j]), fabsf(r[2 * j + 1])));
      }
      amax = fmaxf(amax, mv);
#pragma unroll
      for (int o = 2; o > 0; o >>= 1) {
        mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, o));
        mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o));
      }
      const uint32_t sf_l = mx_scale_byte(mr), sf_c = mx_scale_byte(mv);
      const uint32_t l4 = mx_e2m1x8(r, mx_inv_scale(sf_l)), c4 = mx_e2m1x8(y, mx_inv_scale(sf_c));
      const long long e = row * ld_out + col;
      *reinterpret_cast<uint4*>(out + 2 * e) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint32_t*>(out + 2 * plane + (e >> 1)) = l4;
      *reinterpret_cast<uint32_t*>(out + 2 * plane + (plane >> 1) + (e >> 1)) = c4;
      if ((lane & 3) == 0) {
        const int kb = col >> 5;
        uint8_t* sf = out + 3 * plane + (static_cast<long long>(kb >> 1) * row_blocks + (row >> 7)) * 512 +
                      (row & 31) * 16 + ((row >> 5) & 3) * 4 + (kb & 1);
        sf[0] = static_cast<uint8_t>(sf_l);
        sf[2] = static_cast<uint8_t>(sf_c);
      }
    }
  }
  if (sat != nullptr && !(amax <= 65504.0f)) atomicAdd(sat, 1u);   // also counts NaN
}

struct PeerDev {            // device-side copy of AclipPeerGather for one head launch
  int world, rank, width, signal;
  long long rows_per_rank;
  float* rows[8];
  unsigned int* flags[8];
  unsigned int epoch;
  unsigned int* counter;
};

// One warp per grid row (sub-video order).  E <= 256.
__global__ void __launch_bounds__(kWarpsPerCta * 32)
head_kernel(const float* __restrict__ x1, const float* __restrict__ x2, long long rows, int E,
            const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
            const float* __restrict__ w, float bias, const float* __restrict__ sim, int ld_sim,
            int ncls, RowMap map, float* __restrict__ scores, float* __restrict__ sim_out,
            float* __restrict__ probs_out, PeerDev pg) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (row < rows) {
    const int nvec = E >> 2;
    float4 v[2];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 a = reinterpret_cast<const float4*>(x1 + row * E)[c];
        const float4 b = reinterpret_cast<const float4*>(x2 + row * E)[c];
        // torch.stack(x.chunk(2, dim=1)).mean(dim=0): (a + b) / 2
        v[i] = make_float4((a.x + b.x) * 0.5f, (a.y + b.y) * 0.5f, (a.z + b.z) * 0.5f,
                           (a.w + b.w) * 0.5f);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(s) / static_cast<float>(E);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(E) + eps);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c);
        dot += (v[i].x * rstd * g.x + b.x) * ww.x + (v[i].y * rstd * g.y + b.y) * ww.y +
               (v[i].z * rstd * g.z + b.z) * ww.z + (v[i].w * rstd * g.w + b.w) * ww.w;
      }
    }
    const float z = warp_sum(dot) + bias;
    const float score = 1.0f / (1.0f + expf(-z));  // nn.Sigmoid (classification_head.py:14)
    const long long orow = map.caller_row(row);
    // softmax(similarity, dim=1) * score     (anomaly_clip_module.py:473-477)
    const float sv = lane < ncls ? sim[row * ld_sim + lane] : -INFINITY;
    const float mx = warp_max(sv);
    const float e = lane < ncls ? expf(sv - mx) : 0.f;
    const float den = warp_sum(e);
    const float prob = (e / den) * score;
    if (lane == 0) scores[orow] = score;
    if (lane < ncls) {
      if (sim_out != nullptr) sim_out[orow * ncls + lane] = sv;
      if (probs_out != nullptr) probs_out[orow * ncls + lane] = prob;
    }
    if (pg.world > 0) {
      // fused all-gather: row [score | class_probs] straight into every rank's gathered buffer
      // (peer-mapped memory over NVLink; the local rank is just another entry of the table)
      const float shifted = __shfl_up_sync(0xffffffffu, prob, 1);
      const float val = lane == 0 ? score : shifted;
      if (lane < pg.width) {
        const long long dst = (pg.rank * pg.rows_per_rank + orow) * pg.width + lane;
#pragma unroll 1
        for (int r = 0; r < pg.world; ++r) pg.rows[r][dst] = val;
      }
    }
  }
  if (pg.world > 0 && pg.signal) {
    // publish: every CTA makes its peer stores visible system-wide, the last one raises the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int done = atomicAdd(pg.counter, 1u);
      if (done == gridDim.x - 1) {
        *pg.counter = 0u;  // ready for the next launch
        __threadfence_system();
        for (int r = 0; r < pg.world; ++r)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pg.flags[r] + pg.rank), "r"(pg.epoch)
                       : "memory");
      }
    }
  }
}

// Stream-side wait of the fused gather: spin until every rank's flag has reached `epoch`.
// Bounded: after ~5 s of GPU clocks without progress (a peer died) the kernel gives up and records
// the missing rank in flags[world + r] = 1 instead of hanging the device.
__global__ void peer_wait_kernel(unsigned int* __restrict__ flags, int world, unsigned int epoch,
                                 long long max_cycles) {
  const int r = threadIdx.x;
  if (r < world) {
    const long long t0 = clock64();
    unsigned int v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
      if (static_cast<int>(v - epoch) >= 0) break;
      if (clock64() - t0 > max_cycles) { flags[world + r] = 1u; break; }
      __nanosleep(200);
    }
  }
}

__global__ void peer_signal_kernel(PeerDev pg) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int r = 0; r < pg.world; ++r)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pg.flags[r] + pg.rank), "r"(pg.epoch)
                   : "memory");
  }
}

}  // namespace

int layernorm(const float* x, long long rows, int D, long long ldx, const float* gamma,
              const float* beta, float eps, int mode, float* out_f32, long long ld_f32,
              void* out_split, long long ld_split, long long plane_stride, int out_enc,
              cudaStream_t stream) {
  ACLIP_REQUIRE(x != nullptr && gamma != nullptr && beta != nullptr, "layernorm: null pointer");
  ACLIP_REQUIRE(out_enc == 0 || out_enc == 2 ||
                    (out_enc == 1 && ld_split % 16 == 0 && plane_stride % 16 == 0) ||
                    (out_enc == 3 && mode == 0 && D % 256 == 0 && ld_split % 64 == 0 && ld_split >= D &&
                     plane_stride == rows * ld_split && out_f32 == nullptr),
                "layernorm: out_enc=%d unsupported (f16f8 needs 16-element pitches; f16mx: LayerNorm mode, "
                "D %% 256 == 0, pitch %% 64 == 0, plane_stride = rows * pitch, no fp32 output)", out_enc);
  ACLIP_REQUIRE(D > 0 && D % 4 == 0 && D <= 128 * kMaxVec, "layernorm: D=%d unsupported", D);
  ACLIP_REQUIRE(ldx % 4 == 0 && (out_f32 == nullptr || ld_f32 % 4 == 0) &&
                    (out_split == nullptr || ld_split % 4 == 0),
                "layernorm: pitches must be multiples of 4");
  ACLIP_REQUIRE(mode == 0 || mode == 1, "layernorm: mode must be 0 (LayerNorm) or 1 (ChanLayerNorm)");
  ACLIP_REQUIRE(out_f32 != nullptr || out_split != nullptr, "layernorm: no output");
  if (rows <= 0) return ACLIP_OK;
  const unsigned grid = static_cast<unsigned>((rows + kWarpsPerCta - 1) / kWarpsPerCta);
  auto* os = static_cast<__nv_bfloat16*>(out_split);
  timing_begin(KIND_LAYERNORM, stream);
  unsigned int* sat = (out_split != nullptr && out_enc != 0) ? saturation_counter() : nullptr;
#define ACLIP_LN(MODE, ENC)                                                      \
  ACLIP_CUDA_OK(launch_pdl(layernorm_kernel<MODE, ENC>, dim3(grid), dim3(kWarpsPerCta * 32), 0, stream, x, rows, \
                           D, ldx, gamma, beta, eps, out_f32, ld_f32, os, ld_split, plane_stride, sat))
  if (mode == 0) {
    if (out_enc == 3)
      ACLIP_CUDA_OK(launch_serial(1, layernorm_mx_kernel, dim3(grid), dim3(kWarpsPerCta * 32), 0, stream, x, rows, D, ldx,
                               gamma, beta, eps, static_cast<uint8_t*>(out_split), ld_split, sat));
    else if (out_enc == 2) ACLIP_LN(0, 2);
    else if (out_enc == 1) ACLIP_LN(0, 1);
    else ACLIP_LN(0, 0);
  } else {
    if (out_enc == 2) ACLIP_LN(1, 2);
    else if (out_enc == 1) ACLIP_LN(1, 1);
    else ACLIP_LN(1, 0);
  }
#undef ACLIP_LN
  timing_end(KIND_LAYERNORM, stream, 8.0 * rows * D,
             (double)rows * D * (4.0 + (out_f32 ? 4.0 : 0.0) + (out_split ? (out_enc == 2 ? 2.0 : out_enc == 3 ? 3.06 : 4.0) : 0.0)));
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

int score_head(const float* x1, const float* x2, long long rows, int E, const float* gamma,
               const float* beta, float eps, const float* w, float bias, const float* sim,
               int ld_sim, int ncls, const RowMap& map, float* scores, float* sim_out,
               float* probs_out, const AclipPeerGather* gather, int signal, cudaStream_t stream) {
  ACLIP_REQUIRE(x1 && x2 && gamma && beta && w && scores && (sim || ncls == 0), "score_head: null pointer");
  ACLIP_REQUIRE(E % 4 == 0 && E <= 256, "score_head: E=%d unsupported", E);
  ACLIP_REQUIRE(ncls >= 0 && ncls <= 32 && ld_sim >= ncls, "score_head: ncls=%d unsupported", ncls);
  if (rows <= 0) return ACLIP_OK;
  PeerDev pg{};
  if (gather != nullptr) {
    ACLIP_REQUIRE(gather->world >= 1 && gather->world <= 8 && gather->rank >= 0 &&
                      gather->rank < gather->world && gather->width == ncls + 1 && gather->width <= 32 &&
                      gather->epoch > 0 && gather->counter != nullptr,
                  "score_head: bad peer-gather descriptor");
    pg.world = gather->world; pg.rank = gather->rank; pg.width = gather->width;
    pg.signal = signal; pg.rows_per_rank = gather->rows_per_rank;
    pg.epoch = gather->epoch; pg.counter = gather->counter;
    for (int r = 0; r < gather->world; ++r) {
      ACLIP_REQUIRE(gather->rows[r] && gather->flags[r], "score_head: null peer pointer %d", r);
      pg.rows[r] = gather->rows[r];
      pg.flags[r] = gather->flags[r];
    }
  }
  const unsigned grid = static_cast<unsigned>((rows + kWarpsPerCta - 1) / kWarpsPerCta);
  timing_begin(KIND_SCORE_HEAD, stream);
  ACLIP_CUDA_OK(launch_pdl(head_kernel, dim3(grid), dim3(kWarpsPerCta * 32), 0, stream, x1, x2, rows, E, gamma,
                           beta, eps, w, bias, sim, ld_sim, ncls, map, scores, sim_out, probs_out, pg));
  timing_end(KIND_SCORE_HEAD, stream, 12.0 * rows * E, (double)rows * (8.0 * E + 4.0 * ld_sim + 4.0 + 8.0 * ncls));
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

int peer_wait(unsigned int* local_flags, int world, unsigned int epoch, cudaStream_t stream) {
  ACLIP_REQUIRE(local_flags != nullptr && world >= 1 && world <= 8 && epoch > 0, "peer_wait: bad arguments");
  // bound of the wait in SM cycles (~5 s at 1.9 GHz by default); ACLIP_PEER_WAIT_CYCLES overrides it
  static const long long max_cycles = [] {
    const char* e = getenv("ACLIP_PEER_WAIT_CYCLES");
    const long long v = e != nullptr ? atoll(e) : 0;
    return v > 0 ? v : 10000000000ll;
  }();
  peer_wait_kernel<<<1, 32, 0, stream>>>(local_flags, world, epoch, max_cycles);
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}

}  // namespace aclip

extern "C" int aclip_layernorm(const float* x, long long rows, int D, long long ldx,
                               const float* gamma, const float* beta, float eps, int mode,
                               float* out_f32, long long ld_f32, void* out_split,
                               long long ld_split, long long plane_stride, int out_enc,
                               void* stream) {
  return aclip::layernorm(x, rows, D, ldx, gamma, beta, eps, mode, out_f32, ld_f32, out_split,
                          ld_split, plane_stride, out_enc, aclip::as_stream(stream));
}

extern "C" int aclip_peer_wait(unsigned int* local_flags, int world, unsigned int epoch,
                               void* stream) {
  return aclip::peer_wait(local_flags, world, epoch, aclip::as_stream(stream));
}

extern "C" int aclip_peer_signal(const AclipPeerGather* gather, void* stream) {
  using namespace aclip;
  ACLIP_REQUIRE(gather != nullptr && gather->world >= 1 && gather->world <= 8 && gather->rank >= 0 &&
                    gather->rank < gather->world && gather->epoch > 0,
                "peer_signal: bad descriptor");
  PeerDev pg{};
  pg.world = gather->world; pg.rank = gather->rank; pg.epoch = gather->epoch;
  for (int r = 0; r < gather->world; ++r) {
    ACLIP_REQUIRE(gather->flags[r] != nullptr, "peer_signal: null flag pointer %d", r);
    pg.flags[r] = gather->flags[r];
  }
  peer_signal_kernel<<<1, 32, 0, as_stream(stream)>>>(pg);
  ACLIP_CHECK_LAUNCH();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ACLIP_OK;
}
