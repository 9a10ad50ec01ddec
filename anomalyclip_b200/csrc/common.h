// Host-side helpers shared by the launchers: error reporting across the C ABI and device info.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/aclip_b200.h"

namespace aclip {

// Thread-local text of the last failure; exported through aclip_last_error().
std::string& last_error();
int fail(int code, const char* fmt, ...);
int sm_count();
// Device address of this device's fp16 saturation counter (split.cuh), nullptr if unavailable.
unsigned int* saturation_counter();
extern std::atomic<long long> g_launches;  // kernels launched by this library
int gemm(const AclipGemmArgs& g, cudaStream_t stream);
int gemm_mx(const AclipGemmArgs& g, cudaStream_t stream);   // passes == 7
struct RowMap;
int split_f32(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out,
              long long plane_stride, cudaStream_t stream);
int encode_f16f8(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out,
                 long long plane_stride, int e_main, int e_res, int e_coarse, cudaStream_t stream);
int encode_f16mx(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out, int e_main,
                 cudaStream_t stream);
int layernorm(const float* x, long long rows, int D, long long ldx, const float* gamma,
              const float* beta, float eps, int mode, float* out_f32, long long ld_f32,
              void* out_split, long long ld_split, long long plane_stride, int out_enc,
              cudaStream_t stream);
int score_head(const float* x1, const float* x2, long long rows, int E, const float* gamma,
               const float* beta, float eps, const float* w, float bias, const float* sim,
               int ld_sim, int ncls, const RowMap& map, float* scores, float* sim_out,
               float* probs_out, const AclipPeerGather* gather, int signal, cudaStream_t stream);
int peer_wait(unsigned int* local_flags, int world, unsigned int epoch, cudaStream_t stream);
int patchify(const void* frames, int is_u8, int B, int R, int P, const float* mean3,
             const float* std3, void* out_split, long long plane_stride, int out_enc,
             cudaStream_t stream);
int cls_rows(float* x, int B, int tokens, int width, const float* cls, const float* pos,
             cudaStream_t stream);
int center_regroup(const float* feats, long long rows, int D, const float* centroid,
                   const RowMap& map, void* out_split, int ld_out, long long plane_stride,
                   cudaStream_t stream);
int vit_attention(const void* qkv_split, long long in_plane_stride, int ld_in, int B, int L,
                  int heads, void* out_split, long long out_plane_stride, int ld_out, int kernel,
                  int out_enc, cudaStream_t stream);
int vit_attention_tc(const void* qkv_split, long long in_plane_stride, int ld_in, int B, int L,
                     int heads, void* out_split, long long out_plane_stride, int ld_out,
                     cudaStream_t stream, int debug = 0, int out_enc = 0);
int axial_attention(const float* qkv, long long sub_videos, int n, int l, int E, int heads,
                    int axis, void* out_split, long long plane_stride, cudaStream_t stream);

#define ACLIP_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::aclip::fail(ACLIP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

#define ACLIP_CHECK_LAUNCH()                                                             \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return ::aclip::fail(ACLIP_ERR_CUDA, "kernel launch failed: %s (%s:%d)",           \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

#define ACLIP_REQUIRE(cond, ...)                                                         \
  do {                                                                                   \
    if (!(cond)) return ::aclip::fail(ACLIP_ERR_INVALID, __VA_ARGS__);                   \
  } while (0)

#define ACLIP_TRY(expr)                                                                  \
  do {                                                                                   \
    int _rc = (expr);                                                                    \
    if (_rc != ACLIP_OK) return _rc;                                                     \
  } while (0)

// Optional per-kernel-kind device timing (cudaEvent pairs around each launch), used by bench.py
// for the roofline numbers.  Off by default; when off the cost is one relaxed atomic load.
enum KernelKind : int {
  KIND_GEMM = 0, KIND_VIT_ATTENTION, KIND_LAYERNORM, KIND_PATCHIFY, KIND_CLS_ROWS,
  KIND_CENTER_REGROUP, KIND_AXIAL_ATTENTION, KIND_SCORE_HEAD, KIND_SPLIT, KIND_RESIZE, KIND_COUNT
};
void timing_begin(int kind, cudaStream_t stream);
void timing_end(int kind, cudaStream_t stream, double flops, double bytes);

// "Do this once per device" latch for cudaFuncSetAttribute (the attribute is per device; a process
// may drive several GPUs, and several host threads may launch concurrently).
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  // true if the caller must (re)apply the setting for the current device; call mark() afterwards
  bool need(int& dev) {
    dev = 0;
    cudaGetDevice(&dev);
    return dev < 0 || dev >= 64 || ((done.load(std::memory_order_acquire) >> dev) & 1ull) == 0;
  }
  void mark(int dev) {
    if (dev >= 0 && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
  }
};

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Launch with programmatic stream serialization (see ptx.cuh: pdl_wait / pdl_launch_dependents);
// Opt-in: ACLIP_PDL=1 in the environment (default: plain stream-ordered launches, see pdl_enabled()).
bool pdl_enabled();
// ACLIP_MX_PDL=1: programmatic launch also for the f16mx kernels (see launch_serial); =ln / =gemm:
// only for layernorm_mx_kernel / only for gemm2mx_tcgen05_kernel (experiments)
bool pdl_mx_enabled(int kind);   // kind: 0 = GEMM, 1 = LayerNorm
template <typename... KArgs, typename... Args>
inline cudaError_t launch_with(bool programmatic, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                               cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = programmatic ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  return launch_with(pdl_enabled(), kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}
// Plain stream-ordered launch (the kernel starts when its predecessor has completed); its own
// griddepcontrol instructions still let the SUCCESSOR start early.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_serial(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args&&... args) {
  return launch_with(pdl_enabled() && pdl_mx_enabled(kind), kernel, grid, block, smem, stream,
                     static_cast<Args&&>(args)...);
}

}  // namespace aclip
