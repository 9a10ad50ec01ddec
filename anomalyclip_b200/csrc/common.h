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
extern std::atomic<long long> g_launches;  // kernels launched by this library
int gemm(const AclipGemmArgs& g, cudaStream_t stream);

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

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace aclip
