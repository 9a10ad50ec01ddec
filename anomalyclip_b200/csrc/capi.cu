// Library-level entry points of the C ABI: version, error text, launch counter, kernel timing.
#include <cstring>
#include <mutex>
#include <vector>

#include "common.h"

namespace aclip {

namespace {

const char* const kKindNames[KIND_COUNT] = {"gemm_tcgen05", "vit_attention", "layernorm", "patchify",
                                            "cls_rows", "center_regroup", "axial_attention",
                                            "score_head", "split", "resize_crop"};
struct Sample { int kind; cudaEvent_t a, b; double flops, bytes; };
std::atomic<int> g_timing_on{0};
std::mutex g_timing_mu;
std::vector<Sample> g_samples;
std::vector<cudaEvent_t> g_free_events;
constexpr size_t kMaxSamples = 1 << 16;
thread_local cudaEvent_t t_open = nullptr;

cudaEvent_t take_event() {
  if (!g_free_events.empty()) {
    cudaEvent_t e = g_free_events.back();
    g_free_events.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  return e;
}

}  // namespace

void timing_begin(int, cudaStream_t stream) {
  if (g_timing_on.load(std::memory_order_relaxed) == 0) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (g_samples.size() >= kMaxSamples) return;
  t_open = take_event();
  if (t_open != nullptr) cudaEventRecord(t_open, stream);
}

void timing_end(int kind, cudaStream_t stream, double flops, double bytes) {
  if (t_open == nullptr) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  cudaEvent_t b = take_event();
  if (b != nullptr) {
    cudaEventRecord(b, stream);
    g_samples.push_back(Sample{kind, t_open, b, flops, bytes});
  } else {
    g_free_events.push_back(t_open);
  }
  t_open = nullptr;
}

}  // namespace aclip

extern "C" int aclip_version(void) { return 100; }

extern "C" const char* aclip_last_error(void) { return aclip::last_error().c_str(); }

extern "C" long long aclip_launch_count(void) {
  return aclip::g_launches.load(std::memory_order_relaxed);
}

extern "C" long long aclip_note_launches(long long n) {
  // kernels replayed from a captured CUDA graph never pass through the launchers again: the host
  // side that replays the graph reports how many of this library's kernels it holds
  return aclip::g_launches.fetch_add(n > 0 ? n : 0, std::memory_order_relaxed) + (n > 0 ? n : 0);
}

extern "C" int aclip_timing_enable(int on) {
  aclip::g_timing_on.store(on ? 1 : 0, std::memory_order_relaxed);
  return ACLIP_OK;
}

extern "C" int aclip_timing_collect(AclipTimingRow* rows, int max_rows) {
  using namespace aclip;
  if (rows == nullptr || max_rows < KIND_COUNT)
    return fail(ACLIP_ERR_INVALID, "timing_collect: need room for %d rows", (int)KIND_COUNT);
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (int k = 0; k < KIND_COUNT; ++k) {
    std::memset(&rows[k], 0, sizeof(AclipTimingRow));
    std::strncpy(rows[k].name, kKindNames[k], sizeof(rows[k].name) - 1);
  }
  for (const Sample& s : g_samples) {
    float ms = 0.f;
    ACLIP_CUDA_OK(cudaEventSynchronize(s.b));
    ACLIP_CUDA_OK(cudaEventElapsedTime(&ms, s.a, s.b));
    AclipTimingRow& r = rows[s.kind];
    r.launches += 1;
    r.ms += ms;
    r.flops += s.flops;
    r.bytes += s.bytes;
    g_free_events.push_back(s.a);
    g_free_events.push_back(s.b);
  }
  g_samples.clear();
  return KIND_COUNT;
}
