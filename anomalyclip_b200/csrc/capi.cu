// Library-level entry points of the C ABI (version, error text, launch counter).
#include "common.h"

extern "C" int aclip_version(void) { return 100; }

extern "C" const char* aclip_last_error(void) { return aclip::last_error().c_str(); }

extern "C" long long aclip_launch_count(void) {
  return aclip::g_launches.load(std::memory_order_relaxed);
}
