// abi.cu — the parts of the C ABI that are not kernels: error text, version, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dbs {
static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}
}  // namespace dbs

extern "C" const char* dbs_last_error(void) { return dbs::t_error; }
extern "C" int dbs_abi_version(void) { return 1000; }
extern "C" int dbs_built_sm(void) { return 100; }
extern "C" uint64_t dbs_launch_count(void) { return dbs::g_launches.load(std::memory_order_relaxed); }
