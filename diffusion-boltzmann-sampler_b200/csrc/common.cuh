// common.cuh — error plumbing, launch accounting and small device helpers shared by
// every translation unit of libdbs_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/dbs_b200.h"

namespace dbs {

// Thread-local error text behind dbs_last_error().
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Optional per-kernel-class timing (dbs_profile_enable): CUDA events recorded on the launching
// stream around each launch, summed at dbs_profile_read.  Off by default: zero cost.
enum ProfClass {
  kProfConvTC = 0, kProfConvGeneric, kProfGnStats, kProfGnApply, kProfEm, kProfUnetMisc,
  kProfIsingSweep, kProfIsingObs, kProfIsingMisc, kProfLj, kProfCsnLinear, kProfCsnAttention, kProfNumClasses
};
extern bool g_prof_on;
// free-form label attached to the next recorded launch (per-layer dumps: dbs_profile_dump)
void prof_set_label(const char* fmt, ...);
void prof_begin(ProfClass c, cudaStream_t st);
void prof_end(ProfClass c, cudaStream_t st);
struct ProfScope {
  ProfClass c; cudaStream_t st; bool on;
  ProfScope(ProfClass c_, cudaStream_t st_) : c(c_), st(st_), on(g_prof_on) { if (on) prof_begin(c, st); }
  ~ProfScope() { if (on) prof_end(c, st); }
};

#define DBS_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::dbs::set_error(__VA_ARGS__);           \
      return DBS_ERR_INVALID;                  \
    }                                          \
  } while (0)

#define DBS_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      ::dbs::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),      \
                       __FILE__, __LINE__);                                         \
      return DBS_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// Call after every <<<>>>: counts the launch and surfaces launch-time errors.
#define DBS_LAUNCHED()                                                              \
  do {                                                                              \
    ::dbs::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
    cudaError_t e_ = cudaGetLastError();                                            \
    if (e_ != cudaSuccess) {                                                        \
      ::dbs::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),  \
                       __FILE__, __LINE__);                                         \
      return DBS_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dbs
