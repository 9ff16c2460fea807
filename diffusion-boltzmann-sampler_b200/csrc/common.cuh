// common.cuh — error plumbing, launch accounting and small device helpers shared by
// every translation unit of libdbs_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <vector>

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

// Device memory of a plan (weights, repacked weights, tables): bump allocation out of a few large chunks instead of one
// cudaMalloc per tensor.  A plan's ~100 cudaFree calls — each a device-wide synchronisation, run from a Python finaliser at
// whatever moment the garbage collector picks — become one or two.
struct DevArena {
  std::vector<void*> chunks;
  char* cur = nullptr;
  size_t left = 0;
  cudaError_t take(size_t bytes, void** out) {
    bytes = (bytes ? bytes : 16);
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes > left) {
      const size_t chunk = bytes > ((size_t)8 << 20) ? bytes : ((size_t)8 << 20);
      void* c = nullptr;
      const cudaError_t e = cudaMalloc(&c, chunk);
      if (e != cudaSuccess) return e;
      chunks.push_back(c);
      cur = static_cast<char*>(c);
      left = chunk;
    }
    *out = cur;
    cur += bytes;
    left -= bytes;
    return cudaSuccess;
  }
  void release() {
    for (void* c : chunks) cudaFree(c);
    chunks.clear();
    cur = nullptr;
    left = 0;
  }
};

// Programmatic dependent launch (PDL) for the kernels of a sampler step: a kernel launched through launch_chained may start
// while its predecessor in the stream is still draining — the CTAs of the next grid are scheduled as the SMs free up and run
// their data-independent prologue (barrier init, TMEM allocation, tensor-map prefetch, bias and resident-weight loads)
// behind the predecessor's tail; pdl_wait() then blocks until the predecessor grid has completed and its memory is visible.
// Contract: every kernel launched through launch_chained executes pdl_wait() before its first access to anything another
// kernel of the stream writes or reads-then-overwrites, and pdl_launch_dependents() at its top.  Both are no-ops for a
// kernel launched the ordinary way.  The launch attribute is set only with DBS_PDL=1: measured a gain below ~128 samples
// per step and a loss above (abi.cu, pdl_enabled).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

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
