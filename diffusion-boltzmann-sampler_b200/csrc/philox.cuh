// philox.cuh — Philox4x32-10 counter RNG (Salmon et al., SC'11), host+device.
// The same function, written independently in numpy, lives in oracle/philox.py; both are
// pinned to the Random123 known-answer vectors in tests/test_philox.py.
#pragma once
#include <stdint.h>

namespace dbs {

struct u32x4 {
  uint32_t x, y, z, w;
};

#ifdef __CUDACC__
#define DBS_HD __host__ __device__ __forceinline__
#else
#define DBS_HD inline
#endif

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

DBS_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;  // one IMAD.WIDE.U32 on the device
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}

// 10 rounds; the key schedule (k += W before rounds 2..10) is warp-uniform, so the compiler
// keeps the 20 round keys in uniform registers and a round costs 2 IMAD.WIDE + 2 LOP3.
DBS_HD u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(kPhiloxM0, c.x, hi0, lo0);
    mulhilo(kPhiloxM1, c.z, hi1, lo1);
    u32x4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return c;
}

}  // namespace dbs
