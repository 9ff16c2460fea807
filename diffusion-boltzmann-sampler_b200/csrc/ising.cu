// ising.cu — bit-packed checkerboard Metropolis and lattice observables for sm_100a.
//
// Replaces (reference file:line, /root/reference/backend/ml):
//   MetropolisHastings.step / sweep            samplers/mcmc.py:26-60
//   IsingModel.local_energy_diff               systems/ising.py:68-91
//   IsingModel.energy / magnetization          systems/ising.py:47-66, 93-102
//   IsingModel.random_configuration            systems/ising.py:36-45
//   pair_correlation                           analysis/correlation.py:8-47
//
// State layout in HBM: uint32 words[n_chains][L][W], W = ceil(L/32), 1 bit per spin.
// One thread owns one word (32 columns of one row) for the whole multi-sweep launch; the
// lattice lives in shared memory (double-buffered, one __syncthreads per half-sweep) and
// touches HBM exactly twice per launch.  Per active site the work is one Philox draw
// (2.5 rounds amortised), a 4-bit class extract, one 8-byte table LDS and two compares.
#include <stdlib.h>

#include "common.cuh"
#include "philox.cuh"

namespace dbs {
namespace {

constexpr int kMaxL = 128;

// even L: the bit-packed checkerboard kernel; odd L (>= 5): the three-colour kernel below (an odd torus is not bipartite)
__host__ __device__ inline bool supported_L(int L) { return L >= 4 && L <= kMaxL && ((L & 1) == 0 || L >= 5); }

struct SweepArgs {
  uint32_t* bits;
  const uint32_t* thr;          // [n_tables][16][2]
  const int32_t* table_index;   // [n_chains] or null
  const uint32_t* uniforms;     // [n_sweeps][n_chains][L][L] or null
  int64_t n_chains;
  int64_t chain_offset;
  int64_t chain_stride;         // global id of chain c = chain_offset + c * chain_stride (striped sharding)
  int64_t sweep_offset;
  uint32_t seed_lo, seed_hi;
  int L;
  int n_sweeps;
  int cpb;                      // chains per block
  uint32_t rk[20];              // Philox round keys {seed_lo + r*W0, seed_hi + r*W1}, r = 0..9: read straight from the
                                // kernel-parameter constant bank by the LOP3 of each round (no key-schedule adds in the loop)
  uint32_t one;                 // = 1, a run-time multiplier that keeps the compares on IMAD.WIDE (FMA pipe)
  int antiferro;                // zero-field mode: count disagreeing neighbours (J < 0)
};

// valid-bit mask of word w in a row of L bits
__device__ __forceinline__ int valid_bits(int L, int w) {
  int vb = L - 32 * w;
  return vb > 32 ? 32 : vb;
}
__device__ __forceinline__ uint32_t low_mask(int vb) { return vb >= 32 ? 0xffffffffu : ((1u << vb) - 1u); }

// accept mask of one draw: all ones iff r < thr, computed as the high word of r + (2^64 - thr) with ONE
// IMAD.WIDE.U32 on the FMA pipe (the ALU pipe is the binding one here: ncu, profiles/README.md).
// c = 0xFFFFFFFF00000000 encodes "always" (dE <= 0), c = 0 "never" (thr = 0).
__device__ __forceinline__ uint32_t accept_mask(uint32_t r, uint32_t one, uint64_t c) {
  uint64_t t;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(r), "r"(one), "l"(c));
  return (uint32_t)(t >> 32);
}
__device__ __forceinline__ uint64_t threshold_constant(uint32_t lo, uint32_t hi) {
  return hi ? 0xFFFFFFFF00000000ull : (0ull - (uint64_t)lo);
}

// b |= bit where r < t, as ISETP + one predicated LOP3 (the compiler's own choice is ISETP + SEL + a shared LOP3)
__device__ __forceinline__ void or_if_less(uint32_t& b, uint32_t r, uint32_t t, uint32_t bit) {
  asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(b) : "r"(r), "r"(t), "r"(bit));
}

// Philox4x32-10 with the ten round-key pairs given (same function as philox4x32_10: key schedule precomputed on the host)
__device__ __forceinline__ u32x4 philox4x32_10_rk(u32x4 c, const uint32_t (&rk)[20]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(kPhiloxM0, c.x, hi0, lo0);
    mulhilo(kPhiloxM1, c.z, hi1, lo1);
    u32x4 n;
    n.x = hi1 ^ c.y ^ rk[2 * r];
    n.y = lo1;
    n.z = hi0 ^ c.w ^ rk[2 * r + 1];
    n.w = lo0;
    c = n;
  }
  return c;
}

// kMode 0: any (J, h): per-site class k = 8*s + #up-neighbours, 16-entry table lookup.
// kMode 1: h == 0: the class depends only on the number of AGREEING neighbours (disagreeing for J < 0),
//          counted for all 32 sites of a word at once with a bit-sliced adder; only the two uphill
//          classes (3 and 4 agreeing: dE = 4|J|, 8|J|) need a draw comparison, done against two per-chain
//          constants: no table lookup, no per-site class extraction.
#ifndef DBS_ISING_MINB
#define DBS_ISING_MINB 8    // resident CTAs per SM the register cap is set for (build knob; sweep below)
#endif
template <int LT, bool kInject, int kMode>
// 16/32/64-wide lattices run 128-thread CTAs.  Register cap, A/B-tested on a B200 with the final loop (64x64, 8192
// chains): 12 CTAs/SM (40 registers: the compiler recomputes ~25 loop-invariant address / mask words per half-sweep)
// 1225 Gflips/s, 10 CTAs (48) 1272, 8 CTAs (64 registers, everything hoisted) 1337, 7 / 6 CTAs 1332; 16x16 and 32x32
// lattices gain the same way (1007 -> 1075, 1232 -> 1311).  The earlier loop (key-schedule adds inside) preferred 12
// (930 vs 907 Gflips/s uncapped; a 256-thread CTA was slower, 816).
__global__ void __launch_bounds__((LT == 16 || LT == 32 || LT == 64) ? 128 : 512, (LT == 16 || LT == 32 || LT == 64) ? DBS_ISING_MINB : 1)
ising_sweep_kernel(SweepArgs a) {
  const int L = LT ? LT : a.L;
  const int W = (L + 31) >> 5;
  const int tpc = L * W;
  const int cpb = a.cpb;
  extern __shared__ uint32_t smem[];
  uint32_t* buf0 = smem;
  uint32_t* buf1 = smem + cpb * tpc;
  uint64_t* tbl = reinterpret_cast<uint64_t*>(smem + 2 * cpb * tpc + ((2 * cpb * tpc) & 1));

  const int tid = threadIdx.x;
  const int lc = tid / tpc;
  const int r = tid - lc * tpc;
  const int i = r / W;
  const int w = r - i * W;
  const int64_t chain = (int64_t)blockIdx.x * cpb + lc;
  const bool live = (lc < cpb) && (chain < a.n_chains);

  if (tid < cpb * 16) {
    const int lc2 = tid >> 4, k = tid & 15;
    const int64_t c2 = (int64_t)blockIdx.x * cpb + lc2;
    uint64_t e = 0;
    if (c2 < a.n_chains) {
      const int ti = a.table_index ? a.table_index[c2] : 0;
      e = threshold_constant(a.thr[(ti * 16 + k) * 2 + 0], a.thr[(ti * 16 + k) * 2 + 1]);
    }
    tbl[tid] = e;
  }

  uint32_t self = 0;
  if (live) {
    self = a.bits[chain * tpc + r];
    buf0[lc * tpc + r] = self;
  }
  __syncthreads();

  const int vb = valid_bits(L, w);
  const uint32_t vmask = low_mask(vb);
  const int base = lc * tpc;
  const int row_up = (i == 0 ? L - 1 : i - 1) * W + w;
  const int row_dn = (i == L - 1 ? 0 : i + 1) * W + w;
  const int wl = (w == 0 ? W - 1 : w - 1);
  const int wr = (w == W - 1 ? 0 : w + 1);
  const int top_prev = valid_bits(L, wl) - 1;  // bit index of the last column of the left word
  const int gpr = (L + 7) >> 3;                // Philox calls per row and colour
  const uint64_t* mytbl = tbl + lc * 16;
  // zero-field constants: classes (s=1, 3 or 4 up-neighbours) for J > 0, (s=1, 1 or 0) for J < 0 — same
  // dE as their s=0 mirror images
  // the two thresholds as plain 32-bit compare operands plus an all-ones mask where the threshold rounds to "always"
  // (c = 2^64 - thr, or the always / never codes); only these four words stay live in the sweep loop
  uint32_t t4 = 0, t8 = 0, am4 = 0, am8 = 0;
  if (kMode == 1 && lc < cpb) {
    const uint64_t c4 = mytbl[a.antiferro ? 8 + 1 : 8 + 3];
    const uint64_t c8 = mytbl[a.antiferro ? 8 + 0 : 8 + 4];
    const bool alw4 = (c4 >> 32) == 0xFFFFFFFFull && (uint32_t)c4 == 0u, alw8 = (c8 >> 32) == 0xFFFFFFFFull && (uint32_t)c8 == 0u;
    t4 = alw4 ? 0u : (uint32_t)(0ull - c4); t8 = alw8 ? 0u : (uint32_t)(0ull - c8);
    am4 = alw4 ? 0xffffffffu : 0u; am8 = alw8 ? 0xffffffffu : 0u;
  }
  const uint32_t one = a.one;
  (void)t4; (void)t8; (void)am4; (void)am8; (void)one;
  const uint32_t chain_ctr = (uint32_t)(a.chain_offset + chain * a.chain_stride);

  uint32_t hs = 0;  // half-sweep counter inside this launch
  for (int s = 0; s < a.n_sweeps; ++s) {
#pragma unroll 1
    for (int colour = 0; colour < 2; ++colour, ++hs) {
      const uint32_t* cur = (hs & 1) ? buf1 : buf0;
      uint32_t* nxt = (hs & 1) ? buf0 : buf1;
      if (live) {
        const uint32_t up = cur[base + row_up];
        const uint32_t dn = cur[base + row_dn];
        const uint32_t lw = (W == 1) ? self : cur[base + i * W + wl];
        const uint32_t rw = (W == 1) ? self : cur[base + i * W + wr];
        const uint32_t lf = ((self << 1) | ((lw >> top_prev) & 1u)) & vmask;
        const uint32_t rt = ((self >> 1) | ((rw & 1u) << (vb - 1))) & vmask;
        const int par = (i + colour) & 1;  // column parity of the active sites of this row

        uint32_t KA = 0, KB = 0, is3 = 0, is4 = 0;
        if (kMode == 0) {
          // class nibbles k = 8*s + (#up neighbours), active sites moved to even bit positions
          const uint32_t m1 = 0x11111111u;
          const uint32_t s_ = self >> par, u_ = up >> par, d_ = dn >> par, l_ = lf >> par, r_ = rt >> par;
          KA = (u_ & m1) + (d_ & m1) + (l_ & m1) + (r_ & m1) + ((s_ & m1) << 3);
          KB = ((u_ >> 2) & m1) + ((d_ >> 2) & m1) + ((l_ >> 2) & m1) + ((r_ >> 2) & m1) + (((s_ >> 2) & m1) << 3);
        } else {
          // bit-sliced count of agreeing (J > 0) / disagreeing (J < 0) neighbours for all 32 sites
          const uint32_t fx = a.antiferro ? 0u : 0xffffffffu;
          const uint32_t xu = (self ^ up) ^ fx, xd = (self ^ dn) ^ fx, xl = (self ^ lf) ^ fx, xr = (self ^ rt) ^ fx;
          const uint32_t s1 = xu ^ xd, c1 = xu & xd, s2 = xl ^ xr, c2 = xl & xr;
          is4 = c1 & c2;                               // all four
          is3 = (s1 ^ s2) & (c1 ^ c2 ^ (s1 & s2));     // count = 0b011
        }

        uint32_t flips = 0, b4 = 0, b8 = 0;  // bit 2m = active site m (shifted domain)
        const uint32_t ctr_y = 2u * (uint32_t)(a.sweep_offset + s) + (uint32_t)colour;
#pragma unroll
        for (int call = 0; call < 4; ++call) {
          if (8 * call < vb) {
            uint32_t rr[4];
            if (kInject) {
#pragma unroll
              for (int d = 0; d < 4; ++d) {
                const int j = 32 * w + par + 2 * (4 * call + d);
                rr[d] = (j < L) ? a.uniforms[(((int64_t)s * a.n_chains + chain) * L + i) * L + j] : 0u;
              }
            } else {
              u32x4 c;
              c.x = chain_ctr;
              c.y = ctr_y;
              c.z = (uint32_t)(i * gpr + w * 4 + call);
              c.w = DBS_STREAM_SWEEP;
              const u32x4 o = philox4x32_10_rk(c, a.rk);
              rr[0] = o.x; rr[1] = o.y; rr[2] = o.z; rr[3] = o.w;
            }
#pragma unroll
            for (int d = 0; d < 4; ++d) {
              const int m = 4 * call + d;
              if (kMode == 0) {
                const uint32_t K = (m & 1) ? KB : KA;
                const uint32_t k = (K >> (4 * (m >> 1))) & 15u;
                flips |= accept_mask(rr[d], one, mytbl[k]) & (1u << (2 * m));
              } else {
                // plain compares here: with Philox already filling the FMA pipe with IMAD.WIDE, two more per
                // site measured slower (946 vs 1054 Gflips/s at 64x64) than ISETP + predicated LOP on the ALU pipe
                or_if_less(b4, rr[d], t4, 1u << (2 * m));
                or_if_less(b8, rr[d], t8, 1u << (2 * m));
              }
            }
          }
        }
        if (kMode == 1) {
          // downhill / flat moves (count <= 2) always flip; counts 3 and 4 flip where the draw allowed it
          // (am4 / am8: thresholds that round to "always", p == 1.0 in float64)
          const uint32_t up3 = is3 >> par, up4 = is4 >> par;
          flips = ~(up3 | up4) | (up3 & (b4 | am4)) | (up4 & (b8 | am8));
        }
        const uint32_t amask = (par ? 0xAAAAAAAAu : 0x55555555u) & vmask;
        self ^= (flips << par) & amask;
        nxt[base + r] = self;
      }
      __syncthreads();
    }
  }
  if (live) a.bits[chain * tpc + r] = self;
}

template <int LT, bool kInject, int kMode>
int launch_sweep_t(const SweepArgs& a, cudaStream_t st) {
  const int L = a.L, W = (L + 31) / 32, tpc = L * W;
  const int cpb = a.cpb;
  int threads = cpb * tpc;
  threads = (threads + 31) / 32 * 32;
  if (threads < cpb * 16) threads = cpb * 16;
  const size_t smem = (size_t)(2 * cpb * tpc + 2) * 4 + (size_t)cpb * 16 * 8;  // lattices x2 + 16 x uint64 per chain
  const int64_t blocks = (a.n_chains + cpb - 1) / cpb;
  ProfScope prof(kProfIsingSweep, st);
  ising_sweep_kernel<LT, kInject, kMode><<<(unsigned)blocks, threads, smem, st>>>(a);
  DBS_LAUNCHED();
  return DBS_OK;
}

template <bool kInject, int kMode>
int launch_sweep(const SweepArgs& a, cudaStream_t st) {
  switch (a.L) {
    case 8: return launch_sweep_t<8, kInject, kMode>(a, st);
    case 16: return launch_sweep_t<16, kInject, kMode>(a, st);
    case 32: return launch_sweep_t<32, kInject, kMode>(a, st);
    case 64: return launch_sweep_t<64, kInject, kMode>(a, st);
    case 128: return launch_sweep_t<128, kInject, kMode>(a, st);
    default: return launch_sweep_t<0, kInject, kMode>(a, st);
  }
}

// ---------------------------------------------------------------------------------------------
// Odd L.  A torus with an odd side is not two-colourable: sites (i, L-1) and (i, 0) are neighbours with the same
// (i + j) parity.  The cycle of L sites is three-coloured c(k) = k & 1 for k < L-1, c(L-1) = 2, and
// colour(i, j) = (c(i) + c(j)) mod 3 is a proper colouring of the torus; one sweep = colours 0, 1, 2 in turn, every
// site attempted once with the reference's acceptance rule.  Odd lattices are an API corner (lattice_size is any
// integer in 8..128, routes/sampling.py:26), not a BASELINE configuration: one CTA per chain, one spin per byte in
// shared memory, a Philox call per group of four consecutive sites, the same {lo, hi} class table.
__device__ __forceinline__ int cycle_colour(int k, int L) { return k == L - 1 ? 2 : (k & 1); }

template <bool kInject>
__global__ void __launch_bounds__(256)
ising_sweep3_kernel(SweepArgs a) {
  const int L = a.L, N = L * L;
  const int W = (L + 31) >> 5, tpc = L * W;
  extern __shared__ uint32_t smem[];
  uint64_t* tbl = reinterpret_cast<uint64_t*>(smem);             // 16 class constants
  unsigned char* sp = reinterpret_cast<unsigned char*>(smem + 32);   // [L][L] spin bits
  const int64_t chain = blockIdx.x;
  uint32_t* b = a.bits + chain * tpc;
  if (threadIdx.x < 16) {
    const int ti = a.table_index ? a.table_index[chain] : 0;
    tbl[threadIdx.x] = threshold_constant(a.thr[(ti * 16 + threadIdx.x) * 2 + 0], a.thr[(ti * 16 + threadIdx.x) * 2 + 1]);
  }
  for (int q = threadIdx.x; q < N; q += blockDim.x) {
    const int i = q / L, j = q - i * L;
    sp[q] = (unsigned char)((b[i * W + (j >> 5)] >> (j & 31)) & 1u);
  }
  __syncthreads();
  const uint32_t chain_ctr = (uint32_t)(a.chain_offset + chain * a.chain_stride);
  const int n_groups = (N + 3) >> 2;
  for (int s = 0; s < a.n_sweeps; ++s) {
    for (int colour = 0; colour < 3; ++colour) {
      for (int g = threadIdx.x; g < n_groups; g += blockDim.x) {
        uint32_t rr[4] = {0u, 0u, 0u, 0u};
        bool have = false;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int q = 4 * g + d;
          if (q >= N) break;
          const int i = q / L, j = q - i * L;
          int cs = cycle_colour(i, L) + cycle_colour(j, L);
          cs = cs >= 3 ? cs - 3 : cs;
          if (cs != colour) continue;
          if (kInject) {
            rr[d] = a.uniforms[((int64_t)s * a.n_chains + chain) * N + q];
          } else if (!have) {
            u32x4 c;
            c.x = chain_ctr;
            c.y = (uint32_t)(a.sweep_offset + s);
            c.z = (uint32_t)g;
            c.w = DBS_STREAM_SWEEP3;
            const u32x4 o = philox4x32_10(c, a.seed_lo, a.seed_hi);
            rr[0] = o.x; rr[1] = o.y; rr[2] = o.z; rr[3] = o.w;
            have = true;
          }
          const int iu = i == 0 ? L - 1 : i - 1, id = i == L - 1 ? 0 : i + 1;
          const int jl = j == 0 ? L - 1 : j - 1, jr = j == L - 1 ? 0 : j + 1;
          const int self = sp[q];
          const int k = 8 * self + sp[iu * L + j] + sp[id * L + j] + sp[i * L + jl] + sp[i * L + jr];
          if (accept_mask(rr[d], a.one, tbl[k])) sp[q] = (unsigned char)(self ^ 1);
        }
      }
      __syncthreads();
    }
  }
  for (int r = threadIdx.x; r < tpc; r += blockDim.x) {
    const int i = r / W, w = r - i * W;
    const int vb = valid_bits(L, w);
    uint32_t word = 0;
    for (int bb = 0; bb < vb; ++bb) word |= (uint32_t)sp[i * L + 32 * w + bb] << bb;
    b[r] = word;
  }
}

int launch_sweep3(const SweepArgs& a, cudaStream_t st) {
  const size_t smem = 32 * 4 + (size_t)a.L * a.L;
  ProfScope prof(kProfIsingSweep, st);
  if (a.uniforms) ising_sweep3_kernel<true><<<(unsigned)a.n_chains, 256, smem, st>>>(a);
  else ising_sweep3_kernel<false><<<(unsigned)a.n_chains, 256, smem, st>>>(a);
  DBS_LAUNCHED();
  return DBS_OK;
}

// ---------------------------------------------------------------------------------------------
// hot start: one thread per word
__global__ void ising_init_kernel(uint32_t* bits, int L, int64_t n_words, uint32_t seed_lo,
                                  uint32_t seed_hi, int64_t chain_offset, int64_t chain_stride) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_words) return;
  const int W = (L + 31) >> 5, tpc = L * W;
  const int64_t chain = g / tpc;
  const int r = (int)(g - chain * tpc);
  const int i = r / W, w = r - i * W;
  u32x4 c;
  c.x = (uint32_t)(chain_offset + chain * chain_stride);
  c.y = (uint32_t)i;
  c.z = (uint32_t)w;
  c.w = DBS_STREAM_INIT;
  const u32x4 o = philox4x32_10(c, seed_lo, seed_hi);
  bits[g] = o.x & low_mask(valid_bits(L, w));
}

// fp32 -> bits: one warp per word, coalesced 128 B reads, ballot builds the word
__global__ void ising_pack_kernel(const float* __restrict__ spins, uint32_t* __restrict__ bits, int L,
                                  int64_t n_words) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_words) return;
  const int W = (L + 31) >> 5, tpc = L * W;
  const int64_t chain = g / tpc;
  const int r = (int)(g - chain * tpc);
  const int i = r / W, w = r - i * W;
  const int j = 32 * w + lane;
  float v = 0.f;
  if (j < L) v = spins[(chain * L + i) * L + j];
  const uint32_t word = __ballot_sync(0xffffffffu, v > 0.f);
  if (lane == 0) bits[g] = word;
}

__global__ void ising_unpack_kernel(const uint32_t* __restrict__ bits, float* __restrict__ spins, int L,
                                    int64_t n_words) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_words) return;
  const int W = (L + 31) >> 5, tpc = L * W;
  const int64_t chain = g / tpc;
  const int r = (int)(g - chain * tpc);
  const int i = r / W, w = r - i * W;
  const int j = 32 * w + lane;
  const uint32_t word = bits[g];
  if (j < L) spins[(chain * L + i) * L + j] = ((word >> lane) & 1u) ? 1.0f : -1.0f;
}

// ---------------------------------------------------------------------------------------------
// E, M from packed state: one warp per chain, popcount on word-parallel agreement masks.
__global__ void ising_obs_packed_kernel(const uint32_t* __restrict__ bits, int L, int64_t n_chains,
                                        float negJ, float negh, float* __restrict__ energy,
                                        float* __restrict__ mag) {
  const int64_t chain = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (chain >= n_chains) return;
  const int W = (L + 31) >> 5, tpc = L * W;
  const uint32_t* b = bits + chain * tpc;
  int agree = 0, up = 0;
  for (int r = lane; r < tpc; r += 32) {
    const int i = r / W, w = r - i * W;
    const int vb = valid_bits(L, w);
    const uint32_t vmask = low_mask(vb);
    const uint32_t self = b[r];
    const uint32_t rw = b[i * W + (w == W - 1 ? 0 : w + 1)];
    const uint32_t dn = b[(i == L - 1 ? 0 : i + 1) * W + w];
    const uint32_t rt = ((self >> 1) | ((rw & 1u) << (vb - 1))) & vmask;
    agree += __popc(~(self ^ rt) & vmask) + __popc(~(self ^ dn) & vmask);
    up += __popc(self);
  }
  agree = warp_sum(agree);
  up = warp_sum(up);
  if (lane == 0) {
    const int N = L * L;
    const float s_bond = (float)(2 * agree - 2 * N);  // sum of s*right + s*down  (exact integer)
    const float s_spin = (float)(2 * up - N);
    if (energy) energy[chain] = __fadd_rn(__fmul_rn(negJ, s_bond), __fmul_rn(negh, s_spin));
    if (mag) mag[chain] = __fdiv_rn(s_spin, (float)N);
  }
}

// E, M from an fp32 lattice of arbitrary reals: one block per lattice, fp64 accumulation.
__global__ void ising_obs_f32_kernel(const float* __restrict__ spins, int L, float negJ, float negh,
                                     float* __restrict__ energy, float* __restrict__ mag) {
  const float* x = spins + (int64_t)blockIdx.x * L * L;
  double bond = 0.0, sum = 0.0;
  for (int idx = threadIdx.x; idx < L * L; idx += blockDim.x) {
    const int i = idx / L, j = idx - i * L;
    const float s = x[idx];
    const float rt = x[i * L + (j == L - 1 ? 0 : j + 1)];
    const float dn = x[(i == L - 1 ? 0 : i + 1) * L + j];
    bond += (double)__fadd_rn(__fmul_rn(s, rt), __fmul_rn(s, dn));
    sum += (double)s;
  }
  __shared__ double sb[32], ss[32];
  bond = warp_sum(bond);
  sum = warp_sum(sum);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sb[wid] = bond; ss[wid] = sum; }
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    bond = lane < nw ? sb[lane] : 0.0;
    sum = lane < nw ? ss[lane] : 0.0;
    bond = warp_sum(bond);
    sum = warp_sum(sum);
    if (lane == 0) {
      const float fb = (float)bond, fs = (float)sum;
      if (energy) energy[blockIdx.x] = __fadd_rn(__fmul_rn(negJ, fb), __fmul_rn(negh, fs));
      if (mag) mag[blockIdx.x] = __fdiv_rn(fs, (float)(L * L));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// C(r) sums: one block per chain.  Rows are kept twice-unrolled (bit b of D = s[b mod L]) so a
// periodic column shift is one funnel shift; a displacement costs L*W (xor, and, popc).
constexpr int kCorrThreads = 128;

__global__ void __launch_bounds__(kCorrThreads)
ising_corr_kernel(const uint32_t* __restrict__ bits, int L, const uint32_t* __restrict__ ref_mask,
                  const int32_t* __restrict__ group, long long* __restrict__ acc) {
  const int W = (L + 31) >> 5, tpc = L * W, R = L / 2;
  const int DW = 2 * W + 2;
  extern __shared__ uint32_t sm[];
  uint32_t* S = sm;                 // [L][W]   lattice
  uint32_t* M = S + tpc;            // [L][W]   reference-site mask
  uint32_t* D = M + tpc;            // [L][DW]  doubled rows
  int* sbin = reinterpret_cast<int*>(D + L * DW);  // [R]
  __shared__ int s_refcount;
  const int64_t chain = blockIdx.x;
  const uint32_t* b = bits + chain * tpc;
  if (threadIdx.x == 0) s_refcount = 0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) sbin[r] = 0;
  __syncthreads();
  int myref = 0;
  for (int r = threadIdx.x; r < tpc; r += blockDim.x) {
    const int w = r % W;
    const uint32_t vm = low_mask(valid_bits(L, w));
    S[r] = b[r];
    const uint32_t m = (ref_mask ? ref_mask[r] : 0xffffffffu) & vm;
    M[r] = m;
    myref += __popc(m);
  }
  atomicAdd(&s_refcount, myref);
  __syncthreads();
  for (int q = threadIdx.x; q < L * DW; q += blockDim.x) {
    const int i = q / DW, v = q - i * DW;
    uint32_t word = 0;
    int src = (32 * v) % L;
    for (int bb = 0; bb < 32; ++bb) {
      word |= ((S[i * W + (src >> 5)] >> (src & 31)) & 1u) << bb;
      if (++src == L) src = 0;
    }
    D[q] = word;
  }
  __syncthreads();
  const int side = 2 * R + 1;
  const int refcount = s_refcount;
  for (int d = threadIdx.x; d < side * side; d += blockDim.x) {
    const int di = d / side - R, dj = d % side - R;
    const int x = di * di + dj * dj;
    int rr = (int)sqrtf((float)x);
    while (rr * rr > x) --rr;
    while ((rr + 1) * (rr + 1) <= x) ++rr;
    if (rr <= 0 || rr >= R) continue;
    const int djm = ((dj % L) + L) % L;
    int pc = 0;
    for (int i = 0; i < L; ++i) {
      int i2 = i + di;
      i2 = i2 < 0 ? i2 + L : (i2 >= L ? i2 - L : i2);
      const uint32_t* drow = D + i2 * DW;
      for (int w = 0; w < W; ++w) {
        const int t = 32 * w + djm;
        const uint32_t sh = __funnelshift_r(drow[t >> 5], drow[(t >> 5) + 1], t & 31);
        pc += __popc((S[i * W + w] ^ sh) & M[i * W + w]);
      }
    }
    atomicAdd(&sbin[rr], refcount - 2 * pc);
  }
  __syncthreads();
  const int g = group ? group[chain] : 0;
  for (int r = threadIdx.x; r < R; r += blockDim.x)
    if (sbin[r] != 0) atomicAdd(reinterpret_cast<unsigned long long*>(acc + (int64_t)g * R + r),
                                (unsigned long long)(long long)sbin[r]);
}

// C(r) sums with every site a reference site (the ensemble estimator config 3 uses), L <= 64: ONE WARP per
// chain, the lattice in registers (lane l holds rows l and l + 32 as 64-bit words).  For a row shift di the
// shifted rows are fetched once with shuffles and serve all column shifts; a column shift is a 64-bit rotate
// inside L bits (two funnel shifts, warp-uniform amount), a displacement costs 2 x (xor, popcll) per lane and
// one REDUX when the radial bin changes.  Only the half plane (di > 0, or di = 0 and dj > 0) is visited:
// sum_x s(x) s(x+d) = sum_x s(x) s(x-d) on the torus and every displacement with 0 < r < L/2 differs from its
// negative, so each sum counts twice.  Lane r keeps bin r; nothing is staged in shared memory but the bin table.
constexpr int kCorrWarps = 8;

template <bool k64>
__device__ __forceinline__ uint64_t rot_right_L(uint64_t v, int s, int L, uint64_t maskL) {
  // bit j of the result = bit (j + s) mod L of v; v has no bits at or above L; 0 <= s < L (warp-uniform)
  if (k64) {
    // L = 64: a 64-bit rotate is two funnel shifts of the 32-bit halves (swapped when s >= 32)
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    const uint32_t a = s & 32 ? hi : lo, b = s & 32 ? lo : hi;
    return (uint64_t)__funnelshift_r(a, b, s) | ((uint64_t)__funnelshift_r(b, a, s) << 32);
  }
  if (s == 0) return v;
  return ((v >> s) | (v << (L - s))) & maskL;
}

template <bool k64>
__global__ void __launch_bounds__(32 * kCorrWarps)
ising_corr_warp_kernel(const uint32_t* __restrict__ bits, int L, int64_t n_chains, const int32_t* __restrict__ group,
                       long long* __restrict__ acc) {
  const int W = (L + 31) >> 5, R = L / 2;
  __shared__ unsigned char s_bin[32 * 32];   // radial bin of (di, |dj|), di, |dj| < R; 255 = outside 0 < r < R
  __shared__ int s_acc[32];
  __shared__ int s_group;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = threadIdx.x; q < R * R; q += blockDim.x) {
    const int di = q / R, dj = q - di * R;
    const int x = di * di + dj * dj;
    int rr = (int)sqrtf((float)x);
    while (rr * rr > x) --rr;
    while ((rr + 1) * (rr + 1) <= x) ++rr;
    s_bin[q] = (rr <= 0 || rr >= R) ? 255 : (unsigned char)rr;
  }
  if (threadIdx.x < 32) s_acc[threadIdx.x] = 0;
  const int64_t chain0 = (int64_t)blockIdx.x * kCorrWarps;
  if (threadIdx.x == 0) s_group = group ? group[chain0] : 0;
  __syncthreads();
  const int64_t chain = chain0 + warp;
  int mybin = 0;                     // lane r: sum over the displacements of bin r of (L^2 - 2 * disagreements), x2
  int g = 0;
  if (chain < n_chains) {
    g = group ? group[chain] : 0;
    const uint32_t* b = bits + chain * (int64_t)(L * W);
    const uint64_t maskL = L >= 64 ? ~0ull : ((1ull << L) - 1ull);
    uint64_t lo = 0, hi = 0;
    if (lane < L) lo = W == 2 ? ((uint64_t)b[lane * 2] | ((uint64_t)b[lane * 2 + 1] << 32)) : (uint64_t)b[lane];
    if (lane + 32 < L) hi = (uint64_t)b[(lane + 32) * 2] | ((uint64_t)b[(lane + 32) * 2 + 1] << 32);
    const bool has_hi = L > 32;
    const int LL = L * L;
    for (int di = 0; di < R; ++di) {
      // rows (lane + di) mod L and (lane + 32 + di) mod L
      int r0 = lane + di; if (r0 >= L) r0 -= L;
      int r1 = lane + 32 + di; while (r1 >= L) r1 -= L;
      const uint64_t a_lo = __shfl_sync(0xffffffffu, lo, r0 & 31), a_hi = __shfl_sync(0xffffffffu, hi, r0 & 31);
      const uint64_t b_lo = __shfl_sync(0xffffffffu, lo, r1 & 31), b_hi = __shfl_sync(0xffffffffu, hi, r1 & 31);
      const uint64_t sa = r0 >= 32 ? a_hi : a_lo;     // partner row of my row `lane`
      const uint64_t sb = r1 >= 32 ? b_hi : b_lo;     // partner row of my row `lane + 32`
      const bool v0 = lane < L, v1 = lane + 32 < L;
      for (int sgn = 0; sgn < (di == 0 ? 1 : 2); ++sgn) {
        // sgn 0: dj = (di == 0 ? 1 : 0) .. R-1;  sgn 1: dj = -1 .. -(R-1); the bin grows with |dj|
        int cur_bin = -1, cur = 0, ndisp = 0;
        for (int k = (sgn == 0 ? (di == 0 ? 1 : 0) : 1); k < R; ++k) {
          const int bin = s_bin[di * R + k];
          if (bin == 255) { if (di * di + k * k >= R * R) break; else continue; }   // r = 0 cannot occur here
          if (bin != cur_bin) {
            if (cur_bin >= 0) {
              const int tot = __reduce_add_sync(0xffffffffu, cur);
              if (lane == cur_bin) mybin += 2 * (ndisp * LL - 2 * tot);
            }
            cur_bin = bin; cur = 0; ndisp = 0;
          }
          const int s = sgn == 0 ? k : L - k;       // column shift mod L
          int pc = 0;
          if (k64) {
            pc = __popcll(lo ^ rot_right_L<true>(sa, s, L, maskL)) + __popcll(hi ^ rot_right_L<true>(sb, s, L, maskL));
          } else {
            if (v0) pc += __popcll(lo ^ rot_right_L<false>(sa, s, L, maskL));
            if (has_hi && v1) pc += __popcll(hi ^ rot_right_L<false>(sb, s, L, maskL));
          }
          cur += pc;
          ++ndisp;
        }
        if (cur_bin >= 0) {
          const int tot = __reduce_add_sync(0xffffffffu, cur);
          if (lane == cur_bin) mybin += 2 * (ndisp * LL - 2 * tot);
        }
      }
    }
  }
  // chains of a block usually share their group (temperature): one global atomic per bin per block
  const bool shared_group = chain < n_chains && g == s_group;
  if (shared_group && lane < R && mybin != 0) atomicAdd(&s_acc[lane], mybin);
  if (chain < n_chains && !shared_group && lane < R && mybin != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(acc + (int64_t)g * R + lane), (unsigned long long)(long long)mybin);
  __syncthreads();
  if (threadIdx.x < R && s_acc[threadIdx.x] != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(acc + (int64_t)s_group * R + threadIdx.x),
              (unsigned long long)(long long)s_acc[threadIdx.x]);
}

// Second cut of the warp-per-chain C(r) kernel: same half-plane visit order and the same integer sums, fewer
// instructions per displacement.  (1) The partner rows are rotated by ONE column per step (two funnel shifts for a
// 64-bit row) instead of by a fresh amount; (2) a lane adds its disagreement count straight into its own column of a
// per-warp [bin][lane] table in shared memory (one conflict-free RED.shared) instead of carrying a running sum with a
// warp reduction at every change of the radial bin — at small di the bin changes with every column; the 32 bins are
// reduced once per chain; (3) the number of displacements per bin is a per-block table, not counted in the loop.
template <bool k64>
__device__ __forceinline__ uint64_t rot1_right_L(uint64_t v, int L) {
  return k64 ? ((v >> 1) | (v << 63)) : ((v >> 1) | ((v & 1ull) << (L - 1)));
}
template <bool k64>
__device__ __forceinline__ uint64_t rot1_left_L(uint64_t v, int L, uint64_t maskL) {
  return k64 ? ((v << 1) | (v >> 63)) : (((v << 1) & maskL) | (v >> (L - 1)));
}

// kObs: the same launch also evaluates E and M of every chain (the lattice is in registers anyway) and adds
// {1, E/N, (E/N)^2, |M|, M^2, M^4} to the fp64 observable block of the chain's group — one launch per measurement of the
// temperature sweep instead of three (dbs_ising_measure).  E and M are formed exactly as ising_obs_packed_kernel does.
template <bool k64, bool kObs>
__global__ void __launch_bounds__(32 * kCorrWarps)
ising_corr_warp2_kernel(const uint32_t* __restrict__ bits, int L, int64_t n_chains, const int32_t* __restrict__ group,
                        long long* __restrict__ acc, float negJ, float negh, double* __restrict__ block, int block_stride) {
  const int W = (L + 31) >> 5, R = L / 2;
  __shared__ unsigned char s_bin[32 * 32];   // radial bin of (di, |dj|), di, |dj| < R; 255 = outside 0 < r < R
  __shared__ int s_nd[32];                   // half-plane displacements per bin
  __shared__ int s_part[kCorrWarps][32][32]; // [warp][bin][lane] disagreement counts
  __shared__ int s_acc[32];
  __shared__ int s_group;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 32) { s_acc[threadIdx.x] = 0; s_nd[threadIdx.x] = 0; }
  for (int b = 0; b < 32; ++b) s_part[warp][b][lane] = 0;
  __syncthreads();
  for (int q = threadIdx.x; q < R * R; q += blockDim.x) {
    const int di = q / R, dj = q - di * R;
    const int x = di * di + dj * dj;
    int rr = (int)sqrtf((float)x);
    while (rr * rr > x) --rr;
    while ((rr + 1) * (rr + 1) <= x) ++rr;
    const bool in = rr > 0 && rr < R;
    s_bin[q] = in ? (unsigned char)rr : 255;
    // visits: di = 0: dj = 1..R-1 once; di > 0: dj = 0 once, +-dj twice
    if (in) atomicAdd(&s_nd[rr], (di == 0 || dj == 0) ? 1 : 2);
  }
  const int64_t chain0 = (int64_t)blockIdx.x * kCorrWarps;
  if (threadIdx.x == 0) s_group = group ? group[chain0] : 0;
  __syncthreads();
  const int64_t chain = chain0 + warp;
  int mybin = 0;
  int g = 0;
  if (chain < n_chains) {
    g = group ? group[chain] : 0;
    const uint32_t* b = bits + chain * (int64_t)(L * W);
    const uint64_t maskL = L >= 64 ? ~0ull : ((1ull << L) - 1ull);
    uint64_t lo = 0, hi = 0;
    if (lane < L) lo = W == 2 ? ((uint64_t)b[lane * 2] | ((uint64_t)b[lane * 2 + 1] << 32)) : (uint64_t)b[lane];
    if (lane + 32 < L) hi = (uint64_t)b[(lane + 32) * 2] | ((uint64_t)b[(lane + 32) * 2 + 1] << 32);
    const bool v0 = lane < L, v1 = lane + 32 < L;
    if (kObs) {
      // bonds to the right neighbour (column rotation by one) and to the row below (row shift by one)
      int r0 = lane + 1; if (r0 >= L) r0 -= L;
      int r1 = lane + 33; while (r1 >= L) r1 -= L;
      const uint64_t a_lo = __shfl_sync(0xffffffffu, lo, r0 & 31), a_hi = __shfl_sync(0xffffffffu, hi, r0 & 31);
      const uint64_t b_lo = __shfl_sync(0xffffffffu, lo, r1 & 31), b_hi = __shfl_sync(0xffffffffu, hi, r1 & 31);
      const uint64_t dn0 = v0 ? (r0 >= 32 ? a_hi : a_lo) : 0ull, dn1 = v1 ? (r1 >= 32 ? b_hi : b_lo) : 0ull;
      int dis = __popcll(lo ^ dn0) + __popcll(hi ^ dn1);
      if (v0) dis += __popcll(lo ^ rot1_right_L<k64>(lo, L));
      if (v1) dis += __popcll(hi ^ rot1_right_L<k64>(hi, L));
      int up = __popcll(lo) + __popcll(hi);
      dis = __reduce_add_sync(0xffffffffu, dis);
      up = __reduce_add_sync(0xffffffffu, up);
      if (lane == 0) {
        const int N = L * L;
        const int agree = 2 * N - dis;
        const float s_bond = (float)(2 * agree - 2 * N);
        const float s_spin = (float)(2 * up - N);
        const float energy = __fadd_rn(__fmul_rn(negJ, s_bond), __fmul_rn(negh, s_spin));
        const float mag = __fdiv_rn(s_spin, (float)N);
        const double e = (double)energy / (double)N, m = (double)mag, m2 = m * m;
        double* a = block + (int64_t)g * block_stride;
        atomicAdd(a + 0, 1.0);
        atomicAdd(a + 1, e);
        atomicAdd(a + 2, e * e);
        atomicAdd(a + 3, fabs(m));
        atomicAdd(a + 4, m2);
        atomicAdd(a + 5, m2 * m2);
      }
    }
    int* part = &s_part[warp][0][lane];
    for (int di = 0; di < R; ++di) {
      int r0 = lane + di; if (r0 >= L) r0 -= L;
      int r1 = lane + 32 + di; while (r1 >= L) r1 -= L;
      const uint64_t a_lo = __shfl_sync(0xffffffffu, lo, r0 & 31), a_hi = __shfl_sync(0xffffffffu, hi, r0 & 31);
      const uint64_t b_lo = __shfl_sync(0xffffffffu, lo, r1 & 31), b_hi = __shfl_sync(0xffffffffu, hi, r1 & 31);
      // partner rows; zero where this lane holds no row, so such lanes count no disagreements (their own row is zero too)
      const uint64_t sa = v0 ? (r0 >= 32 ? a_hi : a_lo) : 0ull;
      const uint64_t sb = v1 ? (r1 >= 32 ? b_hi : b_lo) : 0ull;
      const unsigned char* brow = s_bin + di * R;
      // dj = +k (bit j of the shifted row = bit (j + k) mod L: right rotation) and dj = -k (left rotation) fall in the same
      // radial bin, so one pass over k handles both with one bin lookup and one shared-memory add
      {
        const int k0 = di == 0 ? 1 : 0;
        uint64_t ra = sa, rb = sb;
        if (k0) { ra = rot1_right_L<k64>(ra, L); rb = rot1_right_L<k64>(rb, L); }
        uint64_t la = rot1_left_L<k64>(sa, L, maskL), lb = rot1_left_L<k64>(sb, L, maskL);   // k = 1 of the negative side
        if (k0 == 0) {
          const int bin = brow[0];     // dj = 0 (di > 0): counted once
          if (bin != 255) atomicAdd(part + bin * 32, __popcll(lo ^ ra) + __popcll(hi ^ rb));
          ra = rot1_right_L<k64>(ra, L); rb = rot1_right_L<k64>(rb, L);
        }
        for (int k = 1; k < R; ++k) {
          const int bin = brow[k];
          if (bin == 255) break;       // r >= R from here on (r = 0 only at di = dj = 0, which is never visited)
          int pc = __popcll(lo ^ ra) + __popcll(hi ^ rb);
          if (di > 0) pc += __popcll(lo ^ la) + __popcll(hi ^ lb);
          atomicAdd(part + bin * 32, pc);
          ra = rot1_right_L<k64>(ra, L); rb = rot1_right_L<k64>(rb, L);
          la = rot1_left_L<k64>(la, L, maskL); lb = rot1_left_L<k64>(lb, L, maskL);
        }
      }
    }
    __syncwarp();
    const int LL = L * L;
    for (int bn = 1; bn < R; ++bn) {
      const int tot = __reduce_add_sync(0xffffffffu, s_part[warp][bn][lane]);
      if (lane == bn) mybin = 2 * (s_nd[bn] * LL - 2 * tot);
    }
  }
  const bool shared_group = chain < n_chains && g == s_group;
  if (shared_group && lane < R && mybin != 0) atomicAdd(&s_acc[lane], mybin);
  if (chain < n_chains && !shared_group && lane < R && mybin != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(acc + (int64_t)g * R + lane), (unsigned long long)(long long)mybin);
  __syncthreads();
  if (threadIdx.x < R && s_acc[threadIdx.x] != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(acc + (int64_t)s_group * R + threadIdx.x),
              (unsigned long long)(long long)s_acc[threadIdx.x]);
}

__global__ void obs_accumulate_kernel(const float* __restrict__ energy, const float* __restrict__ mag,
                                      double N, int64_t n, const int32_t* __restrict__ group,
                                      double* __restrict__ acc) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int g = group ? group[c] : 0;
  const double e = (double)energy[c] / N;
  const double m = (double)mag[c];
  const double m2 = m * m;
  double* a = acc + (int64_t)g * 5;
  atomicAdd(a + 0, 1.0);
  atomicAdd(a + 1, e);
  atomicAdd(a + 2, fabs(m));
  atomicAdd(a + 3, m2);
  atomicAdd(a + 4, m2 * m2);
}

// Single random-site attempt on an fp32 lattice (the reference's step(), samplers/mcmc.py:26-50):
// same class/threshold arithmetic as the sweep, one thread per lattice of the batch.
__global__ void ising_site_update_kernel(float* __restrict__ spins, int L, int64_t n, int i, int j,
                                         const uint32_t* __restrict__ thr, uint32_t r) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  float* x = spins + b * L * L;
  const int iu = (i == 0 ? L - 1 : i - 1), id = (i == L - 1 ? 0 : i + 1);
  const int jl = (j == 0 ? L - 1 : j - 1), jr = (j == L - 1 ? 0 : j + 1);
  const int s = x[i * L + j] > 0.f;
  const int u = (x[iu * L + j] > 0.f) + (x[id * L + j] > 0.f) + (x[i * L + jl] > 0.f) + (x[i * L + jr] > 0.f);
  const int k = 8 * s + u;
  if ((r < thr[2 * k]) || thr[2 * k + 1]) x[i * L + j] = -x[i * L + j];
}

}  // namespace
}  // namespace dbs

using namespace dbs;

extern "C" int dbs_ising_words_per_chain(int L) {
  if (!supported_L(L)) return -1;
  return L * ((L + 31) / 32);
}

extern "C" int dbs_ising_init_packed_strided(uint32_t* bits, int L, int64_t n_chains, uint64_t seed,
                                             int64_t chain_offset, int64_t chain_stride, void* stream) {
  DBS_REQUIRE(supported_L(L), "ising: lattice size %d unsupported (4..%d, not 3)", L, kMaxL);
  if (n_chains == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(bits && n_chains >= 0 && chain_stride >= 1, "ising_init: null state, negative chain count or stride < 1");
  if (n_chains == 0) return DBS_OK;
  const int64_t n_words = n_chains * dbs_ising_words_per_chain(L);
  ising_init_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, as_stream(stream)>>>(
      bits, L, n_words, (uint32_t)seed, (uint32_t)(seed >> 32), chain_offset, chain_stride);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_ising_init_packed(uint32_t* bits, int L, int64_t n_chains, uint64_t seed,
                                     int64_t chain_offset, void* stream) {
  return dbs_ising_init_packed_strided(bits, L, n_chains, seed, chain_offset, 1, stream);
}

extern "C" int dbs_ising_pack(const float* spins, uint32_t* bits, int L, int64_t n_chains, void* stream) {
  DBS_REQUIRE(supported_L(L), "ising: lattice size %d unsupported (4..%d, not 3)", L, kMaxL);
  if (n_chains == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(spins && bits && n_chains >= 0, "ising_pack: null pointer");
  if (n_chains == 0) return DBS_OK;
  const int64_t n_words = n_chains * dbs_ising_words_per_chain(L);
  ising_pack_kernel<<<(unsigned)((n_words * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(spins, bits, L, n_words);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_ising_unpack(const uint32_t* bits, float* spins, int L, int64_t n_chains, void* stream) {
  DBS_REQUIRE(supported_L(L), "ising: lattice size %d unsupported (4..%d, not 3)", L, kMaxL);
  if (n_chains == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(spins && bits && n_chains >= 0, "ising_unpack: null pointer");
  if (n_chains == 0) return DBS_OK;
  const int64_t n_words = n_chains * dbs_ising_words_per_chain(L);
  ising_unpack_kernel<<<(unsigned)((n_words * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(bits, spins, L, n_words);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_ising_sweeps(uint32_t* bits, int L, int64_t n_chains, int n_sweeps, const uint32_t* thr,
                                int n_tables, const int32_t* table_index, uint64_t seed, int64_t chain_offset,
                                int64_t sweep_offset, const uint32_t* uniforms, int zero_field, void* stream) {
  return dbs_ising_sweeps_strided(bits, L, n_chains, n_sweeps, thr, n_tables, table_index, seed, chain_offset, 1,
                                  sweep_offset, uniforms, zero_field, stream);
}

extern "C" int dbs_ising_sweeps_strided(uint32_t* bits, int L, int64_t n_chains, int n_sweeps, const uint32_t* thr,
                                        int n_tables, const int32_t* table_index, uint64_t seed, int64_t chain_offset,
                                        int64_t chain_stride, int64_t sweep_offset, const uint32_t* uniforms,
                                        int zero_field, void* stream) {
  DBS_REQUIRE(chain_stride >= 1, "ising_sweeps: chain_stride must be >= 1");
  DBS_REQUIRE(supported_L(L), "ising: lattice size %d unsupported (4..%d, not 3)", L, kMaxL);
  if (n_chains == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(bits && thr && n_tables >= 1, "ising_sweeps: null state or threshold table");
  DBS_REQUIRE(n_chains >= 0 && n_sweeps >= 0, "ising_sweeps: negative chain or sweep count");
  DBS_REQUIRE(zero_field >= -1 && zero_field <= 1, "ising_sweeps: zero_field must be -1, 0 or +1");
  if (n_chains == 0 || n_sweeps == 0) return DBS_OK;
  SweepArgs a;
  a.bits = bits; a.thr = thr; a.table_index = table_index; a.uniforms = uniforms;
  a.n_chains = n_chains; a.chain_offset = chain_offset; a.chain_stride = chain_stride; a.sweep_offset = sweep_offset;
  a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    a.rk[2 * r] = a.seed_lo + (uint32_t)r * kPhiloxW0;
    a.rk[2 * r + 1] = a.seed_hi + (uint32_t)r * kPhiloxW1;
  }
  a.L = L; a.n_sweeps = n_sweeps;
  a.one = 1u; a.antiferro = zero_field < 0 ? 1 : 0;
  const int tpc = dbs_ising_words_per_chain(L);
  a.cpb = tpc >= 128 ? 1 : 128 / tpc;
  cudaStream_t st = as_stream(stream);
  if (L & 1) return launch_sweep3(a, st);
  if (zero_field != 0) return uniforms ? launch_sweep<true, 1>(a, st) : launch_sweep<false, 1>(a, st);
  return uniforms ? launch_sweep<true, 0>(a, st) : launch_sweep<false, 0>(a, st);
}

extern "C" int dbs_ising_observables(const uint32_t* bits, int L, int64_t n_chains, float J, float h,
                                     float* energy, float* magnetization, void* stream) {
  DBS_REQUIRE(supported_L(L), "ising: lattice size %d unsupported (4..%d, not 3)", L, kMaxL);
  if (n_chains == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(bits && n_chains >= 0, "ising_observables: null state");
  if (n_chains == 0) return DBS_OK;
  ProfScope prof(kProfIsingObs, as_stream(stream));
  ising_obs_packed_kernel<<<(unsigned)((n_chains * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
      bits, L, n_chains, -J, -h, energy, magnetization);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_ising_observables_f32(const float* spins, int L, int64_t n, float J, float h, float* energy,
                                         float* magnetization, void* stream) {
  if (n == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(L >= 1 && spins && n >= 0, "ising_observables_f32: bad arguments");
  if (n == 0) return DBS_OK;
  const int threads = L * L >= 1024 ? 256 : (L * L >= 128 ? 128 : 32);
  ising_obs_f32_kernel<<<(unsigned)n, threads, 0, as_stream(stream)>>>(spins, L, -J, -h, energy, magnetization);
  DBS_LAUNCHED();
  return DBS_OK;
}

namespace dbs {
// ising_corr_fft.cu: spectral evaluation of the same sums (64 x 64 only)
int ising_corr_fft_launch(const uint32_t* bits, int64_t n_chains, const int32_t* group, long long* acc, float J, float h,
                          double* block, int block_stride, cudaStream_t st);
// 0 = automatic (spectral kernel at L = 64 with every site a reference site, XOR / POPC kernels otherwise),
// 1 = always the XOR / POPC kernels, 2 = spectral where it applies (same as 0 today)
static std::atomic<int> g_corr_method{0};
static bool corr_use_fft(int L, const uint32_t* ref_mask) {
  static const bool env_off = [] { const char* e = getenv("DBS_CORR_FFT"); return e && e[0] == '0'; }();
  return L == 64 && !ref_mask && !env_off && g_corr_method.load(std::memory_order_relaxed) != 1;
}
}  // namespace dbs

extern "C" int dbs_ising_set_correlation_method(int method) {
  if (method < 0 || method > 2) { set_error("ising_set_correlation_method: method %d (0 automatic, 1 direct, 2 spectral)", method); return DBS_ERR_INVALID; }
  return g_corr_method.exchange(method, std::memory_order_relaxed);
}

extern "C" int dbs_ising_correlation(const uint32_t* bits, int L, int64_t n_chains, const uint32_t* ref_mask,
                                     const int32_t* group, long long* acc, void* stream) {
  DBS_REQUIRE(supported_L(L), "ising: lattice size %d unsupported (4..%d, not 3)", L, kMaxL);
  if (n_chains == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(bits && acc && n_chains >= 0, "ising_correlation: null pointer");
  if (n_chains == 0) return DBS_OK;
  const int W = (L + 31) / 32, tpc = L * W;
  const size_t smem = (size_t)(2 * tpc + L * (2 * W + 2)) * 4 + (size_t)(L / 2) * 4;
  ProfScope prof(kProfIsingObs, as_stream(stream));
  if (corr_use_fft(L, ref_mask)) return ising_corr_fft_launch(bits, n_chains, group, acc, 0.f, 0.f, nullptr, 0, as_stream(stream));
  static const bool warp_off = [] { const char* e = getenv("DBS_CORR_WARP"); return e && e[0] == '0'; }();
  if (!ref_mask && L <= 64 && !warp_off) {
    const unsigned grid = (unsigned)((n_chains + kCorrWarps - 1) / kCorrWarps);
    static const bool v1 = [] { const char* e = getenv("DBS_CORR_V2"); return e && e[0] == '0'; }();
    if (v1) {
      if (L == 64) ising_corr_warp_kernel<true><<<grid, 32 * kCorrWarps, 0, as_stream(stream)>>>(bits, L, n_chains, group, acc);
      else ising_corr_warp_kernel<false><<<grid, 32 * kCorrWarps, 0, as_stream(stream)>>>(bits, L, n_chains, group, acc);
    } else {
      if (L == 64) ising_corr_warp2_kernel<true, false><<<grid, 32 * kCorrWarps, 0, as_stream(stream)>>>(bits, L, n_chains, group, acc, 0.f, 0.f, nullptr, 0);
      else ising_corr_warp2_kernel<false, false><<<grid, 32 * kCorrWarps, 0, as_stream(stream)>>>(bits, L, n_chains, group, acc, 0.f, 0.f, nullptr, 0);
    }
    DBS_LAUNCHED();
    return DBS_OK;
  }
  ising_corr_kernel<<<(unsigned)n_chains, kCorrThreads, smem, as_stream(stream)>>>(bits, L, ref_mask, group, acc);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_ising_measure(const uint32_t* bits, int L, int64_t n_chains, float J, float h, const int32_t* group,
                                 double* block, int block_stride, long long* corr_acc, void* stream) {
  DBS_REQUIRE(supported_L(L) && !(L & 1) && L <= 64, "ising_measure: even lattice sizes 4..64 only (got %d)", L);
  if (n_chains == 0) return DBS_OK;
  DBS_REQUIRE(bits && block && corr_acc && n_chains >= 0 && block_stride >= 6, "ising_measure: bad arguments");
  const unsigned grid = (unsigned)((n_chains + kCorrWarps - 1) / kCorrWarps);
  ProfScope prof(kProfIsingObs, as_stream(stream));
  if (corr_use_fft(L, nullptr)) return ising_corr_fft_launch(bits, n_chains, group, corr_acc, J, h, block, block_stride, as_stream(stream));
  if (L == 64) ising_corr_warp2_kernel<true, true><<<grid, 32 * kCorrWarps, 0, as_stream(stream)>>>(bits, L, n_chains, group, corr_acc, -J, -h, block, block_stride);
  else ising_corr_warp2_kernel<false, true><<<grid, 32 * kCorrWarps, 0, as_stream(stream)>>>(bits, L, n_chains, group, corr_acc, -J, -h, block, block_stride);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_observable_accumulate(const float* energy, const float* magnetization, int L, int64_t n_chains,
                                         const int32_t* group, double* acc, void* stream) {
  if (n_chains == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(energy && magnetization && acc && L >= 1 && n_chains >= 0, "observable_accumulate: bad arguments");
  if (n_chains == 0) return DBS_OK;
  obs_accumulate_kernel<<<(unsigned)((n_chains + 255) / 256), 256, 0, as_stream(stream)>>>(
      energy, magnetization, (double)(L * L), n_chains, group, acc);
  DBS_LAUNCHED();
  return DBS_OK;
}

extern "C" int dbs_ising_site_update(float* spins, int L, int64_t n, int i, int j, const uint32_t* thr, uint32_t r,
                                     void* stream) {
  if (n == 0) return DBS_OK;   // empty batch: nothing to do, pointers may be NULL
  DBS_REQUIRE(spins && thr && L >= 1 && n >= 0, "ising_site_update: bad arguments");
  DBS_REQUIRE(i >= 0 && i < L && j >= 0 && j < L, "ising_site_update: site (%d,%d) outside a %dx%d lattice", i, j, L, L);
  if (n == 0) return DBS_OK;
  ising_site_update_kernel<<<(unsigned)((n + 127) / 128), 128, 0, as_stream(stream)>>>(spins, L, n, i, j, thr, r);
  DBS_LAUNCHED();
  return DBS_OK;
}
