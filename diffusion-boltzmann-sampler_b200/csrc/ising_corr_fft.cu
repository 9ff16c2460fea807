// ising_corr_fft.cu — spectral C(r) kernel for 64 x 64 lattices (config 3's measurement): the radial correlation sums of
// dbs_ising_correlation / dbs_ising_measure through |FFT2|^2 -> inverse FFT2 instead of one XOR / POPC pass per
// displacement.  Same integer sums (see ising_corr_fft.cuh for why they are exact); pair_correlation of the reference:
// backend/ml/analysis/correlation.py:8-47.
//
// Work decomposition: a set of 64 threads owns four chains at a time; a CTA is four independent sets (their own named
// barriers), one CTA per SM.  Thread t holds one 64-point line in registers in every pass (row t / column t), lines meet in
// a padded shared-memory buffer between passes; the six passes of a round run through ONE transform body in a loop
// (first cut: six inlined bodies = 206 KB of straight-line code per kernel, ncu: 3.9 of 4.9 stall samples per issue were
// "no instruction" — the kernel was instruction-fetch bound at 25 % issue utilisation).
// Per-thread bin partials live in a [bin][thread] fp32 table (exact small integers), reduced to int64 and added to the
// group's accumulator when the set's temperature group changes (chains of a group are contiguous, so almost never).
// Four chains of different groups are processed as two rounds of single (real-only) lattices, re -> one group, im -> another.
// ~57 KB of shared memory per set.
#include "common.cuh"
#include "ising_corr_fft.cuh"

namespace dbs {
namespace {

using namespace cfft;

constexpr int kSets = 4;             // independent 64-thread sets per CTA, four chains each per iteration
constexpr int kSetThreads = 64;
constexpr int kFlushRounds = 24;     // |q| <= 16384, <= ~20 displacements of a bin per row: 24 rounds stay below 2^24

__device__ const BinTable g_bin_table = make_bin_table();

struct __align__(16) SetSmem {
  cpx buf[64 * kStride];             // lines meet here between passes; its head doubles as the row staging of the E / M step
  float pw[64 * 64];                 // P1, a thread-private column per thread
  float part[32 * 64];               // [bin][thread] partial sums, exact integers in fp32
};

__device__ __forceinline__ void set_barrier(int set) { asm volatile("bar.sync %0, 64;" ::"r"(set + 1) : "memory"); }

template <bool kObs>
__global__ void __launch_bounds__(kSets * kSetThreads, 1)
ising_corr_fft_kernel(const uint32_t* __restrict__ bits, int64_t n_chains, const int32_t* __restrict__ group,
                      long long* __restrict__ acc, float negJ, float negh, double* __restrict__ block, int block_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int set = threadIdx.x >> 6, t = threadIdx.x & 63, lane = t & 31, warp = t >> 5;
  SetSmem& sm = reinterpret_cast<SetSmem*>(smem_raw)[set];
  const unsigned char* bin_tab = g_bin_table.v;
#pragma unroll
  for (int b = 0; b < 32; ++b) sm.part[b * 64 + t] = 0.f;

  // contiguous quads of chains per set: chains of a temperature group are contiguous, so a set rarely changes group
  const int64_t n_quads = (n_chains + 3) >> 2;
  const int64_t unit = (int64_t)blockIdx.x * kSets + set, n_units = (int64_t)gridDim.x * kSets;
  const int64_t q_lo = n_quads * unit / n_units, q_hi = n_quads * (unit + 1) / n_units;
  int cur_g = -1, acc_rounds = 0;
  bool dirty = false;              // the bin table holds sums that belong to cur_g
  float maxdev = 0.f;

  auto flush = [&]() {
    set_barrier(set);
    if (cur_g >= 0 && dirty) {
      if (t > 0 && t < 32) {
        long long s = 0;
#pragma unroll 8
        for (int c = 0; c < 64; ++c) s += (long long)__float2int_rn(sm.part[t * 64 + ((c + t) & 63)]);
        if (s != 0) atomicAdd(reinterpret_cast<unsigned long long*>(acc + (int64_t)cur_g * 32 + t), (unsigned long long)s);
      }
      // a value that was not within 1/4 of an integer would mean the fp32 transforms lost the exactness argument of
      // ising_corr_fft.cuh: bin 0 (never a valid radius) of the group is poisoned and the host refuses the result
      if (maxdev > 0.25f) atomicAdd(reinterpret_cast<unsigned long long*>(acc + (int64_t)cur_g * 32), 1ull);
    }
    maxdev = 0.f;
    set_barrier(set);
#pragma unroll
    for (int b = 0; b < 32; ++b) sm.part[b * 64 + t] = 0.f;
    acc_rounds = 0;
    dirty = false;
  };
  auto ensure_group = [&](int g) {
    if (g != cur_g) {
      if (dirty) flush();
      cur_g = g;
    }
    dirty = true;                  // every call is followed by a binning pass into the table
  };

#pragma unroll 1
  for (int64_t q = q_lo; q < q_hi; ++q) {
    const int64_t c0 = q * 4;
    const int nv = (int)((n_chains - c0) < 4 ? (n_chains - c0) : 4);
    uint64_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    int g0 = group ? group[c0] : 0, g1 = g0, g2 = g0, g3 = g0;
    {
      const uint2* b = reinterpret_cast<const uint2*>(bits + c0 * 128) + t;
      uint2 w = b[0];
      r0 = (uint64_t)w.x | ((uint64_t)w.y << 32);
      if (nv > 1) { w = b[64]; r1 = (uint64_t)w.x | ((uint64_t)w.y << 32); if (group) g1 = group[c0 + 1]; }
      if (nv > 2) { w = b[128]; r2 = (uint64_t)w.x | ((uint64_t)w.y << 32); if (group) g2 = group[c0 + 2]; }
      if (nv > 3) { w = b[192]; r3 = (uint64_t)w.x | ((uint64_t)w.y << 32); if (group) g3 = group[c0 + 3]; }
    }
    const bool same = g1 == g0 && g2 == g0 && g3 == g0;
    if (kObs) {
      // E and M of the four chains, formed exactly as ising_obs_packed_kernel / ising_corr_warp2_kernel do: bonds to the
      // right neighbour (column rotation by one) and to the row below; a warp takes chains warp and warp + 2
      uint64_t (*rows)[64] = reinterpret_cast<uint64_t (*)[64]>(sm.buf);
      rows[0][t] = r0; rows[1][t] = r1; rows[2][t] = r2; rows[3][t] = r3;
      set_barrier(set);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = warp + 2 * cc;
        if (c < nv) {
          const uint64_t lo = rows[c][lane], hi = rows[c][lane + 32];
          const uint64_t dn0 = rows[c][(lane + 1) & 63], dn1 = rows[c][(lane + 33) & 63];
          int dis = __popcll(lo ^ dn0) + __popcll(hi ^ dn1);
          dis += __popcll(lo ^ ((lo >> 1) | (lo << 63))) + __popcll(hi ^ ((hi >> 1) | (hi << 63)));
          int up = __popcll(lo) + __popcll(hi);
          dis = __reduce_add_sync(0xffffffffu, dis);
          up = __reduce_add_sync(0xffffffffu, up);
          if (lane == 0) {
            const int N = 64 * 64;
            const int agree = 2 * N - dis;
            const float s_bond = (float)(2 * agree - 2 * N);
            const float s_spin = (float)(2 * up - N);
            const float energy = __fadd_rn(__fmul_rn(negJ, s_bond), __fmul_rn(negh, s_spin));
            const float mag = __fdiv_rn(s_spin, (float)N);
            const double e = (double)energy / (double)N, m = (double)mag, m2 = m * m;
            const int g = c == 0 ? g0 : c == 1 ? g1 : c == 2 ? g2 : g3;
            double* a = block + (int64_t)g * block_stride;
            atomicAdd(a + 0, 1.0);
            atomicAdd(a + 1, e);
            atomicAdd(a + 2, e * e);
            atomicAdd(a + 3, fabs(m));
            atomicAdd(a + 4, m2);
            atomicAdd(a + 5, m2 * m2);
          }
        }
      }
      set_barrier(set);            // the staging rows live in buf, which pass 0 overwrites
    }
    const int n_rounds = same ? 1 : 2;
#pragma unroll 1
    for (int rd = 0; rd < n_rounds; ++rd) {
      // same group: z1 = chain0 + i chain1, z2 = chain2 + i chain3.  Mixed groups: z1 = chain(2 rd), z2 = chain(2 rd + 1),
      // imaginary parts zero, so Re q and Im q are the two chains' own autocorrelations.
      const uint64_t ra = same ? r0 : (rd ? r2 : r0), rb = same ? r1 : 0ull;
      const uint64_t rc = same ? r2 : (rd ? r3 : r1), rdd = same ? r3 : 0ull;
      const int ia = same ? 0 : 2 * rd, ic = same ? 2 : 2 * rd + 1;
      const float aa = ia < nv ? 1.f : 0.f, ab = (same && nv > 1) ? 1.f : 0.f;
      const float ac = ic < nv ? 1.f : 0.f, ad = (same && nv > 3) ? 1.f : 0.f;
      if (aa == 0.f && ac == 0.f) continue;
      const int g_re = same ? g0 : (rd ? g2 : g0), g_im = same ? g0 : (rd ? g3 : g1);

      // six passes through ONE transform body (see ising_corr_fft.cuh); v stays in registers from pass 3 to pass 4
      cpx v[64];
#pragma unroll 1
      for (int pass = 0; pass < 6; ++pass) {
        if (pass == 0 || pass == 2) {
          const bool first = pass == 0;
          load_bits(v, first ? ra : rc, first ? aa : ac, first ? rb : rdd, first ? ab : ad);
        } else if (pass == 1 || pass == 3) load_col(t, sm.buf, v);
        else if (pass == 5) load_row(t, sm.buf, v);
        if (pass == 3) set_barrier(set);        // every column of buf has been read before pass 4 overwrites it
        fft64<false>(v);
        if (pass == 0 || pass == 2) {
          store_row(t, v, sm.buf);
          set_barrier(set);
        } else if (pass == 1) {
          store_power(t, v, sm.pw);             // thread-private column: no barrier needed for pw
          set_barrier(set);                     // every column of buf has been read before pass 2 overwrites it
        } else if (pass == 3) {
          build_q(t, v, sm.pw);
        } else if (pass == 4) {
          store_col(t, v, sm.buf);
          set_barrier(set);
        } else {
          // row t of buf was read by this thread only: no hazard with the next round's pass 0.  One binning pass when
          // the four chains share their group, else Re q -> the first chain's group and Im q -> the second's
          const int n_bp = same ? 1 : (ac != 0.f ? 2 : 1);
#pragma unroll 1
          for (int bp = 0; bp < n_bp; ++bp) {
            ensure_group(bp ? g_im : g_re);
            bin_row(t, v, (same || bp == 0) ? 1.f : 0.f, (same || bp == 1) ? 1.f : 0.f, bin_tab, sm.part, maxdev);
          }
        }
      }
      if (++acc_rounds >= kFlushRounds) flush();
    }
  }
  flush();
}

}  // namespace

// n_chains > 0, L == 64.  kObs form when block != nullptr.
int ising_corr_fft_launch(const uint32_t* bits, int64_t n_chains, const int32_t* group, long long* acc, float J, float h,
                          double* block, int block_stride, cudaStream_t st) {
  static int n_sm = 0;
  static bool attr_done = false;
  constexpr int smem = (int)(kSets * sizeof(SetSmem));
  if (!attr_done) {
    int dev = 0;
    DBS_CUDA(cudaGetDevice(&dev));
    DBS_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    DBS_CUDA(cudaFuncSetAttribute(ising_corr_fft_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    DBS_CUDA(cudaFuncSetAttribute(ising_corr_fft_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  const int64_t n_quads = (n_chains + 3) / 4;
  int64_t grid = (n_quads + kSets - 1) / kSets;
  if (grid > n_sm) grid = n_sm;
  if (block)
    ising_corr_fft_kernel<true><<<(unsigned)grid, kSets * kSetThreads, smem, st>>>(bits, n_chains, group, acc, -J, -h, block, block_stride);
  else
    ising_corr_fft_kernel<false><<<(unsigned)grid, kSets * kSetThreads, smem, st>>>(bits, n_chains, group, acc, 0.f, 0.f, nullptr, 0);
  DBS_LAUNCHED();
  return DBS_OK;
}

}  // namespace dbs
