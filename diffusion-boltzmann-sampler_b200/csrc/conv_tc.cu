// conv_tc.cu — implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a):
// tcgen05.mma with the accumulator in TMEM, operands staged in shared memory by TMA.
//
// Replaces nn.Conv2d / nn.ConvTranspose2d of the score U-Net (reference
// models/score_network.py:122-131 ConvBlock convs + 1x1 residual, :201 stride-2 downsample,
// :216 ConvTranspose 4x4 stride 2, with the skip concat of :256 read as one K range) and fuses the
// GroupNorm(8) sum / sum-of-squares (:124-125) into the epilogue.
//
// GEMM view (one launch = one ConvGeom):
//   D[128 pixels x Nb couts] += A[128 pixels x KC channels] * B[Nb couts x KC channels]^T
//   summed over k-blocks = (tap, channel chunk).  A is a TMA box {KC, Wt, 1, Ht, Nt} of the NHWC
//   activation shifted by the tap offset — out-of-image rows/columns are zero-filled by TMA, which
//   is the convolution's zero padding — and lands in shared memory as 128 rows of KC bf16, K-major,
//   in the 128/64/32-byte swizzle the UMMA descriptor names.  B is a box {KC, Nb, 1} of the
//   repacked weights [tap][Cout][Cin].  A stride-2 conv reads the same NHWC memory through a
//   space-to-depth view {2C, W/2, 2, H/2, N}; a transposed conv is four stride-1 phase convs whose
//   epilogue scatters to the odd/even output pixels.
//
// CTA = 18 warps, persistent over tiles (static round-robin):
//   warp 0   TMA producer (one lane)   — kStages-deep full/empty mbarrier ring
//   warp 1   MMA issuer  (one lane)    — tcgen05.mma cta_group::1, M=128, N=Nb, K=16 per instruction
//   warps 2-17 epilogue                — tcgen05.ld 32x32b -> +bias -> GN partial sums -> bf16 -> HBM; warp (q, sub)
//                                        drains TMEM lane quadrant q, chunk range sub (a warp's chain per chunk is
//                                        latency-bound, so four warps per quadrant cut the drain time of a tile ~4x)
// Two TMEM accumulator stages (2*Nb columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "tc_host.cuh"
#include "tc_ptx.cuh"
#include "unet.cuh"

namespace dbs {
namespace {

constexpr int kTileM = 128;
constexpr int kEpiWarps = 16;             // four warps per TMEM lane quadrant, each a contiguous range of 16-column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 8;
constexpr uint32_t kSmemBudget = 200 * 1024;
constexpr int kMaxCout = 512;

struct TcParams {
  int Wt, Ht, Nt;
  int tiles_w, tiles_h, tiles_n, n_tiles_m, n_tiles_n;
  int Nb, KC, chunks0, chunks1, n_taps, stages;
  int n_phases;       // 4: merged transposed-conv phases (tile index / tiles-per-phase = phase), else 1
  int B, Hout, Wout, Cout, out_stride, oph, opw;
  int groups, Cg, cg_log2;
  int epi_subs;       // epilogue warps per TMEM lane quadrant that take part (1 or 4), see conv_tc_launch
  // fused GroupNorm apply (whole images inside one tile): out = SiLU(GN(conv)) [+ time bias] [+ residual]
  int fuse;
  const float *gamma, *beta, *tb;
  int64_t tb_stride;   // floats between the time-bias rows of consecutive samples (0: all samples share one row)
  const __nv_bfloat16* res;
  float inv_cnt, eps;
  int tap_dw[16], tap_dh[16], tap_p[16], tap_c0[16], tap_c1[16], tap_w[16];
  const float* bias;
  __nv_bfloat16* out;
  float* stats;
  uint32_t idesc;
  uint32_t a_bytes, b_bytes, stage_bytes;
  int resident;           // all (tap, chunk) weight tiles stay in shared memory (loaded once per CTA); stages carry A only
  uint32_t w_region_bytes, b_pad;
  uint32_t tmem_cols;
  uint64_t desc_hi;  // descriptor bits above the start address (LBO, SBO, version, swizzle)
};

using namespace tc;

__device__ __forceinline__ float silu_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(0.5f, t, 0.5f);
}

// ------------------------------------------------------------------------------------------------
// kRun: 0 = no GroupNorm statistics, else min(channels per group, 16) in {2, 4, 8, 16}
// kPair: two CTAs of a cluster (one TPC) share every MMA (tcgen05 cta_group::2, M = 256): each CTA stages its own
// 128 pixels of A and HALF of the weight tile, so the weight traffic (L2 -> SM and shared-memory operand reads)
// per SM halves — the limiter of the 128/256-channel layers.  The leader CTA (cluster rank 0) issues the MMAs;
// both CTAs' TMA loads signal the leader's full barrier, tcgen05.commit multicasts to both CTAs' empty / tfull
// barriers, and the peer's epilogue warps hand the TMEM stage back on the leader's tempty barrier.
// Pins a loop-invariant value in a register: without it the single-thread roles re-read kernel parameters from the
// constant bank inside their per-k-block loops, a round trip on their critical path each (see conv_strip.cu).
__device__ __forceinline__ uint32_t pin(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ int pin(int v) { asm volatile("" : "+r"(v)); return v; }

template <int kRun, bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_in0, const __grid_constant__ CUtensorMap map_in1,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 5];
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_acc[2][kTileM * 8 * 2];  // GroupNorm partial sums, double-buffered by tile parity
  __shared__ __align__(16) float s_bias[kMaxCout];
  __shared__ __align__(16) float s_gamma[kRun == 16 ? kMaxCout : 4], s_beta[kRun == 16 ? kMaxCout : 4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();          // the next kernel of the step may begin its prologue as SMs free up (common.cuh)
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * kMaxStages + 2]);
  const uint32_t bar_w = smem_u32(&bars[2 * kMaxStages + 4]);     // resident weights landed
  const int rank = kPair ? (int)cluster_ctarank() : 0;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, (kPair ? 8 : 4) * p.epi_subs);
    }
    mbar_init(bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_in0);
    tma_prefetch_desc(&map_in1);
    tma_prefetch_desc(&map_w);
  }
  if (warp == 2) { if (kPair) tmem_alloc2(smem_u32(&s_tmem_base), p.tmem_cols); else tmem_alloc(smem_u32(&s_tmem_base), p.tmem_cols); }
  for (int i = threadIdx.x; i < 2 * kTileM * 8 * 2; i += kThreads) (&s_acc[0][0])[i] = 0.f;
  for (int i = threadIdx.x; i < p.Cout; i += kThreads) s_bias[i] = p.bias[i];
  if (kRun == 16 && p.fuse)
    for (int i = threadIdx.x; i < p.Cout; i += kThreads) { s_gamma[i] = p.gamma[i]; s_beta[i] = p.beta[i]; }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  // everything above touches only this layer's constants (bias, gamma / beta, tensor maps) and on-chip state; from here on the
  // kernel reads activations and statistics buffers that the previous kernel of the stream writes
  pdl_wait();

  // work units: tiles, or pairs of M-adjacent tiles (this CTA takes tile 2k + rank of the pair)
  const int half_m = p.n_tiles_m >> 1;
  const int tiles_per_phase = p.n_tiles_m * p.n_tiles_n;
  const int total_tiles = kPair ? half_m * p.n_tiles_n : tiles_per_phase * p.n_phases;
  const int unit0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tile_of = [&](int u) { return kPair ? (u / half_m) * p.n_tiles_m + 2 * (u % half_m) + rank : u; };
  const int kblocks_per_tap = p.chunks0 + p.chunks1;
  const int n_kblocks = p.n_taps * kblocks_per_tap;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t b_rows_off = kPair ? (uint32_t)rank * (uint32_t)(p.Nb >> 1) : 0u;   // my half of the weight tile
      const uint32_t ring_base = pin(smem_base + p.w_region_bytes);
      const int KC = pin(p.KC), chunks0 = pin(p.chunks0), n_stages = pin(p.stages), n_taps = pin(p.n_taps);
      const int resident = pin(p.resident), Nb = pin(p.Nb);
      const uint32_t stage_bytes = pin(p.stage_bytes), b_off = pin((p.a_bytes + 1023u) & ~1023u);
      const uint32_t tx_bytes = pin(kPair ? 2u * (p.a_bytes + p.b_bytes) : p.a_bytes + (p.resident ? 0u : p.b_bytes));
      if (!kPair && p.resident) {
        // the layer's weights, once: k-block kb = tap * chunks + chunk at w_region + kb * b_pad
        mbar_expect_tx(bar_w, p.b_bytes * (uint32_t)n_kblocks);
        for (int t = 0; t < p.n_taps; ++t)
          for (int ch = 0; ch < kblocks_per_tap; ++ch)
            tma_load_3d(smem_base + (uint32_t)(t * kblocks_per_tap + ch) * p.b_pad, &map_w, bar_w, ch * p.KC, 0, p.tap_w[t]);
      }
      for (int u = unit0; u < total_tiles; u += unit_stride) {
        const int tile_ph = tile_of(u);
        const int ph4 = tile_ph / tiles_per_phase, tile = tile_ph - ph4 * tiles_per_phase;
        const int nt = tile / p.n_tiles_m, mt = tile - nt * p.n_tiles_m;
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.Wt, h0 = th * p.Ht, n0 = tn * p.Nt, co0 = nt * Nb + (int)b_rows_off;
        for (int t = ph4 * n_taps; t < (ph4 + 1) * n_taps; ++t) {
          const int cw = w0 + p.tap_dw[t], chh = h0 + p.tap_dh[t], cp = p.tap_p[t];
          const int tc0 = p.tap_c0[t], tc1 = p.tap_c1[t], tw_ = p.tap_w[t];     // per tap, not per k-block
          for (int ch = 0; ch < kblocks_per_tap; ++ch) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t a_dst = (kPair ? smem_base : ring_base) + stage * stage_bytes;
            const uint32_t b_dst = a_dst + b_off;
            const uint32_t fb = bar_full + 8 * stage;
            if (kPair) {
              // both CTAs' bytes land on the LEADER's barrier (peer bit of the address cleared); the leader posts the
              // expected total once
              if (rank == 0) mbar_expect_tx(fb, tx_bytes);
              const uint32_t fbl = fb & 0xFEFFFFFFu;
              if (ch < chunks0) tma_load_5d_2sm(a_dst, &map_in0, fbl, tc0 + ch * KC, cw, cp, chh, n0);
              else tma_load_5d_2sm(a_dst, &map_in1, fbl, tc1 + (ch - chunks0) * KC, cw, cp, chh, n0);
              tma_load_3d_2sm(b_dst, &map_w, fbl, ch * KC, co0, tw_);
            } else {
              mbar_expect_tx(fb, tx_bytes);
              if (ch < chunks0) tma_load_5d(a_dst, &map_in0, fb, tc0 + ch * KC, cw, cp, chh, n0);
              else tma_load_5d(a_dst, &map_in1, fb, tc1 + (ch - chunks0) * KC, cw, cp, chh, n0);
              if (!resident) tma_load_3d(b_dst, &map_w, fb, ch * KC, co0, tw_);
            }
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===================== MMA issuer (leader CTA only in pair mode) =====================
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const int mmas = pin(p.KC / 16);
      const uint32_t lo_bits = pin((uint32_t)p.desc_hi), hi32 = pin((uint32_t)(p.desc_hi >> 32)), idesc = pin(p.idesc);
      const uint32_t b_off = pin((p.a_bytes + 1023u) & ~1023u);
      const uint32_t ring_base = pin(smem_base + (kPair ? 0u : p.w_region_bytes));
      const uint32_t stage_bytes = pin(p.stage_bytes), b_pad = pin(p.b_pad);
      const int n_stages = pin(p.stages), acc_stride = pin(p.Nb), n_kb = pin(n_kblocks);
      const bool w_resident = !kPair && p.resident;
      if (w_resident) {
        mbar_wait(bar_w, 0);
        tc_fence_after();
      }
      for (int u = unit0; u < total_tiles; u += unit_stride, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * acc_stride;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_addr = ring_base + stage * stage_bytes;
          const uint32_t b_addr = w_resident ? smem_base + (uint32_t)kb * b_pad : a_addr + b_off;
          const uint32_t a_lo = lo_bits | ((a_addr & 0x3FFFFu) >> 4);
          const uint32_t b_lo = lo_bits | ((b_addr & 0x3FFFFu) >> 4);
          const uint32_t acc0 = kb ? 1u : 0u;
          if (kPair) {
            if (mmas == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_lohi_2sm(d_tmem, a_lo + 2 * k, b_lo + 2 * k, hi32, idesc, k ? 1u : acc0);
            } else {
              for (int k = 0; k < mmas; ++k) umma_bf16_lohi_2sm(d_tmem, a_lo + 2 * k, b_lo + 2 * k, hi32, idesc, k ? 1u : acc0);
            }
            umma_commit_2sm(bar_empty + 8 * stage);   // both CTAs' slots
          } else {
            if (mmas == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, hi32, idesc, k ? 1u : acc0);
            } else if (mmas == 2) {
#pragma unroll
              for (int k = 0; k < 2; ++k) umma_bf16_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, hi32, idesc, k ? 1u : acc0);
            } else {
              umma_bf16_lohi(d_tmem, a_lo, b_lo, hi32, idesc, acc0);
            }
            umma_commit(bar_empty + 8 * stage);  // frees the smem slot when these MMAs retire
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        if (kPair) umma_commit_2sm(bar_tfull + 8 * as); else umma_commit(bar_tfull + 8 * as);   // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else if (((warp - 2) >> 2) < p.epi_subs) {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                 // TMEM lane quarter accessible to this warp
    const int sub = (warp - 2) >> 2;        // which of the quadrant's four warps
    const int n_chunks = p.Nb >> 4;
    const int per = (n_chunks + p.epi_subs - 1) / p.epi_subs;    // contiguous chunks per warp
    const int c_lo = sub * per * 16, c_hi = min(p.Nb, (sub + 1) * per * 16);
    const int row = q * 32 + lane;          // tile-local GEMM row = pixel
    const int ppi = p.Ht * p.Wt;            // pixels per image inside a tile
    const int nl = row / ppi, hl = (row / p.Wt) % p.Ht, wl = row % p.Wt;
    const int etid = threadIdx.x - 64;      // 0 .. 32*kEpiWarps-1
    const bool warp_uniform_image = ppi >= 32;
    int it = 0;
    for (int u = unit0; u < total_tiles; u += unit_stride, ++it) {
      const int tile_ph = tile_of(u);
      const int ph4 = tile_ph / tiles_per_phase, tile = tile_ph - ph4 * tiles_per_phase;
      const int oph = p.n_phases > 1 ? (ph4 >> 1) : p.oph, opw = p.n_phases > 1 ? (ph4 & 1) : p.opw;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int nt = tile / p.n_tiles_m, mt = tile - nt * p.n_tiles_m;
      const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
      const int n = tn * p.Nt + nl, a = th * p.Ht + hl, b = tw * p.Wt + wl, co0 = nt * p.Nb;
      const bool valid = n < p.B;
      __nv_bfloat16* optr =
          p.out + (((int64_t)n * p.Hout + (a * p.out_stride + oph)) * p.Wout + (b * p.out_stride + opw)) * p.Cout + co0;
      float* acc = &s_acc[it & 1][nl * 16];

      mbar_wait(bar_tfull + 8 * as, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * p.Nb;
      float run_s = 0.f, run_ss = 0.f;
      uint32_t v[16];
      if (c_lo < c_hi) tmem_ld16(taddr + c_lo, v);
      for (int c = c_lo; c < c_hi; c += 16) {
        tmem_ld_wait();
        float f[16];
        const float4* bq = reinterpret_cast<const float4*>(s_bias + co0 + c);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b4 = bq[j4];
          f[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) + b4.x;
          f[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + b4.y;
          f[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + b4.z;
          f[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + b4.w;
        }
        if (c + 16 < c_hi) tmem_ld16(taddr + c + 16, v);   // next chunk in flight behind this one's arithmetic
        if (kRun > 0) {
          if (!valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = 0.f;
          }
          // GroupNorm partial sums: a run = the columns of this chunk that share a group
          constexpr int R = kRun > 0 ? kRun : 16;
#pragma unroll
          for (int r = 0; r < 16 / R; ++r) {
            float s = 0.f, ss = 0.f;
#pragma unroll
            for (int j = 0; j < R; ++j) {
              s += f[r * R + j];
              ss = fmaf(f[r * R + j], f[r * R + j], ss);
            }
            bool flush = true;
            if (R == 16) {
              run_s += s; run_ss += ss;
              s = run_s; ss = run_ss;
              // warp-uniform: end of a group of >= 16 channels, or of this warp's chunk range (partial sum)
              flush = ((c + 16) & (p.Cg - 1)) == 0 || c + 16 >= c_hi;
              if (flush) { run_s = 0.f; run_ss = 0.f; }
            }
            if (flush) {
              const int gi = ((co0 + c) >> p.cg_log2) + (R == 16 ? 0 : r);
              if (warp_uniform_image) {
                s = warp_sum(s); ss = warp_sum(ss);
                if (lane == 0) { atomicAdd(&acc[gi * 2 + 0], s); atomicAdd(&acc[gi * 2 + 1], ss); }
              } else {
                atomicAdd(&acc[gi * 2 + 0], s); atomicAdd(&acc[gi * 2 + 1], ss);
              }
            }
          }
        }
        if (valid && !(kRun == 16 && p.fuse)) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          st_global_256(optr + c, pk);   // sixteen couts of the pixel: one 32-byte sector (tc_ptx.cuh)
        }
      }
      if (kRun == 16 && p.fuse) {
        // ---- fused GroupNorm apply (whole images inside the tile): after the barrier the statistics of the
        // tile's images are final; a second pass over the TMEM stage normalises the fp32 accumulators, applies
        // SiLU (+ time bias / + residual) and stores the block's activation — neither the raw conv output nor
        // a separate GroupNorm kernel exists for these layers.  (Keeping the first pass's outputs in registers
        // and releasing TMEM early was measured slower: the drain, not the MMA, is the longer side here.)
        asm volatile("bar.sync 1, %0;" ::"r"(128 * p.epi_subs) : "memory");
        const float* tbp = p.tb ? p.tb + (int64_t)(valid ? n : 0) * p.tb_stride + co0 : nullptr;
        const __nv_bfloat16* rptr = p.res ? p.res + (optr - p.out) : nullptr;
        if (c_lo < c_hi) tmem_ld16(taddr + c_lo, v);
        for (int c = c_lo; c < c_hi; c += 16) {
          tmem_ld_wait();
          float f[16];
          const int gi = (co0 + c) >> p.cg_log2;
          const float mean = acc[gi * 2 + 0] * p.inv_cnt;
          const float rstd = rsqrtf(fmaxf(acc[gi * 2 + 1] * p.inv_cnt - mean * mean, 0.f) + p.eps);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[co0 + c + 4 * j4]);
            const float4 g4 = *reinterpret_cast<const float4*>(&s_gamma[co0 + c + 4 * j4]);
            const float4 e4 = *reinterpret_cast<const float4*>(&s_beta[co0 + c + 4 * j4]);
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tbp) t4 = __ldg(reinterpret_cast<const float4*>(tbp + c + 4 * j4));
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
            const float ee[4] = {e4.x, e4.y, e4.z, e4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float a = rstd * gg[k];
              f[4 * j4 + k] = silu_fast(fmaf(__uint_as_float(v[4 * j4 + k]) + bb[k], a, ee[k] - mean * a)) + tt[k];
            }
          }
          if (c + 16 < c_hi) tmem_ld16(taddr + c + 16, v);
          if (valid) {
            if (rptr) {
              const uint4* rq = reinterpret_cast<const uint4*>(rptr + c);
#pragma unroll
              for (int hq = 0; hq < 2; ++hq) {
                const uint4 r4 = rq[hq];
                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[i]);
                  f[8 * hq + 2 * i] = fmaf(1.0f, __low2float(h2), f[8 * hq + 2 * i]);
                  f[8 * hq + 2 * i + 1] = fmaf(1.0f, __high2float(h2), f[8 * hq + 2 * i + 1]);
                }
              }
            }
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            st_global_256(optr + c, pk);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (kPair) mbar_arrive_leader(bar_tempty + 8 * as); else mbar_arrive(bar_tempty + 8 * as); }
        // everyone has read the statistics: clear this buffer for the tile after next
        asm volatile("bar.sync 1, %0;" ::"r"(128 * p.epi_subs) : "memory");
        float* accb = &s_acc[it & 1][0];
        for (int i = etid; i < p.Nt * 16; i += 128 * p.epi_subs) accb[i] = 0.f;
        continue;
      }
      // TMEM stage drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (kPair) mbar_arrive_leader(bar_tempty + 8 * as); else mbar_arrive(bar_tempty + 8 * as); }
      if (kRun > 0) {
        asm volatile("bar.sync 1, %0;" ::"r"(128 * p.epi_subs) : "memory");
        float* accb = &s_acc[it & 1][0];
        for (int i = etid; i < p.Nt * 16; i += 128 * p.epi_subs) {
          const int img = i >> 4, rest = i & 15, gi = rest >> 1, k = rest & 1;
          const float vsum = accb[i];
          accb[i] = 0.f;
          const int ng = tn * p.Nt + img;
          if (ng < p.B && gi < p.groups && vsum != 0.f) atomicAdd(p.stats + ((int64_t)ng * p.groups + gi) * 2 + k, vsum);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();     // the leader's MMAs read the peer's shared memory; remote arrivals need a live CTA
  if (warp == 2) {
    tc_fence_after();
    if (kPair) tmem_dealloc2(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side (shared with conv_strip.cu through tc_host.cuh)
namespace tch {

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// activation map: stride 1 -> {C, W, 1, H, N}; stride 2 -> space-to-depth view {2C, W/2, 2, H/2, N}
int encode_act(EncodeTiledFn enc, CUtensorMap* m, const void* ptr, int C, int H, int W, int B, int stride, int KC, int Wt,
               int Ht, int Nt, CUtensorMapSwizzle sw) {
  cuuint64_t dims[5], strides[4];
  const cuuint64_t e = 2;
  if (stride == 1) {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = B;
    strides[0] = C * e; strides[1] = (cuuint64_t)W * C * e; strides[2] = (cuuint64_t)W * C * e;
    strides[3] = (cuuint64_t)H * W * C * e;
  } else {
    dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = B;
    strides[0] = 2 * C * e; strides[1] = (cuuint64_t)W * C * e; strides[2] = (cuuint64_t)2 * W * C * e;
    strides[3] = (cuuint64_t)H * W * C * e;
  }
  cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)Wt, 1u, (cuuint32_t)Ht, (cuuint32_t)Nt};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_tc: cuTensorMapEncodeTiled(activation C=%d H=%d W=%d B=%d stride=%d) failed with %d", C, H, W, B,
              stride, (int)r);
    return DBS_ERR_CUDA;
  }
  return DBS_OK;
}

int encode_w(EncodeTiledFn enc, CUtensorMap* m, const void* ptr, int Cin, int Cout, int taps, int KC, int Nb,
             CUtensorMapSwizzle sw) {
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * Cout * 2};
  cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)Nb, 1u};
  cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_tc: cuTensorMapEncodeTiled(weights Cin=%d Cout=%d taps=%d) failed with %d", Cin, Cout, taps, (int)r);
    return DBS_ERR_CUDA;
  }
  return DBS_OK;
}

}  // namespace tch

using namespace tch;

int conv_tc_launch(const ConvGeom& g, const __nv_bfloat16* in0, const __nv_bfloat16* in1, const __nv_bfloat16* w,
                   const float* bias, __nv_bfloat16* out, float* stats, int groups, cudaStream_t st, const GnFuse* fuse) {
  // full-resolution small-channel layers go to the strip kernel (conv_strip.cu); it has no fused GroupNorm apply
  if (!fuse) {
    const int rc_strip = conv_strip_launch(g, in0, in1, w, bias, out, stats, groups, st);
    if (rc_strip != DBS_ERR_UNSUPPORTED) return rc_strip;
  }
  // ---- eligibility
  const int Cin = g.C0 + g.C1;
  int KC = 0;
  for (int k : {64, 32, 16})
    if (g.C0 % k == 0 && g.C1 % k == 0) { KC = k; break; }
  if (!KC || g.C0 == 0) return DBS_ERR_UNSUPPORTED;
  if (g.Cout % 16 != 0) return DBS_ERR_UNSUPPORTED;
  const int n_tiles_n = (g.Cout + 255) / 256;
  if (g.Cout % n_tiles_n != 0 || (g.Cout / n_tiles_n) % 16 != 0) return DBS_ERR_UNSUPPORTED;
  const int Nb = g.Cout / n_tiles_n;
  if (g.in_stride == 2 && ((g.Hin & 1) || (g.Win & 1) || g.out_stride != 1)) return DBS_ERR_UNSUPPORTED;
  if (g.in_stride != 1 && g.in_stride != 2) return DBS_ERR_UNSUPPORTED;
  // pixel box {Wt, Ht, Nt} with Wt*Ht*Nt == 128
  int Wt, Ht, Nt;
  if (g.Wm >= kTileM) {
    if (g.Wm % kTileM) return DBS_ERR_UNSUPPORTED;
    Wt = kTileM; Ht = 1; Nt = 1;
  } else {
    if (!is_pow2(g.Wm)) return DBS_ERR_UNSUPPORTED;
    Wt = g.Wm;
    const int rows = kTileM / Wt;
    if (g.Hm >= rows) {
      if (g.Hm % rows) return DBS_ERR_UNSUPPORTED;
      Ht = rows; Nt = 1;
    } else {
      if (!is_pow2(g.Hm)) return DBS_ERR_UNSUPPORTED;
      Ht = g.Hm; Nt = rows / Ht;
    }
  }
  if (groups > 8 || (stats && (groups < 1 || g.Cout % groups))) return DBS_ERR_UNSUPPORTED;
  if (g.Cout > kMaxCout) return DBS_ERR_UNSUPPORTED;
  if (stats && (!is_pow2(g.Cout / groups) || g.Cout / groups < 2)) return DBS_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(in0) | reinterpret_cast<uintptr_t>(in1) | reinterpret_cast<uintptr_t>(w)) & 15)
    return DBS_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(out) & 31) || g.Cout % 16) return DBS_ERR_UNSUPPORTED;   // 32-byte epilogue stores
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled entry point unavailable"); return DBS_ERR_CUDA; }

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.Wt = Wt; p.Ht = Ht; p.Nt = Nt;
  p.tiles_w = g.Wm / Wt; p.tiles_h = g.Hm / Ht; p.tiles_n = (g.B + Nt - 1) / Nt;
  p.n_tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  p.n_tiles_n = n_tiles_n;
  p.Nb = Nb; p.KC = KC; p.chunks0 = g.C0 / KC; p.chunks1 = g.C1 / KC; p.n_taps = g.n_taps;
  p.n_phases = g.n_phases > 1 ? g.n_phases : 1;
  if (p.n_phases > 1 && (p.n_phases != 4 || g.n_taps != 4 || g.out_stride != 2 || g.in_stride != 1 || stats || fuse))
    return DBS_ERR_UNSUPPORTED;
  p.B = g.B; p.Hout = g.Hout; p.Wout = g.Wout; p.Cout = g.Cout;
  p.out_stride = g.out_stride; p.oph = g.oph; p.opw = g.opw;
  p.groups = groups; p.Cg = stats ? g.Cout / groups : g.Cout;
  p.cg_log2 = 0;
  while ((1 << p.cg_log2) < p.Cg) ++p.cg_log2;
  for (int t = 0; t < g.n_taps * p.n_phases; ++t) {
    if (g.in_stride == 1) {
      p.tap_dw[t] = g.dw[t]; p.tap_dh[t] = g.dh[t]; p.tap_p[t] = 0; p.tap_c0[t] = 0; p.tap_c1[t] = 0;
    } else {
      // input row 2a+dh, dh in {-1,0,1}: (row pair, parity) = (a-1, 1) | (a, 0) | (a, 1); same for columns,
      // where the column parity selects the upper or lower C channels of the 2C-wide inner dimension
      p.tap_dh[t] = g.dh[t] < 0 ? -1 : 0;
      p.tap_p[t] = g.dh[t] < 0 ? 1 : g.dh[t];
      p.tap_dw[t] = g.dw[t] < 0 ? -1 : 0;
      const int pw = g.dw[t] < 0 ? 1 : g.dw[t];
      p.tap_c0[t] = pw * g.C0; p.tap_c1[t] = pw * g.C1;
    }
    p.tap_w[t] = g.widx[t];
  }
  // Four warps per quadrant pay off only when a warp's rows span several images (tiny lattices: per-thread
  // shared-memory atomics for the GroupNorm sums dominate the drain); measured slower elsewhere.
  p.epi_subs = (stats && (Ht * Wt < 32 || fuse)) ? 4 : 1;
  if (fuse) {
    // whole images inside one tile, all couts in one N tile, groups of >= 16 channels (kRun == 16 variant)
    if (!stats || Wt != g.Wm || Ht != g.Hm || n_tiles_n != 1 || g.out_stride != 1 || p.Cg < 16 || !fuse->gamma || !fuse->beta)
      return DBS_ERR_UNSUPPORTED;
    if (fuse->res && (reinterpret_cast<uintptr_t>(fuse->res) & 15)) return DBS_ERR_UNSUPPORTED;
    p.fuse = 1; p.gamma = fuse->gamma; p.beta = fuse->beta; p.tb = fuse->tb; p.tb_stride = fuse->tb_stride; p.res = fuse->res;
    p.inv_cnt = 1.0f / (float)((int64_t)g.Hm * g.Wm * p.Cg); p.eps = 1e-5f;
  }
  p.bias = bias; p.out = out; p.stats = stats;
  // 2-CTA mode (cta_group::2): layers with >= 128 couts, where the weight tile dominates the operand traffic
  static const int pair_mode = [] { const char* e = getenv("DBS_CONV_PAIR"); return e ? atoi(e) : 1; }();
  // measured (profiles/pair_sweep.sh, B = 1024): 256-cout 8x8 layers gain 5-14 %; 128-cout layers, 1x1 convs,
  // transposed phases and 4x4 lattices are neutral or slower, so the pair mode is used where the weight tile is the
  // larger operand (DBS_CONV_PAIR=2 forces it wherever it is legal, =0 disables it)
  const bool pair_legal = p.n_phases == 1 && Nb >= 128 && Nb % 32 == 0 && (p.n_tiles_m % 2 == 0) && p.n_tiles_m >= 2;
  const bool pair = pair_legal && (pair_mode == 2 || (pair_mode == 1 && Nb == 256 && g.n_taps == 9 && g.Hm * g.Wm >= 64));
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Nb >> 3) << 17) | ((uint32_t)((pair ? 2 * kTileM : kTileM) >> 4) << 24);
  p.a_bytes = (uint32_t)kTileM * KC * 2;
  p.b_bytes = (uint32_t)(pair ? Nb / 2 : Nb) * KC * 2;
  p.b_pad = (p.b_bytes + 1023u) & ~1023u;
  p.stage_bytes = ((p.a_bytes + 1023u) & ~1023u) + p.b_pad;
  // Resident weights: when every (tap, chunk) weight tile of the layer fits beside >= 4 A stages, the weights are loaded
  // once per CTA and a k-block brings only its activation tile (stride-2 64-channel conv, 1x1 residual convs).
  // Measured gain is small — 59.8 -> 57.8 us for the 32x32 stride-2 conv, 24.5 -> 23.0 us for the 8x8 1x1 conv at
  // B = 1024 — these layers are not bound by operand bytes per k-block; kept because it is never slower.
  static const bool res_off = [] { const char* e = getenv("DBS_TILE_RESIDENT"); return e && e[0] == '0'; }();
  {
    const uint32_t a_pad = (p.a_bytes + 1023u) & ~1023u;
    const uint32_t all_w = p.b_pad * (uint32_t)(g.n_taps * (p.chunks0 + p.chunks1));
    if (!res_off && !pair && n_tiles_n == 1 && p.n_phases == 1 && all_w + 4 * a_pad <= kSmemBudget) {
      p.resident = 1;
      p.w_region_bytes = all_w;
      p.stage_bytes = a_pad;
    }
  }
  int stages = (int)((kSmemBudget - p.w_region_bytes) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return DBS_ERR_UNSUPPORTED;
  p.stages = stages;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * Nb)) cols <<= 1;
  p.tmem_cols = cols;
  const uint32_t row_bytes = (uint32_t)KC * 2;
  const uint64_t layout = KC == 64 ? 2ull : (KC == 32 ? 4ull : 6ull);  // SWIZZLE_128B / 64B / 32B
  p.desc_hi = (1ull << 16) | ((uint64_t)((8 * row_bytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
  const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                         : (KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);

  CUtensorMap m0, m1, mw;
  int rc;
  if ((rc = encode_act(enc, &m0, in0, g.C0, g.Hin, g.Win, g.B, g.in_stride, KC, Wt, Ht, Nt, sw))) return rc;
  if (g.C1 > 0) {
    if ((rc = encode_act(enc, &m1, in1, g.C1, g.Hin, g.Win, g.B, g.in_stride, KC, Wt, Ht, Nt, sw))) return rc;
  } else {
    m1 = m0;
  }
  // weights [taps_total][Cout][Cin]: the tap dimension is addressed by widx, so size it generously
  int taps_total = 0;
  for (int t = 0; t < g.n_taps; ++t) taps_total = g.widx[t] + 1 > taps_total ? g.widx[t] + 1 : taps_total;
  if (g.out_stride == 2) taps_total = 16;
  if ((rc = encode_w(enc, &mw, w, Cin, g.Cout, taps_total, KC, pair ? Nb / 2 : Nb, sw))) return rc;

  const size_t smem = (size_t)p.w_region_bytes + (size_t)stages * p.stage_bytes + 1024;
  const int run = !stats ? 0 : (p.Cg >= 16 ? 16 : p.Cg);
  typedef void (*Kern)(CUtensorMap, CUtensorMap, CUtensorMap, TcParams);
  static const Kern table[2][5] = {
      {conv_tc_kernel<0, false>, conv_tc_kernel<2, false>, conv_tc_kernel<4, false>, conv_tc_kernel<8, false>, conv_tc_kernel<16, false>},
      {conv_tc_kernel<0, true>, conv_tc_kernel<2, true>, conv_tc_kernel<4, true>, conv_tc_kernel<8, true>, conv_tc_kernel<16, true>}};
  const int ri = run == 0 ? 0 : run == 2 ? 1 : run == 4 ? 2 : run == 8 ? 3 : 4;
  Kern kern = table[pair ? 1 : 0][ri];
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 5; ++b) {
        const cudaError_t e = cudaFuncSetAttribute(table[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemBudget + 2048));
        if (e != cudaSuccess) attr_err = e;
      }
  });
  DBS_CUDA(attr_err);
  const int total_units = pair ? (p.n_tiles_m / 2) * p.n_tiles_n : p.n_tiles_m * p.n_tiles_n * p.n_phases;
  const int sms = num_sms();
  prof_set_label("tile%s H%d W%d C%d+%d->%d taps%d s%d os%d KC%d B%d%s%s", pair ? "2" : "", g.Hin, g.Win, g.C0, g.C1, g.Cout,
                 g.n_taps, g.in_stride, g.out_stride, KC, g.B, fuse ? " +gn" : "", p.n_phases > 1 ? " x4 phases" : (p.resident ? " wres" : ""));
  ProfScope prof(kProfConvTC, st);
  if (pair) {
    const int pairs = total_units < sms / 2 ? total_units : sms / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    DBS_CUDA(cudaLaunchKernelEx(&cfg, kern, m0, m1, mw, p));
  } else {
    const int grid = total_units < sms ? total_units : sms;
    DBS_CUDA(launch_chained(kern, dim3((unsigned)grid), dim3(kThreads), smem, st, m0, m1, mw, p));
  }
  DBS_LAUNCHED();
  return DBS_OK;
}

}  // namespace dbs

using namespace dbs;

extern "C" int dbs_conv_tc(const void* in0, int C0, const void* in1, int C1, const void* weight, const float* bias,
                           void* out, float* stats, int64_t B, int H, int W, int C_out, int kind, int ksize,
                           void* stream) {
  DBS_REQUIRE(in0 && weight && bias && out && B >= 1 && H >= 1 && W >= 1, "conv_tc: bad arguments");
  DBS_REQUIRE(kind >= 0 && kind <= 2, "conv_tc: kind must be 0 (stride 1), 1 (3x3 stride 2) or 2 (transposed 4x4)");
  const int groups = C_out < 8 ? C_out : 8;
  int rc = DBS_OK;
  if (kind == 0) {
    DBS_REQUIRE(ksize == 1 || ksize == 3, "conv_tc: ksize must be 1 or 3");
    rc = conv_tc_launch(geom_conv((int)B, H, W, C0, C1, C_out, ksize, 1), (const __nv_bfloat16*)in0,
                        (const __nv_bfloat16*)in1, (const __nv_bfloat16*)weight, bias, (__nv_bfloat16*)out, stats, groups,
                        as_stream(stream));
  } else if (kind == 1) {
    rc = conv_tc_launch(geom_conv((int)B, H, W, C0, C1, C_out, 3, 2), (const __nv_bfloat16*)in0,
                        (const __nv_bfloat16*)in1, (const __nv_bfloat16*)weight, bias, (__nv_bfloat16*)out, stats, groups,
                        as_stream(stream));
  } else {
    DBS_REQUIRE(C1 == 0, "conv_tc: transposed conv takes a single source");
    // same policy as the U-Net executor: strip kernel (horizontal phase pairs, else phase by phase) where it is
    // eligible, else the tile kernel with the four phases merged into one launch (statistics: per-phase launches only)
    const __nv_bfloat16 *xin = (const __nv_bfloat16*)in0, *wq = (const __nv_bfloat16*)weight;
    int first = 0;   // phases [first, 4) still to run one by one
    rc = stats ? DBS_ERR_UNSUPPORTED
               : conv_strip_launch(geom_convT_pair((int)B, H, W, C0, C_out, 0), xin, nullptr, wq, bias, (__nv_bfloat16*)out,
                                   nullptr, groups, as_stream(stream));
    if (rc == DBS_OK) {
      rc = conv_strip_launch(geom_convT_pair((int)B, H, W, C0, C_out, 1), xin, nullptr, wq, bias, (__nv_bfloat16*)out, nullptr,
                             groups, as_stream(stream));
      first = 4;
    } else if (rc == DBS_ERR_UNSUPPORTED) {
      rc = conv_strip_launch(geom_convT_phase((int)B, H, W, C0, C_out, 0, 0), xin, nullptr, wq, bias, (__nv_bfloat16*)out, stats,
                             groups, as_stream(stream));
      if (rc == DBS_OK) {
        first = 1;
      } else if (rc == DBS_ERR_UNSUPPORTED) {
        rc = stats ? DBS_ERR_UNSUPPORTED
                   : conv_tc_launch(geom_convT_all((int)B, H, W, C0, C_out), xin, nullptr, wq, bias, (__nv_bfloat16*)out, nullptr,
                                    groups, as_stream(stream));
        if (rc == DBS_OK) first = 4;
        else if (rc == DBS_ERR_UNSUPPORTED) rc = DBS_OK;
      }
    }
    for (int ph4 = first; ph4 < 4 && rc == DBS_OK; ++ph4)
      rc = conv_tc_launch(geom_convT_phase((int)B, H, W, C0, C_out, ph4 >> 1, ph4 & 1), xin, nullptr, wq, bias,
                          (__nv_bfloat16*)out, stats, groups, as_stream(stream));
  }
  if (rc == DBS_ERR_UNSUPPORTED) set_error("conv_tc: this geometry is not eligible for the tcgen05 kernel");
  return rc;
}

// Parity port for the fused (3x3 conv + 1x1 residual conv of the same input) pass of the strip kernel.
extern "C" int dbs_conv_block_tc(const void* in0, int C0, const void* in1, int C1, const void* w3, const float* bias3,
                                 const void* w1, const float* bias1, void* out3, void* out1, float* stats, int64_t B,
                                 int H, int W, int C_out, void* stream) {
  DBS_REQUIRE(in0 && w3 && bias3 && w1 && bias1 && out3 && out1 && B >= 1, "conv_block_tc: bad arguments");
  const int groups = C_out < 8 ? C_out : 8;
  const int rc = conv_strip_launch(geom_conv((int)B, H, W, C0, C1, C_out, 3, 1), (const __nv_bfloat16*)in0,
                                   (const __nv_bfloat16*)in1, (const __nv_bfloat16*)w3, bias3, (__nv_bfloat16*)out3, stats,
                                   groups, as_stream(stream), (const __nv_bfloat16*)w1, bias1, (__nv_bfloat16*)out1);
  if (rc == DBS_ERR_UNSUPPORTED) set_error("conv_block_tc: this geometry is not eligible for the fused strip kernel");
  return rc;
}
