// conv_strip.cu — "strip" variant of the tcgen05 implicit-GEMM convolution for the full-resolution,
// small-channel layers of the score U-Net (stride-1 convs and transposed-conv phases, Cout <= 64).
//
// Why a second kernel: conv_tc.cu re-fetches the activation tile once per tap and the weights once
// per tile; ncu shows those layers bound by L2->SM request rate (~4 cycles per TMA row, tensor pipe
// < 20 %).  Here a tile is Ht full-width image rows (Ht*W = 128 pixels) and, per channel chunk, ONE
// TMA box brings the rows h0+dh_min .. h0+Ht-1+dh_max (zero-filled outside the image).  Every tap
// then reads the same strip:
//   * a vertical tap offset dh is a start-address offset of dh*W rows in the UMMA descriptor
//     (a multiple of the 8-row swizzle atom because W % 8 == 0, so no descriptor base offset);
//   * a horizontal offset dw in {-1,0,+1} is NOT applied to the operand.  Taps are accumulated
//     into one TMEM accumulator per dw, P_dw[m] = sum_{dh,c} X[m + dh*W][c] * Wt[dh,dw][c], and the
//     epilogue combines  out[m] = P_0[m] + P_{+1}[m+1]*[w<W-1] + P_{-1}[m-1]*[w>0]  with a lane
//     shuffle (rows of the TMEM tile are pixels, so pixel m+-1 is lane m+-1; the mask is the
//     convolution's zero padding at the left/right image border).
// The layer's weights ([chunk][tap] slabs of Nb x KC, K-major, swizzled) are loaded once per CTA and
// stay in shared memory when they fit; otherwise a chunk's slabs travel with its strip.
// Activation traffic drops from 9 tiles to (Ht+2)/Ht strips per chunk; weight traffic to ~zero.
//
// Warp roles, barriers, TMEM double buffering and the GroupNorm epilogue are those of conv_tc.cu.
#include <stdlib.h>

#include <mutex>

#include "tc_host.cuh"
#include "tc_ptx.cuh"
#include "unet.cuh"

namespace dbs {
namespace {

using namespace tc;
using namespace tch;

constexpr int kTileM = 128;
// Timing experiments of round 2 (handshake toggles 128 / 256 / 512 / 1024 / 2048 of DBS_STRIP_DBG, epilogue-side stage
// release DBS_STRIP_EPI_RELEASE) are compiled in only with -DDBS_STRIP_EXPERIMENTS: every extra kernel-parameter read in
// the per-tile loops of the MMA issuer / epilogue is a constant-bank round trip on the critical path (measured: the
// 64 -> 64 layer at 64x64 went from 204 to 247 us with the checks compiled in).
#ifdef DBS_STRIP_EXPERIMENTS
constexpr bool kExp = true;
#else
constexpr bool kExp = false;
#endif
// Epilogue warps: 8 (default) = two per TMEM lane quadrant with up to two 16-column chunks each; 16
// (-DDBS_STRIP_EPI_WARPS=16) = four per quadrant, one chunk each.  Measured equal within 1-2 % in round 1 (conv 1.382 vs
// 1.373 ms per forward at C2): the drain is a latency chain per tile, not issue-bound, so half the warps do the same
// work in the same time.  Eight warps at a 96-register cap leave half of the SM's register file free: the HBM-bound
// kernels of the OTHER half-batch's stream (GroupNorm apply, Euler-Maruyama, init conv) become co-resident with a conv CTA
// and run under it (samplers/diffusion.py: two half-batch programs on two streams).
#ifndef DBS_STRIP_EPI_WARPS
#define DBS_STRIP_EPI_WARPS 16
#endif
constexpr int kEpiWarpsDefault = DBS_STRIP_EPI_WARPS;
// GroupNorm-prologue variants (kPre): 8 epilogue warps + kXW transform warps, 576 threads -> a 112-register cap.  The
// epilogue of this kernel sits at its register cap with sixteen warps (96), so the prologue cannot live in the epilogue
// warps (round-1 experiments, profiles/experiments/); dedicated warps with their own loop keep the per-channel
// scale / shift / time-bias vector of a thread in registers and touch the epilogue only through issue slots.
constexpr int kXW = 8;
constexpr int kEpiWarpsPre = 8;
__host__ __device__ constexpr int strip_threads(int ew, bool pre) { return 64 + 32 * ew + (pre ? 32 * kXW : 0); }
constexpr int kMaxStages = 6;
constexpr uint32_t kSmemBudget = 216 * 1024;   // + 1 KB alignment slack + 9 KB static = the 227 KB a CTA may use
constexpr int kMaxCout = 512;             // Cout > 64 is processed as N tiles of 64 couts

struct StripParams {
  int W, Ht, tiles_h, n_tiles;
  int n_tiles_n;      // N tiles of Nb couts; a CTA's tile order is (image, n-tile, row block)
  int Nb, KC, chunks0, chunks1, n_taps, stages;
  int resident;       // 0: a chunk's weight slabs travel with its strip; 1: the layer's weights stay in shared memory;
                      // 2: N-tiled layer run as one pass per n-tile over the CTA's row blocks, that n-tile's weights
                      //    resident for the pass (tile order (n-tile, image, row block) instead of (image, n-tile, row block))
  int B, Hout, Wout, Cout, out_stride, oph, opw;
  int groups, Cg, cg_log2;
  int dh_min;
  int n_groups;       // distinct vertical offsets dh; every group holds the same nacc horizontal offsets
  int grp_row[4];     // (dh - dh_min): strip row offset of the group
  int n_slabs;        // weight slabs per chunk = n_taps (+1 with a fused 1x1 residual conv)
  int has_res;        // fused residual: an extra slab stacked onto the dh = 0 group, own accumulator slot nacc
  int grp_off[4];     // first slab of each group
  int grp_nslab[4];   // slabs stacked along N in each group
  int slab_tap[16];   // weight tap index of each slab in smem order; -1 = the residual 1x1 weights
  int slot_c, slot_p, slot_m, nacc;  // accumulator slots of dw = 0, +1, -1 (-1 = unused)
  // transposed-conv phase PAIR (ConvGeom::n_phases == 2): the two horizontal output phases pw = 0, 1 of one vertical
  // phase share their A rows, so their four (pw, dw) slabs are stacked along N per dh group (N = 4 * Nb: A is read once
  // for four taps) and the epilogue emits two outputs: output 1 from slots slot2_*, out2_off elements after output 0
  int n_out, slot2_c, slot2_p, slot2_m;
  int64_t out2_off;
  const float* bias;
  __nv_bfloat16* out;
  float* stats;
  const float* bias_res;
  __nv_bfloat16* out_res;
  uint32_t idesc_g[4];  // instruction descriptor per group (N = grp_nslab * Nb)
  uint32_t a_bytes, slab_bytes, stage_bytes, w_region_bytes;
  uint32_t tmem_cols;
  uint64_t desc_hi;
  // GroupNorm prologue (kPre kernels): the input strips are the RAW output of the block's first conv (plus, in the
  // residual form, the block's residual in a second strip behind it); the transform warps rewrite each landed strip in
  // shared memory as  SiLU(GN(y)) + time bias [+ residual]  (score_network.py:143-151) before the MMAs read it, so
  // that activation never exists in HBM and its gn_apply launch disappears.  pre_stats: raw (sum, sumsq) per
  // (sample, group) of the producing conv's output, [B][8][2].
  const float *pre_stats, *pre_gamma, *pre_beta, *pre_tb;
  int64_t pre_tb_stride;
  float pre_inv_cnt;
  int pre_H, pre_cg_log2, pre_rows, pre_res;
  int epi_release;  // 1: the strip stage of a tile's LAST chunk is released by the epilogue (after the tile's TMEM-full
                    // barrier, i.e. after its MMAs have read the operands) instead of by a second tcgen05.commit
  int s2d_cpp;  // > 0: input read through the space-to-depth view, chunks per row parity (= 2C / KC), see ConvGeom::s2d
  int dbg;  // diagnostics (env DBS_STRIP_DBG): 1 no MMA, 2 no epilogue body, 4 no stores, 8 no strip loads, 128 MMA issuer ignores
            // the TMEM-empty barrier, 512 epilogue ignores the TMEM-full barrier, 256 producer ignores the stage-empty barrier (timing experiments only: results are garbage)
};


// Pins a loop-invariant value in a register.  Without it the compiler re-reads kernel parameters from the constant bank
// inside the per-tile loops of the single-thread roles, and every such read is a round trip on the role's critical path
// (measured in round 2: five extra parameter checks in these loops cost 20 % of a layer's time).
__device__ __forceinline__ uint32_t pin(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ int pin(int v) { asm volatile("" : "+r"(v)); return v; }

// MMA issuer role.  The taps of one vertical offset dh share their A rows, so their weight slabs
// are stacked along N in shared memory ([dw = -1 | 0 | +1] x Nb couts) and ONE tcgen05.mma with
// N = nacc * Nb feeds all horizontal-offset accumulators (contiguous TMEM column ranges): the A
// operand is read from shared memory once per (dh, K-slice) instead of once per tap.  Unrolled over
// G groups x MM (= KC/16) MMAs with the per-group offsets in registers (the first generic loop
// measured ~130 cycles of scalar issue overhead per MMA).
template <int G, int MM>
__device__ __forceinline__ void strip_mma_role(const StripParams& p, uint32_t ring, uint32_t w_region, uint32_t tmem_base,
                                               uint32_t bar_full, uint32_t bar_empty, uint32_t bar_tfull,
                                               uint32_t bar_tempty, uint32_t bar_w, uint32_t bar_wfree, int n_it, int n_s,
                                               uint32_t bar_xfull) {
  const int chunks = pin(p.chunks0 + p.chunks1);
  const uint32_t row_bytes = (uint32_t)p.KC * 2;
  const uint32_t chunk_w_bytes = pin(p.slab_bytes * p.n_slabs);
  const int acc_cols = pin((p.nacc + p.has_res) * p.Nb);
  const uint32_t lo_bits = pin((uint32_t)p.desc_hi);            // LBO field
  const uint32_t hi32 = pin((uint32_t)(p.desc_hi >> 32));       // SBO, version, swizzle mode
  const int resident = pin(p.resident), n_stages = pin(p.stages);
  const uint32_t stage_bytes = pin(p.stage_bytes), b_in_stage = pin((p.a_bytes + 1023u) & ~1023u);
  uint32_t aoff[G], boff[G], idesc[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    aoff[g] = pin(((uint32_t)(p.grp_row[g] * p.W) * row_bytes) >> 4);
    boff[g] = pin((p.slab_bytes * (uint32_t)p.grp_off[g]) >> 4);
    idesc[g] = pin(p.idesc_g[g]);
  }
  // Zero-initialisation: the first MMA of a tile (chunk 0, group 0) overwrites the nacc tap slots.  With a
  // fused residual the widest group (dh = 0, nacc + 1 slabs) is issued FIRST so that the same overwrite also
  // covers the residual slot (the host orders the groups accordingly).
  auto acc_flag = [](int ch, int g, int k) -> uint32_t { return (ch == 0 && g == 0 && k == 0) ? 0u : 1u; };
  if (resident == 1) {
    mbar_wait(bar_w, 0);
    tc_fence_after();
  }
  int stage = 0;
  uint32_t phase = 0;
  int pass_left = 0;       // resident == 2: tiles left in the current pass
  uint32_t wphase = 0;
  for (int it = 0; it < n_it; ++it) {
    if (resident == 2 && pass_left == 0) {   // a new pass: its n-tile's weights must have landed
      mbar_wait(bar_w, wphase);
      tc_fence_after();
      wphase ^= 1;
      pass_left = n_s;
    }
    const int as = it & 1;
    const uint32_t aphase = (it >> 1) & 1;
    if (!(kExp && (kExp && (p.dbg & 128)))) mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + as * acc_cols;
    for (int ch = 0; ch < chunks; ++ch) {
      mbar_wait(bar_full + 8 * stage, phase);
      // GroupNorm prologue: the strip is ready once the transform warps have rewritten it (their generic-proxy stores
      // are made visible to the tensor core's async-proxy reads by the fence.proxy.async they execute before arriving)
      if (bar_xfull) mbar_wait(bar_xfull + 8 * stage, phase);
      tc_fence_after();
      const uint32_t a_base = ring + stage * stage_bytes;
      const uint32_t b_base = resident ? w_region + ch * chunk_w_bytes : a_base + b_in_stage;
      const uint32_t a_lo = lo_bits | ((a_base & 0x3FFFFu) >> 4);
      const uint32_t b_lo = lo_bits | ((b_base & 0x3FFFFu) >> 4);
      if (!(kExp && (p.dbg & 1))) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
          for (int k = 0; k < MM; ++k)
            umma_bf16_lohi(d_tmem, a_lo + aoff[g] + 2 * k, b_lo + boff[g] + 2 * k, hi32, idesc[g], acc_flag(ch, g, k));
        }
      }
      if (kExp && (kExp && (p.dbg & 1024))) mbar_arrive(bar_empty + 8 * stage);
      else if (!(kExp && p.epi_release && ch == chunks - 1)) umma_commit(bar_empty + 8 * stage);
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    if (kExp && (kExp && (p.dbg & 2048))) mbar_arrive(bar_tfull + 8 * as); else umma_commit(bar_tfull + 8 * as);
    if (resident == 2 && --pass_left == 0) umma_commit(bar_wfree);   // the pass's MMAs no longer read the weights
  }
}

// kStats: accumulate GroupNorm(8) sums; NCH = Nb / 16 column chunks.  kWide = false: Cout <= 64, group
// width Cg = 2 * NCH <= 8 (several groups per chunk); kWide = true: Cout >= 128, Cg >= 16 (a chunk lies
// inside one group).
__device__ __forceinline__ float strip_silu(float x) {   // same form as gn_apply's bf16 path (unet.cu)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);       // sigmoid(x); the caller multiplies by x (and adds the time bias in the same FMA)
}

// kSpec: the epilogue's per-chunk parameters as compile-time constants for the standard 3x3 layer (one output, accumulator slots
// dw = -1 | 0 | +1 = 0 | 1 | 2): 1 = without, 2 = with the fused 1x1 residual conv; kXchgT = an image row spans several warps
// (W > 32).  0 = every case, parameters read at run time.  The epilogue's instruction stream is latency-critical: with these
// few values folded the 64 -> 64 layer at 64x64 went from 200.8 to 192.2 us (profiles/README.md).
template <bool kStats, int NCH, bool kWide, int kEpiWarps = kEpiWarpsDefault, bool kPre = false, int kSpec = 0, bool kXchgT = false>
__global__ void __launch_bounds__(strip_threads(kEpiWarps, kPre), (kEpiWarps <= 8 && !kPre) ? 2 : 1)   // <= 8 warps: cap at 96 registers
conv_strip_kernel(const __grid_constant__ CUtensorMap map_in0, const __grid_constant__ CUtensorMap map_in1,
                  const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_wr,
                  const __grid_constant__ StripParams p) {
  constexpr int kThreads = strip_threads(kEpiWarps, kPre);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[3 * kMaxStages + 6];
  __shared__ __align__(16) float s_tab[kPre ? 3 : 1][kPre ? 256 : 4];   // kPre: scale / shift / time bias per input channel
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_bias[kMaxCout];
  __shared__ __align__(16) float s_bias_r[kMaxCout];
  __shared__ __align__(16) float s_x[2][4][4][2][16];   // [parity][sub][quadrant][side][col]: cross-warp neighbours (W > 32)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();          // the next kernel of the step may begin its prologue as SMs free up (common.cuh)
  const uint32_t w_region = smem_u32(smem);                       // resident weights (may be empty)
  const uint32_t ring = w_region + p.w_region_bytes;              // strip (+ streamed weight) stages
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * kMaxStages + 2]);
  const uint32_t bar_w = smem_u32(&bars[2 * kMaxStages + 4]);
  const uint32_t bar_wfree = smem_u32(&bars[2 * kMaxStages + 5]);
  const uint32_t bar_xfull = smem_u32(&bars[2 * kMaxStages + 6]);   // kPre: strip rewritten, the MMAs may read it

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
      if (kPre) mbar_init(bar_xfull + 8 * s, kXW);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, NCH <= 2 ? kEpiWarps / 2 : kEpiWarps);
    }
    mbar_init(bar_w, 1);
    mbar_init(bar_wfree, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_in0);
    tma_prefetch_desc(&map_in1);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_wr);
  }
  if (warp == 2) tmem_alloc(smem_u32(&s_tmem_base), p.tmem_cols);
  for (int i = threadIdx.x; i < p.Cout; i += kThreads) {
    s_bias[i] = p.bias[i];
    s_bias_r[i] = p.has_res ? p.bias_res[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  // Everything so far touches only this layer's constants and on-chip state.  The layer's resident weights are constants
  // too: their TMA loads are issued before the wait, so that they travel behind the previous kernel's tail.
  if (warp == 0 && lane == 0 && p.resident == 1) {
    const uint32_t cwb = p.slab_bytes * p.n_slabs;
    const int n_ch = p.chunks0 + p.chunks1;
    mbar_expect_tx(bar_w, cwb * n_ch);
    for (int ch = 0; ch < n_ch; ++ch)
      for (int t = 0; t < p.n_slabs; ++t) {
        const uint32_t dst = w_region + (ch * p.n_slabs + t) * p.slab_bytes;
        if (p.slab_tap[t] >= 0) tma_load_3d(dst, &map_w, bar_w, ch * p.KC, 0, p.slab_tap[t]);
        else tma_load_3d(dst, &map_wr, bar_w, ch * p.KC, 0, 0);
      }
  }
  pdl_wait();                       // from here on: activations / statistics written by the previous kernel of the stream

  // contiguous tile range per CTA: consecutive tiles are rows of the same image, which lets the
  // epilogue keep GroupNorm partial sums in registers until the image changes.  resident == 2: the range is
  // over row blocks (B * tiles_h) and is walked once per n-tile.
  const int n_units = p.resident == 2 ? p.B * p.tiles_h : p.n_tiles;
  const int tile_lo = (int)(((int64_t)n_units * blockIdx.x) / gridDim.x);
  const int tile_hi = (int)(((int64_t)n_units * (blockIdx.x + 1)) / gridDim.x);
  const int n_s = tile_hi - tile_lo;
  const int n_it = p.resident == 2 ? n_s * p.n_tiles_n : n_s;
  const int chunks = p.chunks0 + p.chunks1;
  const uint32_t row_bytes = (uint32_t)p.KC * 2;
  const uint32_t chunk_w_bytes = p.slab_bytes * p.n_slabs;
  const int acc_cols = (p.nacc + p.has_res) * p.Nb;   // TMEM columns of one accumulator stage

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      auto load_weights = [&](int co0) {
        mbar_expect_tx(bar_w, chunk_w_bytes * chunks);
        for (int ch = 0; ch < chunks; ++ch)
          for (int t = 0; t < p.n_slabs; ++t) {
            const uint32_t dst = w_region + (ch * p.n_slabs + t) * p.slab_bytes;
            if (p.slab_tap[t] >= 0) tma_load_3d(dst, &map_w, bar_w, ch * p.KC, co0, p.slab_tap[t]);
            else tma_load_3d(dst, &map_wr, bar_w, ch * p.KC, co0, 0);
          }
      };
      // loop invariants pinned in registers (see pin()): the producer's per-tile work is then address arithmetic only
      const int resident = pin(p.resident), n_stages = pin(p.stages), tiles_h = pin(p.tiles_h), Ht = pin(p.Ht);
      const int n_tiles_n = pin(p.n_tiles_n), Nb = pin(p.Nb), KC = pin(p.KC), chunks0 = pin(p.chunks0), dh_min = pin(p.dh_min);
      const int s2d_cpp = pin(p.s2d_cpp), n_slabs = pin(p.n_slabs);
      const uint32_t stage_bytes = pin(p.stage_bytes), b_in_stage = pin((p.a_bytes + 1023u) & ~1023u), slab_bytes = pin(p.slab_bytes);
      const bool pre_res = kPre && p.pre_res;
      const uint32_t tx_bytes = pin(p.a_bytes * (pre_res ? 2u : 1u) + (p.resident ? 0u : chunk_w_bytes));
      // (resident == 1: the weights were requested before pdl_wait above; n_tiles_n == 1 there)
      int stage = 0;
      uint32_t phase = 0;
      int pass = 0, in_pass = 0;
      // tile -> (image, n-tile, row block) kept incrementally: one division per CTA, none per tile
      int n, nt, hb;
      if (resident == 2) {
        n = tile_lo / tiles_h; hb = tile_lo - n * tiles_h; nt = 0;
      } else {
        const int per_img = n_tiles_n * tiles_h;
        n = tile_lo / per_img;
        const int rem = tile_lo - n * per_img;
        nt = rem / tiles_h;
        hb = rem - nt * tiles_h;
      }
      const int n0 = n, hb0 = hb;
      for (int it = 0; it < n_it; ++it) {
        if (resident == 2 && in_pass == 0) {
          // the previous pass's MMAs must be done with the weight region before it is overwritten
          if (pass > 0) mbar_wait(bar_wfree, (uint32_t)(pass - 1) & 1u);
          load_weights(pass * Nb);
        }
        const int t_n = n, h0 = hb * Ht, co0 = nt * Nb;
        if (resident == 2) {
          if (++hb == tiles_h) { hb = 0; ++n; }
          if (++in_pass == n_s) { in_pass = 0; ++pass; ++nt; n = n0; hb = hb0; }
        } else {
          if (++hb == tiles_h) { hb = 0; if (++nt == n_tiles_n) { nt = 0; ++n; } }
        }
        for (int ch = 0; ch < chunks; ++ch) {
          if (!(kExp && (p.dbg & 256))) mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t a_dst = ring + stage * stage_bytes;
          const uint32_t fb = bar_full + 8 * stage;
          if (kExp && (p.dbg & 8)) {
            mbar_arrive(fb);
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(fb, tx_bytes);
          if (pre_res)   // the block's residual for the same rows, [strip | residual strip] in one stage
            tma_load_5d(a_dst + b_in_stage, &map_in1, fb, ch * KC, 0, 0, h0 + dh_min, t_n);
          if (s2d_cpp) {
            // chunk -> (row parity py, channel offset inside the 2C-wide pixel pair)
            const int py = ch / s2d_cpp;
            tma_load_5d(a_dst, &map_in0, fb, (ch - py * s2d_cpp) * KC, 0, py, h0 + dh_min, t_n);
          } else if (ch < chunks0) tma_load_5d(a_dst, &map_in0, fb, ch * KC, 0, 0, h0 + dh_min, t_n);
          else tma_load_5d(a_dst, &map_in1, fb, (ch - chunks0) * KC, 0, 0, h0 + dh_min, t_n);
          if (!resident) {
            const uint32_t b_dst = a_dst + b_in_stage;
            for (int t = 0; t < n_slabs; ++t) {
              if (p.slab_tap[t] >= 0) tma_load_3d(b_dst + t * slab_bytes, &map_w, fb, ch * KC, co0, p.slab_tap[t]);
              else tma_load_3d(b_dst + t * slab_bytes, &map_wr, fb, ch * KC, co0, 0);
            }
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
#define DBS_STRIP_ROLE(G, MM) \
  strip_mma_role<G, MM>(p, ring, w_region, tmem_base, bar_full, bar_empty, bar_tfull, bar_tempty, bar_w, bar_wfree, n_it, n_s, \
                        kPre ? bar_xfull : 0u)
      const int mm = p.KC / 16;
      if (p.n_groups == 3 && mm == 4) DBS_STRIP_ROLE(3, 4);
      else if (p.n_groups == 3 && mm == 2) DBS_STRIP_ROLE(3, 2);
      else if (p.n_groups == 3) DBS_STRIP_ROLE(3, 1);
      else if (p.n_groups == 2 && mm == 4) DBS_STRIP_ROLE(2, 4);
      else if (p.n_groups == 2 && mm == 2) DBS_STRIP_ROLE(2, 2);
      else if (p.n_groups == 2) DBS_STRIP_ROLE(2, 1);
      else if (mm == 4) DBS_STRIP_ROLE(1, 4);
      else if (mm == 2) DBS_STRIP_ROLE(1, 2);
      else DBS_STRIP_ROLE(1, 1);
#undef DBS_STRIP_ROLE
    }
    __syncwarp();
  } else if (kPre && warp >= 2 + kEpiWarps) {
    // ===================== transform warps (GroupNorm prologue) =====================
    // A thread owns ONE logical 16-byte channel vector (8 channels) of every pixel row it visits, so the vector's
    // scale / shift / time bias stay in registers for a whole chunk; the strip is rewritten in place (the residual form
    // reads the second strip of the stage).  Rows outside the image keep the TMA's zero fill: the conv pads the
    // NORMALISED activation with zeros.
    constexpr int nx = 32 * kXW;
    const int xt = (int)threadIdx.x - 32 * (2 + kEpiWarps);
    const int vpr = p.KC >> 3;                           // 16-byte vectors per pixel row of a chunk: 8 / 4 / 2
    const int vlog = 31 - __clz(vpr), wlog = 31 - __clz(p.W);
    const int v = xt & (vpr - 1);
    const int pix0 = xt >> vlog, pstep = nx >> vlog;
    const int xs = 3 - vlog;                             // swizzle XOR term = (row byte offset >> 7) = pix >> xs
    const int n_pix = p.pre_rows * p.W;
    const int Cin = p.chunks0 * p.KC;
    const uint32_t res_off = (p.a_bytes + 1023u) & ~1023u;
    int stage = 0;
    uint32_t phase = 0;
    int cur_n = -1;
    for (int it = 0; it < n_it; ++it) {
      const int tile = tile_lo + it;                      // resident == 1: tile = image * tiles_h + row block
      const int n_img = tile / p.tiles_h, h0 = (tile - n_img * p.tiles_h) * p.Ht;
      if (n_img != cur_n) {
        // every transform thread is done with the previous image's table before it is overwritten
        asm volatile("bar.sync 8, %0;" ::"n"(nx) : "memory");
        for (int c = xt; c < Cin; c += nx) {
          const int gi = c >> p.pre_cg_log2;
          const float s0 = __ldg(p.pre_stats + ((int64_t)n_img * 8 + gi) * 2), s1 = __ldg(p.pre_stats + ((int64_t)n_img * 8 + gi) * 2 + 1);
          const float mean = s0 * p.pre_inv_cnt;
          const float rstd = rsqrtf(fmaxf(s1 * p.pre_inv_cnt - mean * mean, 0.f) + 1e-5f);
          const float a_ = rstd * __ldg(p.pre_gamma + c);
          s_tab[0][c] = a_;
          s_tab[1][c] = __ldg(p.pre_beta + c) - mean * a_;
          s_tab[2][c] = p.pre_tb ? __ldg(p.pre_tb + (int64_t)n_img * p.pre_tb_stride + c) : 0.f;
        }
        asm volatile("bar.sync 8, %0;" ::"n"(nx) : "memory");
        cur_n = n_img;
      }
      for (int ch = 0; ch < chunks; ++ch) {
        float sc[8], sh[8], tb[8];
        {
          const int c0 = ch * p.KC + v * 8;
          const float4 a0 = *reinterpret_cast<const float4*>(&s_tab[0][c0]), a1 = *reinterpret_cast<const float4*>(&s_tab[0][c0 + 4]);
          const float4 b0 = *reinterpret_cast<const float4*>(&s_tab[1][c0]), b1 = *reinterpret_cast<const float4*>(&s_tab[1][c0 + 4]);
          const float4 t0 = *reinterpret_cast<const float4*>(&s_tab[2][c0]), t1 = *reinterpret_cast<const float4*>(&s_tab[2][c0 + 4]);
          sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
          sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
          tb[0] = t0.x; tb[1] = t0.y; tb[2] = t0.z; tb[3] = t0.w; tb[4] = t1.x; tb[5] = t1.y; tb[6] = t1.z; tb[7] = t1.w;
        }
        mbar_wait(bar_full + 8 * stage, phase);
        uint8_t* base = smem + p.w_region_bytes + (size_t)stage * p.stage_bytes;
#pragma unroll 2
        for (int pix = pix0; pix < n_pix; pix += pstep) {
          const int ih = h0 + p.dh_min + (pix >> wlog);
          if (ih < 0 || ih >= p.pre_H) continue;
          const int pc = v ^ ((pix >> xs) & (vpr - 1));    // physical position of my logical vector (hardware swizzle)
          uint4* ptr = reinterpret_cast<uint4*>(base + ((((size_t)pix << vlog) + pc) << 4));
          const uint4 q = *ptr;
          const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
          uint32_t r4[4] = {0u, 0u, 0u, 0u};
          if (p.pre_res) {
            const uint4 rq = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(ptr) + res_off);
            r4[0] = rq.x; r4[1] = rq.y; r4[2] = rq.z; r4[3] = rq.w;
          }
          uint32_t o4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x0 = fmaf(__uint_as_float(w4[i] << 16), sc[2 * i], sh[2 * i]);
            const float x1 = fmaf(__uint_as_float(w4[i] & 0xffff0000u), sc[2 * i + 1], sh[2 * i + 1]);
            float y0 = fmaf(x0, strip_silu(x0), tb[2 * i]);
            float y1 = fmaf(x1, strip_silu(x1), tb[2 * i + 1]);
            if (p.pre_res) { y0 += __uint_as_float(r4[i] << 16); y1 += __uint_as_float(r4[i] & 0xffff0000u); }
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
            o4[i] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          *ptr = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
        // generic-proxy writes -> visible to the tensor core's async-proxy operand reads
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_xfull + 8 * stage);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // A warp's work per tile is one 16-column chunk of one TMEM lane quadrant, a latency-bound chain
    // (tcgen05.ld -> shuffles -> convert -> store).  With 64 couts the sixteen warps cover the four chunks
    // of a tile in lock step (measured: splitting them into two sets that each take two chunks is 15 %
    // slower).  With <= 32 couts there are only two chunks: the warps then form two sets of eight that
    // drain alternate tiles (set s owns accumulator stage s), so two epilogues overlap instead of half of
    // the warps idling.
    constexpr bool kSplit = NCH <= 2;
    const int q = warp & 3;                 // TMEM lane quadrant
    constexpr int kSubs = kEpiWarps / 4;    // warps per TMEM lane quadrant
    const int half = (warp - 2) >> 2;       // 0..kSubs-1
    constexpr int kSetSubs = kSplit ? kSubs / 2 : kSubs;   // warps of a quadrant that work on the same tile
    const int my_set = kSplit ? half / kSetSubs : 0;
    const int hsel = half % kSetSubs;
    constexpr int MYCH = (NCH + kSetSubs - 1) / kSetSubs;   // chunks this warp owns: hsel * MYCH .. + MYCH - 1
    static_assert(NCH <= 4, "one chunk per warp");
    const int row = q * 32 + lane;          // tile-local pixel
    const int hl = row / p.W, wl = row - hl * p.W;
    const float m_right = wl < p.W - 1 ? 1.f : 0.f;   // pixel m+1 is in the same image row
    const float m_left = wl > 0 ? 1.f : 0.f;
    const bool xchg = kSpec ? kXchgT : (p.W > 32);   // an image row spans several warps
    constexpr int CG = kWide ? 16 : 2 * NCH;  // channels of a group inside one chunk
    constexpr int GPC = 16 / CG;              // groups per 16-column chunk (1 when kWide)
    float st_s[MYCH * GPC], st_ss[MYCH * GPC];
#pragma unroll
    for (int i = 0; i < MYCH * GPC; ++i) { st_s[i] = 0.f; st_ss[i] = 0.f; }
    int cur_n = -1, cur_nt = 0;
    auto flush_stats = [&](int n_img, int nt_img) {
#pragma unroll
      for (int i = 0; i < MYCH * GPC; ++i) {
        const int chunk = hsel * MYCH + i / GPC;
        if (chunk < NCH) {
          const float s_ = warp_sum(st_s[i]), ss_ = warp_sum(st_ss[i]);
          const int gi = kWide ? ((nt_img * p.Nb + chunk * 16) >> p.cg_log2) : chunk * GPC + (i % GPC);
          if (lane == 0) {
            atomicAdd(p.stats + ((int64_t)n_img * 8 + gi) * 2 + 0, s_);
            atomicAdd(p.stats + ((int64_t)n_img * 8 + gi) * 2 + 1, ss_);
          }
        }
        st_s[i] = 0.f; st_ss[i] = 0.f;
      }
    };
    // tile -> (image, n-tile, row block) kept incrementally: one division per CTA, none per tile
    const int per_img = p.n_tiles_n * p.tiles_h;
    int n, nt, hb;
    if (p.resident == 2) {
      n = tile_lo / p.tiles_h; hb = tile_lo - n * p.tiles_h; nt = 0;
    } else {
      n = tile_lo / per_img;
      const int rem = tile_lo - n * per_img;
      nt = rem / p.tiles_h;
      hb = rem - nt * p.tiles_h;
    }
    int in_pass = 0;
    int xb = 0;  // exchange-buffer parity, toggled per chunk (one named barrier per chunk suffices)
    int rel_stage = -1;      // strip stage of the current tile's last chunk (epi_release)
    for (int it = 0; it < n_it; ++it) {
      if (kExp && p.epi_release) {
        rel_stage += chunks;
        while (rel_stage >= p.stages) rel_stage -= p.stages;
      }
      const int t_n = n, t_nt = nt, t_hb = hb;
      if (p.resident == 2) {
        if (++hb == p.tiles_h) { hb = 0; ++n; }
        if (++in_pass == n_s) { in_pass = 0; ++nt; n = tile_lo / p.tiles_h; hb = tile_lo - n * p.tiles_h; }
      } else {
        if (++hb == p.tiles_h) { hb = 0; if (++nt == p.n_tiles_n) { nt = 0; ++n; } }
      }
      if (kSplit && (it & 1) != my_set) continue;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int a = t_hb * p.Ht + hl, co0 = t_nt * p.Nb;
      if (kStats && (t_n != cur_n || t_nt != cur_nt)) {
        if (cur_n >= 0) flush_stats(cur_n, cur_nt);
        cur_n = t_n; cur_nt = t_nt;
      }
      __nv_bfloat16* optr =
          p.out + (((int64_t)t_n * p.Hout + (a * p.out_stride + p.oph)) * p.Wout + (wl * p.out_stride + p.opw)) * p.Cout + co0;

      if (!(kExp && (kExp && (p.dbg & 512)))) mbar_wait(bar_tfull + 8 * as, aphase);
      tc_fence_after();
      // the tile's MMAs are complete: its last strip stage can be refilled (one elected thread of the warps on this tile)
      if (kExp && p.epi_release && q == 0 && hsel == 0 && lane == 0) mbar_arrive(bar_empty + 8 * rel_stage);
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + as * acc_cols;
      // outputs of this tile: 1, or 2 for a transposed-conv phase pair (never with statistics); idle warps: 0
      const int n_out = (kExp && (p.dbg & 2)) ? 0 : (kSpec == 3 ? 2 : ((kStats || kSpec) ? 1 : p.n_out));
#pragma unroll
      for (int mc = 0; mc < MYCH; ++mc) {
      const int chunk = hsel * MYCH + mc;
      if (chunk >= NCH) break;                 // idle warps / chunks beyond the tile's columns
#pragma unroll(kSpec == 3 ? 2 : 1)
      for (int o = 0; o < n_out; ++o) {
        // (with MYCH = 2, requesting both chunks from TMEM before using the first was measured: no gain)
        const int c = chunk * 16;
        constexpr int kNb = NCH * 16;          // = p.Nb (Nb is Cout up to 64 couts, 64 beyond)
        // kSpec 1 / 2: slots dw = -1 | 0 | +1 = 0 | 1 | 2.  kSpec 3 (transposed-conv phase pair): output 0 from slots
        // (dw = -1, 0) = (0, 1), output 1 from (dw = 0, +1) = (2, 3); the loop is unrolled so that the slots fold per output
        const int slot_c = kSpec == 3 ? (o ? 2 : 1) : (kSpec ? 1 : (o ? p.slot2_c : p.slot_c));
        const int slot_p = kSpec == 3 ? (o ? 3 : -1) : (kSpec ? 2 : (o ? p.slot2_p : p.slot_p));
        const int slot_m = kSpec == 3 ? (o ? -1 : 0) : (kSpec ? 0 : (o ? p.slot2_m : p.slot_m));
        __nv_bfloat16* const optr_o = (kSpec == 1 || kSpec == 2) ? optr : optr + o * p.out2_off;
        uint32_t v0[16], vp[16], vm[16];
        tmem_ld16(tbase + slot_c * kNb + c, v0);
        if (slot_p >= 0 && !(kExp && (p.dbg & 64))) tmem_ld16(tbase + slot_p * kNb + c, vp);
        if (slot_m >= 0 && !(kExp && (p.dbg & 64))) tmem_ld16(tbase + slot_m * kNb + c, vm);
        tmem_ld_wait();
        if (xchg) {
          // publish the edge lanes: my lane 0's P_{+1} is pixel m+1 of the previous warp's lane 31,
          // my lane 31's P_{-1} is pixel m-1 of the next warp's lane 0
          // (as four 16-byte accesses per side: the exchange used to be 64 scalar shared-memory instructions per chunk on
          // top of the 32 shuffles, all through the same MIO queue)
          if (lane == 0 && slot_p >= 0) {
            uint4* d = reinterpret_cast<uint4*>(&s_x[xb][half][q][0][0]);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) d[j4] = make_uint4(vp[4 * j4], vp[4 * j4 + 1], vp[4 * j4 + 2], vp[4 * j4 + 3]);
          }
          if (lane == 31 && slot_m >= 0) {
            uint4* d = reinterpret_cast<uint4*>(&s_x[xb][half][q][1][0]);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) d[j4] = make_uint4(vm[4 * j4], vm[4 * j4 + 1], vm[4 * j4 + 2], vm[4 * j4 + 3]);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
        }
        // neighbour columns: all shuffles of a side are issued back to back, in place (no shared temporary, so
        // the sixteen of them are in flight together instead of one SHFL -> FFMA round trip per column)
        if (slot_p >= 0 && !(kExp && (p.dbg & 16))) {
#pragma unroll
          for (int j = 0; j < 16; ++j) vp[j] = __shfl_down_sync(0xffffffffu, vp[j], 1);
          if (xchg && lane == 31) {
            const uint4* sp = reinterpret_cast<const uint4*>(&s_x[xb][half][(q + 1) & 3][0][0]);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 t4 = sp[j4];
              vp[4 * j4] = t4.x; vp[4 * j4 + 1] = t4.y; vp[4 * j4 + 2] = t4.z; vp[4 * j4 + 3] = t4.w;
            }
          }
        }
        if (slot_m >= 0 && !(kExp && (p.dbg & 16))) {
#pragma unroll
          for (int j = 0; j < 16; ++j) vm[j] = __shfl_up_sync(0xffffffffu, vm[j], 1);
          if (xchg && lane == 0) {
            const uint4* sp = reinterpret_cast<const uint4*>(&s_x[xb][half][(q + 3) & 3][1][0]);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 t4 = sp[j4];
              vm[4 * j4] = t4.x; vm[4 * j4 + 1] = t4.y; vm[4 * j4 + 2] = t4.z; vm[4 * j4 + 3] = t4.w;
            }
          }
        }
        float f[16];
        const float4* bq = reinterpret_cast<const float4*>(s_bias + co0 + c);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b4 = bq[j4];
          f[4 * j4 + 0] = __uint_as_float(v0[4 * j4 + 0]) + b4.x;
          f[4 * j4 + 1] = __uint_as_float(v0[4 * j4 + 1]) + b4.y;
          f[4 * j4 + 2] = __uint_as_float(v0[4 * j4 + 2]) + b4.z;
          f[4 * j4 + 3] = __uint_as_float(v0[4 * j4 + 3]) + b4.w;
        }
        if (slot_p >= 0 && !(kExp && (p.dbg & 16))) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaf(m_right, __uint_as_float(vp[j]), f[j]);
        }
        if (slot_m >= 0 && !(kExp && (p.dbg & 16))) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaf(m_left, __uint_as_float(vm[j]), f[j]);
        }
        if (kStats && !(kExp && (p.dbg & 32))) {
#pragma unroll
          for (int r = 0; r < GPC; ++r) {
#pragma unroll
            for (int j = 0; j < CG; ++j) {
              st_s[mc * GPC + r] += f[r * CG + j];
              st_ss[mc * GPC + r] = fmaf(f[r * CG + j], f[r * CG + j], st_ss[mc * GPC + r]);
            }
          }
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        if (!(kExp && (p.dbg & 4))) {
          if (kSpec == 1 && NCH <= 2) {
            // the plain 3x3 layer with <= 32 couts is the one place where the 32-byte store measured slower (191 -> 197 us
            // at B = 1024, 64x64: two warp sets at the register cap); it keeps the two 16-byte stores
            uint4* dst = reinterpret_cast<uint4*>(optr_o + c);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
            st_global_256(optr_o + c, pk);   // the pixel's sixteen couts of this chunk: one 32-byte sector
          }
        }
        if (kSpec == 2 || (kSpec == 0 && p.has_res)) {   // (never with kSpec 1 / 3)
          // fused 1x1 residual conv of the same input (score_network.py:131,151): own slot, own bias, no GN
          uint32_t vr[16];
          tmem_ld16(tbase + (kSpec ? 3 : p.nacc) * kNb + c, vr);
          tmem_ld_wait();
          const float4* br = reinterpret_cast<const float4*>(s_bias_r + co0 + c);
          uint32_t pr[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = br[j4];
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(vr[4 * j4 + 0]) + b4.x,
                                                             __uint_as_float(vr[4 * j4 + 1]) + b4.y);
            const __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(vr[4 * j4 + 2]) + b4.z,
                                                             __uint_as_float(vr[4 * j4 + 3]) + b4.w);
            pr[2 * j4] = *reinterpret_cast<const uint32_t*>(&h0);
            pr[2 * j4 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
          }
          st_global_256(p.out_res + (optr - p.out) + c, pr);
        }
        xb ^= 1;
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
    }
    if (kStats && cur_n >= 0) flush_stats(cur_n, cur_nt);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace

int conv_strip_launch(const ConvGeom& g, const __nv_bfloat16* in0, const __nv_bfloat16* in1, const __nv_bfloat16* w,
                      const float* bias, __nv_bfloat16* out, float* stats, int groups, cudaStream_t st,
                      const __nv_bfloat16* w_res, const float* bias_res, __nv_bfloat16* out_res, const GnFuse* pre,
                      const float* pre_stats) {
  const bool fuse_res = w_res != nullptr;
  // Measured (round 2, profiles/README.md): parity-green but break-even at best — the strip kernel is bound by shared-memory
  // traffic (MMA operand reads + epilogue shuffles), and an in-place rewrite of every strip adds 2 x strip bytes to it —
  // so the prologue is opt-in (DBS_GN_PROLOGUE=1); GroupNorm apply is instead overlapped with the other half-batch's convs.
  static const bool pre_disabled = [] { const char* e = getenv("DBS_GN_PROLOGUE"); return !(e && e[0] == '1'); }();
  if (pre && (pre_disabled || !pre_stats || !pre->gamma || !pre->beta || fuse_res || g.C1 != 0 || in1 != nullptr ||
              g.out_stride != 1 || g.n_phases > 1 || g.s2d || g.n_taps != 9 || g.C0 > 256 || g.C0 % 16 ||
              !is_pow2(g.C0 / 8)))
    return DBS_ERR_UNSUPPORTED;
  const bool pairw = g.n_phases == 2;   // transposed-conv phase pair: taps [0, n_taps) = (ph, pw = 0), [n_taps, 2 n_taps) = (ph, 1)
  if (g.n_phases > 2) return DBS_ERR_UNSUPPORTED;   // all four phases merged: tile kernel only
  if (pairw && (fuse_res || stats || g.out_stride != 2 || g.n_taps != 4)) return DBS_ERR_UNSUPPORTED;
  const int n_taps_all = pairw ? 2 * g.n_taps : g.n_taps;
  if (fuse_res && (!bias_res || !out_res || g.out_stride != 1 || g.n_taps != 9)) return DBS_ERR_UNSUPPORTED;
  static const bool disabled = [] { const char* e = getenv("DBS_CONV_STRIP"); return e && e[0] == '0'; }();
  if (disabled) return DBS_ERR_UNSUPPORTED;
  // ---- eligibility
  if (g.in_stride != 1 || g.Wm != g.Win || g.Hm != g.Hin) return DBS_ERR_UNSUPPORTED;
  const int W = g.Win, H = g.Hin;
  if (!(W == 8 || W == 16 || W == 32 || W == 64 || W == 128)) return DBS_ERR_UNSUPPORTED;
  const int Ht = kTileM / W;
  if (H < Ht || H % Ht) return DBS_ERR_UNSUPPORTED;
  if (g.Cout > kMaxCout || g.Cout % 16 || g.C0 == 0) return DBS_ERR_UNSUPPORTED;
  if (g.Cout > 64 && g.Cout % 64) return DBS_ERR_UNSUPPORTED;
  int KC = 0;
  for (int k : {64, 32, 16})
    if (g.C0 % k == 0 && g.C1 % k == 0) { KC = k; break; }
  if (!KC) return DBS_ERR_UNSUPPORTED;
  if (stats && groups != 8) return DBS_ERR_UNSUPPORTED;     // GroupNorm(8): Cg = Cout / 8 = 2 * (Cout / 16)
  if (g.Cout == 48) return DBS_ERR_UNSUPPORTED;             // 3 chunks: group width 6 straddles no chunk, not built
  if ((reinterpret_cast<uintptr_t>(in0) | reinterpret_cast<uintptr_t>(in1) | reinterpret_cast<uintptr_t>(w)) & 15)
    return DBS_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out_res)) & 31)   // 32-byte epilogue stores
    return DBS_ERR_UNSUPPORTED;
  int dh_min = 0, dh_max = 0;
  for (int t = 0; t < n_taps_all; ++t) {
    if (g.dw[t] < -1 || g.dw[t] > 1) return DBS_ERR_UNSUPPORTED;
    dh_min = g.dh[t] < dh_min ? g.dh[t] : dh_min;
    dh_max = g.dh[t] > dh_max ? g.dh[t] : dh_max;
  }
  const int rows = Ht + dh_max - dh_min;
  if (rows > 256) return DBS_ERR_UNSUPPORTED;
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("conv_strip: cuTensorMapEncodeTiled entry point unavailable"); return DBS_ERR_CUDA; }

  StripParams p;
  memset(&p, 0, sizeof(p));
  const int Cin = g.C0 + g.C1, Nb = g.Cout > 64 ? 64 : g.Cout;
  p.W = W; p.Ht = Ht; p.tiles_h = H / Ht; p.n_tiles_n = g.Cout / Nb; p.n_tiles = g.B * p.tiles_h * p.n_tiles_n;
  p.Nb = Nb; p.KC = KC; p.chunks0 = g.C0 / KC; p.chunks1 = g.C1 / KC; p.n_taps = g.n_taps;
  p.B = g.B; p.Hout = g.Hout; p.Wout = g.Wout; p.Cout = g.Cout;
  p.out_stride = g.out_stride; p.oph = g.oph; p.opw = g.opw;
  p.groups = groups; p.Cg = stats ? g.Cout / groups : g.Cout;
  while ((1 << p.cg_log2) < p.Cg) ++p.cg_log2;
  p.dh_min = dh_min;
  // taps must form a full (dh x dw) grid: groups = distinct dh, slots = distinct dw (ascending), and
  // the dw = 0 slot must exist (the epilogue starts from it)
  int dhs[4], dws[3], G = 0, nacc = 0;
  for (int v = dh_min; v <= dh_max; ++v) {
    bool present = false;
    for (int t = 0; t < n_taps_all; ++t) present |= g.dh[t] == v;
    if (present) { if (G == 4) return DBS_ERR_UNSUPPORTED; dhs[G++] = v; }
  }
  int slot_of[3] = {-1, -1, -1};  // index dw + 1
  if (!pairw) {
    for (int v = -1; v <= 1; ++v) {
      bool present = false;
      for (int t = 0; t < g.n_taps; ++t) present |= g.dw[t] == v;
      if (present) { slot_of[v + 1] = nacc; dws[nacc++] = v; }
    }
    if (slot_of[1] < 0 || G > 3 || G * nacc != g.n_taps) return DBS_ERR_UNSUPPORTED;
  } else {
    nacc = 4;   // slots [pw0: dw = -1 | pw0: dw = 0 | pw1: dw = 0 | pw1: dw = +1]
    if (G != 2) return DBS_ERR_UNSUPPORTED;
  }
  if ((nacc + (fuse_res ? 1 : 0)) * Nb > 256) return DBS_ERR_UNSUPPORTED;
  // issue order of the groups: with a fused residual the dh = 0 group (one slab wider) goes first, so the
  // tile's zero-initialising MMA covers the residual slot too
  if (fuse_res) {
    int z = -1;
    for (int gi = 0; gi < G; ++gi) if (dhs[gi] == 0) z = gi;
    if (z < 0) return DBS_ERR_UNSUPPORTED;
    const int tmp = dhs[0]; dhs[0] = dhs[z]; dhs[z] = tmp;
  }
  int slab = 0;
  for (int gi = 0; gi < G; ++gi) {
    p.grp_row[gi] = dhs[gi] - dh_min;
    p.grp_off[gi] = slab;
    if (!pairw) {
      for (int si = 0; si < nacc; ++si) {
        int found = -1;
        for (int t = 0; t < g.n_taps; ++t)
          if (g.dh[t] == dhs[gi] && g.dw[t] == dws[si]) { if (found >= 0) return DBS_ERR_UNSUPPORTED; found = t; }
        if (found < 0) return DBS_ERR_UNSUPPORTED;
        p.slab_tap[slab++] = g.widx[found];
      }
    } else {
      const int want_dw[4] = {-1, 0, 0, 1};
      for (int si = 0; si < 4; ++si) {
        const int t0 = (si >> 1) * g.n_taps;   // taps of pw = si / 2
        int found = -1;
        for (int t = t0; t < t0 + g.n_taps; ++t)
          if (g.dh[t] == dhs[gi] && g.dw[t] == want_dw[si]) { if (found >= 0) return DBS_ERR_UNSUPPORTED; found = t; }
        if (found < 0) return DBS_ERR_UNSUPPORTED;
        p.slab_tap[slab++] = g.widx[found];
      }
    }
    if (fuse_res && dhs[gi] == 0) p.slab_tap[slab++] = -1;
    p.grp_nslab[gi] = slab - p.grp_off[gi];
    p.idesc_g[gi] = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((p.grp_nslab[gi] * Nb) >> 3) << 17) |
                    ((uint32_t)(kTileM >> 4) << 24);
  }
  p.n_slabs = slab;
  p.has_res = fuse_res ? 1 : 0;
  p.bias_res = bias_res; p.out_res = out_res;
  p.n_groups = G;
  p.nacc = nacc;
  if (!pairw) {
    p.slot_c = slot_of[1]; p.slot_p = slot_of[2]; p.slot_m = slot_of[0];
    p.n_out = 1; p.slot2_c = p.slot2_p = p.slot2_m = -1; p.out2_off = 0;
  } else {
    p.slot_m = 0; p.slot_c = 1; p.slot_p = -1;
    p.slot2_c = 2; p.slot2_p = 3; p.slot2_m = -1;
    p.n_out = 2; p.out2_off = g.Cout;     // output pixel (2a + ph, 2b + 1) follows (2a + ph, 2b)
    p.opw = 0;
  }
  p.bias = bias; p.out = out; p.stats = stats;
  { const char* e = getenv("DBS_STRIP_DBG"); p.dbg = e ? atoi(e) : 0; }
  { static const int er = [] { const char* e = getenv("DBS_STRIP_EPI_RELEASE"); return e ? atoi(e) : 0; }(); p.epi_release = er; }
  p.a_bytes = (uint32_t)rows * W * KC * 2;
  p.slab_bytes = (uint32_t)Nb * KC * 2;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * (nacc + (fuse_res ? 1 : 0)) * Nb)) cols <<= 1;
  if (cols > 512) return DBS_ERR_UNSUPPORTED;
  p.tmem_cols = cols;
  const uint32_t a_pad = (p.a_bytes + 1023u) & ~1023u;
  const uint32_t chunk_w = p.slab_bytes * (uint32_t)p.n_slabs;
  const uint32_t all_w = chunk_w * (uint32_t)(p.chunks0 + p.chunks1);
  if (p.slab_bytes % 1024u && KC == 64) return DBS_ERR_UNSUPPORTED;   // slab bases must stay atom-aligned
  if (p.slab_bytes % (8u * KC * 2u)) return DBS_ERR_UNSUPPORTED;
  // resident weights if they leave room for >= 3 strip stages; N-tiled layers keep one n-tile's weights resident
  // per pass (>= 2 stages: the alternative streams chunk_w bytes of weights with every a_pad-byte strip)
  static const int pass_env = [] { const char* e = getenv("DBS_STRIP_PASS"); return e ? atoi(e) : 1; }();
  const int sp_tiles = g.B * p.tiles_h;   // row blocks: the unit CTAs split in pass mode
  // (the prologue variants carry 3 KB more static shared memory: the per-channel table and a third barrier set)
  const uint32_t budget = pre ? kSmemBudget - 4096 : kSmemBudget;
  const uint32_t strip_pad = pre && pre->res ? 2 * a_pad : a_pad;   // residual form: [raw strip | residual strip] per stage
  if (p.n_tiles_n == 1 && all_w + 3 * strip_pad <= budget) {
    p.resident = 1;
    p.w_region_bytes = (all_w + 1023u) & ~1023u;
    p.stage_bytes = strip_pad;
  } else if (pass_env && p.n_tiles_n > 1 && all_w + (uint32_t)(pass_env >= 2 ? 2 : 3) * a_pad <= kSmemBudget &&
             sp_tiles >= num_sms()) {
    p.resident = 2;
    p.w_region_bytes = (all_w + 1023u) & ~1023u;
    p.stage_bytes = a_pad;
  } else {
    p.resident = 0;
    p.w_region_bytes = 0;
    p.stage_bytes = a_pad + ((chunk_w + 1023u) & ~1023u);
  }
  // space-to-depth stride-2 convs pay 78 % more MMA work for their zero taps: only worth it with resident weights
  // (64 channels: 45 vs 58 us; with streamed weights at 128 channels the tile kernel is faster, 37 vs 60 us)
  if (g.s2d && p.resident != 1) return DBS_ERR_UNSUPPORTED;
  if (pre) {
    // GroupNorm prologue: layer weights resident (tile = image * tiles_h + row block)
    if (p.resident != 1) return DBS_ERR_UNSUPPORTED;
    p.pre_stats = pre_stats; p.pre_gamma = pre->gamma; p.pre_beta = pre->beta; p.pre_tb = pre->tb;
    p.pre_tb_stride = pre->tb_stride;
    p.pre_inv_cnt = 1.0f / (float)((int64_t)H * W * (g.C0 / 8));
    p.pre_H = H; p.pre_rows = rows; p.pre_res = pre->res ? 1 : 0;
    while ((1 << p.pre_cg_log2) < g.C0 / 8) ++p.pre_cg_log2;
  }
  int stages = (int)((budget - p.w_region_bytes) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return DBS_ERR_UNSUPPORTED;
  p.stages = stages;
  const uint32_t row_bytes = (uint32_t)KC * 2;
  const uint64_t layout = KC == 64 ? 2ull : (KC == 32 ? 4ull : 6ull);
  p.desc_hi = (1ull << 16) | ((uint64_t)((8 * row_bytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
  const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                         : (KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap m0, m1, mw;
  int rc;
  if (g.s2d) {
    // {2C, W, 2, H, N} view of the [N][2H][2W][C] tensor (encode_act's stride-2 form takes the FULL sizes)
    if (g.C1 != 0 || (g.C0 / 2) % KC) return DBS_ERR_UNSUPPORTED;
    p.s2d_cpp = (g.C0 / 2) / KC;
    if ((rc = encode_act(enc, &m0, in0, g.C0 / 4, 2 * H, 2 * W, g.B, 2, KC, W, rows, 1, sw))) return rc;
  } else if ((rc = encode_act(enc, &m0, in0, g.C0, H, W, g.B, 1, KC, W, rows, 1, sw))) return rc;
  if (g.C1 > 0) {
    if ((rc = encode_act(enc, &m1, in1, g.C1, H, W, g.B, 1, KC, W, rows, 1, sw))) return rc;
  } else if (pre && pre->res) {
    if ((rc = encode_act(enc, &m1, pre->res, g.C0, H, W, g.B, 1, KC, W, rows, 1, sw))) return rc;
  } else {
    m1 = m0;
  }
  int taps_total = 0;
  for (int t = 0; t < g.n_taps; ++t) taps_total = g.widx[t] + 1 > taps_total ? g.widx[t] + 1 : taps_total;
  if (g.out_stride == 2) taps_total = 16;
  if ((rc = encode_w(enc, &mw, w, Cin, g.Cout, taps_total, KC, Nb, sw))) return rc;
  CUtensorMap mwr = mw;
  if (fuse_res && (rc = encode_w(enc, &mwr, w_res, Cin, g.Cout, 1, KC, Nb, sw))) return rc;

  const size_t smem = (size_t)p.w_region_bytes + (size_t)stages * p.stage_bytes + 1024;
  typedef void (*Kern)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, StripParams);
  // [stats][NCH index] for Cout <= 64, plus the wide-group pair for N-tiled layers (Nb = 64, Cg >= 16)
  static const Kern table[2][4] = {
      {conv_strip_kernel<false, 1, false>, conv_strip_kernel<false, 2, false>, conv_strip_kernel<false, 4, false>,
       conv_strip_kernel<false, 4, true>},
      {conv_strip_kernel<true, 1, false>, conv_strip_kernel<true, 2, false>, conv_strip_kernel<true, 4, false>,
       conv_strip_kernel<true, 4, true>}};
  // specialised epilogues (kSpec) for the standard 3x3 layers with statistics: [NCH 2 | 4 | 4 wide][res][xchg]
  static const Kern spec_table[3][2][2] = {
      {{conv_strip_kernel<true, 2, false, kEpiWarpsDefault, false, 1, false>, conv_strip_kernel<true, 2, false, kEpiWarpsDefault, false, 1, true>},
       {conv_strip_kernel<true, 2, false, kEpiWarpsDefault, false, 2, false>, conv_strip_kernel<true, 2, false, kEpiWarpsDefault, false, 2, true>}},
      {{conv_strip_kernel<true, 4, false, kEpiWarpsDefault, false, 1, false>, conv_strip_kernel<true, 4, false, kEpiWarpsDefault, false, 1, true>},
       {conv_strip_kernel<true, 4, false, kEpiWarpsDefault, false, 2, false>, conv_strip_kernel<true, 4, false, kEpiWarpsDefault, false, 2, true>}},
      {{conv_strip_kernel<true, 4, true, kEpiWarpsDefault, false, 1, false>, conv_strip_kernel<true, 4, true, kEpiWarpsDefault, false, 1, true>},
       {conv_strip_kernel<true, 4, true, kEpiWarpsDefault, false, 2, false>, conv_strip_kernel<true, 4, true, kEpiWarpsDefault, false, 2, true>}}};
  // transposed-conv phase pairs (kSpec 3): [NCH 2 | 4 | 4 wide][xchg]
  static const Kern pair_table[3][2] = {
      {conv_strip_kernel<false, 2, false, kEpiWarpsDefault, false, 3, false>, conv_strip_kernel<false, 2, false, kEpiWarpsDefault, false, 3, true>},
      {conv_strip_kernel<false, 4, false, kEpiWarpsDefault, false, 3, false>, conv_strip_kernel<false, 4, false, kEpiWarpsDefault, false, 3, true>},
      {conv_strip_kernel<false, 4, true, kEpiWarpsDefault, false, 3, false>, conv_strip_kernel<false, 4, true, kEpiWarpsDefault, false, 3, true>}};
  const int nch = Nb / 16;
  const bool wide = g.Cout > 64;
  if (wide && stats && (!is_pow2(p.Cg) || p.Cg < 16)) return DBS_ERR_UNSUPPORTED;
  Kern kern = table[stats ? 1 : 0][wide ? 3 : (nch == 1 ? 0 : (nch == 2 ? 1 : 2))];
  static const bool spec_off = [] { const char* e = getenv("DBS_STRIP_SPEC"); return e && e[0] == '0'; }();
  const bool standard3x3 = stats && !pairw && !pre && !g.s2d && g.n_taps == 9 && nacc == 3 && p.n_out == 1 && p.slot_m == 0 &&
                           p.slot_c == 1 && p.slot_p == 2 && (!fuse_res || p.nacc == 3) && (nch == 2 || nch == 4);
  if (standard3x3 && !spec_off) kern = spec_table[wide ? 2 : (nch == 2 ? 0 : 1)][fuse_res ? 1 : 0][W > 32 ? 1 : 0];
  const bool standard_pair = pairw && !stats && !pre && !fuse_res && p.n_out == 2 && p.slot_m == 0 && p.slot_c == 1 && p.slot_p == -1 &&
                             p.slot2_c == 2 && p.slot2_p == 3 && p.slot2_m == -1 && (nch == 2 || nch == 4);
  if (standard_pair && !spec_off) kern = pair_table[wide ? 2 : (nch == 2 ? 0 : 1)][W > 32 ? 1 : 0];
  // GroupNorm-prologue variants: [stats][NCH index], 8 epilogue warps + kXW transform warps
  static const Kern pre_table[2][3] = {
      {conv_strip_kernel<false, 1, false, kEpiWarpsPre, true>, conv_strip_kernel<false, 2, false, kEpiWarpsPre, true>,
       conv_strip_kernel<false, 4, false, kEpiWarpsPre, true>},
      {conv_strip_kernel<true, 1, false, kEpiWarpsPre, true>, conv_strip_kernel<true, 2, false, kEpiWarpsPre, true>,
       conv_strip_kernel<true, 4, false, kEpiWarpsPre, true>}};
  if (pre) {
    if (wide || (nch != 1 && nch != 2 && nch != 4)) return DBS_ERR_UNSUPPORTED;
    kern = pre_table[stats ? 1 : 0][nch == 1 ? 0 : (nch == 2 ? 1 : 2)];
  }
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 4; ++b) {
        const cudaError_t e = cudaFuncSetAttribute(table[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(kSmemBudget + 2048));
        if (e != cudaSuccess) attr_err = e;
      }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 2; ++b) {
        const cudaError_t e = cudaFuncSetAttribute(pair_table[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(kSmemBudget + 2048));
        if (e != cudaSuccess) attr_err = e;
      }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 2; ++c) {
          const cudaError_t e = cudaFuncSetAttribute(spec_table[a][b][c], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)(kSmemBudget + 2048));
          if (e != cudaSuccess) attr_err = e;
        }
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 3; ++b) {
        const cudaError_t e = cudaFuncSetAttribute(pre_table[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(kSmemBudget - 4096 + 2048));
        if (e != cudaSuccess) attr_err = e;
      }
  });
  DBS_CUDA(attr_err);
  const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();   // resident == 2 has sp_tiles >= num_sms()
  prof_set_label("strip H%d W%d C%d+%d->%d taps%d os%d KC%d B%d res%d st%d%s%s", H, W, g.C0, g.C1, g.Cout, g.n_taps,
                 g.out_stride, KC, g.B, p.resident, p.stages, fuse_res ? " +res1x1" : (pre ? (pre->res ? " gn+res>" : " gn>") : ""),
                 pairw ? " x2 phases" : "");
  ProfScope prof(kProfConvTC, st);
  DBS_CUDA(launch_chained(kern, dim3((unsigned)grid), dim3(strip_threads(pre ? kEpiWarpsPre : kEpiWarpsDefault, pre != nullptr)), smem, st,
                          m0, m1, mw, mwr, p));
  DBS_LAUNCHED();
  return DBS_OK;
}

}  // namespace dbs
