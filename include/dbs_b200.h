/*
 * dbs_b200.h — C ABI of libdbs_b200.so, the sm_100a sampling engine behind the
 * Python API of Sakeeb91/diffusion-boltzmann-sampler's hot path.
 *
 * The reference has no FFI seam: its boundary is the Python class API of five
 * modules.  Each entry point below names the reference symbol
 * (file:line under /root/reference) whose arithmetic it replaces; the Python
 * modules under diffusion-boltzmann-sampler_b200/backend/ml/ bind these through
 * ctypes (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the ABI.
 *   - every function returns int: 0 = ok, <0 = dbs_status; dbs_last_error()
 *     returns a thread-local message.  No exception crosses the boundary.
 *   - all buffers are CALLER-OWNED DEVICE pointers (torch.Tensor.data_ptr()).
 *     The library allocates nothing except the opaque dbs_unet_plan (repacked
 *     weights + TMA descriptors); workspaces are passed in by the caller.
 *   - every launch takes a cudaStream_t as `void* stream` and is asynchronous;
 *     no hidden device synchronisation, no global mutable state.
 */
#ifndef DBS_B200_H
#define DBS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum dbs_status {
  DBS_OK = 0,
  DBS_ERR_INVALID = -1,     /* bad argument (odd L, unsupported shape, null ptr) */
  DBS_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed             */
  DBS_ERR_UNSUPPORTED = -3, /* valid in the reference, not built here (documented) */
  DBS_ERR_NO_DEVICE = -4    /* no sm_100 device visible                        */
} dbs_status;

/* Thread-local text of the last failure on this thread ("" if none). */
const char* dbs_last_error(void);
/* ABI version (major*1000+minor) and the SM architecture the kernels were built for (100). */
int dbs_abi_version(void);
int dbs_built_sm(void);
/* Number of kernels this library has launched since load (monotonic; for bench.py's gpu_launches). */
uint64_t dbs_launch_count(void);

/* Optional per-kernel-class device timing for bench.py's roofline: when enabled, every launch is
 * bracketed by CUDA events on its own stream (no synchronisation); dbs_profile_read waits for the
 * recorded events and returns accumulated milliseconds and launch counts per class. */
int dbs_profile_enable(int on);
int dbs_profile_num_classes(void);
const char* dbs_profile_class_name(int cls);
int dbs_profile_read(double* ms, uint64_t* launches, int n, int reset);
/* Per-launch detail ("class label<TAB>ms" per line) of the labelled launches read so far. */
int dbs_profile_dump(const char* path);

/* ------------------------------------------------------------------------- *
 * Empty batches: every batched entry point returns DBS_OK for n_chains / n / B == 0 without touching its
 * pointers (they may be NULL), as the reference's tensor code does for empty tensors.
 * ------------------------------------------------------------------------- */

/* ------------------------------------------------------------------------- *
 * Ising lattice state: bit-packed, 1 bit per spin (1 = +1, 0 = -1).
 * Layout: uint32 words[n_chains][L][W], W = ceil(L/32); bit (j%32) of word
 * (j/32) of row i is spin (i,j); unused high bits are zero.  4<=L<=128 (odd L from 5).
 * ------------------------------------------------------------------------- */

/* words per chain for lattice size L (L*ceil(L/32)); <0 if L unsupported. */
int dbs_ising_words_per_chain(int L);

/* Hot start.  Replaces IsingModel.random_configuration (systems/ising.py:36-45):
 * spin(i,j) of chain c = bit (j%32) of Philox4x32-10(key=seed,
 * ctr={chain_offset+c, i, j/32, DBS_STREAM_INIT}).x */
int dbs_ising_init_packed(uint32_t* bits, int L, int64_t n_chains, uint64_t seed,
                          int64_t chain_offset, void* stream);

/* Same with global chain id = chain_offset + c * chain_stride: a rank of a striped sharding (rank r of W takes
 * the ids r, r + W, r + 2W, ...: SURVEY.md 8e, config 3) initialises its chains in ONE launch. */
int dbs_ising_init_packed_strided(uint32_t* bits, int L, int64_t n_chains, uint64_t seed,
                                  int64_t chain_offset, int64_t chain_stride, void* stream);

/* fp32 (+/-1, sign test x>0) <-> bits.  Replaces the float lattice representation
 * of systems/ising.py:36-45 / samplers/mcmc.py:82-97 at the API boundary. */
int dbs_ising_pack(const float* spins, uint32_t* bits, int L, int64_t n_chains, void* stream);
int dbs_ising_unpack(const uint32_t* bits, float* spins, int L, int64_t n_chains, void* stream);

/* Checkerboard Metropolis sweeps.  Replaces MetropolisHastings.step/sweep
 * (samplers/mcmc.py:26-60) + IsingModel.local_energy_diff (systems/ising.py:68-91).
 *
 * One sweep = colour-0 half-sweep then colour-1 half-sweep, colour = (i+j)&1.
 * Site class k = 8*s + u, s = spin bit, u = number of +1 neighbours (0..4).
 * accept  <=>  (r < thr[k].lo) || thr[k].hi   with r the site's 32-bit draw, i.e.
 * the reference's  dE<=0 || u < exp(-beta*dE)  with u = r*2^-32
 * (thr[k].lo = ceil(exp(-beta*dE_k)*2^32), thr[k].hi = 1 iff dE_k <= 0).
 *
 *   thr          uint32 [n_tables][16][2]  ({lo,hi} pairs), built on the host
 *   table_index  int32 [n_chains] or NULL (all chains use table 0): per-chain temperature
 *   draw source  uniforms == NULL: Philox4x32-10(key=seed, ctr={chain_offset+c,
 *                2*(sweep_offset+s)+colour, i*GPR + q/4, DBS_STREAM_SWEEP}).v[q%4],
 *                q = rank of the site among its colour in row i, GPR = ceil(L/8);
 *                uniforms != NULL (test-mode injection port): uint32
 *                [n_sweeps][n_chains][L][L], r of site (i,j) in sweep s.
 *   odd L        a torus with an odd side is not bipartite: one sweep = THREE colour classes in turn, colour(i,j) =
 *                (c(i)+c(j)) mod 3 with c(k) = k&1 for k < L-1 and c(L-1) = 2; same classes, table and acceptance;
 *                draw of site q = i*L+j: Philox4x32-10(key=seed, ctr={chain_offset+c, sweep_offset+s, q/4,
 *                DBS_STREAM_SWEEP3}).v[q%4]; the injection port keeps its [n_sweeps][n_chains][L][L] layout.
 *   zero_field   0: general (J, h) table path.  +1 / -1: the caller guarantees h == 0 with J > 0 / J < 0;
 *                the kernel then counts agreeing (disagreeing) neighbours for 32 sites at once and
 *                compares draws only against the two uphill thresholds of the table — the same
 *                decisions bit for bit, ~1.6x faster.
 */
int dbs_ising_sweeps(uint32_t* bits, int L, int64_t n_chains, int n_sweeps,
                     const uint32_t* thr, int n_tables, const int32_t* table_index,
                     uint64_t seed, int64_t chain_offset, int64_t sweep_offset,
                     const uint32_t* uniforms, int zero_field, void* stream);

/* dbs_ising_sweeps with global chain id = chain_offset + c * chain_stride in the Philox counters (striped
 * sharding: the draws of a chain do not depend on which rank, or how many ranks, run it). */
int dbs_ising_sweeps_strided(uint32_t* bits, int L, int64_t n_chains, int n_sweeps,
                             const uint32_t* thr, int n_tables, const int32_t* table_index,
                             uint64_t seed, int64_t chain_offset, int64_t chain_stride,
                             int64_t sweep_offset, const uint32_t* uniforms, int zero_field,
                             void* stream);

/* One random-site Metropolis attempt at (i,j) on fp32 lattices [n][L][L] in place, with draw r:
 * the reference's step() (samplers/mcmc.py:26-50) for API parity (flips at most one spin per
 * lattice).  Any L >= 1. */
int dbs_ising_site_update(float* spins, int L, int64_t n, int i, int j, const uint32_t* thr,
                          uint32_t r, void* stream);

/* Per-chain energy and magnetisation from packed state.  Replaces
 * IsingModel.energy / magnetization (systems/ising.py:47-66, 93-102):
 * E = fl32(fl32(-J)*S_bond) + fl32(fl32(-h)*S_spin), M = fl32(S_spin)/fl32(L*L). */
int dbs_ising_observables(const uint32_t* bits, int L, int64_t n_chains, float J, float h,
                          float* energy, float* magnetization, void* stream);

/* Same observables from an fp32 (n, L, L) tensor of arbitrary real values
 * (the reference API accepts continuous samples: samplers/diffusion.py:571-575).
 * Any L >= 1 here (no packing). */
int dbs_ising_observables_f32(const float* spins, int L, int64_t n, float J, float h,
                              float* energy, float* magnetization, void* stream);

/* Spin-spin correlation sums.  Replaces pair_correlation (analysis/correlation.py:8-47).
 * For every chain c and displacement (di,dj) in [-R,R]^2, R = L/2, with
 * r = floor(sqrt(di^2+dj^2)), 0 < r < R:
 *   acc[group(c)][r] += sum over reference sites (i,j) of s(i,j)*s(i+di,j+dj)
 * ref_mask: uint32 [L][W] bit mask of reference sites (NULL = all sites);
 * group: int32 [n_chains] or NULL (all 0).  acc: int64 [n_groups][R], accumulated
 * atomically (caller zeroes it).  Counts are data-independent: host computes them. */
int dbs_ising_correlation(const uint32_t* bits, int L, int64_t n_chains,
                          const uint32_t* ref_mask, const int32_t* group,
                          long long* acc, void* stream);

/* One launch per measurement of a temperature sweep (even L <= 64): E and M of every chain
 * (as dbs_ising_observables forms them), block[group][0..5] += {1, E/N, (E/N)^2, |M|, M^2, M^4} (fp64 atomics, the
 * specific heat and susceptibility sums of api/routes/sampling.py:522-527 included) and the C(r) sums of
 * dbs_ising_correlation with every site a reference site into corr_acc[group][r] (int64, exact).  block rows are
 * block_stride doubles apart (>= 6): the caller appends the C(r) columns and all-reduces ONE block. */
int dbs_ising_measure(const uint32_t* bits, int L, int64_t n_chains, float J, float h, const int32_t* group,
                      double* block, int block_stride, long long* corr_acc, void* stream);

/* How dbs_ising_correlation (ref_mask == NULL) and dbs_ising_measure evaluate the C(r) sums.  0 (default) = automatic:
 * at L = 64 the spectral kernel (|FFT2|^2 -> inverse FFT2 of four chains at a time, every value rounded to the integer it
 * is; csrc/ising_corr_fft.cu), the XOR / POPC displacement kernels at every other size; 1 = always XOR / POPC; 2 = spectral
 * where it applies.  Both produce the same int64 sums bit for bit.  The spectral kernel leaves acc[group][0] (r = 0 is
 * never a valid radius) at zero unless a value was not within 1/4 of an integer — callers treat a non-zero bin 0 as an
 * error.  Process-wide; returns the previous method, or DBS_ERR_INVALID. */
int dbs_ising_set_correlation_method(int method);

/* Block accumulation of per-chain observables into the fp64 all-reduce block
 * (SURVEY.md §8e): acc[group][0..4] += {1, E/N, |M|, M^2, M^4}. */
int dbs_observable_accumulate(const float* energy, const float* magnetization, int L,
                              int64_t n_chains, const int32_t* group, double* acc,
                              void* stream);

/* ------------------------------------------------------------------------- *
 * Reverse-SDE Euler–Maruyama update.  Replaces the inner step of
 * DiffusionSampler.sample (samplers/diffusion.py:232-243):
 *   x <- (x + (c1*score)*dt) + ((c2*z)*temperature)      (fp32, no FMA contraction)
 * c1 = fl32(sigma^2), c2 = fl32(sigma*sqrt(dt)); add_noise=0 on the last step.
 * noise != NULL: z read from the tensor (test-mode injection port);
 * noise == NULL: z = Box–Muller of Philox4x32-10(key=seed, ctr={sample_offset+n,
 *                e/4, step, DBS_STREAM_NOISE}), e = element index inside the sample.
 * ------------------------------------------------------------------------- */
int dbs_em_step(float* x, const float* score, const float* noise, int64_t n_samples,
                int64_t elems_per_sample, float c1, float dt, float c2, float temperature,
                int add_noise, uint64_t seed, int64_t sample_offset, int64_t step,
                void* stream);

/* x ~ N(0,1) start of the reverse process (torch.randn at samplers/diffusion.py:216)
 * from the same Philox stream with step = DBS_STEP_X0. */
int dbs_gaussian_fill(float* x, int64_t n_samples, int64_t elems_per_sample, uint64_t seed,
                      int64_t sample_offset, int64_t step, void* stream);

/* out[i] = sum_t a[t] * f_t(x[t][i]), 1..5 terms, f_0 = torch.sign when sign_first != 0, else identity.
 * The state combinations of the deterministic integrators and correctors around the score network
 * (samplers/diffusion.py:441-447 corrector, :488-519 Heun / RK4, :762-768 heuristic anneal) as one pass.
 * x: HOST array of n_terms DEVICE pointers (16-byte aligned, may alias out); a: HOST array of coefficients. */
int dbs_lincomb(float* out, int64_t n, int n_terms, const float* const* x, const float* a,
                int sign_first, void* stream);

/* Same with device-resident coefficients (lincomb_dev) — see the step tables below. */
int dbs_lincomb_dev(float* out, int64_t n, int n_terms, const float* const* x, const float* coef_dev,
                    int sign_first, void* stream);

/* ------------------------------------------------------------------------- *
 * Step tables for graph-captured sampler loops (the loops of samplers/diffusion.py:222-243,
 * :425-449, :476-519).  The scalars of step k — times, sigma^2, noise scale, Philox step id ... — are row k of a
 * DEVICE table; one captured CUDA graph of a step then serves the whole trajectory with no host-side parameter
 * update and no host synchronisation.
 *   dbs_step_table_next: cur[0..stride) = table[*counter * stride ...], then ++*counter (one tiny launch at the head
 *                        of the captured step; the step's kernels read `cur`, the U-Net its time from cur + offset).
 *   dbs_em_step_dev:     dbs_em_step with coef = {c1, dt, c2, temperature, add_noise (0/1), step id (int32 bits)} and
 *                        seed_offset = {seed, sample_offset} (int64) read on the device.
 *   dbs_corrector_coef:  Langevin corrector step size of samplers/diffusion.py:441-445 as a device reduction:
 *                        g = mean_b ||score_b||_2, step = 2 * ((snr * sqrt(elems)) / g)^2; coef_out = {step, sqrt(2 step)};
 *                        norms: scratch fp32 [B].
 *   dbs_corrector_update: x <- (x + step*score) + ((sqrt(2 step)*z)*temperature), :447; step_size2 = coef_out above,
 *                        coef = {temperature, active (0 = leave x untouched), step id (int32 bits)}; z from `noise`
 *                        (injection port) or Philox(seed_offset).
 * ------------------------------------------------------------------------- */
int dbs_step_table_next(float* cur, const float* table, int* counter, int stride, void* stream);
int dbs_em_step_dev(float* x, const float* score, int64_t n_samples, int64_t elems_per_sample,
                    const float* coef, const int64_t* seed_offset, void* stream);
int dbs_corrector_coef(const float* score, int64_t B, int64_t elems_per_sample, double snr, float* norms,
                       float* coef_out, void* stream);
int dbs_corrector_update(float* x, const float* score, const float* noise, int64_t n_samples,
                         int64_t elems_per_sample, const float* step_size2, const float* coef,
                         const int64_t* seed_offset, void* stream);

/* ------------------------------------------------------------------------- *
 * Lennard-Jones pair energy / forces.  Replaces LennardJonesSystem.energy /
 * forces / pairwise_distances / minimum_image (systems/lennard_jones.py:58-155).
 * pos fp32 [B][N][2]; energy fp32 [B] (may be NULL); forces fp32 [B][N][2] (may be NULL).
 * Same masks and epsilons as the reference: d = sqrt(dx^2+dy^2+1e-10),
 * pair counted iff 0.1 < d < cutoff, unit = dr/(d+1e-10), E = 0.5*sum_ij.
 * ------------------------------------------------------------------------- */
int dbs_lj_energy_forces(const float* pos, int64_t B, int N, float box, float epsilon,
                         float sigma, float cutoff, float energy_shift, float* energy,
                         float* forces, void* stream);

/* ------------------------------------------------------------------------- *
 * Batched Langevin samplers on the pair kernel and the RDF histogram (f-2).
 * pos/vel/force: fp32 [B][N][2]; energy fp32 [B].  The state carries its cached force and energy,
 * so one MALA step = propose, ONE dbs_lj_energy_forces on the proposal, accept.
 * noise / uniforms != NULL are the test-mode injection ports; otherwise Philox streams keyed by
 * (seed, cfg_offset + b, step).
 * ------------------------------------------------------------------------- */

/* x' = ((x + (dt*F)/T) + z*noise_scale) mod box.  Replaces samplers/langevin.py:56-63. */
int dbs_mala_propose(const float* pos, const float* force, const float* noise, float* pos_new,
                     int64_t B, int N, float dt, float temperature, float noise_scale, float box,
                     uint64_t seed, int64_t cfg_offset, int64_t step, void* stream);

/* Metropolis-Hastings test of the proposal, samplers/langevin.py:65-87: accept iff
 * log(u) < -beta (E'-E) + log q(x|x') - log q(x'|x); accepted configurations take the proposal's
 * positions, force and energy in place.  accepted: int32 [B] or NULL. */
int dbs_mala_accept(float* pos, float* force, float* energy, const float* pos_new,
                    const float* force_new, const float* energy_new, const float* uniforms,
                    int64_t B, int N, float dt, float temperature, uint64_t seed, int64_t cfg_offset,
                    int64_t step, int* accepted, void* stream);

/* BAOAB (samplers/langevin.py:192-225): first = B A O A + periodic wrap (uses the cached force of
 * the old positions), then the caller evaluates forces at the new positions, then second = B. */
int dbs_baoab_first(float* pos, float* vel, const float* force, const float* noise, int64_t B, int N,
                    float dt, float mass, float c1, float c2, float box, uint64_t seed,
                    int64_t cfg_offset, int64_t step, void* stream);
int dbs_baoab_second(float* vel, const float* force, int64_t B, int N, float dt, float mass,
                     void* stream);

/* Pair-distance histogram behind compute_rdf (analysis/rdf.py:49-68): i<j pairs, minimum image,
 * np.histogram's bin rule on float64 edges [n_bins+1]; hist uint64 [n_bins] accumulated atomically. */
int dbs_rdf_histogram(const float* pos, int64_t B, int N, float box, const double* edges, int n_bins,
                      unsigned long long* hist, void* stream);

/* ------------------------------------------------------------------------- *
 * Score U-Net (models/score_network.py:154-268).
 * A plan owns device copies of the weights repacked for the kernels
 * ([tap][C_out][C_in], fp32 and bf16) and the layer schedule.
 * ------------------------------------------------------------------------- */
typedef struct dbs_unet_plan dbs_unet_plan;

typedef struct dbs_unet_config {
  int in_channels;      /* only 1 is built (Ising)                         */
  int base_channels;    /* ScoreNetwork(base_channels=...)                  */
  int time_embed_dim;
  int num_blocks;
  float output_scale;
  int normalize_output; /* 0/1                                              */
} dbs_unet_config;

/* One named parameter of the reference state_dict: fp32, contiguous, DEVICE or
 * HOST pointer (is_device says which); name as in ScoreNetwork.state_dict(). */
typedef struct dbs_param {
  const char* name;
  const float* data;
  int64_t numel;
  int is_device;
} dbs_param;

/* precision: 0 = fp32 CUDA-core path (parity mode), 1 = bf16 tcgen05 path. */
int dbs_unet_plan_create(const dbs_unet_config* cfg, const dbs_param* params, int n_params,
                         const float* freqs_host, dbs_unet_plan** out);
void dbs_unet_plan_destroy(dbs_unet_plan* plan);

/* Bytes of caller-provided workspace needed for a forward of (B,1,H,W). */
int64_t dbs_unet_workspace_bytes(const dbs_unet_plan* plan, int64_t B, int H, int W,
                                 int precision);

/* score = ScoreNetwork.forward(x, t)  (score_network.py:224-268).
 * x, score: fp32 [B][1][H][W]; t: fp32 [n_t], n_t = 1 (all samples share t, the
 * sampler's case: samplers/diffusion.py:223) or n_t = B. */
int dbs_unet_forward(dbs_unet_plan* plan, const float* x, const float* t, int64_t n_t,
                     float* score, int64_t B, int H, int W, int precision, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* Debug/parity port: the same forward, additionally copying one named intermediate as fp32
 * NCHW into tap_out (tap_numel must match).  Names: "init", "enc<i>", "enc<i>.conv1", "down<i>",
 * "mid", "mid.conv1", "up<i>", "dec<i>", "dec<i>.conv1".  DBS_ERR_INVALID if unknown. */
int dbs_unet_forward_tap(dbs_unet_plan* plan, const float* x, const float* t, int64_t n_t,
                         float* score, int64_t B, int H, int W, int precision, void* workspace,
                         int64_t workspace_bytes, const char* tap, float* tap_out,
                         int64_t tap_numel, void* stream);

/* Stand-alone implicit-GEMM convolution on the tensor cores (tcgen05/TMEM/TMA);
 * parity port for the kernel itself.  Replaces nn.Conv2d / nn.ConvTranspose2d as
 * used at score_network.py:122-131, 201, 216, 243-259.
 *   kind: 0 = conv kxk stride 1 pad k/2 (k = 1 or 3); 1 = conv 3x3 stride 2 pad 1;
 *         2 = conv-transpose 4x4 stride 2 pad 1.
 *   in0/in1: bf16 NHWC sources, channels C0 + C1 (in1 may be NULL, C1 = 0): the
 *            skip concat of score_network.py:256 read as one K range.
 *   weight:  bf16 [taps][C_out][C0+C1]; bias fp32 [C_out].
 *   out:     bf16 NHWC [B][Ho][Wo][C_out]; stats (may be NULL): fp32 [B][8][2]
 *            per-(sample, group) sum / sum of squares of the fp32 outputs
 *            (GroupNorm(8) statistics, score_network.py:124-125), accumulated atomically.
 */
int dbs_conv_tc(const void* in0, int C0, const void* in1, int C1, const void* weight,
                const float* bias, void* out, float* stats, int64_t B, int H, int W,
                int C_out, int kind, int ksize, void* stream);

/* Fused first half of a ConvBlock on the tensor cores (parity port): out3 = conv3x3(x) + bias3 with
 * GroupNorm(8) partial sums, out1 = conv1x1(x) + bias1 (the block's residual projection,
 * score_network.py:131,151), both from ONE pass over x = [in0 | in1].  w3: bf16 [9][C_out][C],
 * w1: bf16 [1][C_out][C].  DBS_ERR_UNSUPPORTED when the shape is not eligible for the strip kernel. */
int dbs_conv_block_tc(const void* in0, int C0, const void* in1, int C1, const void* w3,
                      const float* bias3, const void* w1, const float* bias1, void* out3, void* out1,
                      float* stats, int64_t B, int H, int W, int C_out, void* stream);

/* ------------------------------------------------------------------------- *
 * Particle score network for continuous systems.  Replaces
 * ContinuousScoreNetwork.forward and its sub-modules
 * (models/score_network_continuous.py:36-75, 85-101, 123-136, 155-180, 256-282).
 * A plan owns device copies of the fp32 weights (names as in the module's state_dict:
 * time_mlp.{0,2}, input_proj, blocks.i.{fc1,fc2,norm1,norm2,time_mlp},
 * attention_layers.i.{norm,qkv,proj}, output_proj.{0,1,3}).
 * ------------------------------------------------------------------------- */
typedef struct dbs_csn_plan dbs_csn_plan;

typedef struct dbs_csn_config {
  int n_particles;    /* informational: forward() takes N per call, as the reference does */
  int hidden_dim;     /* multiple of 8, <= 256                                           */
  int time_embed_dim;
  int pos_embed_dim;  /* = 4 * n_frequencies                                             */
  int num_blocks;
  int num_heads;      /* head_dim = hidden_dim / num_heads in {8, 16, 32, 64}            */
  int use_attention;  /* 0/1                                                             */
} dbs_csn_config;

/* pos_freqs_host: the module's `pos_embed.freqs` buffer [pos_embed_dim/4];
 * t_freqs_host: exp(-ln(1e4) * k / (half-1)), k < time_embed_dim/2 (computed by the caller with the
 * reference's torch expression so the values are bit-identical). */
int dbs_csn_plan_create(const dbs_csn_config* cfg, const dbs_param* params, int n_params,
                        const float* pos_freqs_host, const float* t_freqs_host,
                        dbs_csn_plan** out);
void dbs_csn_plan_destroy(dbs_csn_plan* plan);
int64_t dbs_csn_workspace_bytes(const dbs_csn_plan* plan, int64_t B, int N, int64_t n_t);

/* score = ContinuousScoreNetwork.forward(x, t)  (score_network_continuous.py:256-282).
 * x, score: fp32 [B][N][2] (positions normalised to [-1,1]); t: fp32 [n_t], n_t = 1 or B.
 * precision: 0 = fp32 CUDA-core path (parity mode, <= 1e-4); 1 = bf16 tcgen05 path for the MLP / attention
 * layers (fc1, fc2, qkv, softmax(QK^T)V, proj; fp32 accumulation, fp32 residual stream), which needs
 * hidden_dim 128, head_dim 32, attention on and N in {16, 32, ..., 256}: DBS_ERR_UNSUPPORTED otherwise. */
int dbs_csn_forward(dbs_csn_plan* plan, const float* x, const float* t, int64_t n_t, float* score,
                    int64_t B, int N, int precision, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* Philox stream tags (ctr.w) and the x0 step tag. */
#define DBS_STREAM_SWEEP 0x53574550u /* 'SWEP' */
#define DBS_STREAM_INIT  0x494e4954u /* 'INIT' */
#define DBS_STREAM_NOISE 0x4e4f4953u /* 'NOIS' */
#define DBS_STREAM_SWEEP3 0x53575033u /* 'SWP3': three-colour sweeps of odd lattices */
#define DBS_STEP_X0 (-1)

#ifdef __cplusplus
}
#endif
#endif /* DBS_B200_H */
