/*
 * smjp_b200 -- C ABI of the B200-native engine for the gauenk/smjp hot path
 * (Rao-Teh uniformization Gibbs sweep over semi-Markov jump process trajectories + the bootstrap
 * particle filter behind pmcmc).
 *
 * The reference is pure Python: it has no FFI, so "the reference's FFI for this path" is the set
 * of Python call sites listed below.  Each entry point names the reference function it replaces
 * (paths relative to the reference root); INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add at each of those call sites.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - "dev" entry points take DEVICE pointers and enqueue work on the context's stream without
 *     synchronising (except where a size must come back to the host, documented per call).
 *   - "host" entry points take HOST pointers, stage through pinned buffers owned by the context,
 *     run the same kernels and return after the results are in the caller's buffers.
 *   - chains are ragged and batch-first: chain c owns elements [offs[c], offs[c+1]) of an array.
 *     Grid arrays (W) hold n_c+1 times per chain (W[0]=0 ... W[n_c]=t_end); every per-grid-step
 *     array (uniforms, logZ, aug_idx) and the path arrays reuse the SAME offsets (one spare slot).
 *   - states are 0-based indices into the model's state space; the Python layer maps labels.
 *   - every function returns 0 on success, a negative SMJP_E_* code otherwise;
 *     smjp_last_error() gives the message.  Per-chain problems are reported in status[] (bit mask).
 */
#ifndef SMJP_B200_H
#define SMJP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMJP_ABI_VERSION 1

/* return codes */
#define SMJP_OK 0
#define SMJP_E_INVALID (-1)   /* bad argument (the reference raises ValueError / IndexError) */
#define SMJP_E_CUDA (-2)      /* CUDA runtime failure, see smjp_last_error() */
#define SMJP_E_NOMEM (-3)     /* workspace does not fit on the device */
#define SMJP_E_UNSUPPORTED (-4)

/* per-chain status bits */
#define SMJP_ST_GRID_NOT_INCREASING 1 /* smjp_utils.py:398-399 "We can not have time_p >= time_c" */
#define SMJP_ST_DEGENERATE 2          /* zero / non-finite normaliser (hidden_markov_model.py:170-185) */
#define SMJP_ST_OVERFLOW 4            /* an output or workspace capacity was too small */
#define SMJP_ST_BAD_INPUT 8           /* an observation symbol outside [0, Kobs) (checked on the device: the kernels never index the emission table out of bounds) */

/* message precision */
#define SMJP_F64 0
#define SMJP_F32 1

/* candidate position law */
#define SMJP_CAND_REFERENCE 0 /* unit-scale Weibull density truncated to the sojourn (distributions.py:295) */
#define SMJP_CAND_EXACT 1     /* density proportional to the hazard: tau * U^(1/k) */

/* holding-time family of the particle filter ("hazard_family" option; the filter kernels are Weibull only) */
#define SMJP_HAZARD_WEIBULL 0 /* the reference: smjp_utils.py:419-423 "hazard function is weibull for now" */
#define SMJP_HAZARD_GAMMA 1   /* new: shape = a, scale = theta (BASELINE configs[2]) */

/* particle-filter resampler */
#define SMJP_RESAMPLE_MULTINOMIAL 0 /* utils.py:7-8 sampleDiscrete, pmcmc.py:302 */
#define SMJP_RESAMPLE_SYSTEMATIC 1  /* new (north_star) */

typedef struct smjp_ctx smjp_ctx_t;

int smjp_abi_version(void);
const char* smjp_last_error(void);

/* One context per (device, stream).  `stream` is a cudaStream_t (NULL = legacy default stream). */
smjp_ctx_t* smjp_ctx_create(int device, void* stream);
void smjp_ctx_destroy(smjp_ctx_t* ctx);
int smjp_ctx_sync(smjp_ctx_t* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t smjp_ctx_launch_count(smjp_ctx_t* ctx);
/* launch geometry of the last FFBS kernel: "ffbs_grid", "ffbs_threads", "ffbs_smem"; "num_sms"; "pf_cluster";
 * "ffbs_live_pairs": grid pairs (i, j) inside the live window that the FFBS kernels (K <= 10) processed since the
 * last query (synchronises; the nominal count is n(n+1)/2 per chain); "spec_launches" / "spec_misses": speculative
 * sweep launches and how many of them had to be repeated (option "speculate") */
int64_t smjp_ctx_get_stat(smjp_ctx_t* ctx, const char* key);
/* knobs: "ffbs_threads" (0 = auto, else threads per CTA of the FFBS kernel);
 *        "time_kernels" (0/1, see smjp_ctx_kernel_time);
 *        "hazard_family" (SMJP_HAZARD_*; gamma is honoured by smjp_pf_host only -- the FFBS / candidate /
 *        prior entry points return SMJP_E_UNSUPPORTED for it);
 *        "pf_cluster" (thread-block-cluster size of one particle-filter run: 0 = auto, 1, 2, 4, 8, 16);
 *        "pf_threads" (threads per CTA of the particle filter: 0 = auto, 256, 512, 1024);
 *        "window_log2" (K <= 3 FFBS kernel, scratch mode.  0, the default: EXACT live window -- only grid points
 *        whose states have underflowed to exact zeros are skipped, results unchanged to the last bit.  -L: states
 *        below 2^-L of their row's sum leave the window too (live-window pruning, an approximation; -70 is far
 *        below fp64 rounding of the row sums));
 *        "ffbs_warp" (-1 auto (default), 0, 1: run K <= 3 scratch-mode FFBS with one warp per chain and the live
 *        window in a shared-memory ring of "ffbs_ring" slots of 32 grid points (0 = auto); chains whose window
 *        outgrows the ring are redone by the full-size kernel, so results never depend on the ring);
 *        "f32_rescue" (default 1: chains whose fp32 FFBS pass loses its row sum -- a grid step over which every
 *        live state changes by more than 38 orders of magnitude -- are redone by the fp64 kernel);
 *        "speculate" (default 1: smjp_sweep_dev enqueues a sweep's fill and FFBS kernels for the grid size predicted
 *        from the previous sweep BEFORE this sweep's counts reach the host; the kernels check the real counts on
 *        the device and return at once if the prediction was too small, in which case the sweep is launched again
 *        with the real size.  Results never depend on it; stats "spec_launches" / "spec_misses" count both; 2 = test setting
 *        that under-predicts on purpose so that most launches are repeated);
 *        "pf_extend" (0 = the particle-filter path is extended flat from the last observation to t_end
 *        as the reference does, pmcmc.py:307-309; 1 = the segment is simulated);
 *        "lpt_order" (0/1, default 1: the FFBS kernel starts the longest grids first);
 *        "obs_product" (0 = observations that share a grid interval are summed, as the reference does
 *        (smjp_utils.py:496-499); 1 = multiplied, the proper likelihood);
 *        "chain_base" (global index of chain 0 of this context's batches: keeps the Philox streams of
 *        chains sharded over several GPUs distinct and independent of the sharding);
 *        "ffbs_fast" (-1 auto (default), 0, 1: the warp-per-chain sweep kernels ffbs_fast.cuh (K <= 3, fp64, exact
 *        window) and ffbs_fastk.cuh (4 <= K <= 10) with "fast_ring" / "fast_ways" / "fast_cpb" (0 = auto) as their
 *        geometry; chains they cannot finish go to the rescue pass, so results never depend on them);
 *        "dense_tc" (-1 auto (default: >= 32768 chains, 32 <= K <= 64, fp32), 0, 1: tensor-core forward pass of the
 *        dense shape == 1 path);
 *        "fast_split" (-1 auto (default), 0, 1: the K <= 3 sweep kernel as a forward-only and a backward-only launch
 *        with a scratch region per chain instead of per resident chain; falls back to the fused kernel when the rows
 *        of all chains would not fit in a third of the free device memory; stat "fast_split" says what the last
 *        launch did.  Results never depend on it);
 *        "obs_zero_copy" (-1 auto (default), 0, 1: smjp_sweep_host with PINNED input arrays lets the FFBS kernel's bulk
 *        copies read each chain's observations in place and a gather kernel pull only the path entries in use,
 *        instead of uploading both first; stats "obs_in_place" / "paths_gathered" say what the last call did.
 *        Results never depend on it).
 * More stats: "ffbs_kernel" (which FFBS kernel the last launch used: 0 ffbs.cuh, 1 ffbs_item, 2 ffbs_big, 3 ffbs_warp,
 * 4 ffbs_fast, 5 ffbs_fastk), "ffbs_rescued", "ffbs_fast_boost_f64" / "_f32", "dense_tc". */
int smjp_ctx_set_option(smjp_ctx_t* ctx, const char* key, int64_t value);
/* With option "time_kernels" = 1 every launch of the named kernel ("ffbs", "hmm", "pf") is bracketed by CUDA events
 * on the launching stream; this returns and resets their summed duration and count (synchronises). */
int smjp_ctx_kernel_time(smjp_ctx_t* ctx, const char* kernel, double* total_ms, int64_t* count);

/*
 * Model parameters (HOST pointers, copied).  Replaces the parameter reads the reference does through
 * smjpHazardFunction.{shape_mat,scale_mat,omega} (smjp_utils.py:425-429), create_upperbound_scale
 * (:180-187; B = omega*A and A_hat = (omega-1)*A are applied in-kernel), smjpEmission
 * .{likelihood_power, final_grid_time} (:484-491) and its d_create table (:552-557), given here as a
 * general row-stochastic emis[K*Kobs].  Call again after a parameter step (raoteh.py:173-193).
 */
int smjp_set_model(smjp_ctx_t* ctx, int K, int Kobs, const double* shape, const double* scale,
                   double omega, const double* emis, double likelihood_power, double t_end);

/*
 * FFBS for C chains: forward filter over the augmented (v,l) states + backward sampling + path
 * compaction + chain metrics.  Replaces smjp_ffbs / sample_smjp_trajectory_posterior
 * (smjp_utils.py:261-330) = likelihood_and_decoding_hmm (fast_hmm.py:202-335) with
 * smjpTransitionFunction (smjp_utils.py:382-416) and smjpEmission (:506-550) evaluated in-kernel,
 * backward_sampling_hmm (fast_hmm.py:104-191), include_endpoint_in_thinned_samples
 * (smjp_utils.py:338-348) and compute_evaluation_chain_metrics (mcmc_utils.py:40-56).
 *
 *   w_offs[C+1], W[w_offs[C]]           grids
 *   obs_offs[C+1], obs_t[], obs_y[]     observations sorted by time per chain (sMJPDataWrapper)
 *   uniforms[] or NULL                  u[w_offs[c]+i] drives the categorical draw at grid step i
 *                                       (inverse CDF in the reference's state-major order
 *                                       a = v*n + j); NULL = Philox stream (seed, sweep, chain)
 *   precision                           SMJP_F64 | SMJP_F32 (arithmetic and message type)
 *   messages or NULL, msg_offs[C+1]     optional export of the filtered messages: chain c, row i
 *                                       (i = 0..n-1) starts at msg_offs[c] + K*i*(i+1)/2 and holds
 *                                       alpha_i(v, W[j]) at [v*(i+1) + j], j <= i (normalised rows,
 *                                       fast_hmm.py:231,321); element type per `precision`
 *   logZ[] or NULL                      log row normalisers per grid step (discarded by the reference)
 *   aug_idx[] or NULL                   sampled dense index a_i = v*n + j per grid step
 *                                       (= `samples` of fast_hmm.py:191)
 *   path_v[], path_t[], path_len[C]     compacted path (s_states, s_times of smjp_utils.py:327-329)
 *   time_in_state[C*K], jumps[C*K]      or NULL
 *   status[C]
 */
int smjp_ffbs_dev(smjp_ctx_t* ctx, int C, const int64_t* w_offs, const double* W,
                  const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                  const double* uniforms, uint64_t seed, uint32_t sweep, int precision,
                  void* messages, const int64_t* msg_offs, double* logZ, int32_t* aug_idx,
                  int32_t* path_v, double* path_t, int32_t* path_len, double* time_in_state,
                  int32_t* jumps, int32_t* status,
                  /* host copies of the ragged sizes, needed to size the workspace */
                  int64_t n_max, int64_t total_w);

/* Same call with HOST buffers (host<->device copies inside). messages/msg_offs may be NULL. */
int smjp_ffbs_host(smjp_ctx_t* ctx, int C, const int64_t* w_offs, const double* W,
                   const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                   const double* uniforms, uint64_t seed, uint32_t sweep, int precision,
                   void* messages, const int64_t* msg_offs, double* logZ, int32_t* aug_idx,
                   int32_t* path_v, double* path_t, int32_t* path_len, double* time_in_state,
                   int32_t* jumps, int32_t* status);

/*
 * Prior trajectories, one per chain.  Replaces sample_smjp_trajectory_prior (smjp_utils.py:216-254):
 * competing Weibull risks, argmin, last entry overwritten with (v_prev, t_end).  The reference draws
 * unit-scale holding times (distributions.py:295) -- honour_scale != 0 multiplies by lambda.
 * p_offs[C+1] are capacity windows of path_v/path_t (chain c may hold p_offs[c+1]-p_offs[c] entries;
 * SMJP_ST_OVERFLOW if it needs more).  pi0 is a HOST array [K].
 */
int smjp_prior_dev(smjp_ctx_t* ctx, int C, const int64_t* p_offs, const double* pi0, int honour_scale,
                   uint64_t seed, int32_t* path_v, double* path_t, int32_t* path_len, int32_t* status);

/*
 * Synthetic observations y_d ~ emis[v(t_d), :] at the given times.  Replaces sample_data_posterior
 * (smjp_utils.py:350-364) + smjpEmission.sample (:559-563).
 */
int smjp_simulate_obs_dev(smjp_ctx_t* ctx, int C, const int64_t* p_offs, const int32_t* path_v,
                          const double* path_t, const int32_t* path_len, const int64_t* obs_offs,
                          const double* obs_t, uint64_t seed, int32_t* obs_y);

/*
 * Thinned-event candidates: W = sorted(thinned U T) per chain.  Replaces sample_smjp_event_times
 * (smjp_utils.py:194-210) + PoissonProcess.sample (:61-99) with the A_hat = (omega-1)*A process.
 * Two passes over the same counter-based draws:
 *   count: writes w_offs[C+1] (device), soff[] (device scratch, int32, same windows as the path
 *          arrays: entry q of a chain's window = the grid index of jump q, entry path_len-1 = the
 *          end point's) and returns n_max / total_w to the HOST (synchronises the stream);
 *   fill:  writes W[total_w] (device).
 */
int smjp_candidates_count_dev(smjp_ctx_t* ctx, int C, const int64_t* p_offs, const int32_t* path_v,
                              const double* path_t, const int32_t* path_len, uint64_t seed,
                              uint32_t sweep, int64_t* w_offs, int32_t* soff, int64_t* n_max_host,
                              int64_t* total_w_host);
int smjp_candidates_fill_dev(smjp_ctx_t* ctx, int C, const int64_t* p_offs, const int32_t* path_v,
                             const double* path_t, const int32_t* path_len, uint64_t seed,
                             uint32_t sweep, int mode, const int64_t* w_offs, const int32_t* soff,
                             double* W);
/* HOST buffers; W has room for w_capacity doubles. If the grids need more, returns SMJP_E_NOMEM with
 * the required size in *total_w (w_offs is valid) so that the caller can retry. */
int smjp_candidates_host(smjp_ctx_t* ctx, int C, const int64_t* p_offs, const int32_t* path_v,
                         const double* path_t, const int32_t* path_len, uint64_t seed, uint32_t sweep,
                         int mode, int64_t* w_offs, double* W, int64_t w_capacity, int64_t* total_w);

/*
 * One Rao-Teh Gibbs sweep for C chains, the loop body raoteh.py:58-59:
 *     W = sample_smjp_event_times(A_hat, V, T);  V, T = smjp_ffbs(W, data, ...)
 * DEVICE buffers.  The new grids/paths use windows w_offs (written here); W, path_v_out and
 * path_t_out must have room for `capacity` elements, else SMJP_E_NOMEM is returned with the required
 * size in *total_w_host (nothing else written).  Backward uniforms come from the Philox stream
 * (seed, sweep, chain).  Returns when the two sizes are on the host; the FFBS kernel is then still running (or,
 * with option "speculate", already enqueued before the sizes arrived): outputs are valid in stream order.
 */
int smjp_sweep_dev(smjp_ctx_t* ctx, int C, const int64_t* p_offs, const int32_t* path_v,
                   const double* path_t, const int32_t* path_len, const int64_t* obs_offs,
                   const double* obs_t, const int32_t* obs_y, uint64_t seed, uint32_t sweep, int mode,
                   int precision, int64_t capacity, int64_t* w_offs, int32_t* soff, double* W,
                   int32_t* path_v_out, double* path_t_out, int32_t* path_len_out,
                   double* time_in_state, int32_t* jumps, int32_t* status, int64_t* n_max_host,
                   int64_t* total_w_host);
/* Same sweep with HOST buffers (one call = H2D of paths+observations, kernels, D2H of W, V, T and
 * metrics).  W / path outputs have room for `capacity` elements (retry protocol as above). */
int smjp_sweep_host(smjp_ctx_t* ctx, int C, const int64_t* p_offs, const int32_t* path_v,
                    const double* path_t, const int32_t* path_len, const int64_t* obs_offs,
                    const double* obs_t, const int32_t* obs_y, uint64_t seed, uint32_t sweep, int mode,
                    int precision, int64_t capacity, int64_t* w_offs, double* W, int32_t* path_v_out,
                    double* path_t_out, int32_t* path_len_out, double* time_in_state, int32_t* jumps,
                    int32_t* status, int64_t* total_w);
/* The same call in two halves, for callers that keep several batches in flight (one context per batch: while one
 * batch's kernels run, the next batch's inputs are staged and its kernels queue up behind -- the sweep kernels are
 * persistent, so the second batch also fills the slots the first one's tail leaves idle).  _begin returns as soon as
 * this sweep's grid sizes are on the host (total_w is valid; SMJP_E_CAPACITY as above) with the kernels and copies
 * still in flight; nothing in the output buffers may be read, and no input buffer changed, before _end has returned.
 * smjp_sweep_host == _begin + _end.  Small pageable batches complete inside _begin. */
int smjp_sweep_host_begin(smjp_ctx_t* ctx, int C, const int64_t* p_offs, const int32_t* path_v,
                          const double* path_t, const int32_t* path_len, const int64_t* obs_offs,
                          const double* obs_t, const int32_t* obs_y, uint64_t seed, uint32_t sweep, int mode,
                          int precision, int64_t capacity, int64_t* w_offs, double* W, int32_t* path_v_out,
                          double* path_t_out, int32_t* path_len_out, double* time_in_state, int32_t* jumps,
                          int32_t* status, int64_t* total_w);
int smjp_sweep_host_end(smjp_ctx_t* ctx);

/*
 * Dense S-state HMM: scaled forward filter + backward sampler, chains batched (ragged step windows
 * t_offs[C+1]).  Replaces HiddenMarkovModel.likelihood / .backward_sampling
 * (hidden_markov_model.py:213-313, :88-202) and is the K-state form of the sMJP filter when every
 * shape is 1.  HOST buffers.
 *   P          [S*S] when p_steps == 0, else [p_steps][S*S] with P[t] the transition into step t
 *   E          dense emission likelihoods [total_steps][S], or NULL to use the table emis[S*Kobs]
 *              with symbols y[total_steps]
 *   uniforms   per step or NULL (Philox stream (seed, 0, chain))
 *   alpha_hat  [total_steps][S] row-normalised messages (element type per `precision`), logc
 *              [total_steps] log scale factors (log alpha_t = log alpha_hat_t + sum_{u<=t} logc[u]),
 *              loglik[C] = log P(data), states[total_steps] sampled path (NULL = forward only)
 */
int smjp_hmm_dense_host(smjp_ctx_t* ctx, int C, int S, const int64_t* t_offs, const double* P, int p_steps,
                        const double* E, const double* emis, int Kobs, const int32_t* y, const double* pi0,
                        const double* uniforms, uint64_t seed, int precision, void* alpha_hat, double* logc,
                        double* loglik, int32_t* states, int32_t* status);

/* The same on DEVICE pointers, asynchronous on the context's stream (HiddenMarkovModel.likelihood / backward_sampling,
 * hidden_markov_model.py:213-313, :88-202, for batches that are already resident).  logc / loglik are required here;
 * sweep keys the Philox stream of the backward draws when uniforms is NULL. */
int smjp_hmm_dense_dev(smjp_ctx_t* ctx, int C, int S, const int64_t* t_offs, const double* P, int p_steps,
                       const double* E, const double* emis, int Kobs, const int32_t* y, const double* pi0,
                       const double* uniforms, uint64_t seed, uint32_t sweep, int precision, void* alpha_hat,
                       double* logc, double* loglik, int32_t* states, int32_t* status);

/*
 * Bootstrap particle filter, one filter run of N particles per chain.  Replaces smjp_smc
 * (pmcmc.py:224-310) with the two reference defects removed.  Uses the model of
 * smjp_set_model (hazard A = shape/scale, emission table, t_end).  HOST buffers.
 *   pi0[K], N particles, resampler SMJP_RESAMPLE_*
 *   logZ[total_obs] per-observation log normalisers (or NULL), loglik[C] = sum_d Z_d
 *   path of particle 0 after the last resampling, in capacity windows p_offs[C+1]
 */
int smjp_pf_host(smjp_ctx_t* ctx, int C, const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                 const double* pi0, int N, uint64_t seed, int resampler, double* logZ, double* loglik,
                 const int64_t* p_offs, int32_t* path_v, double* path_t, int32_t* path_len, int32_t* status);

/*
 * The same filter on DEVICE pointers, asynchronous on the context's stream (pmcmc over many chains stays resident:
 * nothing is staged, nothing synchronises).  max_obs_per_chain bounds obs_offs[c+1] - obs_offs[c] (workspace
 * sizing); logZ may be NULL.  Replaces smjp_smc (pmcmc.py:224-310) like smjp_pf_host.
 */
int smjp_pf_dev(smjp_ctx_t* ctx, int C, const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                const double* pi0, int N, uint64_t seed, int resampler, int64_t max_obs_per_chain, double* logZ,
                double* loglik, const int64_t* p_offs, int32_t* path_v, double* path_t, int32_t* path_len,
                int32_t* status);

/*
 * FFBS on the dense K-state path: the shape == 1 collapse of the sMJP, K <= 64 (smjp_set_model).
 * Same inputs and path outputs as smjp_ffbs_dev; requires every Weibull shape of the model to be 1
 * (SMJP_E_INVALID otherwise).  m_i ~ e_i o (m_{i-1} B) with the constant
 * B = (1 - 1/omega) I + diag(1/(omega A_v)) A; the backward pass draws the state with u_state[] and,
 * when the state repeats, thinned-vs-self-jump with u_jump[] (both per grid step, same offsets as W;
 * NULL = Philox).  Replaces the same reference lines as smjp_ffbs_dev for exponential holding times and is
 * the K-state form of hidden_markov_model.py:213-313 / :88-202.  DEVICE buffers, asynchronous.
 */
int smjp_ffbs_dense_dev(smjp_ctx_t* ctx, int C, const int64_t* w_offs, const double* W,
                        const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                        const double* u_state, const double* u_jump, uint64_t seed, uint32_t sweep,
                        int precision, double* logZ, int32_t* path_v, double* path_t, int32_t* path_len,
                        double* time_in_state, int32_t* jumps, int32_t* status, int64_t total_w);
/* Rao-Teh sweep (candidates + dense FFBS) for shape == 1 models; arguments as smjp_sweep_dev. */
int smjp_sweep_dense_dev(smjp_ctx_t* ctx, int C, const int64_t* p_offs, const int32_t* path_v,
                         const double* path_t, const int32_t* path_len, const int64_t* obs_offs,
                         const double* obs_t, const int32_t* obs_y, uint64_t seed, uint32_t sweep, int mode,
                         int precision, int64_t capacity, int64_t* w_offs, int32_t* soff, double* W,
                         int32_t* path_v_out, double* path_t_out, int32_t* path_len_out,
                         double* time_in_state, int32_t* jumps, int32_t* status, int64_t* n_max_host,
                         int64_t* total_w_host);

/*
 * Parameter step: per-transition holding-time statistics of the current trajectories
 * and the reference's Weibull shape estimate, for every (v, v') and every group -- a group is one chain
 * (pooled = 0, G = C) or all chains together (pooled = 1, G = 1).  Replaces filter_T_by_state
 * (pkg/raoteh.py:136-151), esimtate_weibull_shape (:153-167: the FIRST k of arange(grid, grid_max, grid)
 * minimising |sum T^k ln T / sum T^k - 1/k - mean ln T|; reference defaults grid = 0.001, grid_max = 5) and
 * esimtate_weibull_shape_mat (:119-134).  khat / gval are NaN where count == 0 (the reference then draws
 * U(0.6, 3) on the host).  Paths in the layout of smjp_sweep_dev; outputs [G, K, K]; DEVICE buffers,
 * asynchronous on the context's stream.  K is passed explicitly (no model needed): path_v holds indices in [0, K).
 */
int smjp_shape_mle_dev(smjp_ctx_t* ctx, int C, int K, const int64_t* p_offs, const int32_t* path_v,
                       const double* path_t, const int32_t* path_len, int pooled, double grid, double grid_max,
                       double* khat, double* gval, int32_t* count, double* sum_log, double* sum_t);

#ifdef __cplusplus
}
#endif
#endif /* SMJP_B200_H */
