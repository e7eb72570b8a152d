/* pzgpu.h -- C ABI of the B200 streaming particle-filter step.
 *
 * This is the drop-in boundary for the one hot path of psg-mit/probzelus-haskell that
 * this repository accelerates: the per-time-step particle filter behind
 *
 *     zparticles   :: Int -> ZStream (Weighted m) a b -> ZStream m a [b]    (haskell/src/Inference.hs:107-108)
 *     zdsparticles :: Int -> ZStream (StateT Heap (Weighted m)) a b -> ...  (haskell/src/Inference.hs:113-114)
 *
 * The reference passes the model as a Haskell closure, which cannot cross to a GPU; here
 * the model is a plain-old-data descriptor (pz_model_desc) for the supported state-space
 * families.  A Haskell host binds these symbols with `foreign import ccall` (the stub is
 * in INTEGRATION.md and haskell/PzGpu.hs); no torch / C++ types appear in any signature.
 *
 * Conventions
 *   - every function returns 0 on success or a negative PZ_E* code; nothing throws;
 *   - pz_last_error() returns a thread-local message for the last failing call;
 *   - the caller owns all host buffers; the library owns device memory, streams, NCCL;
 *   - a pz_filter handle is NOT thread-safe (the reference is single-threaded,
 *     haskell/package.yaml:49-52); calls on one handle must be serialised;
 *   - a call on a handle makes the filter's device the calling thread's current CUDA
 *     device and leaves it current (restoring the previous one would create a context on
 *     a device the host never asked for); handles on different devices may be used from
 *     one thread or from several.
 */
#ifndef PZGPU_H
#define PZGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PZ_ABI_VERSION 2

#if defined(__GNUC__)
#define PZ_API __attribute__((visibility("default")))
#else
#define PZ_API
#endif

#define PZ_MAX_NX 6 /* state scalars per particle for the linear-Gaussian family */
#define PZ_MAX_NY 3 /* observation scalars                                        */

/* error codes */
#define PZ_OK 0
#define PZ_EINVAL (-1)      /* bad argument / unsupported descriptor           */
#define PZ_ECUDA (-2)       /* CUDA runtime failure                            */
#define PZ_ENCCL (-3)       /* NCCL failure / NCCL not loadable                */
#define PZ_EDEGENERATE (-4) /* all weights zero or NaN: nothing to resample    */
#define PZ_ECAPACITY (-5)   /* fixed capacity exceeded (tracks, observations, migration halo) */
#define PZ_ENOMEM (-6)
#define PZ_ESTATE (-7)      /* call not valid in the current filter state      */

/* Model families (a17, a18, a19). */
enum {
  PZ_MODEL_RW1D = 1,     /* Examples/Demo.hs:184-201 (kalman1D*), Inference.hs:143-147 (rw)         */
  PZ_MODEL_LINGAUSS = 2, /* Examples/SingleTargetTracking.hs:28-54; per-track model of MTT :56-73    */
  PZ_MODEL_MTT = 3,      /* Examples/MultiTargetTracking.hs:26-145                                   */
  PZ_MODEL_WALKER = 4,   /* Examples/Walker.hs:20-148: hybrid discrete / continuous walker under the MStream
                            `particles` filter (Inference.hs:73-86): weights carried across resampling          */
  PZ_MODEL_BETA_BERNOULLI = 5, /* Examples/DelayedSamplingAbstract.hs:57-69 betaBernoulli under delayed sampling:
                            Beta prior, Bernoulli observations, conjugate update (DelayedSampling.hs:140,150-151;
                            SymbolicDistr.hs:165-188)                                                         */
  PZ_MODEL_MULTICAM = 6  /* Examples/MultiCamera.hs:40-227: two cameras, box tracks over R^4 (F = H = I), appearance
                            marginals over R^10 shared by the tracks of one identity, per-camera association with the
                            custom proposal (:162-180).  nx = 4, ny = 0 (H = I, R = meas_var I implied).  Uses F / Q / x0 (box prior and clutter mean) and the
                            `mtt` block (clutter_lambda per camera, new_track_lambda, survival_prob, birth_pos_var /
                            birth_vel_var = box prior variances, meas_var; pd = 1).  Observation rows are 16 doubles:
                            camera, pose[10], box[4], appearance reading                                          */
};

enum {
  PZ_RESAMPLE_SYSTEMATIC = 0,     /* canonical fixed-point systematic resampler        */
  PZ_RESAMPLE_MULTINOMIAL_REF = 1 /* N iid categorical draws: resample2, Inference.hs:57-61        */
};

/* Multi-target tracker constants, Examples/MultiTargetTracking.hs:36-73. */
typedef struct pz_mtt_params {
  double clutter_lambda;   /* clutterLambda = 3                                   */
  double new_track_lambda; /* newTrackLambda = birthRate * tdiff = 0.1            */
  double pd;               /* detection probability 0.8                           */
  double survival_prob;    /* exp(-tdiff*deathRate)                               */
  double clutter_var;      /* clutter ~ N(0, clutter_var * I2), 10                */
  double birth_pos_var;    /* birth prior diag(5,5,0.1,0.1)                       */
  double birth_vel_var;
  double meas_var;         /* R = meas_var * I2, 1                                */
  int32_t k_max;           /* fixed per-particle track capacity (<= 16)           */
  int32_t m_max;           /* max observations per step (<= 32)                   */
} pz_mtt_params;

/* Walker constants, Examples/Walker.hs:28-96.  Motion types 0 = Stationary, 1 = Walking, 2 = Running. */
typedef struct pz_walker_params {
  double pos_noise;        /* posNoise = 0.01 (:28-29); coast passes sqrt(dt) * posNoise as the VARIANCE of D.normal (:33-39) */
  double obs_std;          /* positionStdDev = 10 (:85-86); the observation variance is obs_std^2 (:90-91)   */
  double dt;               /* time between measurements, 10 s (:123)                                         */
  double p_init[3];        /* walkerInit: 0.7 / 0.25 / 0.05 (:107-111)                                        */
  double half_life[3][3];  /* motionTypeTransition (:43-49): half_life[from][to] in seconds, 0 = no transition */
  double speed_lo[3];      /* initVelocity (:51-64): speed ~ U(lo, hi) at a uniform angle; Stationary 0..0     */
  double speed_hi[3];
  int32_t exp_mean_as_written; /* 1 = reference behaviour (:69): tTransition = -log(u) * sum(rates), because
                                  D.exponential takes a RATE (Distributions.hs:209-212) and is handed
                                  recip(sum rates); 0 = the intended -log(u) / sum(rates)                   */
  int32_t reserved;
} pz_walker_params;

/* Kernel-side model descriptor.  Matrices are row-major, leading dimension = the
 * logical column count (nx for F/Q/P0, nx for H, ny for R).                                 */
typedef struct pz_model_desc {
  int32_t kind;         /* PZ_MODEL_*                                                        */
  int32_t nx;           /* state dim: 1 (RW1D), 4 (MTT track model), 6 (STT), <= PZ_MAX_NX   */
  int32_t ny;           /* observation dim <= PZ_MAX_NY                                      */
  int32_t scalar_ll_2x; /* 1 = reference's doubled scalar Gaussian score,
                           Distributions.hs:160-161 (only honoured when ny == 1)             */
  int32_t resampler;    /* PZ_RESAMPLE_*                                                     */
  int32_t delayed;      /* 1 = delayed sampling (`delay=True`, SingleTargetTracking.hs:36-38;
                           kalman1DObservedDelayed, Demo.hs:215-242): the state stays a symbolic
                           Gaussian marginal, so every particle carries the exact Kalman
                           posterior (makeMarginal / makeConditional, DelayedSampling.hs:135-152)
                           and all weights are equal.  RW1D / LINGAUSS only.                  */
  int32_t state_f64;    /* 1 = particle state, model arithmetic, log-weights and moment sums in double,
                           the reference's own precision (Inference.hs:28 `type Weight = Log Double`,
                           Distributions.hs:157-166); 0 = fp32 state (half the HBM traffic).  The integer
                           resampler is the same in both.  RW1D / LINGAUSS only.                */
  int32_t reserved;
  double F[PZ_MAX_NX * PZ_MAX_NX];  /* x' ~ N(F x, Q)                                        */
  double Q[PZ_MAX_NX * PZ_MAX_NX];
  double H[PZ_MAX_NY * PZ_MAX_NX];  /* y  ~ N(H x', R)                                       */
  double R[PZ_MAX_NY * PZ_MAX_NY];
  double x0[PZ_MAX_NX];             /* initial state (the reference starts at a point mass)  */
  pz_mtt_params mtt;
  /* ---- ABI 2 ---- */
  double ess_threshold;  /* > 0: adaptive resampling (not in the reference): the population is
                            resampled only when ESS < ess_threshold * N, otherwise particles keep their slot and
                            their log-weights accumulate; after a resample the weights restart from 1.        */
  int32_t carry_weights; /* 1 = the MStream `particles` semantics (Inference.hs:73-86, streamWeights): resample
                            every step, but a child inherits its parent's CUMULATIVE weight.  RW1D / LINGAUSS
                            fp32 / WALKER; runs on the stored-log-weight path.                                */
  int32_t reserved2;
  pz_walker_params walker;
  double beta_a, beta_b; /* PZ_MODEL_BETA_BERNOULLI: the Beta(a, b) prior                                    */
} pz_model_desc;

/* Posterior summary of one step: the weighted particle approximation of
 * p(x_t | y_1..t).  The reference prints `average particles` after resampling
 * (Examples/Demo.hs:207); `mean` is the Rao-Blackwellised version of that estimator
 * (weighted, before resampling: same expectation, lower variance).                          */
typedef struct pz_step_summary {
  int64_t step;             /* 0-based index of the observation just assimilated             */
  double log_evidence_inc;  /* log p^(y_t | y_1..t-1) = log( (1/N) sum_i w_i )                */
  double ess;               /* (sum w)^2 / sum w^2                                           */
  double mean[PZ_MAX_NX];
  double var[PZ_MAX_NX];
  int64_t n_migrated;       /* sharded: this rank's output slots whose ancestor lives on another
                               rank; -1 = some rank's ancestors reach beyond the halo margin
                               (PZ_ECAPACITY, the same verdict on every rank)                 */
  int32_t e_max;            /* canonical resampler diagnostics: global block exponent        */
  int32_t reserved;
  uint64_t total_weight;    /* canonical fixed-point total T                                 */
} pz_step_summary;

typedef struct pz_filter pz_filter;

/* Fill a descriptor with the reference's constants. */
PZ_API int pz_model_rw1d(pz_model_desc* d, double q_var, double r_var);       /* Demo.hs:184-192: (1, 3)  */
PZ_API int pz_model_stt(pz_model_desc* d);                                    /* SingleTargetTracking.hs:39-54 */
PZ_API int pz_model_mtt_track(pz_model_desc* d);                              /* MultiTargetTracking.hs:56-73  */
PZ_API int pz_model_mtt(pz_model_desc* d);                                    /* MultiTargetTracking.hs:36-145 */
PZ_API int pz_model_multicam(pz_model_desc* d);                               /* MultiCamera.hs:40-227 */
PZ_API int pz_model_walker(pz_model_desc* d);                                 /* Walker.hs:20-148 (carry_weights = 1) */
PZ_API int pz_model_beta_bernoulli(pz_model_desc* d, double a, double b);     /* DelayedSamplingAbstract.hs:57-69 (delayed = 1) */

/* zparticles / zdsparticles: a filter over n_particles copies of the model on one device. */
PZ_API int pz_filter_create(const pz_model_desc* model, uint64_t n_particles, uint64_t seed, int device,
                     pz_filter** out);

/* The same over n_devices GPUs driven by ONE host thread of one process (what a single-threaded
 * Haskell host calls): the particles are sharded over the devices, resampling is global, the exchange runs over
 * NVLink peer memory inside the step kernels; pz_filter_step / pz_filter_run / pz_filter_read_state /
 * pz_filter_destroy apply to the returned handle unchanged (the taps that address one device's memory do not).
 * n_particles must be divisible by n_devices * 256.  RW1D / LINGAUSS.
 * PZ_MODEL_MTT / PZ_MODEL_MULTICAM: the trackers shard too (any n_particles that leaves every device at least one
 * 2048-slot resampling tile): each device steps its own particles, parent records are read from their owners over
 * NVLink, the block statistics are all-gathered by peer stores and every device derives the global plan; the run
 * equals the one-device run bit for bit.  pz_filter_read_tracks / pz_filter_read_mcam_tracks apply to the handle.    */
PZ_API int pz_filter_create_multi(const pz_model_desc* model, uint64_t n_particles, uint64_t seed, int n_devices,
                                  const int* device_ids, pz_filter** out);

/* One rank of a filter whose n_global particles are sharded over `world` GPUs (one process
 * per GPU).  nccl_unique_id is the 128-byte ncclUniqueId produced by pz_comm_unique_id() on
 * rank 0 and distributed by the host.                                                        */
PZ_API int pz_comm_unique_id(void* out128);
PZ_API int pz_filter_create_sharded(const pz_model_desc* model, uint64_t n_global, uint64_t seed,
                             int device, int rank, int world, const void* nccl_unique_id,
                             pz_filter** out);

/* ZStream.step (Util/ZStream.hs:16): assimilate one observation, return the posterior.
 * obs is n_obs rows of obs_dim doubles (n_obs == 1 except for MTT).  Blocks until the
 * summary is on the host.                                                                    */
PZ_API int pz_filter_step(pz_filter* f, const double* obs, int32_t n_obs, int32_t obs_dim,
                   pz_step_summary* out);

/* ZS.runStream (Util/ZStream.hs:65-69) over a whole observation stream: n_steps rows of
 * obs_dim doubles are staged on the device once, the steps run back to back, and the
 * summaries (may be NULL) are copied back at the end.                                        */
PZ_API int pz_filter_run(pz_filter* f, const double* obs, int64_t n_steps, int32_t obs_dim,
                  pz_step_summary* out);

/* Device time (ms, CUDA events on the filter's stream) of the kernels enqueued by the last
 * pz_filter_step / pz_filter_run call, and the number of kernel launches it made.           */
PZ_API int pz_filter_last_timing(pz_filter* f, double* ms, int64_t* launches);

/* Diagnostics: run ONE step outside the CUDA graph with an event between its kernels;
 * ms[0] = fused resample+propagate+weight kernel, ms[1] = block-sum scan, ms[2] = partition. */
PZ_API int pz_filter_step_profile(pz_filter* f, const double* obs, int32_t n_obs, int32_t obs_dim,
                                  double ms[3]);

/* The stream's output `[b]`: equally weighted particles AFTER resampling (what the Java
 * visualiser draws).  Returns every `stride`-th output slot, all state dims, slot-major:
 * out[(j/stride)*nx + d].                                                                    */
PZ_API int pz_filter_read_particles(pz_filter* f, int32_t stride, double* out, int64_t out_len);

/* Tracker (PZ_MODEL_MTT) output: for every `stride`-th resampled particle the track map
 * `Map id (mean, cov)` that processObservations returns (MultiTargetTracking.hs:133-139):
 * count_out[i], id_out[i*16+k] (-1 past the count), mean_out[(i*16+k)*4..], cov_out[(i*16+k)*16..]. */
PZ_API int pz_filter_read_tracks(pz_filter* f, int32_t stride, int32_t* count_out, int32_t* id_out, double* mean_out,
                                 double* cov_out, int64_t n_sel_cap);

/* MultiCamera tracker (PZ_MODEL_MULTICAM) output: for every `stride`-th resampled particle the list of `MarginalTrack`
 * that processObservationsStream returns (MultiCamera.hs:208-216): count_out[i]; per track slot k < 16 (-1 / 0 past the
 * count): id_out (trackID), identity_out (index into the particle's appearances), camera_out, start_out (startTime, may
 * be NULL), mean_out[(i*16+k)*4..] and var_out[(i*16+k)*4..] (the box marginal: its covariance is diagonal);
 * n_app_out[i] (may be NULL); app_mean_out[(i*16+a)*10..] and app_cov_out[(i*16+a)*100..] (both NULL or both given).     */
PZ_API int pz_filter_read_mcam_tracks(pz_filter* f, int32_t stride, int32_t* count_out, int32_t* id_out, int32_t* identity_out,
                                      int32_t* camera_out, double* start_out, double* mean_out, double* var_out,
                                      int32_t* n_app_out, double* app_mean_out, double* app_cov_out, int64_t n_sel_cap);

/* Delayed-sampling filters (model.delayed = 1): the Gaussian marginal every particle carries,
 * `MDistr` (mean[nx], cov[nx*nx] row-major) -- what the reference serialises per particle as
 * {"mean":..,"cov":..} (DelayedSampling.hs:69-73).  For PZ_MODEL_BETA_BERNOULLI the marginal is
 * MBeta: mean_out[0] = alpha, mean_out[1] = beta, cov_out[0] = the Beta variance.              */
PZ_API int pz_filter_read_posterior(pz_filter* f, double* mean_out, double* cov_out);

/* Parity / debug taps. */
PZ_API int pz_filter_read_ancestors(pz_filter* f, int32_t* out, int64_t n);      /* ancestors of the pending resample */
PZ_API int pz_filter_read_state(pz_filter* f, float* out, int64_t out_len);       /* pre-resample population, SoA [nx][N] (fp32 filters) */
PZ_API int pz_filter_read_logw(pz_filter* f, float* out, int64_t n);              /* log-weights of that population */
PZ_API int pz_filter_read_state_f64(pz_filter* f, double* out, int64_t out_len);  /* the same, exact for state_f64 filters (widened for fp32 ones) */
PZ_API int pz_filter_read_logw_f64(pz_filter* f, double* out, int64_t n);
PZ_API int64_t pz_filter_n_local(const pz_filter* f);

/* Standalone canonical systematic resampler on stored log-weights (bit-exact tests):
 * lw32_i = (float)logw[i], U = floor(u * 2^32).                                              */
PZ_API int pz_resample_systematic(const double* logw, int64_t n, double u, int32_t* anc_out);
PZ_API int pz_resample_multinomial(const double* logw, int64_t n, uint64_t seed, uint32_t step,
                            int32_t* anc_out);

/* importanceResampling (Inference.hs:68-71): k iid draws from Categorical(exp(logw - max)) on the canonical
 * fixed-point CDF; `resample` (Inference.hs:51-52) is the k = n case = pz_resample_multinomial.             */
PZ_API int pz_importance_resampling(const double* logw, int64_t n, uint64_t seed, uint32_t step, int64_t k,
                                    int32_t* idx_out);

/* The distribution primitives of Distributions.hs as device functions (sample on the GPU from Philox words,
 * score = log-density): parity taps for SURVEY 8a row a11.  params: BERNOULLI {p}; CATEGORICAL {p_0..p_{np-1}};
 * POISSON {lambda}; UNIFORM {a, b}; EXPONENTIAL {rate}; NORMAL {mu, sigma^2} (score doubled as the reference's
 * gaussian_ll', Distributions.hs:160-161).  Samples and arguments of the discrete ones are integers in doubles. */
enum { PZ_DIST_BERNOULLI = 1, PZ_DIST_CATEGORICAL = 2, PZ_DIST_POISSON = 3, PZ_DIST_UNIFORM = 4,
       PZ_DIST_EXPONENTIAL = 5, PZ_DIST_NORMAL = 6 };
PZ_API int pz_dist_sample(int32_t kind, const double* params, int32_t n_params, uint64_t seed, int64_t n, double* out);
PZ_API int pz_dist_score(int32_t kind, const double* params, int32_t n_params, const double* x, int64_t n, double* out);

/* Back to the prior at step 0 (same model, new seed): the way on after PZ_EDEGENERATE / PZ_ECAPACITY, which
 * leave the populations overwritten.  Single-GPU filters.                                                    */
PZ_API int pz_filter_reset(pz_filter* f, uint64_t seed);

PZ_API void pz_filter_destroy(pz_filter* f);
PZ_API const char* pz_last_error(void);
PZ_API int pz_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PZGPU_H */
