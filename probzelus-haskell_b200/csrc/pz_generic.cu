// pz_generic.cu -- the stored-log-weight path of the particle-filter step: a particle carries an fp32
// log-weight in HBM, ancestors are materialised, and the step runs as separate kernels.  It serves the
// variants of the reference's filter that the fused kernel (pz_kernels.cu) does not:
//
//   carry_weights   the MStream filter `particles` (Inference.hs:73-86): streamWeights threads the CUMULATIVE
//                   weight through every yield, particles' resamples by it every step, and a child inherits it;
//   ess_threshold   adaptive resampling (not in the reference): resample only when
//                   ESS < threshold * N, otherwise keep the slots and let the log-weights accumulate;
//   PZ_MODEL_WALKER Examples/Walker.hs:20-148: position/velocity walker whose motion type switches after
//                   exponential dwell times (a data-dependent loop per particle), run by `particles 1000`;
//   pz_dist_*       the primitives of Distributions.hs as device functions (parity taps);
//   beta_bernoulli  Beta prior / Bernoulli observations under delayed sampling (DelayedSampling.hs:140,150-151).
//
// The resampler is the canonical one (pf_stats_from_lw -> pf_scan_blocks -> pf_partition -> pf_ancestors of
// pz_kernels.cu); with carry_weights = 0 and ess_threshold = 0 the linear-Gaussian step here reproduces the
// fused kernel's population bit for bit (tests/test_generic.py).
#include <cmath>
#include <cstring>

#include "pz_dist.cuh"
#include "pz_generic.cuh"
#include "pz_model.cuh"

namespace pz {

constexpr float kLn2f = 0x1.62e43p-1f;  // the carried log-weights are re-based by emax * ln 2 every step

// ------------------------------------------------------------------------------------------
// ancestors: identity when the adaptive rule decided not to resample
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gen_anc_identity(const CtlDev* ctl, int32_t* anc, int64_t n) {
  if (ctl->ess_resampled) return;
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j < n) anc[j] = (int32_t)j;
}

// carried log-weight of the parent of slot j, re-based (0 when the weights restart)
__device__ __forceinline__ float carried_lw(const GenArgs& g, const float* __restrict__ lw_in, int64_t parent, bool carry,
                                            int32_t emax) {
  if (!carry) return 0.0f;
  const float c = __fmul_rn((float)emax, kLn2f);
  return __fsub_rn(lw_in[parent], c);
}

// ------------------------------------------------------------------------------------------
// linear-Gaussian step, fp32 state: one thread per output slot
// ------------------------------------------------------------------------------------------
template <int NX, int NY>
__global__ void __launch_bounds__(kThreads) gen_step_lg(const GenArgs g) {
  __shared__ float4 s_tab[PZ_ICDF_ROWS];
  for (int i = threadIdx.x; i < PZ_ICDF_ROWS; i += kThreads) s_tab[i] = g_icdf_tab[i];
  const float4* tab = s_tab;
  const float4* tab_m8 = tab - 8;   // the sub-segment field of a table index carries the leading one
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= g.n) return;
  const uint32_t t = g.ctl->t;
  const int cur = t & 1, nxt = cur ^ 1;
  const float* __restrict__ xin = reinterpret_cast<const float*>(g.X[cur]);
  float* __restrict__ xout = reinterpret_cast<float*>(g.X[nxt]);
  const bool carry = g.carry_weights || !g.ctl->ess_resampled;
  const int64_t p = g.anc[j];
  float x[NX], z[NX], xn[NX], y[NY];
#pragma unroll
  for (int d = 0; d < NX; ++d) x[d] = xin[(size_t)d * g.ld + p];
#pragma unroll
  for (int k = 0; k < NY; ++k) y[k] = (float)g.obs[k];
  lg_noise1<NX, float>((uint32_t)j, t, g.seed, tab_m8, z);
  lg_propagate<NX>(g.model, x, z, xn);
  const float inc = lg_logweight<NX, NY>(g.model, xn, y);
#pragma unroll
  for (int d = 0; d < NX; ++d) xout[(size_t)d * g.ld + j] = xn[d];
  g.lw[nxt][j] = carry ? __fadd_rn(carried_lw(g, g.lw[cur], p, true, g.plan->emax), inc) : inc;
}

// ------------------------------------------------------------------------------------------
// Walker (Examples/Walker.hs): state (x, y, vx, vy, motion type) in double, SoA [5][ld]
// ------------------------------------------------------------------------------------------
struct WalkerState { double x, y, vx, vy; int mt; };

// initVelocity (Walker.hs:51-64)
__device__ __forceinline__ void walker_init_velocity(const pz_walker_params& w, PhiloxSeq& rng, int mt, double& vx, double& vy) {
  if (w.speed_hi[mt] <= 0.0) { vx = 0.0; vy = 0.0; return; }
  const double speed = sample_uniform(rng, w.speed_lo[mt], w.speed_hi[mt]);
  const double angle = sample_uniform(rng, 0.0, 6.283185307179586);
  vx = speed * cos(angle);
  vy = speed * sin(angle);
}

// coast (Walker.hs:31-41): D.normal takes a VARIANCE and is handed sqrt(dt) * posNoise
__device__ __forceinline__ void walker_coast(const pz_walker_params& w, PhiloxSeq& rng, const float4* tab_m8, double dt, WalkerState& s) {
  const double var = sqrt(dt) * w.pos_noise;
  s.x = sample_normal(rng, s.x + s.vx * dt, var, tab_m8);
  s.y = sample_normal(rng, s.y + s.vy * dt, var, tab_m8);
}

// motion (Walker.hs:67-81): exponential dwell time, then a categorical jump and a fresh velocity, until dt is used up
__device__ __forceinline__ void walker_motion(const pz_walker_params& w, PhiloxSeq& rng, const float4* tab_m8, WalkerState& s) {
  double dt = w.dt;
  for (int it = 0; it < 100000; ++it) {
    double rates[3], sum = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      rates[k] = w.half_life[s.mt][k] > 0.0 ? 0.6931471805599453 / w.half_life[s.mt][k] : 0.0;
      sum += rates[k];
    }
    // D.exponential (recip (sum rates)): -log u / (1 / sum) as written; the intended law divides by the sum
    const double lambda = w.exp_mean_as_written ? 1.0 / sum : sum;
    const double t_tr = sum > 0.0 ? sample_exponential(rng, lambda) : INFINITY;
    if (t_tr > dt) { walker_coast(w, rng, tab_m8, dt, s); return; }
    walker_coast(w, rng, tab_m8, t_tr, s);
    s.mt = sample_categorical(rng, rates, 3);      // labeledLogCategorical transLam
    walker_init_velocity(w, rng, s.mt, s.vx, s.vy);
    dt -= t_tr;
  }
}

__global__ void __launch_bounds__(kThreads) gen_init_walker(const GenArgs g) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= g.n) return;
  PhiloxSeq rng(g.seed, (uint32_t)j, 0u, kStreamWalkInit);
  WalkerState s;
  s.x = 0.0; s.y = 0.0;
  s.mt = sample_categorical(rng, g.walker.p_init, 3);   // walkerInit (Walker.hs:105-111)
  walker_init_velocity(g.walker, rng, s.mt, s.vx, s.vy);
  for (int b = 0; b < 2; ++b) {
    double* x = reinterpret_cast<double*>(g.X[b]);
    x[j] = s.x; x[g.ld + j] = s.y; x[2 * g.ld + j] = s.vx; x[3 * g.ld + j] = s.vy; x[4 * g.ld + j] = (double)s.mt;
    g.lw[b][j] = 0.0f;
  }
}

__global__ void __launch_bounds__(kThreads) gen_step_walker(const GenArgs g) {
  __shared__ float4 s_tab[PZ_ICDF_ROWS];
  for (int i = threadIdx.x; i < PZ_ICDF_ROWS; i += kThreads) s_tab[i] = g_icdf_tab[i];
  const float4* tab = s_tab;
  const float4* tab_m8 = tab - 8;   // the sub-segment field of a table index carries the leading one
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= g.n) return;
  const uint32_t t = g.ctl->t;
  const int cur = t & 1, nxt = cur ^ 1;
  const double* __restrict__ xin = reinterpret_cast<const double*>(g.X[cur]);
  double* __restrict__ xout = reinterpret_cast<double*>(g.X[nxt]);
  const bool carry = g.carry_weights || !g.ctl->ess_resampled;
  const int64_t p = g.anc[j];
  WalkerState s;
  s.x = xin[p]; s.y = xin[g.ld + p]; s.vx = xin[2 * g.ld + p]; s.vy = xin[3 * g.ld + p]; s.mt = (int)xin[4 * g.ld + p];
  PhiloxSeq rng(g.seed, (uint32_t)j, t, kStreamWalker);
  walker_motion(g.walker, rng, tab_m8, s);
  // walkerMeasure (Walker.hs:88-94): two scalar observes, each the doubled Gaussian score
  const double v = g.walker.obs_std * g.walker.obs_std;
  const double inc = score_normal_2x(s.x, v, g.obs[0]) + score_normal_2x(s.y, v, g.obs[1]);
  xout[j] = s.x; xout[g.ld + j] = s.y; xout[2 * g.ld + j] = s.vx; xout[3 * g.ld + j] = s.vy; xout[4 * g.ld + j] = (double)s.mt;
  const float incf = (float)inc;
  g.lw[nxt][j] = carry ? __fadd_rn(carried_lw(g, g.lw[cur], p, true, g.plan->emax), incf) : incf;
}

// ------------------------------------------------------------------------------------------
// posterior summary from the quantised weights (the weights the resampler uses): per-CTA partials in a
// fixed order, then one CTA -- deterministic
// ------------------------------------------------------------------------------------------
constexpr int kGenRec = 2 + 2 * PZ_MAX_NX;   // sw, sw2, sx[6], sxx[6]

__global__ void __launch_bounds__(kThreads) gen_summary_partial(const GenArgs g) {
  __shared__ double s_red[kWarps][kGenRec];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t j = (int64_t)blockIdx.x * kThreads + tid;
  const uint32_t t_new = g.ctl->t + 1u;   // the population just produced lives in buffer (t + 1) & 1
  const int32_t emax = g.plan->emax;
  double v[kGenRec];
#pragma unroll
  for (int r = 0; r < kGenRec; ++r) v[r] = 0.0;
  if (j < g.n) {
    const int32_t eb = g.eb[j / kBlk];
    if (eb != kNZero) {
      const double w = ldexp((double)g.q16[j], eb - emax - 15);
      v[0] = w; v[1] = w * w;
      for (int d = 0; d < g.n_summary; ++d) {
        double x;
        if (g.kind == PZ_MODEL_WALKER) {
          const double* X = reinterpret_cast<const double*>(g.X[t_new & 1]);
          x = d < 4 ? X[(size_t)d * g.ld + j] : ((int)X[(size_t)4 * g.ld + j] == d - 3 ? 1.0 : 0.0);  // d = 4, 5: P(walking), P(running)
        } else {
          x = (double)reinterpret_cast<const float*>(g.X[t_new & 1])[(size_t)d * g.ld + j];
        }
        const double dx = x - g.plan->center64[d];
        v[2 + d] = w * dx; v[2 + PZ_MAX_NX + d] = w * dx * dx;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kGenRec; ++r) {
    double x = v[r];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)__double_as_longlong(x), d);
      const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(__double_as_longlong(x) >> 32), d);
      x += __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    }
    if (lane == 0) s_red[warp][r] = x;
  }
  __syncthreads();
  if (tid < kGenRec) {
    double x = 0;
    for (int w = 0; w < kWarps; ++w) x += s_red[w][tid];
    g.part[(size_t)blockIdx.x * kGenRec + tid] = x;
  }
}

__global__ void __launch_bounds__(kThreads) gen_summary_final(const GenArgs g, int nparts) {
  __shared__ double s_tot[kGenRec];
  __shared__ double s_grp[kThreads / 32][kGenRec];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int r = 0; r < kGenRec; ++r) {
    double x = 0;
    for (int c = tid; c < nparts; c += kThreads) x += g.part[(size_t)c * kGenRec + r];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)__double_as_longlong(x), d);
      const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(__double_as_longlong(x) >> 32), d);
      x += __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    }
    if (lane == 0) s_grp[warp][r] = x;
  }
  __syncthreads();
  if (tid < kGenRec) {
    double x = 0;
    for (int w = 0; w < kThreads / 32; ++w) x += s_grp[w][tid];
    s_tot[tid] = x;
  }
  __syncthreads();
  if (tid != 0) return;
  PlanDev* pl = g.plan;
  CtlDev* ctl = g.ctl;
  pz_step_summary* o = g.summ;
  const double sw = s_tot[0], sw2 = s_tot[1];
  const bool carried = g.carry_weights || !ctl->ess_resampled;   // how the population just weighted got its weights
  for (int d = 0; d < PZ_MAX_NX; ++d) {
    if (d < g.n_summary && sw > 0) {
      const double m = s_tot[2 + d] / sw;
      const double mean = pl->center64[d] + m;
      o->mean[d] = mean;
      o->var[d] = s_tot[2 + PZ_MAX_NX + d] / sw - m * m;
      pl->center64[d] = mean;
      pl->center[d] = (float)mean;
    } else { o->mean[d] = 0.0; o->var[d] = 0.0; }
  }
  // log of the sum of the stored weights; the evidence increment is its ratio to the previous sum (re-based by the
  // same constant the step kernel subtracted) when the weights were carried, to N when they restarted
  const double cur_log = sw > 0 ? log(sw) + (double)pl->emax * 0.6931471805599453 : -INFINITY;
  const double rebase = (double)__fmul_rn((float)g.gen->prev_emax, kLn2f);
  o->log_evidence_inc = carried ? cur_log - (g.gen->prev_log - rebase) : cur_log - log((double)g.n);
  o->log_evidence_inc += g.ll_const;
  o->step = ctl->t;
  o->ess = sw > 0 ? sw * sw / sw2 : 0.0;
  o->n_migrated = 0;
  o->e_max = pl->emax;
  o->reserved = (int32_t)ctl->ess_resampled;   // 1: this step's population came from a resample
  o->total_weight = pl->total;
  g.gen->prev_log = cur_log;
  g.gen->prev_emax = pl->emax;
  // the adaptive rule for the NEXT resample (pf_partition advances ctl->t afterwards)
  ctl->ess_resampled = (g.ess_threshold <= 0.0 || o->ess < g.ess_threshold * (double)g.n) ? 1u : 0u;
}

__global__ void gen_state_init(GenState* st, double n) { st->prev_log = log(n); st->prev_emax = 0; }

// ------------------------------------------------------------------------------------------
// LG initial population (point mass, equal weights)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gen_init_lg(const GenArgs g) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= g.n) return;
  for (int b = 0; b < 2; ++b) {
    float* x = reinterpret_cast<float*>(g.X[b]);
    for (int d = 0; d < g.model.nx; ++d) x[(size_t)d * g.ld + j] = g.model.x0[d];
    g.lw[b][j] = 0.0f;
  }
}

// ------------------------------------------------------------------------------------------
// distribution taps
// ------------------------------------------------------------------------------------------
struct DistArgs { int32_t kind, np; double p[8]; uint64_t seed; int64_t n; const double* x; double* out; };

__global__ void __launch_bounds__(kThreads) dist_sample_kernel(const DistArgs a) {
  __shared__ float4 s_tab[PZ_ICDF_ROWS];
  for (int i = threadIdx.x; i < PZ_ICDF_ROWS; i += kThreads) s_tab[i] = g_icdf_tab[i];
  const float4* tab = s_tab;
  const float4* tab_m8 = tab - 8;   // the sub-segment field of a table index carries the leading one
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n) return;
  PhiloxSeq rng(a.seed, (uint32_t)i, 0u, kStreamDist);
  double v = 0;
  switch (a.kind) {
    case PZ_DIST_BERNOULLI: v = sample_bernoulli(rng, a.p[0]); break;
    case PZ_DIST_CATEGORICAL: v = sample_categorical(rng, a.p, a.np); break;
    case PZ_DIST_POISSON: v = sample_poisson(rng, a.p[0]); break;
    case PZ_DIST_UNIFORM: v = sample_uniform(rng, a.p[0], a.p[1]); break;
    case PZ_DIST_EXPONENTIAL: v = sample_exponential(rng, a.p[0]); break;
    case PZ_DIST_NORMAL: v = sample_normal(rng, a.p[0], a.p[1], tab_m8); break;
  }
  a.out[i] = v;
}

__global__ void __launch_bounds__(kThreads) dist_score_kernel(const DistArgs a) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n) return;
  const double x = a.x[i];
  double v = 0;
  switch (a.kind) {
    case PZ_DIST_BERNOULLI: v = score_bernoulli(a.p[0], x); break;
    case PZ_DIST_CATEGORICAL: v = score_categorical(a.p, a.np, x); break;
    case PZ_DIST_POISSON: v = score_poisson(a.p[0], x); break;
    case PZ_DIST_UNIFORM: v = score_uniform(a.p[0], a.p[1], x); break;
    case PZ_DIST_EXPONENTIAL: v = x >= 0 ? log(a.p[0]) - a.p[0] * x : -INFINITY; break;
    case PZ_DIST_NORMAL: v = score_normal_2x(a.p[0], a.p[1], x); break;
  }
  a.out[i] = v;
}

// ------------------------------------------------------------------------------------------
// Beta-Bernoulli under delayed sampling: every particle carries the same Beta marginal (one thread)
// ------------------------------------------------------------------------------------------
__global__ void beta_bernoulli_step(BetaDev* st, pz_step_summary* out, double n_particles) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double a = st->a, b = st->b;
  const double p = a / (a + b);                 // makeMarginal (MBeta a b) CBernoulli = MBernoulli (a / (a + b))
  const bool y = st->y != 0.0;
  const double ll = y ? log(p) : log(1.0 - p);  // score of the marginal (bernoulli_ll, Distributions.hs:148-149)
  if (y) st->a = a + 1.0; else st->b = b + 1.0; // makeConditional (DelayedSampling.hs:150-151)
  const double a2 = st->a, b2 = st->b, s = a2 + b2;
  memset(out, 0, sizeof(*out));
  out->step = st->t;
  out->log_evidence_inc = ll;
  out->ess = n_particles;
  out->mean[0] = a2 / s;
  out->var[0] = a2 * b2 / (s * s * (s + 1.0));
  out->total_weight = (uint64_t)n_particles << 15;
  st->t += 1;
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
static inline int cdiv64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

int gen_summary_parts(int64_t n) { return cdiv64(n, kThreads); }
int gen_summary_rec() { return kGenRec; }

int launch_gen_init(const GenArgs& g, cudaStream_t s) {
  if (g.kind == PZ_MODEL_WALKER) gen_init_walker<<<cdiv64(g.n, kThreads), kThreads, 0, s>>>(g);
  else gen_init_lg<<<cdiv64(g.n, kThreads), kThreads, 0, s>>>(g);
  gen_state_init<<<1, 1, 0, s>>>(g.gen, (double)g.n);
  return 2;
}

int launch_gen_anc_identity(const GenArgs& g, cudaStream_t s) {
  gen_anc_identity<<<cdiv64(g.n, kThreads), kThreads, 0, s>>>(g.ctl, g.anc, g.n);
  return 1;
}

int launch_gen_step(const GenArgs& g, cudaStream_t s) {
  const int grid = cdiv64(g.n, kThreads);
  if (g.kind == PZ_MODEL_WALKER) gen_step_walker<<<grid, kThreads, 0, s>>>(g);
  else if (g.model.nx == 1 && g.model.ny == 1) gen_step_lg<1, 1><<<grid, kThreads, 0, s>>>(g);
  else if (g.model.nx == 4 && g.model.ny == 2) gen_step_lg<4, 2><<<grid, kThreads, 0, s>>>(g);
  else if (g.model.nx == 6 && g.model.ny == 3) gen_step_lg<6, 3><<<grid, kThreads, 0, s>>>(g);
  else return -1;
  return 1;
}

int launch_gen_summary(const GenArgs& g, cudaStream_t s) {
  const int parts = gen_summary_parts(g.n);
  gen_summary_partial<<<parts, kThreads, 0, s>>>(g);
  gen_summary_final<<<1, kThreads, 0, s>>>(g, parts);
  return 2;
}

int launch_dist(bool sample, int32_t kind, const double* params, int32_t np, uint64_t seed, int64_t n, const double* x, double* out,
                cudaStream_t s) {
  DistArgs a{};
  a.kind = kind; a.np = np; a.seed = seed; a.n = n; a.x = x; a.out = out;
  for (int i = 0; i < np && i < 8; ++i) a.p[i] = params[i];
  if (sample) dist_sample_kernel<<<cdiv64(n, kThreads), kThreads, 0, s>>>(a);
  else dist_score_kernel<<<cdiv64(n, kThreads), kThreads, 0, s>>>(a);
  return 1;
}

int launch_beta_bernoulli(BetaDev* st, pz_step_summary* out, double n, cudaStream_t s) {
  beta_bernoulli_step<<<1, 32, 0, s>>>(st, out, n);
  return 1;
}

}  // namespace pz
