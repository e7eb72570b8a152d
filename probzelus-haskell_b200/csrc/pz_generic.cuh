// pz_generic.cuh -- arguments and launchers of the stored-log-weight path (pz_generic.cu).
#pragma once
#include "pz_kernels.cuh"

namespace pz {

// device-resident state of a generic-path filter that outlives a step
struct GenState {
  double prev_log;     // log of the sum of the stored weights of the pending population
  int32_t prev_emax;   // its global block exponent (the re-basing constant of the carried log-weights)
  int32_t pad;
};

struct GenArgs {
  int32_t kind;            // PZ_MODEL_RW1D / LINGAUSS (fp32 state) or PZ_MODEL_WALKER (double state, 5 scalars)
  int32_t carry_weights;   // descriptor flag: children inherit the cumulative weight (MStream `particles`)
  int32_t n_summary;       // state scalars summarised (LG: nx; walker: x, y, vx, vy, P(walking), P(running))
  double ess_threshold;    // > 0: resample only when ESS < threshold * N
  void* X[2];              // state SoA [nx][ld], double-buffered (buffer = step parity)
  float* lw[2];            // stored log-weights, double-buffered
  const uint16_t* q16;     // quantised weights of the population just weighted (pf_stats_from_lw)
  const int32_t* eb;       // its block exponents
  int32_t* anc;            // ancestors of the pending resample
  const double* obs;       // device [PZ_MAX_NY]
  PlanDev* plan;
  CtlDev* ctl;
  GenState* gen;
  pz_step_summary* summ;   // device, one record
  double* part;            // [gen_summary_parts(n)][gen_summary_rec()]
  int64_t n, ld;
  uint64_t seed;
  double ll_const;
  LgDev model;
  pz_walker_params walker;
};

struct BetaDev { double a, b, y; uint32_t t; uint32_t pad; };

int gen_summary_parts(int64_t n);
int gen_summary_rec();
int launch_gen_init(const GenArgs& g, cudaStream_t s);
int launch_gen_anc_identity(const GenArgs& g, cudaStream_t s);
int launch_gen_step(const GenArgs& g, cudaStream_t s);
int launch_gen_summary(const GenArgs& g, cudaStream_t s);
int launch_dist(bool sample, int32_t kind, const double* params, int32_t np, uint64_t seed, int64_t n, const double* x, double* out,
                cudaStream_t s);
int launch_beta_bernoulli(BetaDev* st, pz_step_summary* out, double n, cudaStream_t s);

}  // namespace pz
