// pz_dist.cuh -- the distribution primitives of haskell/src/Distributions.hs as device functions:
// sample from Philox words, score = log-density (the `Distr{sample, score}` pairs the reference's
// models are written with, Distributions.hs:20-36).  SURVEY.md section 8a row a11.
//
//   bernoulli / bernoulli01   Distributions.hs:127-139   (bernoulli_ll :148-149)
//   categorical               Distributions.hs:141-142   (mwc-random's cumulative scan: first i with cum_i >= u * total)
//   poisson                   Distributions.hs:168-172   (poisson_ll = k log lambda - lambda - logGamma(k + 1))
//   uniform                   Distributions.hs:174-180
//   exponential               Distributions.hs:209-212   (-log u / lambda)
//   gaussian / normal         Distributions.hs:157-166   (sample mu + sqrt(sigma2) z; score gaussian_ll' = TWICE the log-density)
//
// The CPU oracle (oracle/pz_oracle.cpp, "distribution primitives") writes the same formulas on the same Philox
// words; integer-valued samples agree exactly, real-valued ones to the last ulp or two of libm's log / exp.
#pragma once
#include "pz_math.cuh"

namespace pz {

constexpr uint32_t kStreamDist = 0x44495354u;    // "DIST": the pz_dist_sample tap
constexpr uint32_t kStreamWalker = 0x57414C4Bu;  // "WALK": draws of the walker's motion model
constexpr uint32_t kStreamWalkInit = 0x57494E49u;  // "WINI": walkerInit

// Sequential 32-bit words from one Philox counter family: ctr = (slot, call, step, stream), call = 0, 1, ...
struct PhiloxSeq {
  uint32_t slot, step, stream, k0, k1, call;
  int have;
  U4 buf;
  __device__ __forceinline__ PhiloxSeq(uint64_t seed, uint32_t slot_, uint32_t step_, uint32_t stream_)
      : slot(slot_), step(step_), stream(stream_), k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), call(0), have(0), buf{0, 0, 0, 0} {}
  __device__ __forceinline__ uint32_t next() {
    if (have == 0) { buf = philox4x32_10(slot, call++, step, stream, k0, k1); have = 4; }
    const uint32_t w = have == 4 ? buf.x : (have == 3 ? buf.y : (have == 2 ? buf.z : buf.w));
    --have;
    return w;
  }
  // uniform on (0, 1): (r + 1/2) 2^-32, exact in double
  __device__ __forceinline__ double uniform01() { return ((double)next() + 0.5) * 0x1p-32; }
};

__device__ __forceinline__ int sample_bernoulli(PhiloxSeq& g, double p) { return g.uniform01() < p ? 1 : 0; }
__device__ __forceinline__ double score_bernoulli(double p, double b) { return b * log(p) + (1.0 - b) * log(1.0 - p); }

__device__ __forceinline__ int sample_categorical(PhiloxSeq& g, const double* ps, int n) {
  double total = 0;
  for (int i = 0; i < n; ++i) total += ps[i];
  const double u = g.uniform01() * total;
  double c = 0;
  for (int i = 0; i < n; ++i) { c += ps[i]; if (c >= u) return i; }
  return n - 1;
}
__device__ __forceinline__ double score_categorical(const double* ps, int n, double i) {
  const int k = (int)i;
  return (k >= 0 && k < n && (double)k == i) ? log(ps[k]) : -INFINITY;
}

// inversion by sequential search on the pmf recurrence p_k = p_{k-1} lambda / k
__device__ __forceinline__ int sample_poisson(PhiloxSeq& g, double lambda) {
  const double u = g.uniform01();
  double p = exp(-lambda), c = p;
  int k = 0;
  while (u > c && k < 100000) { ++k; p *= lambda / (double)k; c += p; }
  return k;
}
__device__ __forceinline__ double score_poisson(double lambda, double k) { return k * log(lambda) - lambda - lgamma(k + 1.0); }

__device__ __forceinline__ double sample_uniform(PhiloxSeq& g, double a, double b) { return a + (b - a) * g.uniform01(); }
__device__ __forceinline__ double score_uniform(double a, double b, double x) { return (a <= x && x <= b) ? -log(b - a) : -INFINITY; }

__device__ __forceinline__ double sample_exponential(PhiloxSeq& g, double rate) { return -log(g.uniform01()) / rate; }

// mu + sqrt(sigma2) z with the table normal of the canonical spec (one Philox word per draw)
__device__ __forceinline__ double sample_normal(PhiloxSeq& g, double mu, double sigma2, const float4* __restrict__ tab_m8) {
  return mu + sqrt(sigma2) * (double)icdf_normal(g.next(), tab_m8);
}
// gaussian_ll' (Distributions.hs:160-161): - (x - mu)^2 / sigma2 - log(2 pi) - log sigma2 = 2 log N(x; mu, sigma2)
__device__ __forceinline__ double score_normal_2x(double mu, double sigma2, double x) {
  const double d = x - mu;
  return -(d * d) / sigma2 - 1.8378770664093453 - log(sigma2);
}

}  // namespace pz
