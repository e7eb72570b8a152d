// pz_oracle.cpp -- CPU oracle for the streaming particle-filter step.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under probzelus-haskell_b200/ may include, link or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker (never as the thing shipped).
//
// PARITY UNPINNED: the reference (psg-mit/probzelus-haskell) has no tests, golden vectors
// or seeded outputs, and no GHC toolchain exists in this image, so
// this restatement cannot be pinned against outputs of the Haskell binary.  It is pinned
// instead against (i) the Random123 known-answer vectors for Philox4x32-10, (ii) the
// exact Kalman filter written from the reference's own algebra (MVDistributions.hs:75-100,
// DelayedSampling.hs:88-95), and (iii) libm for the math primitives.
//
// What is restated, with the reference lines each part follows:
//   pzo_refpf_*      zparticles' + resample2 + normalize + labeledLogCategorical,
//                    Inference.hs:57-66,101-108, in double, multinomial, O(N^2): the
//                    reference-faithful filter (CPU-ref row of BASELINE.md section 4).
//   pzo_filter_*     the SAME filter step with the build's canonical fixed-point
//                    systematic resampler and fp32 state: the bit-exact twin of the CUDA
//                    path (every rounding below is part of the spec in DESIGN.md).
//   pzo_kalman_*     kalmanPredict / kalmanUpdate, MVDistributions.hs:75-100.
//   model step       noisyIntegrateGen + observe (normal x 3), Demo.hs:184-198;
//                    SingleTargetTracking.hs:33-36; MultiTargetTracking.hs:56-73.
//   scores           gaussian_ll' (doubled!) Distributions.hs:160-161; mvNormal_ll0
//                    MVDistributions.hs:45-46.
//
// Build: make -C oracle   (g++ -O2 -mfma -ffp-contract=off; FMAs appear only where the
// spec says fmaf).

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

#include "../include/pzgpu.h"
#include "../include/pz_spec_tables.h"

#if defined(_OPENMP)
#include <omp.h>
#endif

typedef unsigned __int128 u128;

namespace {

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; Random123 philox.h).  Replaces the reference's
// MWC256 SamplerIO.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

// counter word 3 selects the stream
constexpr uint32_t kStreamProp = 0x50524F50u;      // "PROP": propagation noise, ctr=(group, call, step, .)
constexpr uint32_t kStreamResample = 0x52455341u;  // "RESA": the one uniform of a systematic resample
constexpr uint32_t kStreamMulti = 0x4D554C54u;     // "MULT": per-slot uniforms of a multinomial resample
constexpr uint32_t kStreamObs = 0x4F425347u;       // "OBSG": synthetic observation generator

inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)kPhiloxM0 * c[0];
    uint64_t p1 = (uint64_t)kPhiloxM1 * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += kPhiloxW0; k1 += kPhiloxW1;
  }
}

inline void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4]) {
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  philox4x32_10(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// ---------------------------------------------------------------------------------------
// Canonical fp32 math (DESIGN.md "Math").  Every operation is an individually rounded
// IEEE fp32 op; fused multiply-adds appear exactly where fmaf is written.
// ---------------------------------------------------------------------------------------
inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

constexpr float kLog2e = 0x1.715476p+0f;
constexpr int32_t kNZero = -(1 << 30);  // exponent marker of an empty block (every weight zero)
constexpr int kGuard = 16;              // guard bits of the global fixed-point scale
constexpr int kBlk = 256;

static const float kIcdfTab[PZ_ICDF_ROWS][4] = PZ_ICDF_TABLE_INIT;

// One 32-bit word -> one N(0,1) draw by the inverse CDF on dyadic segments (spec v2, DESIGN.md
// "Canonical math"): bit 31 is the sign; v = (r << 1) | 1 is the tail probability u = v * 2^-33
// in (0, 1/2); k = clz(v) selects the dyadic segment [2^-(k+2), 2^-(k+1)), the next 3 bits the
// sub-segment, the following 23 bits the cubic's argument tf in [1, 2).
inline float icdf_normal(uint32_t r) {
  const uint32_t v = (r << 1) | 1u;
  const int k = __builtin_clz(v);
  const uint32_t vn = v << k;
  const uint32_t row = ((uint32_t)k << PZ_ICDF_SUB_BITS) + ((vn >> 28) & 7u);
  const float tf = u2f(0x3F800000u | ((vn >> 5) & 0x7FFFFFu));
  const float* c = kIcdfTab[row];
  const float z = fmaf(fmaf(fmaf(c[3], tf, c[2]), tf, c[1]), tf, c[0]);
  return u2f(f2u(z) ^ (r & 0x80000000u));
}

// Block exponent n_b from the block's largest log-weight (-inf: empty block).
inline int32_t block_exponent(float maxlw) {
  if (!(maxlw > -INFINITY)) return kNZero;
  float e = maxlw * kLog2e;
  e = fminf(fmaxf(e, -262144.0f), 262144.0f);
  return (int32_t)rintf(e);
}

// Weight of a particle relative to 2^n_b: w = 2^d, d = max(lw * log2(e) - n_b, -32), by the
// degree-5 polynomial of include/pz_spec_tables.h, and its 16-bit fixed-point value
// q = rint(w * 2^15) (<= 46990).  NaN log-weights give the smallest weight (q = 0).
inline void weight_q_from_arg(float d, float* w, uint32_t* q) {
  d = fminf(fmaxf(d, -32.0f), 1.0f);
  const float t = d + 12582912.0f;
  const float r = t - 12582912.0f;
  const float f = d - r;
  float p = PZ_EXP2_C5;
  p = fmaf(p, f, PZ_EXP2_C4);
  p = fmaf(p, f, PZ_EXP2_C3);
  p = fmaf(p, f, PZ_EXP2_C2);
  p = fmaf(p, f, PZ_EXP2_C1);
  p = fmaf(p, f, 1.0f);
  const float wv = u2f(f2u(p) + (f2u(t) << 23));
  *w = wv;
  *q = f2u(fmaf(wv, 32768.0f, 8388608.0f)) & 0x7FFFFFu;
}
inline void weight_q(float lw, float nbf, float* w, uint32_t* q) { weight_q_from_arg(fmaf(lw, kLog2e, -nbf), w, q); }
// f64 filters (descriptor flag state_f64): the log-weight is evaluated in double and rounded to fp32 once;
// block exponent, weight and quantisation are the fp32 spec applied to that value.
// (the block exponent of an f64 filter is the fp32 definition applied to the log-weights rounded to fp32)
inline float lw_key(float lw) { return lw; }
inline float lw_key(double lw) { return (float)lw; }
inline void weight_q(double lw, double nbf, float* w, uint32_t* q) { weight_q((float)lw, (float)nbf, w, q); }

// stored log-weights may hold anything: +inf and NaN count as -inf (zero weight)
inline float sanitize_lw(float lw) { return (lw == lw && lw < INFINITY) ? lw : -INFINITY; }
inline double sanitize_lw(double lw) { return (lw == lw && lw < INFINITY) ? lw : -INFINITY; }
inline float t_fma(float a, float b, float c) { return fmaf(a, b, c); }
inline double t_fma(double a, double b, double c) { return fma(a, b, c); }

// CDF order inside a canonical block: position p <-> memory offset (two interleaved halves of
// 128, four particles per lane): p = lane * 8 + half * 4 + k  <->  off = half * 128 + lane * 4 + k.
inline int cdf_offset(int p) { return ((p >> 2) & 1) * 128 + (p >> 3) * 4 + (p & 3); }

// ---------------------------------------------------------------------------------------
// Linear-Gaussian model family: derived fp32 parameters.
// ---------------------------------------------------------------------------------------
template <class T>
struct LgParamsT {
  int nx = 0, ny = 0;
  T F[PZ_MAX_NX][PZ_MAX_NX];
  T Lq[PZ_MAX_NX][PZ_MAX_NX];  // lower Cholesky factor of Q (sampleNormalDist uses hmatrix's
                                   // UPPER factor, MVDistributions.hs:40; identical for the diagonal
                                   // Q of every in-scope model)
  T H[PZ_MAX_NY][PZ_MAX_NX];
  T Rinv[PZ_MAX_NY][PZ_MAX_NY];
  T x0[PZ_MAX_NX];
  T wscale;         // -0.5 (mvNormal_ll0) or -1.0 (doubled scalar score, Distributions.hs:160-161)
  double ll_const;  // additive constant of the score dropped from the per-particle weight
};
typedef LgParamsT<float> LgParams;

template <class T>
bool derive_lg(const pz_model_desc* d, LgParamsT<T>* p) {
  int nx = d->nx, ny = d->ny;
  if (nx < 1 || nx > PZ_MAX_NX || ny < 1 || ny > PZ_MAX_NY) return false;
  *p = LgParamsT<T>();
  p->nx = nx; p->ny = ny;
  double L[PZ_MAX_NX][PZ_MAX_NX] = {};
  for (int j = 0; j < nx; ++j) {
    double s = d->Q[j * nx + j];
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    double ljj = s > 0 ? sqrt(s) : 0.0;
    L[j][j] = ljj;
    for (int i = j + 1; i < nx; ++i) {
      double t = d->Q[i * nx + j];
      for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
      L[i][j] = ljj > 0 ? t / ljj : 0.0;
    }
  }
  // R^-1 and log det R by Gauss-Jordan without pivoting (R is SPD)
  double A[PZ_MAX_NY][2 * PZ_MAX_NY] = {};
  for (int i = 0; i < ny; ++i) {
    for (int j = 0; j < ny; ++j) A[i][j] = d->R[i * ny + j];
    A[i][ny + i] = 1.0;
  }
  double logdet = 0;
  for (int c = 0; c < ny; ++c) {
    double piv = A[c][c];
    if (!(piv > 0)) return false;
    logdet += log(piv);
    for (int j = 0; j < 2 * ny; ++j) A[c][j] /= piv;
    for (int r = 0; r < ny; ++r) {
      if (r == c) continue;
      double f = A[r][c];
      for (int j = 0; j < 2 * ny; ++j) A[r][j] -= f * A[c][j];
    }
  }
  for (int i = 0; i < nx; ++i) {
    for (int j = 0; j < nx; ++j) { p->F[i][j] = (T)d->F[i * nx + j]; p->Lq[i][j] = (T)L[i][j]; }
    p->x0[i] = (T)d->x0[i];
  }
  for (int i = 0; i < ny; ++i) {
    for (int j = 0; j < nx; ++j) p->H[i][j] = (T)d->H[i * nx + j];
    for (int j = 0; j < ny; ++j) p->Rinv[i][j] = (T)A[i][ny + j];
  }
  bool dbl = (ny == 1 && d->scalar_ll_2x);
  p->wscale = dbl ? (T)-1.0 : (T)-0.5;
  p->ll_const = (dbl ? -1.0 : -0.5) * (logdet + ny * log(2 * M_PI));
  return true;
}

template <class T>
inline void lg_propagate(const LgParamsT<T>& p, const T* x, const T* z, T* xn) {
  for (int r = 0; r < p.nx; ++r) {
    T acc = p.F[r][0] * x[0];
    for (int c = 1; c < p.nx; ++c) acc = t_fma(p.F[r][c], x[c], acc);
    for (int c = 0; c <= r; ++c) acc = t_fma(p.Lq[r][c], z[c], acc);
    xn[r] = acc;
  }
}

template <class T>
inline T lg_logweight(const LgParamsT<T>& p, const T* xn, const T* y) {
  T d[PZ_MAX_NY];
  for (int k = 0; k < p.ny; ++k) {
    T acc = p.H[k][0] * xn[0];
    for (int c = 1; c < p.nx; ++c) acc = t_fma(p.H[k][c], xn[c], acc);
    d[k] = y[k] - acc;
  }
  T quad = 0;
  for (int k = 0; k < p.ny; ++k) {
    T t = p.Rinv[k][0] * d[0];
    for (int l = 1; l < p.ny; ++l) t = t_fma(p.Rinv[k][l], d[l], t);
    quad = (k == 0) ? d[0] * t : t_fma(d[k], t, quad);
  }
  return p.wscale * quad;
}

// normals for output slot j at step t (counter keyed on the GLOBAL slot, never on a thread id):
// one Philox word per normal.  nx = 1: four consecutive slots share one call; nx = 6: two
// consecutive slots share three calls (even slot words 0..5, odd slot words 6..11); otherwise
// ceil(nx / 4) calls per slot.
inline void lg_noise(int nx, uint64_t j, uint32_t t, uint64_t seed, float* z) {
  uint32_t r[4];
  if (nx == 1) {
    philox((uint32_t)(j >> 2), 0, t, kStreamProp, seed, r);
    z[0] = icdf_normal(r[j & 3]);
    return;
  }
  if (nx == 6) {
    uint32_t w[12];
    for (int k = 0; k < 3; ++k) {
      philox((uint32_t)(j >> 1), (uint32_t)k, t, kStreamProp, seed, r);
      for (int i = 0; i < 4; ++i) w[4 * k + i] = r[i];
    }
    const int o = (j & 1) ? 6 : 0;
    for (int k = 0; k < 6; ++k) z[k] = icdf_normal(w[o + k]);
    return;
  }
  for (int k = 0; 4 * k < nx; ++k) {
    philox((uint32_t)j, (uint32_t)k, t, kStreamProp, seed, r);
    for (int i = 0; i < 4 && 4 * k + i < nx; ++i) z[4 * k + i] = icdf_normal(r[i]);
  }
}

// ---------------------------------------------------------------------------------------
// Canonical fixed-point systematic resampler "PZ-SYS-256/q16" (DESIGN.md "Resampler").
// ---------------------------------------------------------------------------------------
struct ResamplePlan {
  int64_t n = 0, nblk = 0;
  std::vector<uint16_t> q;     // per particle 16-bit weight at its block's scale 2^(n_b - 15)
  std::vector<float> w;        // per particle weight relative to 2^n_b
  std::vector<int32_t> eb;     // per block exponent n_b (kNZero: empty)
  std::vector<uint32_t> sb;    // per block sum of q
  std::vector<uint64_t> pb;    // per block exclusive prefix at the global scale 2^(E - 31) (nblk+1 entries)
  int32_t emax = kNZero;
  uint64_t total = 0;
  uint64_t rho = 0;   // floor(N_out * 2^(32+L) / T) + 1
  int L = 0;          // bit length of T
  uint32_t Mq = 0;    // 30-bit slope multiplier, floor(rho / 2^lam) + 1
  int h0 = 0;         // L - kGuard - lam: a block at exponent distance ksh shifts by h0 + ksh
};

inline int bitlen64(uint64_t v) { return v ? 64 - __builtin_clzll(v) : 0; }

// block sum at the global scale: s * 2^(kGuard - ksh), truncated
inline uint64_t shift_guard(uint64_t s, int64_t ksh) {
  if (ksh <= kGuard) return s << (kGuard - ksh);
  const int64_t sh = ksh - kGuard;
  return sh >= 64 ? 0 : (s >> sh);
}

// returns false when every weight is zero (degenerate)
template <class T>
bool build_plan(const T* lw, int64_t n, int64_t n_out, ResamplePlan* pl) {
  pl->n = n;
  pl->nblk = (n + kBlk - 1) / kBlk;
  pl->q.assign(n, 0); pl->w.assign(n, 0.0f);
  pl->eb.assign(pl->nblk, kNZero); pl->sb.assign(pl->nblk, 0); pl->pb.assign(pl->nblk + 1, 0);
  pl->emax = kNZero;
  for (int64_t b = 0; b < pl->nblk; ++b) {
    const int64_t hi = std::min(n, (b + 1) * kBlk);
    float mx = -INFINITY;
    for (int64_t i = b * kBlk; i < hi; ++i) mx = fmaxf(mx, lw_key(sanitize_lw(lw[i])));
    const int32_t nb = block_exponent(mx);
    pl->eb[b] = nb;
    if (nb == kNZero) continue;
    uint32_t s = 0;
    for (int64_t i = b * kBlk; i < hi; ++i) {
      uint32_t qi; float wi;
      weight_q(sanitize_lw(lw[i]), (T)nb, &wi, &qi);
      pl->q[i] = (uint16_t)qi; pl->w[i] = wi;
      s += qi;
    }
    pl->sb[b] = s;
    pl->emax = std::max(pl->emax, nb);
  }
  if (pl->emax == kNZero) return false;
  uint64_t run = 0;
  for (int64_t b = 0; b < pl->nblk; ++b) {
    pl->pb[b] = run;
    if (pl->eb[b] != kNZero) run += shift_guard(pl->sb[b], (int64_t)pl->emax - pl->eb[b]);
  }
  pl->pb[pl->nblk] = run;
  pl->total = run;
  if (run == 0) return false;
  const int L = bitlen64(run);
  pl->L = L;
  pl->rho = (uint64_t)((((u128)(uint64_t)n_out) << (32 + L)) / run) + 1;
  const int lam = bitlen64(pl->rho) - 30;
  pl->Mq = (uint32_t)(pl->rho >> lam) + 1u;
  pl->h0 = L - kGuard - lam;
  return true;
}

// position of CDF value C in units of 2^-32 output slots: floor(C * rho / 2^L) <= N * 2^32
inline uint64_t pos_of(const ResamplePlan& pl, uint64_t C) {
  return (uint64_t)(((u128)(C << (64 - pl.L)) * pl.rho) >> 64);
}

// threshold-shifted position of a block boundary: its high word is the number of output slots
// whose threshold j * 2^32 + U lies below pos(C)
inline uint64_t block_pos(const ResamplePlan& pl, uint64_t C, uint32_t U) { return pos_of(pl, C) + (0xFFFFFFFFull - U); }

inline int64_t slot_end(const ResamplePlan& pl, uint64_t C, uint32_t U, int64_t n_out) {
  return (int64_t)std::min<uint64_t>(block_pos(pl, C, U) >> 32, (uint64_t)n_out);
}

// Slots between the end of a particle's children and the end of its block's children, negated:
// floor((A * 2^h - r * M) / 2^(h+32)), r = block mass AFTER the particle, A = low word of the
// block-end position, M / 2^h = slots (in 2^-32 units) per unit of block-scale weight.
inline int64_t end_delta(uint64_t r, uint32_t M, uint64_t A, int64_t h) {
  const u128 rm = (u128)r * M;
  if (h >= 0) {
    u128 t;  // ceil(r * M / 2^h)
    if (h >= 100) t = rm ? 1 : 0;
    else t = (rm + (((u128)1) << h) - 1) >> h;
    const __int128 num = (__int128)A - (__int128)t;
    return (int64_t)(num >> 32);  // arithmetic shift: floor
  }
  const __int128 num = (__int128)A - (__int128)(rm << (-h));
  return (int64_t)(num >> 32);
}

// ancestors of output slots [0, n_out); parent indices refer to [0, n)
void systematic_ancestors(const ResamplePlan& pl, uint32_t U, int64_t n_out, int32_t* anc) {
  int64_t prev = 0;
  for (int64_t b = 0; b < pl.nblk; ++b) {
    const uint64_t X0 = block_pos(pl, pl.pb[b], U), X1 = block_pos(pl, pl.pb[b + 1], U);
    const int64_t o0 = (int64_t)(X0 >> 32), o1 = (int64_t)(X1 >> 32);
    if (o1 == o0) continue;  // no children (includes every empty block)
    const int64_t h = (int64_t)pl.h0 + ((int64_t)pl.emax - pl.eb[b]);
    const uint64_t A = X1 & 0xFFFFFFFFull;
    uint64_t c = 0;
    for (int p = 0; p < kBlk; ++p) {
      const int64_t i = b * kBlk + cdf_offset(p);
      if (i >= pl.n) continue;  // padding carries no weight
      c += pl.q[i];
      const int64_t e = std::max(o1 + end_delta(pl.sb[b] - c, pl.Mq, A, h), o0);
      for (int64_t j = prev; j < e; ++j) anc[j] = (int32_t)i;
      if (e > prev) prev = e;
    }
  }
  // by construction prev == n_out (coverage lemma)
  for (int64_t j = prev; j < n_out; ++j) anc[j] = -1;
}

// multinomial on the same fixed-point CDF: N iid categorical draws (resample2,
// Inference.hs:57-61), slot j uses the 64-bit uniform of Philox ctr=(j,0,t,"MULT").
void multinomial_ancestors(const ResamplePlan& pl, uint64_t seed, uint32_t t, int64_t n_out, int32_t* anc) {
  // full CDF in CDF order (oracle can afford it): entry (b * 256 + p)
  std::vector<uint64_t> cdf((size_t)pl.nblk * kBlk);
  for (int64_t b = 0; b < pl.nblk; ++b) {
    const int64_t k = (int64_t)pl.emax - (int64_t)pl.eb[b];
    uint64_t c = 0;
    for (int p = 0; p < kBlk; ++p) {
      const int64_t i = b * kBlk + cdf_offset(p);
      if (i < pl.n && pl.eb[b] != kNZero) c += pl.q[i];
      cdf[(size_t)b * kBlk + p] = pl.pb[b] + (pl.eb[b] == kNZero ? 0 : shift_guard(c, k));
    }
  }
  for (int64_t j = 0; j < n_out; ++j) {
    uint32_t r[4];
    philox((uint32_t)j, 0, t, kStreamMulti, seed, r);
    uint64_t u = ((uint64_t)r[0] << 32) | r[1];
    uint64_t tau = (uint64_t)(((u128)u * pl.total) >> 64);  // [0, T)
    int64_t pos = std::upper_bound(cdf.begin(), cdf.end(), tau) - cdf.begin();  // min{pos: C > tau}
    anc[j] = (int32_t)((pos / kBlk) * kBlk + cdf_offset((int)(pos % kBlk)));
  }
}

inline uint32_t resample_uniform(uint64_t seed, uint32_t t) {
  uint32_t r[4];
  philox(0, 0, t, kStreamResample, seed, r);
  return r[0];
}

// ---------------------------------------------------------------------------------------
// The canonical filter (bit-exact twin of the CUDA path).
// ---------------------------------------------------------------------------------------
struct OFilterBase {
  virtual ~OFilterBase() {}
  virtual int step(const double* obs, pz_step_summary* out) = 0;
  virtual void state32(float* x, float* lw, int32_t* anc) = 0;
  virtual void state64(double* x, double* lw, int32_t* anc) = 0;
  virtual int block_stats(int32_t* eb, uint64_t* sb, int32_t* emax, uint64_t* total) = 0;
  virtual void derived(float* F, float* Lq, float* H, float* Rinv, float* wscale, double* ll_const) = 0;
};

template <class T>
struct OFilterT : OFilterBase {
  pz_model_desc desc;
  LgParamsT<T> p;
  int64_t n = 0;
  uint64_t seed = 0;
  uint32_t t = 0;
  std::vector<T> x;    // SoA [nx][n], population the pending weights refer to
  std::vector<T> lw;   // its log-weights
  std::vector<int32_t> anc;  // ancestors used by the last step
  double center[PZ_MAX_NX];  // centring constants of the moment sums (previous posterior mean)
  // stored-log-weight variants (descriptor flags carry_weights / ess_threshold; pz_generic.cu)
  bool resample_next = true;   // the adaptive rule's decision for the next resample
  double prev_log = 0;         // log of the sum of the stored weights of the pending population
  int32_t prev_emax = 0;

  bool generic() const { return desc.carry_weights || desc.ess_threshold > 0; }

  // summary of the stored-log-weight path: the QUANTISED weights, centred on the previous mean
  void summarise_generic(const ResamplePlan& pl, bool carried, pz_step_summary* s) {
    double sw = 0, sw2 = 0, sx[PZ_MAX_NX] = {}, sxx[PZ_MAX_NX] = {};
    for (int64_t i = 0; i < n; ++i) {
      const int32_t eb = pl.eb[i / kBlk];
      if (eb == kNZero) continue;
      const double w = ldexp((double)pl.q[i], eb - pl.emax - 15);
      sw += w; sw2 += w * w;
      for (int d = 0; d < p.nx; ++d) {
        const double dx = (double)x[(size_t)d * n + i] - center[d];
        sx[d] += w * dx; sxx[d] += w * dx * dx;
      }
    }
    s->ess = sw * sw / sw2;
    for (int d = 0; d < p.nx; ++d) {
      const double m = sx[d] / sw;
      s->mean[d] = center[d] + m;
      s->var[d] = sxx[d] / sw - m * m;
    }
    const double cur_log = log(sw) + pl.emax * 0.6931471805599453;
    const double rebase = (double)((float)prev_emax * 0x1.62e43p-1f);
    s->log_evidence_inc = (carried ? cur_log - (prev_log - rebase) : cur_log - log((double)n)) + p.ll_const;
    s->e_max = pl.emax;
    s->total_weight = pl.total;
    s->n_migrated = 0;
    prev_log = cur_log; prev_emax = pl.emax;
  }

  void summarise(const ResamplePlan& pl, pz_step_summary* s) const {
    // the same fp32 weights the resampler quantises, at scale 2^emax, in double
    double sw = 0, sw2 = 0, sx[PZ_MAX_NX] = {}, sxx[PZ_MAX_NX] = {};
    for (int64_t i = 0; i < n; ++i) {
      const int32_t eb = pl.eb[i / kBlk];
      if (eb == kNZero) continue;
      double w = ldexp((double)pl.w[i], eb - pl.emax);
      sw += w; sw2 += w * w;
      for (int d = 0; d < p.nx; ++d) {
        double dx = (double)(T)(x[(size_t)d * n + i] - (T)center[d]);
        sx[d] += w * dx; sxx[d] += w * dx * dx;
      }
    }
    s->ess = sw * sw / sw2;
    for (int d = 0; d < p.nx; ++d) {
      double m = sx[d] / sw;
      s->mean[d] = (double)(T)center[d] + m;
      s->var[d] = sxx[d] / sw - m * m;
    }
    s->log_evidence_inc = log(sw) + pl.emax * M_LN2 - log((double)n) + p.ll_const;
    s->e_max = pl.emax;
    s->total_weight = pl.total;
    s->n_migrated = 0;
  }

  // One step of zparticles' (Inference.hs:101-104) in the cross-step order of the CUDA path:
  // resample the pending population by its weights, propagate every child, weight it by the
  // new observation.  (The reference weights first and resamples at the end of the same step;
  // the populations and weights visited are identical, only the step boundary moves.)
  int step(const double* obs, pz_step_summary* out) override {
    ResamplePlan pl;
    if (!build_plan(lw.data(), n, n, &pl)) return PZ_EDEGENERATE;
    const bool resampled = !(desc.ess_threshold > 0) || resample_next;
    if (!resampled) {
      for (int64_t j = 0; j < n; ++j) anc[j] = (int32_t)j;   // the adaptive rule keeps every particle in its slot
    } else if (desc.resampler == PZ_RESAMPLE_MULTINOMIAL_REF) {
      multinomial_ancestors(pl, seed, t, n, anc.data());
    } else {
      uint32_t U = resample_uniform(seed, t);
      systematic_ancestors(pl, U, n, anc.data());
    }
    // MStream `particles` (Inference.hs:73-86): a child inherits its parent's cumulative weight; without a
    // resample the weights accumulate as well.  The carried fp32 log-weight is re-based by emax * ln 2.
    const bool carried = generic() && (desc.carry_weights || !resampled);
    const T rebase = (T)((float)pl.emax * 0x1.62e43p-1f);
    std::vector<T> lw_old;
    if (carried) lw_old = lw;
    T y[PZ_MAX_NY];
    for (int k = 0; k < p.ny; ++k) y[k] = (T)obs[k];
    std::vector<T> xn((size_t)p.nx * n);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) {
      T xp[PZ_MAX_NX], z[PZ_MAX_NX + 2], xo[PZ_MAX_NX];
      float zf[PZ_MAX_NX + 2];
      int64_t a = anc[j];
      for (int d = 0; d < p.nx; ++d) xp[d] = x[(size_t)d * n + a];
      lg_noise(p.nx, (uint64_t)j, t, seed, zf);
      for (int d = 0; d < p.nx; ++d) z[d] = (T)zf[d];   // a double filter uses the same fp32 table normals
      lg_propagate(p, xp, z, xo);
      for (int d = 0; d < p.nx; ++d) xn[(size_t)d * n + j] = xo[d];
      const T inc = lg_logweight(p, xo, y);
      lw[j] = carried ? (T)((lw_old[a] - rebase) + inc) : inc;
    }
    x.swap(xn);
    ResamplePlan pl2;
    bool ok = build_plan(lw.data(), n, n, &pl2);
    pz_step_summary tmp;
    if (!out) out = &tmp;
    memset(out, 0, sizeof(*out));
    out->step = t;
    if (ok) {
      if (generic()) {
        summarise_generic(pl2, carried, out);
        out->reserved = resampled ? 1 : 0;
        resample_next = !(desc.ess_threshold > 0) || out->ess < desc.ess_threshold * (double)n;
      } else {
        summarise(pl2, out);
      }
      for (int d = 0; d < p.nx; ++d) center[d] = out->mean[d];
    }
    t += 1;
    return ok ? 0 : PZ_EDEGENERATE;
  }
  void state32(float* xo, float* lwo, int32_t* ao) override {
    if (xo) for (size_t i = 0; i < x.size(); ++i) xo[i] = (float)x[i];
    if (lwo) for (size_t i = 0; i < lw.size(); ++i) lwo[i] = (float)lw[i];
    if (ao) memcpy(ao, anc.data(), anc.size() * sizeof(int32_t));
  }
  void state64(double* xo, double* lwo, int32_t* ao) override {
    if (xo) for (size_t i = 0; i < x.size(); ++i) xo[i] = (double)x[i];
    if (lwo) for (size_t i = 0; i < lw.size(); ++i) lwo[i] = (double)lw[i];
    if (ao) memcpy(ao, anc.data(), anc.size() * sizeof(int32_t));
  }
  int block_stats(int32_t* eb, uint64_t* sb, int32_t* emax, uint64_t* total) override {
    ResamplePlan pl;
    bool ok = build_plan(lw.data(), n, n, &pl);
    if (eb) memcpy(eb, pl.eb.data(), pl.nblk * sizeof(int32_t));
    if (sb) for (int64_t b = 0; b < pl.nblk; ++b) sb[b] = pl.sb[b];
    if (emax) *emax = pl.emax;
    if (total) *total = pl.total;
    return ok ? 0 : PZ_EDEGENERATE;
  }
  void derived(float* F, float* Lq, float* H, float* Rinv, float* wscale, double* ll_const) override {
    for (int i = 0; i < PZ_MAX_NX; ++i) for (int j = 0; j < PZ_MAX_NX; ++j) { F[i * PZ_MAX_NX + j] = (float)p.F[i][j]; Lq[i * PZ_MAX_NX + j] = (float)p.Lq[i][j]; }
    for (int i = 0; i < PZ_MAX_NY; ++i) for (int j = 0; j < PZ_MAX_NX; ++j) H[i * PZ_MAX_NX + j] = (float)p.H[i][j];
    for (int i = 0; i < PZ_MAX_NY; ++i) for (int j = 0; j < PZ_MAX_NY; ++j) Rinv[i * PZ_MAX_NY + j] = (float)p.Rinv[i][j];
    *wscale = (float)p.wscale; *ll_const = p.ll_const;
  }
};

template <class T>
OFilterBase* make_filter(const pz_model_desc* d, int64_t n, uint64_t seed) {
  OFilterT<T>* f = new OFilterT<T>();
  f->desc = *d;
  if (!derive_lg(d, &f->p) || n < 1) { delete f; return nullptr; }
  f->n = n; f->seed = seed; f->t = 0;
  f->x.assign((size_t)f->p.nx * n, (T)0);
  for (int k = 0; k < f->p.nx; ++k) {
    for (int64_t i = 0; i < n; ++i) f->x[(size_t)k * n + i] = f->p.x0[k];
    f->center[k] = d->x0[k];
  }
  f->lw.assign(n, (T)0);  // equal weights before the first observation
  f->anc.assign(n, 0);
  f->prev_log = log((double)n);
  return f;
}

}  // namespace

// =========================================================================================
// extern "C" surface for ctypes
// =========================================================================================
extern "C" {

void pzo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox4x32_10(c, key[0], key[1]);
  memcpy(out, c, sizeof(c));
}

// per-block weights of a log-weight vector: q (16-bit fixed point at the block scale), w (fp32,
// relative to 2^n_b), eb[nblk] (block exponents)
void pzo_block_weights(const float* lw, int64_t n, uint16_t* q, float* w, int32_t* eb) {
  ResamplePlan pl;
  build_plan(lw, n, n, &pl);
  if (q) memcpy(q, pl.q.data(), n * sizeof(uint16_t));
  if (w) memcpy(w, pl.w.data(), n * sizeof(float));
  if (eb) memcpy(eb, pl.eb.data(), pl.nblk * sizeof(int32_t));
}

void pzo_icdf_normals(const uint32_t* r, int64_t n, float* z) {
  for (int64_t i = 0; i < n; ++i) z[i] = icdf_normal(r[i]);
}

uint32_t pzo_resample_uniform(uint64_t seed, uint32_t t) { return resample_uniform(seed, t); }

// canonical systematic ancestors from fp32 log-weights; optional diagnostics
int pzo_resample_systematic(const float* lw, int64_t n, uint32_t U, int32_t* anc, int32_t* eb_out,
                            uint64_t* sb_out, uint64_t* pb_out, int32_t* emax, uint64_t* total) {
  ResamplePlan pl;
  if (!build_plan(lw, n, n, &pl)) return PZ_EDEGENERATE;
  systematic_ancestors(pl, U, n, anc);
  if (eb_out) memcpy(eb_out, pl.eb.data(), pl.nblk * sizeof(int32_t));
  if (sb_out) for (int64_t b = 0; b < pl.nblk; ++b) sb_out[b] = pl.sb[b];
  if (pb_out) memcpy(pb_out, pl.pb.data(), (pl.nblk + 1) * sizeof(uint64_t));
  if (emax) *emax = pl.emax;
  if (total) *total = pl.total;
  return 0;
}

int pzo_resample_multinomial(const float* lw, int64_t n, uint64_t seed, uint32_t t, int32_t* anc) {
  ResamplePlan pl;
  if (!build_plan(lw, n, n, &pl)) return PZ_EDEGENERATE;
  multinomial_ancestors(pl, seed, t, n, anc);
  return 0;
}

// Reference semantics of systematic resampling in double (monad-bayes
// Control.Monad.Bayes.Population.systematic lineage: thresholds (u+j)/N, sequential CDF
// walk).  Used to check the fixed-point resampler statistically, not bitwise.
void pzo_resample_systematic_f64(const double* logw, int64_t n, double u, int32_t* anc) {
  double mx = -INFINITY;
  for (int64_t i = 0; i < n; ++i) if (logw[i] > mx) mx = logw[i];
  double tot = 0;
  for (int64_t i = 0; i < n; ++i) tot += exp(logw[i] - mx);
  double c = 0; int64_t j = 0;
  for (int64_t i = 0; i < n; ++i) {
    c += exp(logw[i] - mx) / tot;
    while (j < n && (j + u) / n < c) anc[j++] = (int32_t)i;
  }
  while (j < n) anc[j++] = (int32_t)(n - 1);
}

void* pzo_filter_create(const pz_model_desc* d, int64_t n, uint64_t seed) {
  return d->state_f64 ? make_filter<double>(d, n, seed) : make_filter<float>(d, n, seed);
}

void pzo_filter_destroy(void* h) { delete (OFilterBase*)h; }

void pzo_filter_derived(void* h, float* F, float* Lq, float* H, float* Rinv, float* wscale, double* ll_const) {
  ((OFilterBase*)h)->derived(F, Lq, H, Rinv, wscale, ll_const);
}

int pzo_filter_step(void* h, const double* obs, pz_step_summary* out) { return ((OFilterBase*)h)->step(obs, out); }

void pzo_filter_state(void* h, float* x, float* lw, int32_t* anc) { ((OFilterBase*)h)->state32(x, lw, anc); }
void pzo_filter_state_f64(void* h, double* x, double* lw, int32_t* anc) { ((OFilterBase*)h)->state64(x, lw, anc); }

// block statistics of the pending population (parity tap for the fused kernel's metadata)
int pzo_filter_block_stats(void* h, int32_t* eb, uint64_t* sb, int32_t* emax, uint64_t* total) {
  return ((OFilterBase*)h)->block_stats(eb, sb, emax, total);
}

}  // extern "C"

// =========================================================================================
// Distribution primitives (Distributions.hs:127-212) on Philox words: the CPU twin of
// probzelus-haskell_b200/csrc/pz_dist.cuh.
// =========================================================================================
namespace {

constexpr uint32_t kStreamDist = 0x44495354u, kStreamWalker = 0x57414C4Bu, kStreamWalkInit = 0x57494E49u;

struct PhiloxSeq {  // sequential words of one counter family: ctr = (slot, call, step, stream)
  uint64_t seed; uint32_t slot, step, stream, call = 0; int have = 0; uint32_t r[4];
  PhiloxSeq(uint64_t s, uint32_t sl, uint32_t st, uint32_t strm) : seed(s), slot(sl), step(st), stream(strm) {}
  uint32_t next() {
    if (!have) { philox(slot, call++, step, stream, seed, r); have = 4; }
    return r[4 - have--];
  }
  double uniform01() { return ((double)next() + 0.5) * 0x1p-32; }
};

int sample_bernoulli(PhiloxSeq& g, double p) { return g.uniform01() < p ? 1 : 0; }            // B.bernoulli
double score_bernoulli(double p, double b) { return b * log(p) + (1.0 - b) * log(1.0 - p); }   // bernoulli_ll :148-149
int sample_categorical(PhiloxSeq& g, const double* ps, int n) {   // mwc-random categorical: first i with cum_i >= u * total
  double total = 0;
  for (int i = 0; i < n; ++i) total += ps[i];
  const double u = g.uniform01() * total;
  double c = 0;
  for (int i = 0; i < n; ++i) { c += ps[i]; if (c >= u) return i; }
  return n - 1;
}
double score_categorical(const double* ps, int n, double i) {     // log (ps !! i) :141-142
  const int k = (int)i;
  return (k >= 0 && k < n && (double)k == i) ? log(ps[k]) : -INFINITY;
}
int sample_poisson(PhiloxSeq& g, double lambda) {
  const double u = g.uniform01();
  double p = exp(-lambda), c = p;
  int k = 0;
  while (u > c && k < 100000) { ++k; p *= lambda / (double)k; c += p; }
  return k;
}
double score_poisson(double lambda, double k) { return k * log(lambda) - lambda - lgamma(k + 1.0); }   // poisson_ll :168-169
double sample_uniform(PhiloxSeq& g, double a, double b) { return a + (b - a) * g.uniform01(); }
double score_uniform(double a, double b, double x) { return (a <= x && x <= b) ? -log(b - a) : -INFINITY; }   // :174-180
double sample_exponential(PhiloxSeq& g, double rate) { return -log(g.uniform01()) / rate; }   // :209-212
double sample_normal(PhiloxSeq& g, double mu, double sigma2) { return mu + sqrt(sigma2) * (double)icdf_normal(g.next()); }   // :163-164
double score_normal_2x(double mu, double sigma2, double x) {    // gaussian_ll' :160-161
  const double d = x - mu;
  return -(d * d) / sigma2 - 1.8378770664093453 - log(sigma2);
}

// ---- Walker (Examples/Walker.hs) ----
struct WalkerState { double x, y, vx, vy; int mt; };

void walker_init_velocity(const pz_walker_params& w, PhiloxSeq& rng, int mt, double& vx, double& vy) {   // :51-64
  if (w.speed_hi[mt] <= 0.0) { vx = 0.0; vy = 0.0; return; }
  const double speed = sample_uniform(rng, w.speed_lo[mt], w.speed_hi[mt]);
  const double angle = sample_uniform(rng, 0.0, 6.283185307179586);
  vx = speed * cos(angle);
  vy = speed * sin(angle);
}
void walker_coast(const pz_walker_params& w, PhiloxSeq& rng, double dt, WalkerState& s) {   // :31-41 (variance = sqrt dt * posNoise)
  const double var = sqrt(dt) * w.pos_noise;
  s.x = sample_normal(rng, s.x + s.vx * dt, var);
  s.y = sample_normal(rng, s.y + s.vy * dt, var);
}
void walker_motion(const pz_walker_params& w, PhiloxSeq& rng, WalkerState& s) {   // :67-81
  double dt = w.dt;
  for (int it = 0; it < 100000; ++it) {
    double rates[3], sum = 0.0;
    for (int k = 0; k < 3; ++k) {
      rates[k] = w.half_life[s.mt][k] > 0.0 ? 0.6931471805599453 / w.half_life[s.mt][k] : 0.0;
      sum += rates[k];
    }
    const double lambda = w.exp_mean_as_written ? 1.0 / sum : sum;
    const double t_tr = sum > 0.0 ? sample_exponential(rng, lambda) : INFINITY;
    if (t_tr > dt) { walker_coast(w, rng, dt, s); return; }
    walker_coast(w, rng, t_tr, s);
    s.mt = sample_categorical(rng, rates, 3);
    walker_init_velocity(w, rng, s.mt, s.vx, s.vy);
    dt -= t_tr;
  }
}

struct OWalker {
  pz_model_desc desc;
  int64_t n; uint64_t seed; uint32_t t = 0;
  std::vector<WalkerState> s;
  std::vector<float> lw;
  std::vector<int32_t> anc;
  bool resample_next = true;
};

}  // namespace

extern "C" {

int pzo_dist_sample(int32_t kind, const double* p, int32_t np, uint64_t seed, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    PhiloxSeq g(seed, (uint32_t)i, 0u, kStreamDist);
    switch (kind) {
      case PZ_DIST_BERNOULLI: out[i] = sample_bernoulli(g, p[0]); break;
      case PZ_DIST_CATEGORICAL: out[i] = sample_categorical(g, p, np); break;
      case PZ_DIST_POISSON: out[i] = sample_poisson(g, p[0]); break;
      case PZ_DIST_UNIFORM: out[i] = sample_uniform(g, p[0], p[1]); break;
      case PZ_DIST_EXPONENTIAL: out[i] = sample_exponential(g, p[0]); break;
      case PZ_DIST_NORMAL: out[i] = sample_normal(g, p[0], p[1]); break;
      default: return PZ_EINVAL;
    }
  }
  return 0;
}

int pzo_dist_score(int32_t kind, const double* p, int32_t np, const double* x, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    switch (kind) {
      case PZ_DIST_BERNOULLI: out[i] = score_bernoulli(p[0], x[i]); break;
      case PZ_DIST_CATEGORICAL: out[i] = score_categorical(p, np, x[i]); break;
      case PZ_DIST_POISSON: out[i] = score_poisson(p[0], x[i]); break;
      case PZ_DIST_UNIFORM: out[i] = score_uniform(p[0], p[1], x[i]); break;
      case PZ_DIST_EXPONENTIAL: out[i] = x[i] >= 0 ? log(p[0]) - p[0] * x[i] : -INFINITY; break;
      case PZ_DIST_NORMAL: out[i] = score_normal_2x(p[0], p[1], x[i]); break;
      default: return PZ_EINVAL;
    }
  }
  return 0;
}

// importanceResampling (Inference.hs:68-71) / resample (:51-52): k iid categorical draws on the canonical CDF
int pzo_importance_resampling(const float* lw, int64_t n, uint64_t seed, uint32_t t, int64_t k, int32_t* idx) {
  ResamplePlan pl;
  if (!build_plan(lw, n, n, &pl)) return PZ_EDEGENERATE;
  std::vector<int32_t> anc(n);
  multinomial_ancestors(pl, seed, t, n, anc.data());   // slot j uses its own counter: the first k slots are the k draws
  for (int64_t j = 0; j < k; ++j) idx[j] = anc[j];
  return 0;
}

void* pzo_walker_create(const pz_model_desc* d, int64_t n, uint64_t seed) {
  OWalker* f = new OWalker();
  f->desc = *d; f->n = n; f->seed = seed;
  f->s.resize(n); f->lw.assign(n, 0.0f); f->anc.assign(n, 0);
  for (int64_t j = 0; j < n; ++j) {   // walkerInit (Walker.hs:105-111)
    PhiloxSeq rng(seed, (uint32_t)j, 0u, kStreamWalkInit);
    WalkerState& s = f->s[j];
    s.x = 0; s.y = 0;
    s.mt = sample_categorical(rng, d->walker.p_init, 3);
    walker_init_velocity(d->walker, rng, s.mt, s.vx, s.vy);
  }
  return f;
}
void pzo_walker_destroy(void* h) { delete (OWalker*)h; }

// one step of `particles N (walkerSimulate measurements)` (Walker.hs:113-131; Inference.hs:73-86)
int pzo_walker_step(void* h, const double* obs, double* ess_out, int32_t* resampled_out) {
  OWalker* f = (OWalker*)h;
  const int64_t n = f->n;
  ResamplePlan pl;
  if (!build_plan(f->lw.data(), n, n, &pl)) return PZ_EDEGENERATE;
  const bool resampled = !(f->desc.ess_threshold > 0) || f->resample_next;
  if (!resampled) for (int64_t j = 0; j < n; ++j) f->anc[j] = (int32_t)j;
  else if (f->desc.resampler == PZ_RESAMPLE_MULTINOMIAL_REF) multinomial_ancestors(pl, f->seed, f->t, n, f->anc.data());
  else systematic_ancestors(pl, resample_uniform(f->seed, f->t), n, f->anc.data());
  const bool carried = f->desc.carry_weights || !resampled;
  const float rebase = (float)pl.emax * 0x1.62e43p-1f;
  std::vector<WalkerState> sn(n);
  std::vector<float> lwn(n);
  const double v = f->desc.walker.obs_std * f->desc.walker.obs_std;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t j = 0; j < n; ++j) {
    const int64_t a = f->anc[j];
    WalkerState s = f->s[a];
    PhiloxSeq rng(f->seed, (uint32_t)j, f->t, kStreamWalker);
    walker_motion(f->desc.walker, rng, s);
    const double inc = score_normal_2x(s.x, v, obs[0]) + score_normal_2x(s.y, v, obs[1]);   // walkerMeasure :88-94
    sn[j] = s;
    lwn[j] = carried ? (f->lw[a] - rebase) + (float)inc : (float)inc;
  }
  f->s.swap(sn); f->lw.swap(lwn);
  ResamplePlan pl2;
  const bool ok = build_plan(f->lw.data(), n, n, &pl2);
  double sw = 0, sw2 = 0;
  if (ok) for (int64_t i = 0; i < n; ++i) {
    const int32_t eb = pl2.eb[i / kBlk];
    if (eb == kNZero) continue;
    const double w = ldexp((double)pl2.q[i], eb - pl2.emax - 15);
    sw += w; sw2 += w * w;
  }
  const double ess = ok ? sw * sw / sw2 : 0;
  if (ess_out) *ess_out = ess;
  if (resampled_out) *resampled_out = resampled ? 1 : 0;
  f->resample_next = !(f->desc.ess_threshold > 0) || ess < f->desc.ess_threshold * (double)n;
  f->t += 1;
  return ok ? 0 : PZ_EDEGENERATE;
}

// state as SoA [5][n] (x, y, vx, vy, motion type) and the stored log-weights
void pzo_walker_state(void* h, double* x, float* lw, int32_t* anc) {
  OWalker* f = (OWalker*)h;
  const int64_t n = f->n;
  if (x) for (int64_t j = 0; j < n; ++j) {
    x[j] = f->s[j].x; x[n + j] = f->s[j].y; x[2 * n + j] = f->s[j].vx; x[3 * n + j] = f->s[j].vy; x[4 * n + j] = f->s[j].mt;
  }
  if (lw) memcpy(lw, f->lw.data(), n * sizeof(float));
  if (anc) memcpy(anc, f->anc.data(), n * sizeof(int32_t));
}

// observation stream of the walker's generative model (generateWalker, Walker.hs:113-123): (dt, measured position)
int pzo_walker_generate(const pz_model_desc* d, uint64_t seed, int64_t T, double* obs, double* truth) {
  PhiloxSeq rng(seed, 0u, 0u, kStreamObs);
  WalkerState s;
  s.x = 0; s.y = 0;
  s.mt = sample_categorical(rng, d->walker.p_init, 3);
  walker_init_velocity(d->walker, rng, s.mt, s.vx, s.vy);
  for (int64_t t = 0; t < T; ++t) {
    walker_motion(d->walker, rng, s);
    obs[2 * t] = sample_normal(rng, s.x, d->walker.obs_std);      // walkerGenMeasurement: D.normal x positionStdDev (variance = 10)
    obs[2 * t + 1] = sample_normal(rng, s.y, d->walker.obs_std);
    if (truth) { truth[5 * t] = s.x; truth[5 * t + 1] = s.y; truth[5 * t + 2] = s.vx; truth[5 * t + 3] = s.vy; truth[5 * t + 4] = s.mt; }
  }
  return 0;
}

}  // extern "C"

extern "C" {

// ---------------------------------------------------------------------------------------
// Exact Kalman filter, kalmanPredict / kalmanUpdate (MVDistributions.hs:75-100), double.
// mean[nx], cov[nx*nx] row-major, updated in place.  For ny == 1 with scalar_ll_2x the
// exact target of the reference's doubled score is R_eff = R/2.
// ---------------------------------------------------------------------------------------
int pzo_kalman_step(const pz_model_desc* d, const double* y, double* mean, double* cov, double* loglik) {
  int nx = d->nx, ny = d->ny;
  double Rm[PZ_MAX_NY * PZ_MAX_NY];
  bool dbl = (ny == 1 && d->scalar_ll_2x);
  for (int i = 0; i < ny * ny; ++i) Rm[i] = d->R[i] * (dbl ? 0.5 : 1.0);
  // predict: (F x, F P F^T + Q)
  double xm[PZ_MAX_NX], FP[PZ_MAX_NX * PZ_MAX_NX], Pm[PZ_MAX_NX * PZ_MAX_NX];
  for (int i = 0; i < nx; ++i) { xm[i] = 0; for (int k = 0; k < nx; ++k) xm[i] += d->F[i * nx + k] * mean[k]; }
  for (int i = 0; i < nx; ++i) for (int j = 0; j < nx; ++j) {
    double s = 0; for (int k = 0; k < nx; ++k) s += d->F[i * nx + k] * cov[k * nx + j]; FP[i * nx + j] = s; }
  for (int i = 0; i < nx; ++i) for (int j = 0; j < nx; ++j) {
    double s = 0; for (int k = 0; k < nx; ++k) s += FP[i * nx + k] * d->F[j * nx + k]; Pm[i * nx + j] = s + d->Q[i * nx + j]; }
  // update: S = R + H P H^T, K = P H^T S^-1
  double PHt[PZ_MAX_NX * PZ_MAX_NY], S[PZ_MAX_NY][2 * PZ_MAX_NY] = {}, yt[PZ_MAX_NY];
  for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) {
    double s = 0; for (int k = 0; k < nx; ++k) s += Pm[i * nx + k] * d->H[j * nx + k]; PHt[i * ny + j] = s; }
  for (int i = 0; i < ny; ++i) {
    for (int j = 0; j < ny; ++j) {
      double s = 0; for (int k = 0; k < nx; ++k) s += d->H[i * nx + k] * PHt[k * ny + j]; S[i][j] = s + Rm[i * ny + j]; }
    S[i][ny + i] = 1.0;
    double s = 0; for (int k = 0; k < nx; ++k) s += d->H[i * nx + k] * xm[k]; yt[i] = y[i] - s;
  }
  double logdet = 0;
  for (int c = 0; c < ny; ++c) {
    double piv = S[c][c]; if (!(piv > 0)) return PZ_EINVAL; logdet += log(piv);
    for (int j = 0; j < 2 * ny; ++j) S[c][j] /= piv;
    for (int r = 0; r < ny; ++r) { if (r == c) continue; double f = S[r][c]; for (int j = 0; j < 2 * ny; ++j) S[r][j] -= f * S[c][j]; }
  }
  double Siy[PZ_MAX_NY], quad = 0;
  for (int i = 0; i < ny; ++i) { double s = 0; for (int j = 0; j < ny; ++j) s += S[i][ny + j] * yt[j]; Siy[i] = s; quad += yt[i] * s; }
  if (loglik) *loglik = -0.5 * (quad + logdet + ny * log(2 * M_PI));  // correct m-dim constant
  double K[PZ_MAX_NX * PZ_MAX_NY];
  for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) {
    double s = 0; for (int k = 0; k < ny; ++k) s += PHt[i * ny + k] * S[k][ny + j]; K[i * ny + j] = s; }
  for (int i = 0; i < nx; ++i) { double s = 0; for (int j = 0; j < ny; ++j) s += K[i * ny + j] * yt[j]; mean[i] = xm[i] + s; }
  for (int i = 0; i < nx; ++i) for (int j = 0; j < nx; ++j) {
    double s = 0; for (int k = 0; k < ny; ++k) s += K[i * ny + k] * PHt[j * ny + k]; cov[i * nx + j] = Pm[i * nx + j] - s; }
  (void)Siy;
  return 0;
}

// ---------------------------------------------------------------------------------------
// Synthetic observation stream from the generative model (Demo.hs:203-207,
// SingleTargetTracking.hs:61-65): x_t ~ N(F x_{t-1}, Q), y_t ~ N(H x_t, R), in double with
// libm Box-Muller on Philox stream "OBSG".  truth may be NULL.
// ---------------------------------------------------------------------------------------
static void std_normals(uint64_t seed, uint32_t t, uint32_t call, double z[4]) {
  uint32_t r[4];
  philox(call, 0, t, kStreamObs, seed, r);
  for (int k = 0; k < 2; ++k) {
    double u1 = ((double)r[2 * k] + 0.5) * 0x1p-32, u2 = ((double)r[2 * k + 1] + 0.5) * 0x1p-32;
    double R = sqrt(-2.0 * log(u1));
    z[2 * k] = R * cos(2 * M_PI * u2); z[2 * k + 1] = R * sin(2 * M_PI * u2);
  }
}

int pzo_generate_observations(const pz_model_desc* d, uint64_t seed, int64_t T, double* obs, double* truth) {
  int nx = d->nx, ny = d->ny;
  double Lq[PZ_MAX_NX][PZ_MAX_NX] = {}, Lr[PZ_MAX_NY][PZ_MAX_NY] = {};
  for (int j = 0; j < nx; ++j) {
    double s = d->Q[j * nx + j]; for (int k = 0; k < j; ++k) s -= Lq[j][k] * Lq[j][k];
    double l = s > 0 ? sqrt(s) : 0; Lq[j][j] = l;
    for (int i = j + 1; i < nx; ++i) { double t = d->Q[i * nx + j]; for (int k = 0; k < j; ++k) t -= Lq[i][k] * Lq[j][k]; Lq[i][j] = l > 0 ? t / l : 0; }
  }
  for (int j = 0; j < ny; ++j) {
    double s = d->R[j * ny + j]; for (int k = 0; k < j; ++k) s -= Lr[j][k] * Lr[j][k];
    double l = s > 0 ? sqrt(s) : 0; Lr[j][j] = l;
    for (int i = j + 1; i < ny; ++i) { double t = d->R[i * ny + j]; for (int k = 0; k < j; ++k) t -= Lr[i][k] * Lr[j][k]; Lr[i][j] = l > 0 ? t / l : 0; }
  }
  double x[PZ_MAX_NX];
  for (int i = 0; i < nx; ++i) x[i] = d->x0[i];
  for (int64_t t = 0; t < T; ++t) {
    double z[12];
    std_normals(seed, (uint32_t)t, 0, z); std_normals(seed, (uint32_t)t, 1, z + 4); std_normals(seed, (uint32_t)t, 2, z + 8);
    double xn[PZ_MAX_NX];
    for (int i = 0; i < nx; ++i) {
      double s = 0; for (int k = 0; k < nx; ++k) s += d->F[i * nx + k] * x[k];
      for (int k = 0; k <= i; ++k) s += Lq[i][k] * z[k];
      xn[i] = s;
    }
    for (int i = 0; i < nx; ++i) x[i] = xn[i];
    for (int i = 0; i < ny; ++i) {
      double s = 0; for (int k = 0; k < nx; ++k) s += d->H[i * nx + k] * x[k];
      for (int k = 0; k <= i; ++k) s += Lr[i][k] * z[8 + k];
      obs[t * ny + i] = s;
    }
    if (truth) for (int i = 0; i < nx; ++i) truth[t * nx + i] = x[i];
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// Reference-faithful particle filter in double (CPU-ref row of BASELINE.md section 4):
//   zparticles' (Inference.hs:101-104): N weighted model steps, then resample2;
//   normalize (Inference.hs:63-66): total mass by a left fold of log-domain (+);
//   labeledLogCategorical (Inference.hs:57-58): per draw, exponentiate all N normalised
//   weights and walk the CDF -- O(N) per draw, O(N^2) per step, exactly the reference's
//   cost profile.  Output per step: `average particles` after resampling (Demo.hs:207).
// mode 0 = faithful O(N^2) walk; mode 1 = same draws by binary search on one CDF
// (identical ancestors, O(N log N)), used where the faithful walk would not finish.
// ---------------------------------------------------------------------------------------
static inline double logaddexp(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  double m = a > b ? a : b;
  return m + log1p(exp(-fabs(a - b)));
}

int pzo_refpf_run(const pz_model_desc* d, int64_t n, uint64_t seed, const double* obs, int64_t T, int mode,
                  double* mean_out, double* var_out, int threads) {
  int nx = d->nx, ny = d->ny;
  LgParams lp;
  if (!derive_lg(d, &lp)) return PZ_EINVAL;
  // double-precision model constants
  double Lq[PZ_MAX_NX][PZ_MAX_NX] = {}, Rinv[PZ_MAX_NY][PZ_MAX_NY];
  for (int i = 0; i < nx; ++i) for (int j = 0; j < nx; ++j) Lq[i][j] = lp.Lq[i][j];  // diagonal in all in-scope models
  {
    // recompute in double for fidelity
    for (int j = 0; j < nx; ++j) {
      double s = d->Q[j * nx + j]; for (int k = 0; k < j; ++k) s -= Lq[j][k] * Lq[j][k];
      double l = s > 0 ? sqrt(s) : 0; Lq[j][j] = l;
      for (int i = j + 1; i < nx; ++i) { double t = d->Q[i * nx + j]; for (int k = 0; k < j; ++k) t -= Lq[i][k] * Lq[j][k]; Lq[i][j] = l > 0 ? t / l : 0; }
    }
    double A[PZ_MAX_NY][2 * PZ_MAX_NY] = {};
    for (int i = 0; i < ny; ++i) { for (int j = 0; j < ny; ++j) A[i][j] = d->R[i * ny + j]; A[i][ny + i] = 1; }
    for (int c = 0; c < ny; ++c) {
      double piv = A[c][c]; for (int j = 0; j < 2 * ny; ++j) A[c][j] /= piv;
      for (int r = 0; r < ny; ++r) { if (r == c) continue; double f = A[r][c]; for (int j = 0; j < 2 * ny; ++j) A[r][j] -= f * A[c][j]; }
    }
    for (int i = 0; i < ny; ++i) for (int j = 0; j < ny; ++j) Rinv[i][j] = A[i][ny + j];
  }
  double wscale = lp.wscale, llc = lp.ll_const;
#if defined(_OPENMP)
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
  std::vector<double> x((size_t)nx * n), xn((size_t)nx * n), lw(n), w(n), cdf(n);
  std::vector<int64_t> anc(n);
  for (int k = 0; k < nx; ++k) for (int64_t i = 0; i < n; ++i) x[(size_t)k * n + i] = d->x0[k];
  for (int64_t t = 0; t < T; ++t) {
    const double* y = obs + t * ny;
    // res <- sequence [ runWeighted (f a) | f <- fs ]
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      uint32_t r[4]; double z[8];
      for (int c = 0; 4 * c < nx; ++c) {
        philox((uint32_t)i, (uint32_t)c, (uint32_t)t, kStreamProp, seed ^ 0x5EEDull, r);
        for (int k = 0; k < 2; ++k) {
          double u1 = ((double)r[2 * k] + 0.5) * 0x1p-32, u2 = ((double)r[2 * k + 1] + 0.5) * 0x1p-32;
          double R = sqrt(-2.0 * log(u1));
          z[4 * c + 2 * k] = R * cos(2 * M_PI * u2); z[4 * c + 2 * k + 1] = R * sin(2 * M_PI * u2);
        }
      }
      double xo[PZ_MAX_NX];
      for (int a = 0; a < nx; ++a) {
        double s = 0; for (int k = 0; k < nx; ++k) s += d->F[a * nx + k] * x[(size_t)k * n + i];
        for (int k = 0; k <= a; ++k) s += Lq[a][k] * z[k];
        xo[a] = s; xn[(size_t)a * n + i] = s;
      }
      double dd[PZ_MAX_NY], quad = 0;
      for (int a = 0; a < ny; ++a) { double s = 0; for (int k = 0; k < nx; ++k) s += d->H[a * nx + k] * xo[k]; dd[a] = y[a] - s; }
      for (int a = 0; a < ny; ++a) { double s = 0; for (int k = 0; k < ny; ++k) s += Rinv[a][k] * dd[k]; quad += dd[a] * s; }
      lw[i] = wscale * quad + llc;  // score accumulated by `observe`
    }
    if (mode == 2) {
      // "optimised CPU" row of SURVEY.md 8(d)(ii): not the reference's algorithm but the best a host
      // does with the same step -- max-shifted weights (Util/Numeric.hs:12-17), one O(N) systematic
      // resample from a chunked parallel prefix sum, parallel gather; every loop over all cores.
      double mx = -INFINITY;
#pragma omp parallel for schedule(static) reduction(max : mx)
      for (int64_t i = 0; i < n; ++i) mx = lw[i] > mx ? lw[i] : mx;
      if (!(mx > -INFINITY)) return PZ_EDEGENERATE;
      int nchunk = 1;
#if defined(_OPENMP)
      nchunk = omp_get_max_threads();
#endif
      if ((int64_t)nchunk > n) nchunk = (int)n;
      std::vector<double> csum((size_t)nchunk + 1, 0.0);
      const int64_t per = (n + nchunk - 1) / nchunk;
#pragma omp parallel for schedule(static, 1)
      for (int c = 0; c < nchunk; ++c) {
        const int64_t lo = c * per, hi = std::min<int64_t>(n, lo + per);
        double acc = 0;
        for (int64_t i = lo; i < hi; ++i) { acc += exp(lw[i] - mx); cdf[i] = acc; }
        csum[c + 1] = acc;
      }
      for (int c = 0; c < nchunk; ++c) csum[c + 1] += csum[c];
      const double total = csum[nchunk];
#pragma omp parallel for schedule(static, 1)
      for (int c = 1; c < nchunk; ++c) {
        const int64_t lo = c * per, hi = std::min<int64_t>(n, lo + per);
        for (int64_t i = lo; i < hi; ++i) cdf[i] += csum[c];
      }
      uint32_t r[4];
      philox(0u, 0u, (uint32_t)t, kStreamMulti, seed ^ 0x5EEDull, r);
      const double u = ((double)r[0] + 0.5) * 0x1p-32, stepw = total / (double)n;
#pragma omp parallel for schedule(static, 1)
      for (int c = 0; c < nchunk; ++c) {  // slots [lo, hi): one search, then a merge walk
        const int64_t lo = c * per, hi = std::min<int64_t>(n, lo + per);
        if (lo >= hi) continue;
        int64_t k = std::upper_bound(cdf.begin(), cdf.end(), ((double)lo + u) * stepw) - cdf.begin();
        for (int64_t j = lo; j < hi; ++j) {
          const double pos = ((double)j + u) * stepw;
          while (k < n - 1 && cdf[k] <= pos) ++k;
          anc[j] = k < n ? k : n - 1;
        }
      }
      for (int a = 0; a < nx; ++a) {
        double s1 = 0, s2 = 0;
        double* dst = &x[(size_t)a * n]; const double* src = &xn[(size_t)a * n];
#pragma omp parallel for schedule(static) reduction(+ : s1, s2)
        for (int64_t j = 0; j < n; ++j) { const double v = src[anc[j]]; dst[j] = v; s1 += v; s2 += v * v; }
        if (mean_out) mean_out[t * nx + a] = s1 / n;
        if (var_out) var_out[t * nx + a] = s2 / n - (s1 / n) * (s1 / n);
      }
      continue;
    }
    // normalize: totalMass = sum (map snd xs)  -- left fold of Log's (+).  mode 0 keeps the literal serial
    // fold; mode 1 (the all-cores port) folds per-thread chunks with the same logaddexp and combines the
    // chunk totals in order (the same sum up to rounding).
    double tot = -INFINITY;
    int nchunk = 1;
#if defined(_OPENMP)
    if (mode == 1) nchunk = omp_get_max_threads();
#endif
    if ((int64_t)nchunk > n) nchunk = (int)n;
    const int64_t per = (n + nchunk - 1) / nchunk;
    std::vector<double> ctot((size_t)nchunk, -INFINITY);
#pragma omp parallel for schedule(static, 1) if (mode == 1)
    for (int c = 0; c < nchunk; ++c) {
      const int64_t lo = c * per, hi = std::min<int64_t>(n, lo + per);
      double acc = -INFINITY;
      for (int64_t i = lo; i < hi; ++i) acc = logaddexp(acc, lw[i]);
      ctot[c] = acc;
    }
    for (int c = 0; c < nchunk; ++c) tot = logaddexp(tot, ctot[c]);
    if (!(tot > -INFINITY)) return PZ_EDEGENERATE;
    // replicateM N (labeledLogCategorical normalized)
    if (mode == 0) {
#pragma omp parallel for schedule(static)
      for (int64_t j = 0; j < n; ++j) {
        uint32_t r[4];
        philox((uint32_t)j, 0, (uint32_t)t, kStreamMulti, seed ^ 0x5EEDull, r);
        double u = ((double)(((uint64_t)r[0] << 21) ^ (r[1] >> 11)) + 0.5) * 0x1p-53;
        double c = 0; int64_t k = n - 1;
        for (int64_t i = 0; i < n; ++i) { c += exp(lw[i] - tot); if (u < c) { k = i; break; } }
        anc[j] = k;
      }
    } else {
      // the CDF every draw of logCategorical walks, built once (chunked over the cores), then one binary
      // search per draw instead of the O(N) walk: identical ancestors up to the rounding of the partial sums
      std::vector<double> csum((size_t)nchunk + 1, 0.0);
#pragma omp parallel for schedule(static, 1)
      for (int c = 0; c < nchunk; ++c) {
        const int64_t lo = c * per, hi = std::min<int64_t>(n, lo + per);
        double acc = 0;
        for (int64_t i = lo; i < hi; ++i) { acc += exp(lw[i] - tot); cdf[i] = acc; }
        csum[c + 1] = acc;
      }
      for (int c = 0; c < nchunk; ++c) csum[c + 1] += csum[c];
#pragma omp parallel for schedule(static, 1)
      for (int c = 1; c < nchunk; ++c) {
        const int64_t lo = c * per, hi = std::min<int64_t>(n, lo + per);
        for (int64_t i = lo; i < hi; ++i) cdf[i] += csum[c];
      }
#pragma omp parallel for schedule(static)
      for (int64_t j = 0; j < n; ++j) {
        uint32_t r[4];
        philox((uint32_t)j, 0, (uint32_t)t, kStreamMulti, seed ^ 0x5EEDull, r);
        double u = ((double)(((uint64_t)r[0] << 21) ^ (r[1] >> 11)) + 0.5) * 0x1p-53;
        int64_t k = std::upper_bound(cdf.begin(), cdf.end(), u) - cdf.begin();
        anc[j] = k < n ? k : n - 1;
      }
    }
    // unzip <$> resample2 res ; output = average particles (Util/Numeric.hs:7-10)
    for (int a = 0; a < nx; ++a) {
      double s = 0, s2 = 0;
      double* dst = &x[(size_t)a * n]; const double* src = &xn[(size_t)a * n];
#pragma omp parallel for schedule(static) reduction(+ : s, s2) if (mode == 1)
      for (int64_t j = 0; j < n; ++j) { double v = src[anc[j]]; dst[j] = v; s += v; s2 += v * v; }
      if (mean_out) mean_out[t * nx + a] = s / n;
      if (var_out) var_out[t * nx + a] = s2 / n - (s / n) * (s / n);
    }
  }
  return 0;
}

int pzo_num_threads(void) {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"

// =========================================================================================
// Multi-target tracker (Examples/MultiTargetTracking.hs:26-145), restated in double.
//
// Per particle: a map id -> Rao-Blackwellised track (Gaussian marginal over R^4), the heap of
// zdsparticles reduced to its closed-form Kalman recursions:
//   trackMotion     (:56-67)  x' ~ N(F x, Q)            -> Kalman predict (makeMarginal, DelayedSampling.hs:135-139)
//   trackMeasurement(:69-73)  z  ~ N(H x', R)           -> score' (:124-128) / kalmanUpdate (MVDistributions.hs:83-100)
//   proposeAssocs   (:111-124) sequential categorical over {Clutter < NewTrack < Track id asc}
//   observingWithProposal (Metaprob.hs:157-162): w *= p(assocs, obs | state) / q(assocs)
//   sampleStep      (:89-99)  count prior shuffleWithRepeats' (Distributions.hs:224-233)
//   updateWithAssocs(:75-87)  births, coasting Bernoulli((1-pd)/(1-ps*pd)) as written
// The weight is accumulated term by term as the reference does; the algebraic simplification
// used by the CUDA kernel (sum_j log L_j + n_undetected log(1 - ps pd) + const) is checked
// against it on every call (pzo_mtt_max_simplification_error).
// =========================================================================================
namespace {

constexpr uint32_t kStreamMtt = 0x4D545453u;     // "MTTS": association / coasting uniforms
constexpr uint32_t kStreamMttGen = 0x4D545447u;  // "MTTG": ground-truth generator

struct MttTrack { int id; double m[4]; double P[4][4]; };

struct MttConst {
  double F[4][4], Q[4][4];
  double lc, lb, pd, ps, cv, bp, bv, r;
  int kmax;
};

MttConst mtt_const(const pz_model_desc* d) {
  MttConst c;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { c.F[i][j] = d->F[i * 4 + j]; c.Q[i][j] = d->Q[i * 4 + j]; }
  c.lc = d->mtt.clutter_lambda; c.lb = d->mtt.new_track_lambda; c.pd = d->mtt.pd; c.ps = d->mtt.survival_prob;
  c.cv = d->mtt.clutter_var; c.bp = d->mtt.birth_pos_var; c.bv = d->mtt.birth_vel_var; c.r = d->mtt.meas_var;
  c.kmax = d->mtt.k_max;
  return c;
}

struct U01 {  // sequential uniforms from one Philox counter family
  uint64_t seed; uint32_t a, t, stream; uint32_t call = 0; int have = 0; uint32_t r[4];
  U01(uint64_t s, uint32_t a_, uint32_t t_, uint32_t st) : seed(s), a(a_), t(t_), stream(st) {}
  double next() {
    if (!have) { philox(a, call++, t, stream, seed, r); have = 4; }
    return ((double)r[4 - have--] + 0.5) * 0x1p-32;
  }
  double normal() { double u1 = next(), u2 = next(); return sqrt(-2.0 * log(u1)) * cos(2 * M_PI * u2); }
};

void mtt_predict(const MttConst& c, MttTrack& t) {
  double m[4], FP[4][4], P[4][4];
  for (int i = 0; i < 4; ++i) { m[i] = 0; for (int k = 0; k < 4; ++k) m[i] += c.F[i][k] * t.m[k]; }
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += c.F[i][k] * t.P[k][j]; FP[i][j] = s; }
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += FP[i][k] * c.F[j][k]; P[i][j] = s + c.Q[i][j]; }
  memcpy(t.m, m, sizeof(m)); memcpy(t.P, P, sizeof(P));
}

// log N(z; mu2, S) for a 2x2 S = [[a, b], [b, d]]  (mvNormal_ll0, MVDistributions.hs:45-46)
double ll2(const double* z, double mu0, double mu1, double a, double b, double d) {
  double det = a * d - b * b, d0 = z[0] - mu0, d1 = z[1] - mu1;
  double quad = (d * d0 * d0 - 2 * b * d0 * d1 + a * d1 * d1) / det;
  return -0.5 * (quad + log(det) + 2 * log(2 * M_PI));
}

// kalmanUpdate with H = [I2 0], R = r I2  (MVDistributions.hs:83-100)
void mtt_update(MttTrack& t, const double* z, double r) {
  double a = t.P[0][0] + r, b = t.P[0][1], d = t.P[1][1] + r, det = a * d - b * b;
  double Si[2][2] = {{d / det, -b / det}, {-b / det, a / det}};
  double y0 = z[0] - t.m[0], y1 = z[1] - t.m[1];
  double K[4][2];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 2; ++j) K[i][j] = t.P[i][0] * Si[0][j] + t.P[i][1] * Si[1][j];
  double P[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) P[i][j] = t.P[i][j] - (K[i][0] * t.P[0][j] + K[i][1] * t.P[1][j]);
  for (int i = 0; i < 4; ++i) t.m[i] += K[i][0] * y0 + K[i][1] * y1;
  memcpy(t.P, P, sizeof(P));
}

double poisson_ll(double lambda, int k) { return k * log(lambda) - lambda - lgamma(k + 1.0); }
double log_fact(int n) { return lgamma(n + 1.0); }

struct OMtt {
  pz_model_desc desc; MttConst c;
  int64_t n; uint64_t seed; uint32_t t = 0;
  std::vector<std::vector<MttTrack>> tr;  // per particle, sorted by id
  std::vector<int> next_id;
  std::vector<float> lw;
  std::vector<int32_t> anc;
  double max_simpl_err = 0;
  int64_t overflow = 0;
};

MttTrack mtt_birth(const MttConst& c, int id) {
  MttTrack t; t.id = id;
  for (int i = 0; i < 4; ++i) { t.m[i] = 0; for (int j = 0; j < 4; ++j) t.P[i][j] = 0; }
  t.P[0][0] = t.P[1][1] = c.bp; t.P[2][2] = t.P[3][3] = c.bv;
  return t;
}

// one particle, one step; returns the log-weight increment
double mtt_particle_step(OMtt& f, std::vector<MttTrack>& tracks, int& next_id, const double* obs, int M, U01& rng) {
  const MttConst& c = f.c;
  const double pspd = c.ps * c.pd;
  std::vector<MttTrack> pred = tracks;
  for (auto& t : pred) mtt_predict(c, t);
  const int K = (int)pred.size();
  std::vector<int> assoc_of_track(K, -1);   // obs index or -1
  std::vector<int> assoc(M);                // -1 clutter, -2 new, k >= 0 track slot
  double q_log = 0, obs_ll = 0, sum_logL = 0;
  int n_c = 0, n_n = 0;
  std::vector<double> like(K + 2), llv(K + 2);
  for (int j = 0; j < M; ++j) {
    const double* z = obs + 2 * j;
    llv[0] = ll2(z, 0, 0, c.cv, 0, c.cv); like[0] = c.lc * exp(llv[0]);
    llv[1] = ll2(z, 0, 0, c.bp + c.r, 0, c.bp + c.r); like[1] = c.lb * exp(llv[1]);
    double tot = like[0] + like[1];
    for (int k = 0; k < K; ++k) {
      if (assoc_of_track[k] >= 0) { like[2 + k] = 0; continue; }
      llv[2 + k] = ll2(z, pred[k].m[0], pred[k].m[1], pred[k].P[0][0] + c.r, pred[k].P[0][1], pred[k].P[1][1] + c.r);
      like[2 + k] = pspd * exp(llv[2 + k]);
      tot += like[2 + k];
    }
    double u = rng.next() * tot, cacc = 0; int pick = -1;
    for (int a = 0; a < K + 2; ++a) {
      if (a >= 2 && assoc_of_track[a - 2] >= 0) continue;
      cacc += like[a]; pick = a;
      if (u < cacc) break;
    }
    q_log += log(like[pick] / tot); obs_ll += llv[pick]; sum_logL += log(tot);
    if (pick == 0) { assoc[j] = -1; ++n_c; }
    else if (pick == 1) { assoc[j] = -2; ++n_n; }
    else { assoc[j] = pick - 2; assoc_of_track[pick - 2] = j; }
  }
  // count prior: shuffleWithRepeats' score (Distributions.hs:230-233)
  double prior = poisson_ll(c.lc, n_c) + poisson_ll(c.lb, n_n);
  int n_undet = 0;
  for (int k = 0; k < K; ++k) { if (assoc_of_track[k] >= 0) prior += log(pspd); else { prior += log(1 - pspd); ++n_undet; } }
  prior -= (log_fact(M) - log_fact(n_c) - log_fact(n_n));
  const double lw_inc = -q_log + prior + obs_ll;
  const double simpl = sum_logL + n_undet * log(1 - pspd) - (c.lc + c.lb) - log_fact(M);
  f.max_simpl_err = std::max(f.max_simpl_err, fabs(simpl - lw_inc));
  // update: observed tracks conditioned, unobserved coast with Bernoulli((1-pd)/(1-ps pd)), births appended
  std::vector<MttTrack> out;
  const double p_coast = (1 - c.pd) / (1 - pspd);
  for (int k = 0; k < K; ++k) {
    if (assoc_of_track[k] >= 0) { mtt_update(pred[k], obs + 2 * assoc_of_track[k], c.r); out.push_back(pred[k]); }
    else if (rng.next() < p_coast) out.push_back(pred[k]);
  }
  for (int j = 0; j < M; ++j)
    if (assoc[j] == -2) {
      MttTrack t = mtt_birth(c, next_id++);
      mtt_update(t, obs + 2 * j, c.r);
      if ((int)out.size() < c.kmax) out.push_back(t); else ++f.overflow;
    }
  tracks.swap(out);
  return lw_inc;
}

}  // namespace

extern "C" {

void* pzo_mtt_create(const pz_model_desc* d, int64_t n, uint64_t seed) {
  OMtt* f = new OMtt();
  f->desc = *d; f->c = mtt_const(d); f->n = n; f->seed = seed;
  f->tr.resize(n); f->next_id.assign(n, 0); f->lw.assign(n, 0.0f); f->anc.assign(n, 0);
  return f;
}
void pzo_mtt_destroy(void* h) { delete (OMtt*)h; }

// one filter step: canonical systematic resample of the pending weights, then every child
// runs observeStep on the observations (n_obs rows of 2).  Outputs the weighted mean / variance
// of the track count and the log-evidence increment.
int pzo_mtt_step(void* h, const double* obs, int32_t n_obs, double* mean_cnt, double* var_cnt, double* logz) {
  OMtt* f = (OMtt*)h;
  ResamplePlan pl;
  if (!build_plan(f->lw.data(), f->n, f->n, &pl)) return PZ_EDEGENERATE;
  systematic_ancestors(pl, pzo_resample_uniform(f->seed, f->t), f->n, f->anc.data());
  std::vector<std::vector<MttTrack>> ntr(f->n);
  std::vector<int> nid(f->n);
  std::vector<double> lwd(f->n);
  for (int64_t j = 0; j < f->n; ++j) {
    ntr[j] = f->tr[f->anc[j]]; nid[j] = f->next_id[f->anc[j]];
    U01 rng(f->seed, (uint32_t)j, f->t, kStreamMtt);
    lwd[j] = mtt_particle_step(*f, ntr[j], nid[j], obs, n_obs, rng);
    f->lw[j] = (float)lwd[j];
  }
  f->tr.swap(ntr); f->next_id.swap(nid);
  double mx = -INFINITY; for (double v : lwd) mx = std::max(mx, v);
  double sw = 0, sc = 0, scc = 0;
  for (int64_t j = 0; j < f->n; ++j) { double w = exp(lwd[j] - mx), c = (double)f->tr[j].size(); sw += w; sc += w * c; scc += w * c * c; }
  if (mean_cnt) *mean_cnt = sc / sw;
  if (var_cnt) *var_cnt = scc / sw - (sc / sw) * (sc / sw);
  if (logz) *logz = mx + log(sw / f->n);
  f->t += 1;
  return 0;
}

double pzo_mtt_max_simplification_error(void* h) { return ((OMtt*)h)->max_simpl_err; }
void pzo_mtt_logw(void* h, float* out) { OMtt* f = (OMtt*)h; memcpy(out, f->lw.data(), f->n * sizeof(float)); }

// tracks of particle i: returns the count; ids[k], mean[k*4..], cov[k*16..]
int pzo_mtt_particle(void* h, int64_t i, int32_t* ids, double* mean, double* cov) {
  OMtt* f = (OMtt*)h;
  const auto& v = f->tr[i];
  for (size_t k = 0; k < v.size(); ++k) {
    ids[k] = v[k].id;
    for (int a = 0; a < 4; ++a) { mean[k * 4 + a] = v[k].m[a]; for (int b = 0; b < 4; ++b) cov[k * 16 + a * 4 + b] = v[k].P[a][b]; }
  }
  return (int)v.size();
}

// weighted posterior mean of the state of track `id` over the particles that carry it, and the
// weighted fraction of particles carrying it
int pzo_mtt_track_posterior(void* h, int32_t id, double* mean4, double* frac) {
  OMtt* f = (OMtt*)h;
  double mx = -INFINITY; for (float v : f->lw) mx = std::max(mx, (double)v);
  double sw = 0, swt = 0, m[4] = {0, 0, 0, 0};
  for (int64_t j = 0; j < f->n; ++j) {
    double w = exp((double)f->lw[j] - mx); sw += w;
    for (const auto& t : f->tr[j]) if (t.id == id) { swt += w; for (int a = 0; a < 4; ++a) m[a] += w * t.m[a]; }
  }
  for (int a = 0; a < 4; ++a) mean4[a] = swt > 0 ? m[a] / swt : 0;
  *frac = swt / sw;
  return 0;
}

// Ground truth + observations: sampleStep under MP.sim (MultiTargetTracking.hs:126-131).
// obs_out: up to m_max rows of 2 per step (row-major [T][m_max][2]); n_obs_out[T];
// truth_out: [T][k_max][5] = (id, x0..x3), n_truth_out[T].
int pzo_mtt_generate(const pz_model_desc* d, uint64_t seed, int64_t T, int32_t m_max, double* obs_out, int32_t* n_obs_out,
                     double* truth_out, int32_t* n_truth_out) {
  MttConst c = mtt_const(d);
  struct Tr { int id; double x[4]; };
  std::vector<Tr> tracks; int next_id = 0;
  const double pspd = c.ps * c.pd, p_coast = (1 - c.pd) / (1 - pspd);
  double Lq[4][4] = {};
  for (int j = 0; j < 4; ++j) {
    double s = c.Q[j][j]; for (int k = 0; k < j; ++k) s -= Lq[j][k] * Lq[j][k];
    double l = s > 0 ? sqrt(s) : 0; Lq[j][j] = l;
    for (int i = j + 1; i < 4; ++i) { double t = c.Q[i][j]; for (int k = 0; k < j; ++k) t -= Lq[i][k] * Lq[j][k]; Lq[i][j] = l > 0 ? t / l : 0; }
  }
  auto poisson = [](double lambda, U01& r) { double L = exp(-lambda), p = 1; int k = 0; do { ++k; p *= r.next(); } while (p > L); return k - 1; };
  for (int64_t t = 0; t < T; ++t) {
    U01 rng(seed, 0, (uint32_t)t, kStreamMttGen);
    for (auto& tr : tracks) {  // trackMotion
      double z[4], xn[4];
      for (int i = 0; i < 4; ++i) z[i] = rng.normal();
      for (int i = 0; i < 4; ++i) { double s = 0; for (int k = 0; k < 4; ++k) s += c.F[i][k] * tr.x[k]; for (int k = 0; k <= i; ++k) s += Lq[i][k] * z[k]; xn[i] = s; }
      memcpy(tr.x, xn, sizeof(xn));
    }
    int n_c = poisson(c.lc, rng), n_n = poisson(c.lb, rng);
    std::vector<int> keys;  // -1 clutter, -2 new, k track slot
    for (int i = 0; i < n_c; ++i) keys.push_back(-1);
    for (int i = 0; i < n_n; ++i) keys.push_back(-2);
    std::vector<char> det(tracks.size(), 0);
    for (size_t k = 0; k < tracks.size(); ++k) if (rng.next() < pspd) { det[k] = 1; keys.push_back((int)k); }
    for (size_t i = keys.size(); i > 1; --i) { size_t j = (size_t)(rng.next() * i); std::swap(keys[i - 1], keys[j < i ? j : i - 1]); }  // uniform interleave
    std::vector<Tr> born;
    int M = 0;
    for (int key : keys) {
      if (M >= m_max) break;
      double z0, z1;
      if (key == -1) { z0 = sqrt(c.cv) * rng.normal(); z1 = sqrt(c.cv) * rng.normal(); }
      else if (key == -2) {
        Tr nt; nt.id = next_id++;
        nt.x[0] = sqrt(c.bp) * rng.normal(); nt.x[1] = sqrt(c.bp) * rng.normal(); nt.x[2] = sqrt(c.bv) * rng.normal(); nt.x[3] = sqrt(c.bv) * rng.normal();
        born.push_back(nt);
        z0 = nt.x[0] + sqrt(c.r) * rng.normal(); z1 = nt.x[1] + sqrt(c.r) * rng.normal();
      } else { z0 = tracks[key].x[0] + sqrt(c.r) * rng.normal(); z1 = tracks[key].x[1] + sqrt(c.r) * rng.normal(); }
      obs_out[(t * m_max + M) * 2] = z0; obs_out[(t * m_max + M) * 2 + 1] = z1; ++M;
    }
    n_obs_out[t] = M;
    std::vector<Tr> next;
    for (size_t k = 0; k < tracks.size(); ++k) if (det[k] || rng.next() < p_coast) next.push_back(tracks[k]);
    for (auto& b : born) if ((int)next.size() < c.kmax) next.push_back(b);
    tracks.swap(next);
    n_truth_out[t] = (int)tracks.size();
    for (size_t k = 0; k < tracks.size(); ++k) {
      truth_out[(t * c.kmax + k) * 5] = tracks[k].id;
      for (int a = 0; a < 4; ++a) truth_out[(t * c.kmax + k) * 5 + 1 + a] = tracks[k].x[a];
    }
  }
  return 0;
}

}  // extern "C"

// =========================================================================================
// MultiCamera tracker (Examples/MultiCamera.hs:40-227), restated in double with FULL matrices (the CUDA kernel
// uses the diagonal structure F = H = I, diagonal Q / R / prior imply; this file does not).
//
// Per particle: a list of tracks (box marginal over R^4, trackID, identity, camera, startTime) and a growing list of
// appearance marginals over R^10 shared by the tracks of one identity: the heap of zdsparticles (:224) reduced to
// Kalman recursions:
//   trackSurvivalMotion (:118-126)  survive ~ Bernoulli(exp(-deathRate tdiff)); posWH' ~ N(I posWH, diag(1,1,.1,.1)) -> predict
//   tracksMotion        (:183-195)  births ~ Poisson(birthRate tdiff); newTrack (:96-108): camera ~ U{0,1},
//                                   new appearance with probability 1/(A+1) (prior N(0, I10)) else identity ~ U{0..A-1};
//                                   box prior N((0,0,1,1), diag(1,1,.2,.2)) (:88-94)
//   trackMeasurement    (:128-132)  box ~ N(posWH, I4), pose ~ uniformNSphere, appear ~ N(pose . appearance, 1)
//   assocWithClutterCustomProposal (:162-180), per camera, tracks in list order: lls over the remaining observations,
//                                   i ~ categorical(lls / total), condition on observation i, weight *= total
//                                   (likelihood times the proposal correction 1 / adjLL_i); no observation left for a
//                                   track -> passert False (weight 0); the rest is clutter: count prior
//                                   -(lgamma(nT + nC + 1) - lgamma(nC + 1)), Poisson(1) count, clutter densities
//                                   box ~ N((0,0,1,1), I4), pose ~ N(0, I10), appear ~ N(0, 1) (:169-176, 182-189)
// Observation row: 16 doubles (camera, pose[10], box[4], appear).
// =========================================================================================
namespace {

constexpr uint32_t kStreamMcam = 0x4D43414Du;     // "MCAM"
constexpr uint32_t kStreamMcamGen = 0x4D434147u;  // "MCAG"
constexpr int kMcD = 10, kMcObsW = 16, kMcCams = 2;

struct McTrack { int id, identity, camera; double start; double m[4]; double P[4][4]; };
struct McApp { double m[kMcD]; double P[kMcD][kMcD]; };
struct McConst {
  double Q[4], ps, lb, lc, prior_m[4], prior_v[4], clutter_m[4], r_box, r_app;
  int kmax, amax;
};
McConst mcam_const(const pz_model_desc* d) {
  McConst c;
  for (int i = 0; i < 4; ++i) { c.Q[i] = d->Q[i * 4 + i]; c.prior_m[i] = d->x0[i]; c.clutter_m[i] = d->x0[i]; }
  c.prior_v[0] = c.prior_v[1] = d->mtt.birth_pos_var; c.prior_v[2] = c.prior_v[3] = d->mtt.birth_vel_var;
  c.ps = d->mtt.survival_prob; c.lb = d->mtt.new_track_lambda; c.lc = d->mtt.clutter_lambda;
  c.r_box = d->mtt.meas_var; c.r_app = 1.0;
  c.kmax = d->mtt.k_max; c.amax = 16;
  return c;
}

// log N(x; mu, S) for a dense symmetric positive definite S (mvNormal_ll0, MVDistributions.hs:45-46) by Cholesky
template <int N>
double ll_dense(const double* x, const double* mu, const double (*S)[N]) {
  double L[N][N] = {};
  double logdet = 0;
  for (int j = 0; j < N; ++j) {
    double s = S[j][j]; for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    L[j][j] = sqrt(s); logdet += 2 * log(L[j][j]);
    for (int i = j + 1; i < N; ++i) { double t = S[i][j]; for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k]; L[i][j] = t / L[j][j]; }
  }
  double y[N], quad = 0;
  for (int i = 0; i < N; ++i) { double t = x[i] - mu[i]; for (int k = 0; k < i; ++k) t -= L[i][k] * y[k]; y[i] = t / L[i][i]; quad += y[i] * y[i]; }
  return -0.5 * (quad + logdet + N * log(2 * M_PI));
}

// logNSphereVolume (MVDistributions.hs:64-66)
double log_nsphere_volume(int n) { return log((double)n) + n / 2.0 * log(M_PI) - lgamma(n / 2.0 + 1); }

double mc_score(const McConst& c, const McTrack& t, const McApp& a, const double* row) {
  const double* h = row + 1; const double* box = row + 11; const double app = row[15];
  double S[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) S[i][j] = t.P[i][j] + (i == j ? c.r_box : 0.0);
  double ll = ll_dense<4>(box, t.m, S);
  ll += -log_nsphere_volume(kMcD);
  double hm = 0, hPh = 0;
  for (int i = 0; i < kMcD; ++i) { hm += h[i] * a.m[i]; for (int j = 0; j < kMcD; ++j) hPh += h[i] * a.P[i][j] * h[j]; }
  const double s = hPh + c.r_app, d = app - hm;
  ll += -0.5 * (d * d / s + log(s) + log(2 * M_PI));
  return ll;
}

// kalmanUpdate (MVDistributions.hs:83-100) with H = I4, R = r I4: gain by Gaussian elimination on S = P + R
void mc_update_box(const McConst& c, McTrack& t, const double* box) {
  double S[4][8];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { S[i][j] = t.P[i][j] + (i == j ? c.r_box : 0.0); S[i][4 + j] = (i == j); }
  for (int p = 0; p < 4; ++p) {
    const double piv = S[p][p];
    for (int j = 0; j < 8; ++j) S[p][j] /= piv;
    for (int i = 0; i < 4; ++i) if (i != p) { const double f = S[i][p]; for (int j = 0; j < 8; ++j) S[i][j] -= f * S[p][j]; }
  }
  double K[4][4], y[4], Pn[4][4];
  for (int i = 0; i < 4; ++i) { y[i] = box[i] - t.m[i]; for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += t.P[i][k] * S[k][4 + j]; K[i][j] = s; } }
  for (int i = 0; i < 4; ++i) { double s = 0; for (int j = 0; j < 4; ++j) s += K[i][j] * y[j]; t.m[i] += s; }
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += K[i][k] * t.P[k][j]; Pn[i][j] = t.P[i][j] - s; }
  memcpy(t.P, Pn, sizeof(Pn));
}
void mc_update_app(const McConst& c, McApp& a, const double* h, double obs) {
  double Ph[kMcD], hm = 0, s = c.r_app;
  for (int i = 0; i < kMcD; ++i) { double t = 0; for (int j = 0; j < kMcD; ++j) t += a.P[i][j] * h[j]; Ph[i] = t; hm += h[i] * a.m[i]; }
  for (int i = 0; i < kMcD; ++i) s += h[i] * Ph[i];
  const double d = obs - hm;
  for (int i = 0; i < kMcD; ++i) a.m[i] += Ph[i] / s * d;
  for (int i = 0; i < kMcD; ++i) for (int j = 0; j < kMcD; ++j) a.P[i][j] -= Ph[i] * Ph[j] / s;
}

int mc_poisson(double lambda, double u) {   // inverse CDF on one uniform
  int k = 0; double p = exp(-lambda), F = p;
  while (u >= F && k < 64) { ++k; p *= lambda / k; F += p; }
  return k;
}

struct OMcam {
  pz_model_desc desc; McConst c;
  int64_t n; uint64_t seed; uint32_t t = 0;
  std::vector<std::vector<McTrack>> tr;
  std::vector<std::vector<McApp>> apps;
  std::vector<int> num_tracks;
  std::vector<float> lw;
  std::vector<int32_t> anc;
  int64_t overflow = 0;
};

double mcam_particle_step(const McConst& c, std::vector<McTrack>& tracks, std::vector<McApp>& apps, int& num_tracks, double t_now,
                          const double* obs, int M, U01& rng, int64_t* overflow) {
  std::vector<McTrack> live;
  for (auto& t : tracks)
    if (rng.next() < c.ps) { for (int i = 0; i < 4; ++i) t.P[i][i] += c.Q[i]; live.push_back(t); }   // F = I
  const int n_new = mc_poisson(c.lb, rng.next());
  for (int i = 0; i < n_new; ++i) {
    const double u_cam = rng.next(), u_new = rng.next(), u_id = rng.next();
    int cam = (int)(u_cam * kMcCams); cam = cam < kMcCams ? cam : kMcCams - 1;
    const int A = (int)apps.size();
    bool fresh = u_new < 1.0 / (A + 1);
    if (fresh && A >= c.amax) { fresh = false; ++*overflow; }
    int identity;
    if (fresh) {
      McApp a; for (int x = 0; x < kMcD; ++x) { a.m[x] = 0; for (int y = 0; y < kMcD; ++y) a.P[x][y] = (x == y); }
      apps.push_back(a); identity = A;
    } else { identity = (int)(u_id * A); identity = identity < A ? identity : A - 1; }
    if ((int)live.size() < c.kmax) {
      McTrack t; t.id = num_tracks + 1 + i; t.identity = identity; t.camera = cam; t.start = t_now;
      for (int x = 0; x < 4; ++x) { t.m[x] = c.prior_m[x]; for (int y = 0; y < 4; ++y) t.P[x][y] = (x == y) ? c.prior_v[x] : 0.0; }
      live.push_back(t);
    } else ++*overflow;
  }
  num_tracks += n_new;
  double lw = 0;
  bool dead = false;
  for (int cam = 0; cam < kMcCams && !dead; ++cam) {
    std::vector<int> idx;
    for (int j = 0; j < M; ++j) if ((int)obs[j * kMcObsW] == cam) idx.push_back(j);
    int nT = 0;
    for (auto& t : live) {
      if (t.camera != cam) continue;
      ++nT;
      if (idx.empty()) { dead = true; break; }   // passert False (:168-169)
      std::vector<double> ll(idx.size());
      double mx = -INFINITY;
      for (size_t q = 0; q < idx.size(); ++q) { ll[q] = mc_score(c, t, apps[t.identity], obs + idx[q] * kMcObsW); mx = std::max(mx, ll[q]); }
      double total = 0; for (double v : ll) total += exp(v - mx);
      const double target = rng.next() * total;
      size_t pick = idx.size() - 1; double cum = 0;
      for (size_t q = 0; q < idx.size(); ++q) { cum += exp(ll[q] - mx); if (cum > target) { pick = q; break; } }
      lw += mx + log(total);   // likelihood of the pick times the proposal correction total / like_pick
      const double* row = obs + idx[pick] * kMcObsW;
      mc_update_box(c, t, row + 11);
      mc_update_app(c, apps[t.identity], row + 1, row[15]);
      idx.erase(idx.begin() + pick);
    }
    if (dead) break;
    const int nC = (int)idx.size();
    lw += -(lgamma(nT + nC + 1.0) - lgamma(nC + 1.0));
    lw += -c.lc + nC * log(c.lc) - lgamma(nC + 1.0);
    for (int j : idx) {
      const double* row = obs + j * kMcObsW;
      double q4 = 0, q10 = 0;
      for (int i = 0; i < 4; ++i) { const double d = row[11 + i] - c.clutter_m[i]; q4 += d * d; }
      for (int i = 0; i < kMcD; ++i) q10 += row[1 + i] * row[1 + i];
      lw += -0.5 * (q4 + 4 * log(2 * M_PI)) - 0.5 * (q10 + kMcD * log(2 * M_PI)) - 0.5 * (row[15] * row[15] + log(2 * M_PI));
    }
  }
  tracks.swap(live);
  return dead ? -INFINITY : lw;
}

}  // namespace

extern "C" {

void* pzo_mcam_create(const pz_model_desc* d, int64_t n, uint64_t seed) {
  OMcam* f = new OMcam();
  f->desc = *d; f->c = mcam_const(d); f->n = n; f->seed = seed;
  f->tr.resize(n); f->apps.resize(n); f->num_tracks.assign(n, 0); f->lw.assign(n, 0.0f); f->anc.assign(n, 0);
  return f;
}
void pzo_mcam_destroy(void* h) { delete (OMcam*)h; }

// one filter step (obs: n_obs rows of 16); outputs as pzo_mtt_step
int pzo_mcam_step(void* h, const double* obs, int32_t n_obs, double* mean_cnt, double* var_cnt, double* logz) {
  OMcam* f = (OMcam*)h;
  ResamplePlan pl;
  if (!build_plan(f->lw.data(), f->n, f->n, &pl)) return PZ_EDEGENERATE;
  systematic_ancestors(pl, pzo_resample_uniform(f->seed, f->t), f->n, f->anc.data());
  std::vector<std::vector<McTrack>> ntr(f->n);
  std::vector<std::vector<McApp>> nap(f->n);
  std::vector<int> nnt(f->n);
  std::vector<double> lwd(f->n);
  for (int64_t j = 0; j < f->n; ++j) {
    ntr[j] = f->tr[f->anc[j]]; nap[j] = f->apps[f->anc[j]]; nnt[j] = f->num_tracks[f->anc[j]];
    U01 rng(f->seed, (uint32_t)j, f->t, kStreamMcam);
    lwd[j] = mcam_particle_step(f->c, ntr[j], nap[j], nnt[j], (double)f->t, obs, n_obs, rng, &f->overflow);
    f->lw[j] = (float)lwd[j];
  }
  f->tr.swap(ntr); f->apps.swap(nap); f->num_tracks.swap(nnt);
  double mx = -INFINITY; for (double v : lwd) mx = std::max(mx, v);
  if (!(mx > -INFINITY)) { f->t += 1; return PZ_EDEGENERATE; }
  double sw = 0, sc = 0, scc = 0;
  for (int64_t j = 0; j < f->n; ++j) { double w = exp(lwd[j] - mx), c = (double)f->tr[j].size(); sw += w; sc += w * c; scc += w * c * c; }
  if (mean_cnt) *mean_cnt = sc / sw;
  if (var_cnt) *var_cnt = scc / sw - (sc / sw) * (sc / sw);
  if (logz) *logz = mx + log(sw / f->n);
  f->t += 1;
  return 0;
}

void pzo_mcam_logw(void* h, float* out) { OMcam* f = (OMcam*)h; memcpy(out, f->lw.data(), f->n * sizeof(float)); }

// tracks of particle i: returns the count; ids / identity / camera [k], mean[k*4..], cov[k*16..]; *n_app = its appearance count
int pzo_mcam_particle(void* h, int64_t i, int32_t* ids, int32_t* identity, int32_t* camera, double* mean, double* cov, int32_t* n_app) {
  OMcam* f = (OMcam*)h;
  const auto& v = f->tr[i];
  for (size_t k = 0; k < v.size(); ++k) {
    ids[k] = v[k].id; identity[k] = v[k].identity; camera[k] = v[k].camera;
    for (int a = 0; a < 4; ++a) { mean[k * 4 + a] = v[k].m[a]; for (int b = 0; b < 4; ++b) cov[k * 16 + a * 4 + b] = v[k].P[a][b]; }
  }
  if (n_app) *n_app = (int32_t)f->apps[i].size();
  return (int)v.size();
}
// appearance a of particle i: mean[10], cov[100]
int pzo_mcam_appearance(void* h, int64_t i, int32_t a, double* mean, double* cov) {
  OMcam* f = (OMcam*)h;
  if (a < 0 || a >= (int)f->apps[i].size()) return -1;
  const McApp& ap = f->apps[i][a];
  for (int x = 0; x < kMcD; ++x) { mean[x] = ap.m[x]; for (int y = 0; y < kMcD; ++y) cov[x * kMcD + y] = ap.P[x][y]; }
  return 0;
}

// Ground truth + observations: zstepGen under MP.sim (generateGroundTruth, MultiCamera.hs:197-219).
// obs_out [T][m_max][16]; n_obs_out[T]; truth_out [T][k_max][7] = (id, identity, camera, box[4]); n_truth_out[T].
// (The reference shuffles each camera's list; the filter's result does not depend on the order only in
// distribution, so the shuffle is kept.)
int pzo_mcam_generate(const pz_model_desc* d, uint64_t seed, int64_t T, int32_t m_max, double* obs_out, int32_t* n_obs_out,
                      double* truth_out, int32_t* n_truth_out) {
  const McConst c = mcam_const(d);
  struct Tr { int id, identity, camera; double x[4]; };
  std::vector<Tr> tracks; std::vector<std::vector<double>> apps; int num_tracks = 0;
  for (int64_t t = 0; t < T; ++t) {
    U01 rng(seed, 0, (uint32_t)t, kStreamMcamGen);
    std::vector<Tr> live;
    for (auto& tr : tracks) if (rng.next() < c.ps) { for (int i = 0; i < 4; ++i) tr.x[i] += sqrt(c.Q[i]) * rng.normal(); live.push_back(tr); }
    const int n_new = mc_poisson(c.lb, rng.next());
    for (int i = 0; i < n_new; ++i) {
      Tr nt; nt.id = num_tracks + 1 + i;
      int cam = (int)(rng.next() * kMcCams); nt.camera = cam < kMcCams ? cam : kMcCams - 1;
      const int A = (int)apps.size();
      if (rng.next() < 1.0 / (A + 1) && A < c.amax) { std::vector<double> a(kMcD); for (auto& v : a) v = rng.normal(); apps.push_back(a); nt.identity = A; }
      else { int id = (int)(rng.next() * A); nt.identity = id < A ? id : A - 1; }
      for (int x = 0; x < 4; ++x) nt.x[x] = c.prior_m[x] + sqrt(c.prior_v[x]) * rng.normal();
      if ((int)live.size() < c.kmax) live.push_back(nt);
    }
    num_tracks += n_new;
    tracks.swap(live);
    int M = 0;
    for (int cam = 0; cam < kMcCams; ++cam) {
      std::vector<std::vector<double>> rows;
      for (auto& tr : tracks) {
        if (tr.camera != cam) continue;
        std::vector<double> row(kMcObsW); row[0] = cam;
        double nrm = 0; for (int i = 0; i < kMcD; ++i) { row[1 + i] = rng.normal(); nrm += row[1 + i] * row[1 + i]; }
        nrm = sqrt(nrm); double hm = 0;
        for (int i = 0; i < kMcD; ++i) { row[1 + i] /= nrm; hm += row[1 + i] * apps[tr.identity][i]; }
        for (int i = 0; i < 4; ++i) row[11 + i] = tr.x[i] + sqrt(c.r_box) * rng.normal();
        row[15] = hm + sqrt(c.r_app) * rng.normal();
        rows.push_back(row);
      }
      const int nC = mc_poisson(c.lc, rng.next());
      for (int q = 0; q < nC; ++q) {
        std::vector<double> row(kMcObsW); row[0] = cam;
        for (int i = 0; i < kMcD; ++i) row[1 + i] = rng.normal();
        for (int i = 0; i < 4; ++i) row[11 + i] = c.clutter_m[i] + rng.normal();
        row[15] = rng.normal();
        rows.push_back(row);
      }
      for (size_t i = rows.size(); i > 1; --i) { size_t j = (size_t)(rng.next() * i); std::swap(rows[i - 1], rows[j < i ? j : i - 1]); }
      for (auto& row : rows) { if (M >= m_max) break; memcpy(obs_out + (t * m_max + M) * kMcObsW, row.data(), kMcObsW * sizeof(double)); ++M; }
    }
    n_obs_out[t] = M;
    n_truth_out[t] = (int)tracks.size();
    for (size_t k = 0; k < tracks.size(); ++k) {
      double* o = truth_out + (t * c.kmax + k) * 7;
      o[0] = tracks[k].id; o[1] = tracks[k].identity; o[2] = tracks[k].camera;
      for (int a = 0; a < 4; ++a) o[3 + a] = tracks[k].x[a];
    }
  }
  return 0;
}

}  // extern "C"
