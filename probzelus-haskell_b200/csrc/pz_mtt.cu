// pz_mtt.cu -- the multi-target tracker step (Examples/MultiTargetTracking.hs:26-145) on sm_100a.
//
// One particle = up to 16 Rao-Blackwellised tracks (Gaussian marginal over R^4), the heap of
// zdsparticles (Inference.hs:110-114) reduced to closed-form Kalman recursions:
//   trackMotion (:56-67) -> predict; trackMeasurement (:69-73) -> marginal score / update
//   (MVDistributions.hs:75-100, DelayedSampling.hs:135-152); proposeAssocs (:111-124) ->
//   sequential categorical over {Clutter < NewTrack < Track id ascending};
//   observingWithProposal (Metaprob.hs:157-162) + sampleStep (:89-99) -> log-weight
//       sum_j log L_j + n_undetected log(1 - ps pd)   [+ constants added on the host]
//   (the algebraic reduction of -log q + log prior + log likelihood; the CPU oracle accumulates
//   the reference's terms one by one and checks the reduction on every call);
//   updateWithAssocs (:75-87) -> Kalman update, coasting Bernoulli((1-pd)/(1-ps pd)), births.
//
// Mapping: half a warp per particle, one lane per track slot.  The 16 candidate likelihoods of
// an observation are evaluated in parallel, the categorical draw is a 16-lane inclusive scan +
// ballot, survivors are compacted with a ballot / popc.  A particle is a 1040-byte record
// (16-byte header + 16 x 64-byte track records), read through the stored ancestor index and
// written contiguously, so both directions are coalesced.
//
// Unlike the linear-Gaussian kernels this translation unit is NOT bound by the bit-exact fp32
// spec (parity for MTT is statistical): it uses fast intrinsics and fmad.
#include "pz_kernels.cuh"
#include "pz_math.cuh"

namespace pz {

constexpr uint32_t kStreamMtt = 0x4D545453u;  // "MTTS"

struct MttDev {
  float F[4][4], Q[4][4];
  float lc, lb, pspd, p_coast, log1m_pspd;
  float cv, bp, bv, r;
  float cl_norm, nb_norm;  // log normalisers of the clutter / new-track predictive densities
};

__device__ __forceinline__ float u01(uint32_t w) { return __fmaf_rn(__uint2float_rn(w), 0x1p-32f, 0x1p-33f); }

struct Track {
  int id;
  float m[4];
  float c[10];  // symmetric 4x4, upper triangle row-major: 00 01 02 03 11 12 13 22 23 33
};
__device__ __forceinline__ int tri(int i, int j) {  // index of (i, j), i <= j
  return i * 4 - (i * (i - 1)) / 2 + (j - i);
}
__device__ __forceinline__ float sym(const float* c, int i, int j) { return i <= j ? c[tri(i, j)] : c[tri(j, i)]; }

__device__ __forceinline__ void predict(const MttDev& p, Track& t) {
  float m[4], FP[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = p.F[i][0] * t.m[0] + p.F[i][1] * t.m[1] + p.F[i][2] * t.m[2] + p.F[i][3] * t.m[3];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      FP[i][j] = p.F[i][0] * sym(t.c, 0, j) + p.F[i][1] * sym(t.c, 1, j) + p.F[i][2] * sym(t.c, 2, j) + p.F[i][3] * sym(t.c, 3, j);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    t.m[i] = m[i];
#pragma unroll
    for (int j = i; j < 4; ++j)
      t.c[tri(i, j)] = FP[i][0] * p.F[j][0] + FP[i][1] * p.F[j][1] + FP[i][2] * p.F[j][2] + FP[i][3] * p.F[j][3] + p.Q[i][j];
  }
}

// kalmanUpdate with H = [I2 0], R = r I2 (MVDistributions.hs:83-100)
__device__ __forceinline__ void update(Track& t, float z0, float z1, float r) {
  const float a = t.c[0] + r, b = t.c[1], d = t.c[4] + r;
  const float idet = 1.0f / (a * d - b * b);
  const float s00 = d * idet, s01 = -b * idet, s11 = a * idet;
  const float y0 = z0 - t.m[0], y1 = z1 - t.m[1];
  float K[4][2], P0[4], P1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { P0[i] = sym(t.c, i, 0); P1[i] = sym(t.c, i, 1); }
#pragma unroll
  for (int i = 0; i < 4; ++i) { K[i][0] = P0[i] * s00 + P1[i] * s01; K[i][1] = P0[i] * s01 + P1[i] * s11; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    t.m[i] += K[i][0] * y0 + K[i][1] * y1;
#pragma unroll
    for (int j = i; j < 4; ++j) t.c[tri(i, j)] -= K[i][0] * P0[j] + K[i][1] * P1[j];
  }
}

constexpr int kMttWords = 4 + 16 * 16;  // 32-bit words per particle record

struct MttArgs {
  const uint32_t* sin;   // [N][kMttWords]
  uint32_t* sout;
  const int32_t* anc;    // [N]
  float* lw;             // [N]
  const float* obs;      // [m_max][2]
  const CtlDev* ctl;
  unsigned long long* overflow;  // counter of dropped births
  int64_t n;
  int32_t n_obs;
  uint64_t seed;
  MttDev p;
};

__global__ void __launch_bounds__(kThreads) mtt_step(const __grid_constant__ MttArgs a) {
  __shared__ float s_z[32][2], s_lc[32], s_ln[32];
  const int tid = threadIdx.x;
  const int M = a.n_obs;
  if (tid < M) {
    const float z0 = a.obs[2 * tid], z1 = a.obs[2 * tid + 1];
    s_z[tid][0] = z0; s_z[tid][1] = z1;
    const float r2 = z0 * z0 + z1 * z1;
    s_lc[tid] = a.p.lc * __expf(a.p.cl_norm - 0.5f * r2 / a.p.cv);
    s_ln[tid] = a.p.lb * __expf(a.p.nb_norm - 0.5f * r2 / (a.p.bp + a.p.r));
  }
  __syncthreads();
  const int64_t jraw = (int64_t)blockIdx.x * (kThreads / 16) + (tid >> 4);
  const bool valid = jraw < a.n;          // an odd tail keeps its partner half-warp alive for the warp-wide shuffles
  const int64_t j = valid ? jraw : a.n - 1;
  const int k = tid & 15;                       // track slot of this lane
  const int hsh = (tid & 16);                   // bit offset of this half in a warp ballot
  const uint32_t hmask = 0xFFFFu << hsh;        // lanes of this half-warp
  const uint32_t t = a.ctl->t;
  const int64_t par = a.anc[j];
  const uint32_t* __restrict__ src = a.sin + (size_t)par * kMttWords;
  uint32_t* __restrict__ dst = a.sout + (size_t)j * kMttWords;
  const uint4 hdr = *reinterpret_cast<const uint4*>(src);
  const int cnt = (int)hdr.x;
  int next_id = (int)hdr.y;
  const bool live = k < cnt;

  Track tr;
  tr.id = 0;
  if (live) {
    const uint4* rp = reinterpret_cast<const uint4*>(src + 4 + 16 * k);
    const uint4 w0 = rp[0], w1 = rp[1], w2 = rp[2], w3 = rp[3];
    tr.id = (int)w0.x;
    tr.m[0] = __uint_as_float(w0.y); tr.m[1] = __uint_as_float(w0.z); tr.m[2] = __uint_as_float(w0.w);
    tr.m[3] = __uint_as_float(w1.x);
    tr.c[0] = __uint_as_float(w1.y); tr.c[1] = __uint_as_float(w1.z); tr.c[2] = __uint_as_float(w1.w);
    tr.c[3] = __uint_as_float(w2.x); tr.c[4] = __uint_as_float(w2.y); tr.c[5] = __uint_as_float(w2.z);
    tr.c[6] = __uint_as_float(w2.w); tr.c[7] = __uint_as_float(w3.x); tr.c[8] = __uint_as_float(w3.y);
    tr.c[9] = __uint_as_float(w3.z);
    predict(a.p, tr);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) tr.m[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) tr.c[i] = 0.f;
  }
  // predictive density of a measurement of this track: N(z; H m, H P H^T + r I)
  float s00 = 0.f, s01 = 0.f, s11 = 0.f, lnorm = 0.f;
  if (live) {
    const float sa = tr.c[0] + a.p.r, sb = tr.c[1], sd = tr.c[4] + a.p.r;
    const float det = sa * sd - sb * sb, idet = 1.0f / det;
    s00 = sd * idet; s01 = -sb * idet; s11 = sa * idet;
    lnorm = -0.5f * __logf(det) - 1.8378770664093453f;
  }
  // uniforms: lane l of the half draws Philox call l -> words 4l .. 4l+3 (M association draws, then 16 coasting draws)
  const U4 rnd = philox4x32_10((uint32_t)j, (uint32_t)k, t, kStreamMtt, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  auto uniform = [&](int idx) -> float {  // idx < 64
    const int srcl = (idx >> 2) + hsh;
    const uint32_t x = __shfl_sync(0xffffffffu, rnd.x, srcl), y = __shfl_sync(0xffffffffu, rnd.y, srcl);
    const uint32_t z = __shfl_sync(0xffffffffu, rnd.z, srcl), w = __shfl_sync(0xffffffffu, rnd.w, srcl);
    const int q = idx & 3;
    return u01(q == 0 ? x : (q == 1 ? y : (q == 2 ? z : w)));
  };

  bool assigned = false;
  int myobs = -1;
  uint32_t newmask = 0;  // observations that start a new track
  float lw = 0.f;
  for (int o = 0; o < M; ++o) {
    const float z0 = s_z[o][0], z1 = s_z[o][1];
    float like = 0.f;
    if (live && !assigned) {
      const float d0 = z0 - tr.m[0], d1 = z1 - tr.m[1];
      const float quad = s00 * d0 * d0 + 2.0f * s01 * d0 * d1 + s11 * d1 * d1;
      like = a.p.pspd * __expf(lnorm - 0.5f * quad);
    }
    float incl = like;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, d, 16);
      if (k >= d) incl += up;
    }
    const float ttot = __shfl_sync(0xffffffffu, incl, 15, 16);
    const float lc = s_lc[o], ln = s_ln[o];
    const float tot = lc + ln + ttot;
    lw += __logf(tot);
    const float u = uniform(o) * tot;
    const float target = u - lc - ln;
    // warp-wide votes are taken unconditionally: the two half-warps (two particles) may pick differently
    uint32_t hit = (__ballot_sync(0xffffffffu, like > 0.f && incl > target) & hmask) >> hsh;
    const uint32_t cand = (__ballot_sync(0xffffffffu, like > 0.f) & hmask) >> hsh;
    if (u < lc) {
      // clutter
    } else if (u < lc + ln) {
      newmask |= 1u << o;
    } else {
      if (!hit && cand) hit = 1u << (31 - __clz(cand));  // round-off past the last candidate: take the last one
      if (!hit) newmask |= 1u << o;                      // no candidate track at all: start a track
      else if (k == __ffs(hit) - 1) { assigned = true; myobs = o; }
    }
  }
  const uint32_t undet = (__ballot_sync(0xffffffffu, live && !assigned) & hmask) >> hsh;
  lw += (float)__popc(undet) * a.p.log1m_pspd;

  // update / coast
  bool survive = false;
  if (live) {
    if (assigned) { update(tr, s_z[myobs][0], s_z[myobs][1], a.p.r); survive = true; }
  }
  {
    const float uc = uniform(32 + k);  // all lanes take part in the shuffles
    if (live && !assigned) survive = uc < a.p.p_coast;
  }
  const uint32_t smask = (__ballot_sync(0xffffffffu, survive) & hmask) >> hsh;
  const int n_surv = __popc(smask);
  auto store_track = [&](int slot, const Track& x) {
    uint4* wp = reinterpret_cast<uint4*>(dst + 4 + 16 * slot);
    wp[0] = make_uint4((uint32_t)x.id, __float_as_uint(x.m[0]), __float_as_uint(x.m[1]), __float_as_uint(x.m[2]));
    wp[1] = make_uint4(__float_as_uint(x.m[3]), __float_as_uint(x.c[0]), __float_as_uint(x.c[1]), __float_as_uint(x.c[2]));
    wp[2] = make_uint4(__float_as_uint(x.c[3]), __float_as_uint(x.c[4]), __float_as_uint(x.c[5]), __float_as_uint(x.c[6]));
    wp[3] = make_uint4(__float_as_uint(x.c[7]), __float_as_uint(x.c[8]), __float_as_uint(x.c[9]), 0u);
  };
  if (survive && valid) store_track(__popc(smask & ((1u << k) - 1u)), tr);
  // births: lane i takes the i-th new-track observation (newTrackD conditioned on it)
  const int n_new = __popc(newmask);
  const int room = 16 - n_surv;
  const int n_born = n_new < room ? n_new : room;
  if (k < n_born && valid) {
    const int o = __fns(newmask, 0, k + 1);
    Track nt;
    nt.id = next_id + k;
#pragma unroll
    for (int i = 0; i < 4; ++i) nt.m[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) nt.c[i] = 0.f;
    nt.c[0] = a.p.bp; nt.c[4] = a.p.bp; nt.c[7] = a.p.bv; nt.c[9] = a.p.bv;
    update(nt, s_z[o][0], s_z[o][1], a.p.r);
    store_track(n_surv + k, nt);
  }
  if (k == 0 && valid) {
    *reinterpret_cast<uint4*>(dst) = make_uint4((uint32_t)(n_surv + n_born), (uint32_t)(next_id + n_new), n_new > room ? 1u : 0u, 0u);
    a.lw[j] = lw;
    if (n_new > room) atomicAdd(a.overflow, (unsigned long long)(n_new - room));
  }
}

// weighted track-count moments: sum w, sum w^2, sum w c, sum w c^2 at scale 2^emax
__global__ void __launch_bounds__(kThreads) mtt_summary_reduce(const float* __restrict__ lw, const uint32_t* __restrict__ s,
                                                               int64_t n, PlanDev* plan) {
  __shared__ double s_red[kWarps][4];
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[4] = {0, 0, 0, 0};
  if (i < n) {
    const float l = sanitize_lw(lw[i]);
    if (l > -INFINITY) {
      const double w = exp2((double)l * 1.4426950408889634 - (double)plan->emax);  // scale 2^emax
      const double c = (double)s[(size_t)i * kMttWords];
      v[0] = w; v[1] = w * w; v[2] = w * c; v[3] = w * c * c;
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double x = v[r];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const long long y = __double_as_longlong(x);
      const int lo = __shfl_xor_sync(0xffffffffu, (int)y, d), hi = __shfl_xor_sync(0xffffffffu, (int)(y >> 32), d);
      x += __longlong_as_double(((long long)hi << 32) | (unsigned int)lo);
    }
    if (lane == 0) s_red[warp][r] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double x = 0;
    for (int w = 0; w < kWarps; ++w) x += s_red[w][threadIdx.x];
    double* dstp = threadIdx.x == 0 ? &plan->sw : (threadIdx.x == 1 ? &plan->sw2 : (threadIdx.x == 2 ? &plan->sx[0] : &plan->sxx[0]));
    atomicAdd(dstp, x);
  }
}

__global__ void mtt_summary_final(PlanDev* pl, const CtlDev* ctl, pz_step_summary* summ, int64_t n, double ll_const,
                                  const unsigned long long* overflow) {
  pz_step_summary* o = summ + (ctl->t % kObsRing);
  const double sw = pl->sw, m = pl->sx[0] / sw;
  for (int d = 0; d < PZ_MAX_NX; ++d) { o->mean[d] = 0; o->var[d] = 0; }
  o->mean[0] = m;
  o->var[0] = pl->sxx[0] / sw - m * m;
  o->step = ctl->t;
  o->log_evidence_inc = log(sw) + (double)pl->emax * 0.6931471805599453 - log((double)n) + ll_const;
  o->ess = sw * sw / pl->sw2;
  o->n_migrated = 0;
  o->e_max = pl->emax;
  o->reserved = (int32_t)(*overflow > 0x7fffffffull ? 0x7fffffff : *overflow);
  o->total_weight = pl->total;
  pl->sw = 0; pl->sw2 = 0; pl->sx[0] = 0; pl->sxx[0] = 0;
}

// thinned read of the resampled population: every stride-th output slot's parent record
__global__ void mtt_gather_thin(const uint32_t* __restrict__ s, const int32_t* __restrict__ anc, int32_t stride, int64_t n_sel,
                                uint32_t* __restrict__ out) {
  const int64_t sel = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (sel >= n_sel) return;
  const int64_t p = anc[sel * stride];
  for (int w = threadIdx.x & 31; w < kMttWords; w += 32) out[(size_t)sel * kMttWords + w] = s[(size_t)p * kMttWords + w];
}

// ---- host side ---------------------------------------------------------------------------------
struct MttLaunch {
  MttArgs args;
};

int mtt_words_per_particle() { return kMttWords; }

void mtt_fill_params(const pz_model_desc* d, void* out_mttdev) {
  MttDev& p = *reinterpret_cast<MttDev*>(out_mttdev);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { p.F[i][j] = (float)d->F[i * 4 + j]; p.Q[i][j] = (float)d->Q[i * 4 + j]; }
  const double pspd = d->mtt.survival_prob * d->mtt.pd;
  p.lc = (float)d->mtt.clutter_lambda; p.lb = (float)d->mtt.new_track_lambda; p.pspd = (float)pspd;
  p.p_coast = (float)((1.0 - d->mtt.pd) / (1.0 - pspd));
  p.log1m_pspd = (float)log(1.0 - pspd);
  p.cv = (float)d->mtt.clutter_var; p.bp = (float)d->mtt.birth_pos_var; p.bv = (float)d->mtt.birth_vel_var; p.r = (float)d->mtt.meas_var;
  p.cl_norm = (float)(-log(d->mtt.clutter_var) - log(2 * M_PI));
  p.nb_norm = (float)(-log(d->mtt.birth_pos_var + d->mtt.meas_var) - log(2 * M_PI));
}

size_t mtt_params_size() { return sizeof(MttDev); }

int launch_mtt_step(const void* params, const uint32_t* sin, uint32_t* sout, const int32_t* anc, float* lw, const float* obs,
                    const CtlDev* ctl, unsigned long long* overflow, int64_t n, int32_t n_obs, uint64_t seed, cudaStream_t s) {
  MttArgs a;
  a.sin = sin; a.sout = sout; a.anc = anc; a.lw = lw; a.obs = obs; a.ctl = ctl; a.overflow = overflow; a.n = n; a.n_obs = n_obs;
  a.seed = seed; a.p = *reinterpret_cast<const MttDev*>(params);
  const int per_cta = kThreads / 16;
  mtt_step<<<(unsigned)((n + per_cta - 1) / per_cta), kThreads, 0, s>>>(a);
  return 1;
}

int launch_mtt_summary(const float* lw, const uint32_t* sdata, int64_t n, PlanDev* plan, const CtlDev* ctl, pz_step_summary* summ,
                       double ll_const, const unsigned long long* overflow, cudaStream_t s) {
  mtt_summary_reduce<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, s>>>(lw, sdata, n, plan);
  mtt_summary_final<<<1, 1, 0, s>>>(plan, ctl, summ, n, ll_const, overflow);
  return 2;
}

int launch_mtt_gather(const uint32_t* sdata, const int32_t* anc, int32_t stride, int64_t n_sel, uint32_t* out, cudaStream_t s) {
  const int per_cta = kThreads / 32;
  mtt_gather_thin<<<(unsigned)((n_sel + per_cta - 1) / per_cta), kThreads, 0, s>>>(sdata, anc, stride, n_sel, out);
  return 1;
}

}  // namespace pz
