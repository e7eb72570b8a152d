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
// Mapping: one thread per particle (see mtt_step).  A particle is a 1040-byte record (16-byte
// header + 16 x 64-byte track records) read through the stored ancestor index; only the live
// track records move (every 64-byte record is two full 32-byte sectors).  Array-of-structures on purpose:
// the gather is by ANCESTOR, and the 32 parents of a warp are whichever particles survived, so a layout that
// interleaves neighbouring particles fetches lines it uses half of.  Measured at N = 10^6, 8 targets: this
// layout 0.42 ms per step; S[field][particle] with scalar loads 0.53 ms; S[group of 32][field][lane] 0.51 ms;
// S[group of 32][16-byte quad][lane] with 128-bit loads 0.45 ms.
//
// Unlike the linear-Gaussian kernels this translation unit is NOT bound by the bit-exact fp32
// spec (parity for MTT is statistical): it uses fast intrinsics and fmad.
#include <atomic>
#include <cstdlib>

#include "pz_kernels.cuh"
#include "pz_math.cuh"

namespace pz {

constexpr uint32_t kStreamMtt = 0x4D545453u;  // "MTTS"

struct MttDev {
  float F[4][4], Q[4][4];
  int kron;             // F = [[a, b], [c, d]] (x) I2 and diagonal Q (the reference's constant-velocity model)
  float fa, fb, fc, fd;
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

// position/velocity structured predict: with P = [[A, B], [B^T, D]] (2x2 blocks; c[] holds A = {0,1,4},
// B = {2,3,5,6}, D = {7,8,9}),  F P F^T = [[a^2 A + ab (B + B^T) + b^2 D, ac A + ad B + bc B^T + bd D], [., c^2 A + cd (B + B^T) + d^2 D]]
__device__ __forceinline__ void predict_kron(const MttDev& p, Track& t) {
  const float a = p.fa, b = p.fb, c = p.fc, d = p.fd;
  const float m0 = a * t.m[0] + b * t.m[2], m1 = a * t.m[1] + b * t.m[3];
  const float m2 = c * t.m[0] + d * t.m[2], m3 = c * t.m[1] + d * t.m[3];
  t.m[0] = m0; t.m[1] = m1; t.m[2] = m2; t.m[3] = m3;
  const float A00 = t.c[0], A01 = t.c[1], A11 = t.c[4];
  const float B00 = t.c[2], B01 = t.c[3], B10 = t.c[5], B11 = t.c[6];
  const float D00 = t.c[7], D01 = t.c[8], D11 = t.c[9];
  const float aa = a * a, ab = a * b, bb = b * b, ac = a * c, ad = a * d, bc = b * c, bd = b * d, cc = c * c, cd = c * d, dd = d * d;
  t.c[0] = aa * A00 + ab * (B00 + B00) + bb * D00 + p.Q[0][0];
  t.c[1] = aa * A01 + ab * (B01 + B10) + bb * D01;
  t.c[4] = aa * A11 + ab * (B11 + B11) + bb * D11 + p.Q[1][1];
  t.c[2] = ac * A00 + ad * B00 + bc * B00 + bd * D00;
  t.c[3] = ac * A01 + ad * B01 + bc * B10 + bd * D01;
  t.c[5] = ac * A01 + ad * B10 + bc * B01 + bd * D01;
  t.c[6] = ac * A11 + ad * B11 + bc * B11 + bd * D11;
  t.c[7] = cc * A00 + cd * (B00 + B00) + dd * D00 + p.Q[2][2];
  t.c[8] = cc * A01 + cd * (B01 + B10) + dd * D01;
  t.c[9] = cc * A11 + cd * (B11 + B11) + dd * D11 + p.Q[3][3];
}

__device__ __forceinline__ void predict(const MttDev& p, Track& t) {
  if (p.kron) { predict_kron(p, t); return; }
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
  double* part;          // [gridDim.x][5]: per-CTA (max lw, sum w, sum w^2, sum w c, sum w c^2), w = exp(lw - max)
  int64_t n;             // end of this launch's particle range (global index)
  int64_t j0;            // its start; sout and part are indexed relative to it
  int64_t n_per;         // multi-device: particles per device (parents are read from their owners); 0: single device
  const uint32_t* sin_all[kMttMaxDevices];
  int32_t n_obs;
  uint64_t seed;
  MttDev p;
};

__device__ __forceinline__ void load_track(const uint32_t* __restrict__ rec, Track& tr) {
  const uint4* rp = reinterpret_cast<const uint4*>(rec);
  const uint4 w0 = rp[0], w1 = rp[1], w2 = rp[2], w3 = rp[3];
  tr.id = (int)w0.x;
  tr.m[0] = __uint_as_float(w0.y); tr.m[1] = __uint_as_float(w0.z); tr.m[2] = __uint_as_float(w0.w);
  tr.m[3] = __uint_as_float(w1.x);
  tr.c[0] = __uint_as_float(w1.y); tr.c[1] = __uint_as_float(w1.z); tr.c[2] = __uint_as_float(w1.w);
  tr.c[3] = __uint_as_float(w2.x); tr.c[4] = __uint_as_float(w2.y); tr.c[5] = __uint_as_float(w2.z);
  tr.c[6] = __uint_as_float(w2.w); tr.c[7] = __uint_as_float(w3.x); tr.c[8] = __uint_as_float(w3.y);
  tr.c[9] = __uint_as_float(w3.z);
}
__device__ __forceinline__ void store_track(uint32_t* __restrict__ rec, const Track& x) {
  uint4* wp = reinterpret_cast<uint4*>(rec);
  wp[0] = make_uint4((uint32_t)x.id, __float_as_uint(x.m[0]), __float_as_uint(x.m[1]), __float_as_uint(x.m[2]));
  wp[1] = make_uint4(__float_as_uint(x.m[3]), __float_as_uint(x.c[0]), __float_as_uint(x.c[1]), __float_as_uint(x.c[2]));
  wp[2] = make_uint4(__float_as_uint(x.c[3]), __float_as_uint(x.c[4]), __float_as_uint(x.c[5]), __float_as_uint(x.c[6]));
  wp[3] = make_uint4(__float_as_uint(x.c[7]), __float_as_uint(x.c[8]), __float_as_uint(x.c[9]), 0u);
}

// One THREAD per particle (the track count of a particle is small and nearly equal across the
// particles of a warp, so the loops over tracks stay converged; a lane-per-track-slot mapping leaves
// most lanes idle when few targets are alive).  Three passes over the particle's tracks:
//   A  load + Kalman predict -> the 6 numbers per track the association needs (predicted position,
//      inverse innovation covariance, log normaliser), kept in shared memory [track][thread];
//   B  the sequential association proposal over the M detections (proposeAssocs, :111-124), weight
//      sum_j log L_j + n_undetected log(1 - ps pd) (observingWithProposal, Metaprob.hs:157-162);
//   C  load + predict again (the parent record is an L2 hit), Kalman update of detected tracks,
//      coasting Bernoulli for the others, survivors compacted in id order, then births.
// HBM traffic is the parent record read once and the child record written once.
constexpr int kMttThreads = 128;
constexpr int kMttDerived = 7;   // m0, m1, A, B, C, L (likelihood = 2^(L + A d0^2 + B d0 d1 + C d1^2)), likelihood scratch
constexpr int kMttFast = 8;      // track slots whose derived data live in shared memory; the rest (rare) in local memory

// derived-data accessors: slots [0, 8) in shared memory [field][slot pair][thread] as float2 (two slots ride in one packed
// FFMA2 / FADD2 / FMUL2 and one 64-bit shared-memory access), slots [8, 16) in a per-thread array
struct DerShared {
  float* base; int tid;
  __device__ __forceinline__ float& operator()(int f, int k) const {
    return base[(((f * (kMttFast / 2) + (k >> 1)) * kMttThreads + tid) << 1) + (k & 1)];
  }
  __device__ __forceinline__ float2& pair(int f, int kp) const {
    return reinterpret_cast<float2*>(base)[(f * (kMttFast / 2) + kp) * kMttThreads + tid];
  }
};
struct DerLocal {
  float* loc;
  __device__ __forceinline__ float& operator()(int f, int k) const { return loc[f * (16 - kMttFast) + (k - kMttFast)]; }
};

// (Parking the predicted track in slot k of the child's record so that pass C need not load and predict the parent's track a
// second time was measured: 0.649 vs 0.465 ms per step -- the extra 64-byte stores and L2 read-backs cost far more than the
// ~50 FMAs of the second predict.)
template <class Der>
__device__ __forceinline__ void mtt_derive(const MttDev& p, const uint32_t* __restrict__ src, int k, const Der& der) {
  Track tr;
  load_track(src + 4 + 16 * k, tr);
  predict(p, tr);
  const float sa = tr.c[0] + p.r, sb = tr.c[1], sd = tr.c[4] + p.r;
  const float det = sa * sd - sb * sb, idet = 1.0f / det;
  // p_s p_d N(z; H m, S) = 2^(L + A d0^2 + B d0 d1 + C d1^2), d = z - H m: every constant folded here, once per track
  constexpr float kL2e = 1.4426950408889634f;
  der(0, k) = tr.m[0]; der(1, k) = tr.m[1];
  der(2, k) = -0.5f * kL2e * sd * idet; der(3, k) = kL2e * sb * idet; der(4, k) = -0.5f * kL2e * sa * idet;
  der(5, k) = kL2e * (-0.5f * __logf(det) - 1.8378770664093453f + __logf(p.pspd));
}
template <class Der>
__device__ __forceinline__ float mtt_like(const MttDev& p, float z0, float z1, uint32_t assigned, int k, const Der& der) {
  float like = 0.f;
  if (!((assigned >> k) & 1u)) {
    const float d0 = z0 - der(0, k), d1 = z1 - der(1, k);
    like = exp2f(fmaf(der(4, k) * d1, d1, fmaf(fmaf(der(2, k), d0, der(3, k) * d1), d0, der(5, k))));
  }
  der(6, k) = like;
  return like;
}
// two shared-memory slots (2 kp, 2 kp + 1) at once; slot 2 kp + 1 may be past the track count
__device__ __forceinline__ float mtt_like_pair(float z0, float z1, uint32_t assigned, int kp, int cnt, const DerShared& der) {
  const float2 m0 = der.pair(0, kp), m1 = der.pair(1, kp), A = der.pair(2, kp), B = der.pair(3, kp), C = der.pair(4, kp), L = der.pair(5, kp);
  const float2 d0 = __fadd2_rn(make_float2(z0, z0), make_float2(-m0.x, -m0.y));
  const float2 d1 = __fadd2_rn(make_float2(z1, z1), make_float2(-m1.x, -m1.y));
  const float2 t = __ffma2_rn(A, d0, __fmul2_rn(B, d1));
  const float2 q = __ffma2_rn(__fmul2_rn(C, d1), d1, __ffma2_rn(t, d0, L));
  const int k = 2 * kp;
  float2 like;
  like.x = ((assigned >> k) & 1u) ? 0.f : exp2f(q.x);
  like.y = (((assigned >> (k + 1)) & 1u) || k + 1 >= cnt) ? 0.f : exp2f(q.y);
  der.pair(6, kp) = like;
  return like.x + like.y;
}
template <class Der>
__device__ __forceinline__ void mtt_pick(float target, int k, const Der& der, float& cum, int& pick, int& last) {
  const float like = der(6, k);
  if (like > 0.f) {
    last = k;
    cum += like;
    if (pick < 0 && cum > target) pick = k;
  }
}

// per-CTA partial of the weighted track-count moments (no separate pass over the particles): every thread brings its
// particle's log-weight and track count (-inf / 0 past the end); row `blockIdx.x` of `part` receives
// (max lw, sum w, sum w^2, sum w c, sum w c^2), w = exp(lw - max)
__device__ __forceinline__ void tracker_block_summary(float out_lw, float out_cnt, double* __restrict__ part) {
  const int tid = threadIdx.x;
  __shared__ float s_mx[kMttThreads / 32];
  __shared__ double s_sum[kMttThreads / 32][4];
  const int lane = tid & 31, warp = tid >> 5;
  float mx = out_lw;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  if (lane == 0) s_mx[warp] = mx;
  __syncthreads();
  mx = s_mx[0];
#pragma unroll
  for (int w = 1; w < kMttThreads / 32; ++w) mx = fmaxf(mx, s_mx[w]);
  double v[4] = {0, 0, 0, 0};
  if (out_lw > -INFINITY) {
    const double w = exp((double)out_lw - (double)mx), c = (double)out_cnt;
    v[0] = w; v[1] = w * w; v[2] = w * c; v[3] = w * c * c;
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
    if (lane == 0) s_sum[warp][r] = x;
  }
  __syncthreads();
  if (tid < 5) {
    double x = (double)mx;
    if (tid > 0) { x = 0; for (int w = 0; w < kMttThreads / 32; ++w) x += s_sum[w][tid - 1]; }
    part[(size_t)blockIdx.x * 5 + tid] = x;
  }
}

__global__ void __launch_bounds__(kMttThreads) mtt_step(const __grid_constant__ MttArgs a) {
  extern __shared__ float s_der[];  // [kMttDerived][kMttFast][kMttThreads]
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
  const int64_t j = a.j0 + (int64_t)blockIdx.x * kMttThreads + tid;
  float out_lw = -INFINITY, out_cnt = 0.f;  // this particle's log-weight and track count (summary below)
  do {
  if (j >= a.n) break;
  float loc[kMttDerived * (16 - kMttFast)];
  const DerShared ds{s_der, tid};
  const DerLocal dl{loc};
  const uint32_t t = a.ctl->t;
  const int64_t par = a.anc[j];
  const uint32_t* __restrict__ src;
  if (a.n_per) {   // the parent's record lives on the device that owns it (peer load over NVLink when it is not this one)
    const int64_t owner = par / a.n_per;
    src = a.sin_all[owner] + (size_t)(par - owner * a.n_per) * kMttWords;
  } else src = a.sin + (size_t)par * kMttWords;
  uint32_t* __restrict__ dst = a.sout + (size_t)(j - a.j0) * kMttWords;
  const uint4 hdr = *reinterpret_cast<const uint4*>(src);
  const int cnt = (int)hdr.x;
  PZ_ASSERT(par >= 0 && cnt >= 0 && cnt <= 16 && (a.n_per == 0 || par / a.n_per < kMttMaxDevices));
  const int cnt_fast = cnt < kMttFast ? cnt : kMttFast;
  const int next_id = (int)hdr.y;
  const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);

  // ---- pass A: predictive density of a measurement of every track: N(z; H m, H P H^T + r I)
  for (int k = 0; k < cnt_fast; ++k) mtt_derive(a.p, src, k, ds);
  for (int k = kMttFast; k < cnt; ++k) mtt_derive(a.p, src, k, dl);

  // ---- pass B: association proposal; uniforms: detection o takes word o & 3 of Philox call o >> 2
  uint32_t assigned = 0;   // bit k: track k has a detection
  uint32_t newmask = 0;    // bit o: detection o starts a track
  unsigned long long obs_lo = 0, obs_hi = 0;  // detection index of track k, 5 bits each (k < 12 | k >= 12)
  float lw = 0.f;
  U4 rnd = U4{0, 0, 0, 0};
  for (int o = 0; o < M; ++o) {
    if ((o & 3) == 0) rnd = philox4x32_10((uint32_t)j, (uint32_t)(o >> 2), t, kStreamMtt, k0, k1);
    const uint32_t word = (o & 3) == 0 ? rnd.x : ((o & 3) == 1 ? rnd.y : ((o & 3) == 2 ? rnd.z : rnd.w));
    const float z0 = s_z[o][0], z1 = s_z[o][1];
    float ttot = 0.f;
    for (int kp = 0; 2 * kp < cnt_fast; ++kp) ttot += mtt_like_pair(z0, z1, assigned, kp, cnt_fast, ds);
    for (int k = kMttFast; k < cnt; ++k) ttot += mtt_like(a.p, z0, z1, assigned, k, dl);
    const float lc = s_lc[o], ln = s_ln[o];
    const float tot = lc + ln + ttot;
    lw += __logf(tot);
    const float u = u01(word) * tot;
    if (u < lc) {
      // clutter
    } else if (u < lc + ln) {
      newmask |= 1u << o;
    } else {
      const float target = u - lc - ln;
      float cum = 0.f;
      int pick = -1, last = -1;
      for (int k = 0; k < cnt_fast; ++k) mtt_pick(target, k, ds, cum, pick, last);
      for (int k = kMttFast; k < cnt; ++k) mtt_pick(target, k, dl, cum, pick, last);
      if (pick < 0) pick = last;          // round-off past the last candidate: take the last one
      if (pick < 0) newmask |= 1u << o;   // no candidate track at all: start a track
      else {
        assigned |= 1u << pick;
        if (pick < 12) obs_lo |= (unsigned long long)o << (5 * pick);
        else obs_hi |= (unsigned long long)o << (5 * (pick - 12));
      }
    }
  }
  const int n_undet = cnt - __popc(assigned);
  lw += (float)n_undet * a.p.log1m_pspd;

  // ---- pass C: update / coast, survivors compacted in order; coasting draw of track k: word k & 3 of call 8 + (k >> 2)
  int n_surv = 0;
  for (int k = 0; k < cnt; ++k) {
    if ((k & 3) == 0) rnd = philox4x32_10((uint32_t)j, (uint32_t)(8 + (k >> 2)), t, kStreamMtt, k0, k1);
    const uint32_t word = (k & 3) == 0 ? rnd.x : ((k & 3) == 1 ? rnd.y : ((k & 3) == 2 ? rnd.z : rnd.w));
    const bool det = (assigned >> k) & 1u;
    if (!det && !(u01(word) < a.p.p_coast)) continue;  // undetected and not coasting: the track dies
    Track tr;
    load_track(src + 4 + 16 * k, tr);
    predict(a.p, tr);
    if (det) {
      const int o = (int)((k < 12 ? obs_lo >> (5 * k) : obs_hi >> (5 * (k - 12))) & 31ull);
      update(tr, s_z[o][0], s_z[o][1], a.p.r);
    }
    store_track(dst + 4 + 16 * n_surv, tr);
    ++n_surv;
  }
  // births: the i-th new-track detection (newTrackD conditioned on it)
  const int n_new = __popc(newmask);
  const int room = 16 - n_surv;
  const int n_born = n_new < room ? n_new : room;
  uint32_t nm = newmask;
  for (int i = 0; i < n_born; ++i) {
    const int o = __ffs(nm) - 1;
    nm &= nm - 1;
    Track nt;
    nt.id = next_id + i;
#pragma unroll
    for (int q = 0; q < 4; ++q) nt.m[q] = 0.f;
#pragma unroll
    for (int q = 0; q < 10; ++q) nt.c[q] = 0.f;
    nt.c[0] = a.p.bp; nt.c[4] = a.p.bp; nt.c[7] = a.p.bv; nt.c[9] = a.p.bv;
    update(nt, s_z[o][0], s_z[o][1], a.p.r);
    store_track(dst + 4 + 16 * (n_surv + i), nt);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4((uint32_t)(n_surv + n_born), (uint32_t)(next_id + n_new), n_new > room ? 1u : 0u, 0u);
  a.lw[j] = lw;
  if (n_new > room) atomicAdd(a.overflow, (unsigned long long)(n_new - room));
  out_lw = sanitize_lw(lw);
  out_cnt = (float)(n_surv + n_born);
  } while (0);

  tracker_block_summary(out_lw, out_cnt, a.part);
}

// =================================================================================================
// MultiCamera tracker step (Examples/MultiCamera.hs:40-227), one thread per particle.
//
// A particle = up to 16 tracks (box marginal over R^4: F = H = I and diagonal Q / R / prior keep its covariance
// DIAGONAL, so a track is 4 means + 4 variances; the CPU oracle carries the full matrices and agrees) and up to 16
// appearance marginals over R^10 (mean + upper triangle of the covariance) shared by the tracks of one identity.
//   tracksMotion (:183-195): survival Bernoulli(exp(-0.02)) and predict per track, Poisson(0.1) births, each with a
//     camera ~ U{0,1}, a new appearance with probability 1 / (A + 1) or the identity of an existing one, the box
//     prior N((0,0,1,1), diag(1,1,.2,.2));
//   assocWithClutterCustomProposal (:162-180) per camera, tracks in list order: scores of the remaining observations
//     (box predictive x uniform-sphere pose density x scalar appearance predictive), categorical pick, Kalman update
//     of the box and rank-one update of the appearance, weight *= sum of the likelihoods; a track without an
//     observation left kills the particle (passert False); the rest is clutter with the count prior of :166-167.
// Record (1252 words): header {tracks, numTracks, appearances, flags}; 16 x 12 words {trackID, identity, camera,
// startTime, mean[4], var[4]}; 16 x 66 words {mean[10], cov upper triangle[55], pad}.  The child's record is built in
// place in global memory (it is this thread's own; only live parts move).  Statistical parity with the oracle (different
// uniforms), exact agreement where the association is forced (tests/test_mcam.py).
constexpr uint32_t kStreamMcam = 0x4D43414Du;  // "MCAM"
constexpr int kMcK = 16, kMcA = 16, kMcD = 10, kMcTrackW = 12, kMcAppW = 66;
constexpr int kMcApps0 = 4 + kMcK * kMcTrackW;            // 196
constexpr int kMcWords = kMcApps0 + kMcA * kMcAppW;       // 1252
constexpr int kMcObsW = 16;                               // camera, pose[10], box[4], appearance reading

struct McamDev {
  float q[4], prior_m[4], prior_v[4], clutter_m[4];
  float ps, lb, lc, r_box, r_app, lsphere;
};

struct McamArgs {
  const uint32_t* sin; uint32_t* sout; const int32_t* anc; float* lw; const float* obs; const CtlDev* ctl;
  unsigned long long* overflow; double* part;
  int64_t n, j0, n_per;
  const uint32_t* sin_all[kMttMaxDevices];
  int32_t n_obs; uint64_t seed;
  McamDev p;
};

struct McRng {   // sequential uniforms of one particle: word (i & 3) of Philox call (i >> 2)
  uint32_t j, t, k0, k1, call; U4 r; int have;
  __device__ __forceinline__ float next() {
    if (!have) { r = philox4x32_10(j, call++, t, kStreamMcam, k0, k1); have = 4; }
    const uint32_t w = have == 4 ? r.x : (have == 3 ? r.y : (have == 2 ? r.z : r.w));
    --have;
    return u01(w);
  }
};

__device__ __forceinline__ int mc_tri(int i, int j) { return i * kMcD - (i * (i - 1)) / 2 + (j - i); }  // i <= j

__global__ void __launch_bounds__(kMttThreads) mcam_step(const __grid_constant__ McamArgs a) {
  __shared__ float s_obs[32][kMcObsW];
  __shared__ float s_cl[32], s_lfact[64];
  const int tid = threadIdx.x;
  const int M = a.n_obs;
  const McamDev& p = a.p;
  for (int i = tid; i < M * kMcObsW; i += kMttThreads) s_obs[i / kMcObsW][i % kMcObsW] = a.obs[i];
  if (tid < 64) s_lfact[tid] = lgammaf((float)tid + 1.0f);
  __syncthreads();
  if (tid < M) {   // clutter density of observation tid: box ~ N(clutter_m, I4), pose ~ N(0, I10), reading ~ N(0, 1)
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float d = s_obs[tid][11 + i] - p.clutter_m[i]; q += d * d; }
#pragma unroll
    for (int i = 0; i < kMcD; ++i) q += s_obs[tid][1 + i] * s_obs[tid][1 + i];
    q += s_obs[tid][15] * s_obs[tid][15];
    s_cl[tid] = -0.5f * q - 0.5f * 15.0f * 1.8378770664093453f;
  }
  __syncthreads();
  const int64_t j = a.j0 + (int64_t)blockIdx.x * kMttThreads + tid;
  float out_lw = -INFINITY, out_cnt = 0.f;
  do {
  if (j >= a.n) break;
  const uint32_t t = a.ctl->t;
  const int64_t par = a.anc[j];
  const uint32_t* __restrict__ src;
  if (a.n_per) { const int64_t owner = par / a.n_per; src = a.sin_all[owner] + (size_t)(par - owner * a.n_per) * kMcWords; }
  else src = a.sin + (size_t)par * kMcWords;
  uint32_t* __restrict__ dst = a.sout + (size_t)(j - a.j0) * kMcWords;
  const uint4 hdr = *reinterpret_cast<const uint4*>(src);
  const int cnt = (int)hdr.x;
  int num_tracks = (int)hdr.y, A = (int)hdr.z;
  PZ_ASSERT(par >= 0 && cnt >= 0 && cnt <= kMcK && A >= 0 && A <= kMcA);
  McRng rng{(uint32_t)j, t, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), 0u, U4{0, 0, 0, 0}, 0};
  // the parent's appearances
  for (int ap = 0; ap < A; ++ap) {
    const uint2* s2 = reinterpret_cast<const uint2*>(src + kMcApps0 + ap * kMcAppW);
    uint2* d2 = reinterpret_cast<uint2*>(dst + kMcApps0 + ap * kMcAppW);
#pragma unroll 11
    for (int w = 0; w < kMcAppW / 2; ++w) d2[w] = s2[w];
  }
  // tracksMotion: survival + predict, survivors compacted in order
  int n_live = 0, ovf = 0;
  for (int k = 0; k < cnt; ++k) {
    if (!(rng.next() < p.ps)) continue;
    const uint4* sp = reinterpret_cast<const uint4*>(src + 4 + kMcTrackW * k);
    const uint4 w0 = sp[0], w1 = sp[1];
    uint4 w2 = sp[2];
    w2.x = __float_as_uint(__uint_as_float(w2.x) + p.q[0]); w2.y = __float_as_uint(__uint_as_float(w2.y) + p.q[1]);
    w2.z = __float_as_uint(__uint_as_float(w2.z) + p.q[2]); w2.w = __float_as_uint(__uint_as_float(w2.w) + p.q[3]);
    uint4* dp = reinterpret_cast<uint4*>(dst + 4 + kMcTrackW * n_live);
    dp[0] = w0; dp[1] = w1; dp[2] = w2;
    ++n_live;
  }
  // births (Poisson by the inverse CDF on one uniform)
  int n_new = 0;
  {
    const float u = rng.next();
    float pk = __expf(-p.lb), F = pk;
    while (u >= F && n_new < 64) { ++n_new; pk *= p.lb / (float)n_new; F += pk; }
  }
  for (int i = 0; i < n_new; ++i) {
    const float u_cam = rng.next(), u_new = rng.next(), u_id = rng.next();
    int cam = (int)(u_cam * 2.0f); cam = cam < 2 ? cam : 1;
    bool fresh = u_new < 1.0f / (float)(A + 1);
    if (fresh && A >= kMcA) { fresh = false; ++ovf; }
    int identity;
    if (fresh) {
      uint32_t* ap = dst + kMcApps0 + A * kMcAppW;
      for (int w = 0; w < kMcAppW; ++w) ap[w] = 0u;
      for (int d = 0; d < kMcD; ++d) ap[kMcD + mc_tri(d, d)] = __float_as_uint(1.0f);
      identity = A; ++A;
    } else { identity = (int)(u_id * (float)A); identity = identity < A ? identity : A - 1; }
    if (n_live < kMcK) {
      uint4* dp = reinterpret_cast<uint4*>(dst + 4 + kMcTrackW * n_live);
      dp[0] = make_uint4((uint32_t)(num_tracks + 1 + i), (uint32_t)identity, (uint32_t)cam, __float_as_uint((float)t));
      dp[1] = make_uint4(__float_as_uint(p.prior_m[0]), __float_as_uint(p.prior_m[1]), __float_as_uint(p.prior_m[2]), __float_as_uint(p.prior_m[3]));
      dp[2] = make_uint4(__float_as_uint(p.prior_v[0]), __float_as_uint(p.prior_v[1]), __float_as_uint(p.prior_v[2]), __float_as_uint(p.prior_v[3]));
      ++n_live;
    } else ++ovf;
  }
  num_tracks += n_new;
  // association, camera by camera
  float lw = 0.f;
  bool dead = false;
  for (int cam = 0; cam < 2 && !dead; ++cam) {
    uint32_t rem = 0;
    for (int o = 0; o < M; ++o) if ((int)s_obs[o][0] == cam) rem |= 1u << o;
    int nT = 0;
    for (int k = 0; k < n_live; ++k) {
      uint32_t* tw = dst + 4 + kMcTrackW * k;
      if ((int)tw[2] != cam) continue;
      ++nT;
      if (!rem) { dead = true; break; }   // passert False: more tracks than observations in this camera
      const int identity = (int)tw[1];
      PZ_ASSERT(identity >= 0 && identity < A);
      float m[4], v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { m[i] = __uint_as_float(tw[4 + i]); v[i] = __uint_as_float(tw[8 + i]); }
      float* ap = reinterpret_cast<float*>(dst + kMcApps0 + identity * kMcAppW);
      float am[kMcD], ac[55];
#pragma unroll
      for (int i = 0; i < kMcD; ++i) am[i] = ap[i];
#pragma unroll
      for (int i = 0; i < 55; ++i) ac[i] = ap[kMcD + i];
      float lbox0 = -0.5f * 4.0f * 1.8378770664093453f;
#pragma unroll
      for (int i = 0; i < 4; ++i) lbox0 -= 0.5f * __logf(v[i] + p.r_box);
      float ll[32];
      float mx = -INFINITY;
      for (uint32_t rr = rem; rr; rr &= rr - 1) {
        const int o = __ffs(rr) - 1;
        const float* row = s_obs[o];
        float quad = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float d = row[11 + i] - m[i]; quad += d * d / (v[i] + p.r_box); }
        float hm = 0.f, hPh = 0.f;
#pragma unroll
        for (int i = 0; i < kMcD; ++i) {
          const float hi = row[1 + i];
          hm += hi * am[i];
          float acc = 0.5f * ac[mc_tri(i, i)] * hi;
#pragma unroll
          for (int jj = i + 1; jj < kMcD; ++jj) acc += ac[mc_tri(i, jj)] * row[1 + jj];
          hPh += 2.0f * hi * acc;
        }
        const float sv = hPh + p.r_app, d = row[15] - hm;
        const float l = lbox0 - 0.5f * quad + p.lsphere - 0.5f * (d * d / sv + __logf(sv) + 1.8378770664093453f);
        ll[o] = l;
        mx = fmaxf(mx, l);
      }
      float total = 0.f;
      for (uint32_t rr = rem; rr; rr &= rr - 1) total += __expf(ll[__ffs(rr) - 1] - mx);
      const float target = rng.next() * total;
      int pick = -1, last = -1;
      float cum = 0.f;
      for (uint32_t rr = rem; rr; rr &= rr - 1) {
        const int o = __ffs(rr) - 1;
        last = o;
        cum += __expf(ll[o] - mx);
        if (pick < 0 && cum > target) pick = o;
      }
      if (pick < 0) pick = last;
      PZ_ASSERT(pick >= 0 && pick < M && ((rem >> pick) & 1u));
      lw += mx + __logf(total);
      rem &= ~(1u << pick);
      const float* row = s_obs[pick];
      // condition the box (diagonal Kalman update, H = I) and the appearance (rank-one update) on the picked observation
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float K = v[i] / (v[i] + p.r_box);
        m[i] += K * (row[11 + i] - m[i]);
        v[i] -= K * v[i];
        tw[4 + i] = __float_as_uint(m[i]); tw[8 + i] = __float_as_uint(v[i]);
      }
      float Ph[kMcD], hm = 0.f, sv = p.r_app;
#pragma unroll
      for (int i = 0; i < kMcD; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int jj = 0; jj < kMcD; ++jj) acc += ac[i <= jj ? mc_tri(i, jj) : mc_tri(jj, i)] * row[1 + jj];
        Ph[i] = acc;
        hm += row[1 + i] * am[i];
      }
#pragma unroll
      for (int i = 0; i < kMcD; ++i) sv += row[1 + i] * Ph[i];
      const float isv = 1.0f / sv, d = row[15] - hm;
#pragma unroll
      for (int i = 0; i < kMcD; ++i) {
        ap[i] = am[i] + Ph[i] * isv * d;
#pragma unroll
        for (int jj = i; jj < kMcD; ++jj) ap[kMcD + mc_tri(i, jj)] = ac[mc_tri(i, jj)] - Ph[i] * Ph[jj] * isv;
      }
    }
    if (dead) break;
    const int nC = __popc(rem);
    lw += -(s_lfact[nT + nC] - s_lfact[nC]) + (-p.lc + (float)nC * __logf(p.lc) - s_lfact[nC]);
    for (uint32_t rr = rem; rr; rr &= rr - 1) lw += s_cl[__ffs(rr) - 1];
  }
  if (dead) lw = -INFINITY;
  *reinterpret_cast<uint4*>(dst) = make_uint4((uint32_t)n_live, (uint32_t)num_tracks, (uint32_t)A, ovf ? 1u : 0u);
  a.lw[j] = lw;
  if (ovf) atomicAdd(a.overflow, (unsigned long long)ovf);
  out_lw = sanitize_lw(lw);
  out_cnt = (float)n_live;
  } while (0);
  tracker_block_summary(out_lw, out_cnt, a.part);
}


// combine the per-CTA partials (log-sum-exp over their maxima) into the step summary
constexpr int kFinalThreads = 1024, kFinalWarps = kFinalThreads / 32;
__global__ void __launch_bounds__(kFinalThreads) mtt_summary_final(PlanDev* pl, const CtlDev* ctl, pz_step_summary* summ, int64_t n,
                                                              double ll_const, const unsigned long long* overflow, int n_ovf,
                                                              const double* __restrict__ part, int32_t ncta) {
  __shared__ double s_m[kFinalWarps], s_v[kFinalWarps][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto shfl_d = [](double x, int d) {
    const long long y = __double_as_longlong(x);
    const int lo = __shfl_xor_sync(0xffffffffu, (int)y, d), hi = __shfl_xor_sync(0xffffffffu, (int)(y >> 32), d);
    return __longlong_as_double(((long long)hi << 32) | (unsigned int)lo);
  };
  double M = -INFINITY;
  for (int c = tid; c < ncta; c += kFinalThreads) M = fmax(M, part[(size_t)c * 5]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) M = fmax(M, shfl_d(M, d));
  if (lane == 0) s_m[warp] = M;
  __syncthreads();
  M = s_m[0];
  for (int w = 1; w < kFinalWarps; ++w) M = fmax(M, s_m[w]);
  double v[4] = {0, 0, 0, 0};
  for (int c = tid; c < ncta; c += kFinalThreads) {
    const double m = part[(size_t)c * 5];
    if (m > -INFINITY) {
      const double f = exp(m - M);
      v[0] += f * part[(size_t)c * 5 + 1];
      v[1] += f * f * part[(size_t)c * 5 + 2];
      v[2] += f * part[(size_t)c * 5 + 3];
      v[3] += f * part[(size_t)c * 5 + 4];
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double x = v[r];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += shfl_d(x, d);
    if (lane == 0) s_v[warp][r] = x;
  }
  __syncthreads();
  if (tid != 0) return;
  double sw = 0, sw2 = 0, sc = 0, scc = 0;
  for (int w = 0; w < kFinalWarps; ++w) { sw += s_v[w][0]; sw2 += s_v[w][1]; sc += s_v[w][2]; scc += s_v[w][3]; }
  pz_step_summary* o = summ + (ctl->t % kObsRing);
  const double m = sc / sw;
  for (int d = 0; d < PZ_MAX_NX; ++d) { o->mean[d] = 0; o->var[d] = 0; }
  o->mean[0] = m;
  o->var[0] = scc / sw - m * m;
  o->step = ctl->t;
  o->log_evidence_inc = M + log(sw) - log((double)n) + ll_const;
  o->ess = sw * sw / sw2;
  o->n_migrated = 0;
  o->e_max = pl->emax;
  unsigned long long ovf = 0;
  for (int i = 0; i < n_ovf; ++i) ovf += overflow[i];
  o->reserved = (int32_t)(ovf > 0x7fffffffull ? 0x7fffffff : ovf);
  o->total_weight = pl->total;
}

// thinned read of the resampled population: every stride-th output slot's parent record
__global__ void mtt_gather_thin(const uint32_t* __restrict__ s, const int32_t* __restrict__ anc, int32_t stride, int64_t n_sel,
                                uint32_t* __restrict__ out, int W) {
  const int64_t sel = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (sel >= n_sel) return;
  const int64_t p = anc[sel * stride];
  for (int w = threadIdx.x & 31; w < W; w += 32) out[(size_t)sel * W + w] = s[(size_t)p * W + w];
}

// the same on a multi-device tracker: the parent's record is read from the device that owns it
__global__ void mtt_gather_thin_multi(MttShard sh, const int32_t* __restrict__ anc, int32_t stride, int64_t n_sel,
                                      uint32_t* __restrict__ out, int W) {
  const int64_t sel = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (sel >= n_sel) return;
  const int64_t p = anc[sel * stride];
  const int64_t owner = p / sh.n_per;
  const uint32_t* s = sh.sin_all[owner] + (size_t)(p - owner * sh.n_per) * W;
  for (int w = threadIdx.x & 31; w < W; w += 32) out[(size_t)sel * W + w] = s[w];
}

// ---- host side ---------------------------------------------------------------------------------
struct MttLaunch {
  MttArgs args;
};

int mtt_words_per_particle() { return kMttWords; }

void mtt_fill_params(const pz_model_desc* d, void* out_mttdev) {
  MttDev& p = *reinterpret_cast<MttDev*>(out_mttdev);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { p.F[i][j] = (float)d->F[i * 4 + j]; p.Q[i][j] = (float)d->Q[i * 4 + j]; }
  p.fa = p.F[0][0]; p.fb = p.F[0][2]; p.fc = p.F[2][0]; p.fd = p.F[2][2];
  p.kron = 1;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const float fexp = (i % 2 == j % 2) ? (i < 2 ? (j < 2 ? p.fa : p.fb) : (j < 2 ? p.fc : p.fd)) : 0.0f;
      if (p.F[i][j] != fexp || (i != j && p.Q[i][j] != 0.0f)) p.kron = 0;
    }
  if (getenv("PZ_NO_KRON")) p.kron = 0;
  const double pspd = d->mtt.survival_prob * d->mtt.pd;
  p.lc = (float)d->mtt.clutter_lambda; p.lb = (float)d->mtt.new_track_lambda; p.pspd = (float)pspd;
  p.p_coast = (float)((1.0 - d->mtt.pd) / (1.0 - pspd));
  p.log1m_pspd = (float)log(1.0 - pspd);
  p.cv = (float)d->mtt.clutter_var; p.bp = (float)d->mtt.birth_pos_var; p.bv = (float)d->mtt.birth_vel_var; p.r = (float)d->mtt.meas_var;
  p.cl_norm = (float)(-log(d->mtt.clutter_var) - log(2 * M_PI));
  p.nb_norm = (float)(-log(d->mtt.birth_pos_var + d->mtt.meas_var) - log(2 * M_PI));
}

size_t mtt_params_size() { return sizeof(MttDev); }

// ---- MultiCamera host side -------------------------------------------------------------------
int mcam_words_per_particle() { return kMcWords; }
size_t mcam_params_size() { return sizeof(McamDev); }
void mcam_fill_params(const pz_model_desc* d, void* out) {
  McamDev& p = *reinterpret_cast<McamDev*>(out);
  for (int i = 0; i < 4; ++i) { p.q[i] = (float)d->Q[i * 4 + i]; p.prior_m[i] = (float)d->x0[i]; p.clutter_m[i] = (float)d->x0[i]; }
  p.prior_v[0] = p.prior_v[1] = (float)d->mtt.birth_pos_var; p.prior_v[2] = p.prior_v[3] = (float)d->mtt.birth_vel_var;
  p.ps = (float)d->mtt.survival_prob; p.lb = (float)d->mtt.new_track_lambda; p.lc = (float)d->mtt.clutter_lambda;
  p.r_box = (float)d->mtt.meas_var; p.r_app = 1.0f;
  p.lsphere = (float)(-(log((double)kMcD) + kMcD / 2.0 * log(M_PI) - lgamma(kMcD / 2.0 + 1.0)));  // -logNSphereVolume, MVDistributions.hs:64-66
}
int launch_mcam_step(const void* params, const uint32_t* sin, uint32_t* sout, const int32_t* anc, float* lw, const float* obs,
                     const CtlDev* ctl, unsigned long long* overflow, double* part, int64_t n, int32_t n_obs, uint64_t seed,
                     cudaStream_t s, const MttShard* shard) {
  McamArgs a;
  a.sin = sin; a.sout = sout; a.anc = anc; a.lw = lw; a.obs = obs; a.ctl = ctl; a.overflow = overflow; a.part = part; a.n = n;
  a.j0 = 0; a.n_per = 0;
  for (int i = 0; i < kMttMaxDevices; ++i) a.sin_all[i] = nullptr;
  if (shard) {
    a.j0 = shard->j0; a.n = shard->j_end; a.n_per = shard->n_per;
    for (int i = 0; i < kMttMaxDevices; ++i) a.sin_all[i] = shard->sin_all[i];
  }
  a.n_obs = n_obs; a.seed = seed; a.p = *reinterpret_cast<const McamDev*>(params);
  mcam_step<<<(unsigned)((a.n - a.j0 + kMttThreads - 1) / kMttThreads), kMttThreads, 0, s>>>(a);
  return 1;
}

int mtt_step_ctas(int64_t n) { return (int)((n + kMttThreads - 1) / kMttThreads); }

int launch_mtt_step(const void* params, const uint32_t* sin, uint32_t* sout, const int32_t* anc, float* lw, const float* obs,
                    const CtlDev* ctl, unsigned long long* overflow, double* part, int64_t n, int32_t n_obs, uint64_t seed,
                    cudaStream_t s, const MttShard* shard) {
  MttArgs a;
  a.sin = sin; a.sout = sout; a.anc = anc; a.lw = lw; a.obs = obs; a.ctl = ctl; a.overflow = overflow; a.part = part; a.n = n;
  a.j0 = 0; a.n_per = 0;
  for (int i = 0; i < kMttMaxDevices; ++i) a.sin_all[i] = nullptr;
  if (shard) {
    a.j0 = shard->j0; a.n = shard->j_end; a.n_per = shard->n_per;
    for (int i = 0; i < kMttMaxDevices; ++i) a.sin_all[i] = shard->sin_all[i];
  }
  a.n_obs = n_obs;
  a.seed = seed; a.p = *reinterpret_cast<const MttDev*>(params);
  const size_t smem = (size_t)kMttDerived * kMttFast * kMttThreads * sizeof(float);  // 28 KB
  static std::atomic<unsigned long long> attr_done{0};  // per device (function attributes are per context)
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(attr_done.load(std::memory_order_acquire) & bit)) {
    cudaFuncSetAttribute(mtt_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  mtt_step<<<(unsigned)((a.n - a.j0 + kMttThreads - 1) / kMttThreads), kMttThreads, smem, s>>>(a);
  return 1;
}

int launch_mtt_summary(const double* part, int64_t n, PlanDev* plan, const CtlDev* ctl, pz_step_summary* summ, double ll_const,
                       const unsigned long long* overflow, cudaStream_t s, int n_ovf) {
  mtt_summary_final<<<1, kFinalThreads, 0, s>>>(plan, ctl, summ, n, ll_const, overflow, n_ovf, part, mtt_step_ctas(n));
  return 1;
}

int launch_mtt_gather_multi(const MttShard& sh, const int32_t* anc, int32_t stride, int64_t n_sel, uint32_t* out, cudaStream_t s, int words) {
  const int per_cta = kThreads / 32;
  mtt_gather_thin_multi<<<(unsigned)((n_sel + per_cta - 1) / per_cta), kThreads, 0, s>>>(sh, anc, stride, n_sel, out, words);
  return 1;
}

int launch_mtt_gather(const uint32_t* sdata, const int32_t* anc, int32_t stride, int64_t n_sel, uint32_t* out, cudaStream_t s, int words) {
  const int per_cta = kThreads / 32;
  mtt_gather_thin<<<(unsigned)((n_sel + per_cta - 1) / per_cta), kThreads, 0, s>>>(sdata, anc, stride, n_sel, out, words);
  return 1;
}

}  // namespace pz
