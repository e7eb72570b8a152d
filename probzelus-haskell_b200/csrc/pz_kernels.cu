// pz_kernels.cu -- hand-written sm_100a kernels of the streaming particle-filter step.
//
// One time step of zparticles' (haskell/src/Inference.hs:101-104) is three launches:
//
//   pf_step_tile<NX,NY,TO,MODE>   per output tile of TO slots:
//       input side  -- merge-path locate the input blocks whose cumulative weight covers the
//                      tile, recompute their weights from state + previous observation
//                      (score/factor/observe, Distributions.hs:29-36,160-161;
//                      MVDistributions.hs:45-46), warp-scan them in fixed point, and stage
//                      every parent at the shared-memory slot of its FIRST child (head mark);
//       output side -- a warp max-scan of the head marks gives every slot its parent
//                      (resample2 + `fst . (xs !!)`, Inference.hs:57-61, as a systematic
//                      resampler), then per child: Philox normals, x' = F x + L z
//                      (noisyIntegrateGen Demo.hs:184-187; SingleTargetTracking.hs:34;
//                      MultiTargetTracking.hs:56-67), the new log-weight, 128-bit coalesced
//                      stores of the state, the block statistics (max exponent, fixed-point
//                      sum) and the centred moment sums of the new population (average /
//                      weightedAverage, Util/Numeric.hs:7-22).
//   pf_scan_blocks     decoupled look-back exclusive scan of the block sums rescaled to the
//       global exponent (normalize / shiftByMax, Inference.hs:63-66, Util/Numeric.hs:12-17),
//       and, in the last CTA, the resampling plan and the posterior summary.
//   pf_partition       merge-path split: first input block of every output tile.
//
// No tensor cores: the path is a streaming scan / gather; see DESIGN.md for the roofline.
#include <cstdlib>
#include <cstring>

#include "pz_kernels.cuh"
#include "pz_math.cuh"

namespace pz {

typedef uint64_t u64;

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
  return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_up_u64(u64 v, int d) {
  uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, d);
  uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), d);
  return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  return __longlong_as_double((long long)shfl_xor_u64((u64)__double_as_longlong(v), m));
}
__device__ __forceinline__ u64 warp_inclusive_scan_u64(u64 v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u64 o = shfl_up_u64(v, d);
    if (lane >= d) v += o;
  }
  return v;
}
__device__ __forceinline__ uint32_t warp_inclusive_scan_u32(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}
__device__ __forceinline__ uint32_t warp_inclusive_max_u32(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v = max(v, o);
  }
  return v;
}
// 2^k as a double (0 when it would underflow)
__device__ __forceinline__ double pow2d(int k) {
  if (k < -1022) return 0.0;
  return __hiloint2double((1023 + k) << 20, 0);
}

// Output-slot position of a CDF value (DESIGN.md "Resampler"):
//   pos(C) = floor(C * rho / 2^L)                 in units of 2^-32 slots, pos(T) = N * 2^32
//   slot_end(C) = #{ j : j * 2^32 + U < pos(C) }  = hi32(pos + 2^32 - 1 - U)
struct SlotMap {
  u64 rho;
  u64 kadd;      // 0xFFFFFFFF - U
  int lsh;       // 64 - L
  uint32_t n_out;
  __device__ __forceinline__ SlotMap(const PlanDev* pl, int64_t n_global)
      : rho(pl->rho), kadd(0xFFFFFFFFull - pl->U), lsh(64 - pl->L), n_out((uint32_t)n_global) {}
  __device__ __forceinline__ uint32_t slot_end(u64 C) const {
    const u64 pos = __umul64hi(C << lsh, rho);
    return min((uint32_t)((pos + kadd) >> 32), n_out);
  }
};

// ------------------------------------------------------------------------------------------
// linear-Gaussian model step (same operation order as the oracle's lg_propagate / lg_logweight)
// ------------------------------------------------------------------------------------------
template <int NX>
__device__ __forceinline__ void lg_propagate(const LgDev& p, const float* x, const float* z, float* xn) {
#pragma unroll
  for (int r = 0; r < NX; ++r) {
    float acc = __fmul_rn(p.F[r][0], x[0]);
#pragma unroll
    for (int c = 1; c < NX; ++c) acc = __fmaf_rn(p.F[r][c], x[c], acc);
#pragma unroll
    for (int c = 0; c <= r; ++c) acc = __fmaf_rn(p.Lq[r][c], z[c], acc);
    xn[r] = acc;
  }
}

template <int NX, int NY>
__device__ __forceinline__ float lg_logweight(const LgDev& p, const float* xn, const float* y) {
  float d[NY];
#pragma unroll
  for (int k = 0; k < NY; ++k) {
    float acc = __fmul_rn(p.H[k][0], xn[0]);
#pragma unroll
    for (int c = 1; c < NX; ++c) acc = __fmaf_rn(p.H[k][c], xn[c], acc);
    d[k] = __fsub_rn(y[k], acc);
  }
  float quad = 0.0f;
#pragma unroll
  for (int k = 0; k < NY; ++k) {
    float t = __fmul_rn(p.Rinv[k][0], d[0]);
#pragma unroll
    for (int l = 1; l < NY; ++l) t = __fmaf_rn(p.Rinv[k][l], d[l], t);
    quad = (k == 0) ? __fmul_rn(d[0], t) : __fmaf_rn(d[k], t, quad);
  }
  return __fmul_rn(p.wscale, quad);
}

// position/velocity structured variants (LgDev::kron): same operations as the dense code minus the
// terms whose coefficient is zero
template <int NX>
__device__ __forceinline__ void kron_propagate(const LgDev& p, const float* x, const float* z, float* xn) {
  constexpr int P = NX / 2;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    float acc = __fmul_rn(p.ka, x[i]);
    acc = __fmaf_rn(p.kb, x[P + i], acc);
    xn[i] = __fmaf_rn(p.klp, z[i], acc);
  }
#pragma unroll
  for (int i = 0; i < P; ++i) {
    float acc;
    if (p.kc != 0.0f) { acc = __fmul_rn(p.kc, x[i]); acc = __fmaf_rn(p.kd, x[P + i], acc); }
    else acc = __fmul_rn(p.kd, x[P + i]);
    xn[P + i] = __fmaf_rn(p.klv, z[P + i], acc);
  }
}

template <int NX, int NY>
__device__ __forceinline__ float kron_logweight(const LgDev& p, const float* xn, const float* y) {
  float quad = 0.0f;
#pragma unroll
  for (int k = 0; k < NY; ++k) {
    const float d = __fsub_rn(y[k], xn[k]);
    const float t = __fmul_rn(p.krinv, d);
    quad = (k == 0) ? __fmul_rn(d, t) : __fmaf_rn(d, t, quad);
  }
  return __fmul_rn(p.wscale, quad);
}

template <int NX, bool KRON>
__device__ __forceinline__ void model_propagate(const LgDev& p, const float* x, const float* z, float* xn) {
  if constexpr (KRON) kron_propagate<NX>(p, x, z, xn); else lg_propagate<NX>(p, x, z, xn);
}
template <int NX, int NY, bool KRON>
__device__ __forceinline__ float model_logweight(const LgDev& p, const float* xn, const float* y) {
  if constexpr (KRON) return kron_logweight<NX, NY>(p, xn, y); else return lg_logweight<NX, NY>(p, xn, y);
}

// normals of the 4 consecutive slots j..j+3 (j % 4 == 0): z[d][k]
template <int NX>
__device__ __forceinline__ void lg_noise4(uint32_t j, uint32_t t, uint64_t seed, float (&z)[NX][4]) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  if constexpr (NX == 1) {
    U4 r = philox4x32_10(j >> 2, 0u, t, kStreamProp, k0, k1);
    normal_pair(r.x, r.y, z[0][0], z[0][1]);
    normal_pair(r.z, r.w, z[0][2], z[0][3]);
  } else if constexpr (NX == 6) {
    // two consecutive slots share three Philox calls: even slot words 0..5, odd slot words 6..11
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      const U4 r0 = philox4x32_10((j >> 1) + pr, 0u, t, kStreamProp, k0, k1);
      const U4 r1 = philox4x32_10((j >> 1) + pr, 1u, t, kStreamProp, k0, k1);
      const U4 r2 = philox4x32_10((j >> 1) + pr, 2u, t, kStreamProp, k0, k1);
      const int s = 2 * pr;
      normal_pair(r0.x, r0.y, z[0][s], z[1][s]);
      normal_pair(r0.z, r0.w, z[2][s], z[3][s]);
      normal_pair(r1.x, r1.y, z[4][s], z[5][s]);
      normal_pair(r1.z, r1.w, z[0][s + 1], z[1][s + 1]);
      normal_pair(r2.x, r2.y, z[2][s + 1], z[3][s + 1]);
      normal_pair(r2.z, r2.w, z[4][s + 1], z[5][s + 1]);
    }
  } else {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
#pragma unroll
      for (int c = 0; 4 * c < NX; ++c) {
        U4 r = philox4x32_10(j + s, (uint32_t)c, t, kStreamProp, k0, k1);
        float zz[4];
        normal_pair(r.x, r.y, zz[0], zz[1]);
        normal_pair(r.z, r.w, zz[2], zz[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (4 * c + i < NX) z[4 * c + i][s] = zz[i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// pf_init: the point-mass population before the first observation, with equal weights
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) pf_init(StepArgs a) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < a.nblk * kBlk) {
    for (int d = 0; d < a.model.nx; ++d) {
      a.X[0][(size_t)d * a.ld + i] = a.model.x0[d];
      a.X[1][(size_t)d * a.ld + i] = a.model.x0[d];
    }
  }
  if (i < a.nblk) {
    int64_t cnt = a.n_local - i * kBlk;
    cnt = cnt > kBlk ? kBlk : cnt;
    a.eb[0][i] = 0;
    a.eb[1][i] = kNZero;
    a.sb[i] = (u64)cnt << 23;  // q of lw = 0 is 2^23
  }
  if (i < a.nscan) a.scan_status[i] = 0ull;
  if (i == 0) {
    a.ctl->t = 0;
    a.ctl->emax_acc = 0;
    a.ctl->scan_counter = 0;
    a.ctl->scan_done = 0;
    a.ctl->uniform_in = 1;
    for (int d = 0; d < 8; ++d) a.plan->center[d] = d < a.model.nx ? a.model.x0[d] : 0.0f;
  }
}

// ------------------------------------------------------------------------------------------
// pf_scan_blocks: single-pass decoupled look-back scan over the block sums
// ------------------------------------------------------------------------------------------
constexpr u64 kFlagAgg = 1ull << 62, kFlagIncl = 2ull << 62, kValMask = (1ull << 62) - 1;

struct ScanArgs {
  const int32_t* eb0;  // block exponents of buffer 0 / 1; the population scanned is (ctl->t + delta) & 1
  const int32_t* eb1;
  const u64* sb;
  u64* pb;
  u64* status;
  double* part;      // [nscan][kMaxRec]
  const double* rec; // [nrec][kMaxRec] or nullptr
  PlanDev* plan;
  CtlDev* ctl;
  pz_step_summary* summ;  // ring or nullptr
  int64_t nblk;
  int64_t n_out_global;
  int64_t n_local;
  int32_t nscan;
  int32_t nrec;
  int32_t nx;
  int32_t delta;      // 0: population already current (init / generic); 1: produced by the step kernel
  int32_t explicit_U; // standalone API: use U_value instead of Philox
  uint32_t U_value;
  uint64_t seed;
  double ll_const;
  int32_t world;
  int32_t rank;
  u64* shard_rec;  // sharded: this rank's slot of the all-gathered record
};

__device__ void finalize_plan(PlanDev* pl, u64 total, int64_t n_out_global, uint32_t U) {
  pl->total = total;
  pl->degenerate = (total == 0);
  pl->U = U;
  if (total != 0) {
    const int L = 64 - __clzll((long long)total);
    pl->L = L;
    pl->rho = (u64)((((unsigned __int128)(u64)n_out_global) << (32 + L)) / total) + 1ull;
  } else {
    pl->L = 1;
    pl->rho = 0;
  }
}

// posterior summary from the centred moment sums in the plan (also re-centres for the next step)
__device__ void write_summary(PlanDev* pl, pz_step_summary* o, uint32_t t, int nx, int64_t n_out_global, double ll_const) {
  const double sw = pl->sw, sw2 = pl->sw2;
  for (int d = 0; d < nx; ++d) {
    const float cen = pl->center[d];
    const double m = pl->sx[d] / sw;
    const double mean = (double)cen + m, var = pl->sxx[d] / sw - m * m;
    pl->center[d] = (float)mean;
    if (o) { o->mean[d] = mean; o->var[d] = var; }
  }
  if (o) {
    for (int d = nx; d < PZ_MAX_NX; ++d) { o->mean[d] = 0.0; o->var[d] = 0.0; }
    o->step = t;
    o->log_evidence_inc = log(sw) + (double)(pl->emax - 23) * 0.6931471805599453 - log((double)n_out_global) + ll_const;
    o->ess = sw * sw / sw2;
    o->n_migrated = 0;
    o->e_max = pl->emax;
    o->reserved = 0;
    o->total_weight = pl->total;
  }
}

__global__ void __launch_bounds__(kThreads) pf_scan_blocks(const ScanArgs s) {
  __shared__ u64 s_warp[kWarps];
  __shared__ u64 s_prefix;
  __shared__ uint32_t s_tile, s_last;
  __shared__ double s_red[kWarps][kMaxRec];
  __shared__ double s_tot[kMaxRec];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(&s.ctl->scan_counter, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int32_t emax = s.ctl->emax_acc;
  const int32_t* __restrict__ ebp = ((s.ctl->t + (uint32_t)s.delta) & 1u) ? s.eb1 : s.eb0;

  // ---- local scan of kScanItems block sums per thread
  u64 v[kScanItems];
  u64 tsum = 0;
  const int64_t b0 = (int64_t)tile * kScanTile + (int64_t)tid * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t b = b0 + k;
    u64 S = 0;
    if (b < s.nblk) {
      int32_t e = ebp[b];
      if (e != kNZero) {
        uint32_t sh = (uint32_t)(emax - e);
        S = sh < 64u ? (s.sb[b] >> sh) : 0ull;
      }
    }
    v[k] = S;
    tsum += S;
  }
  u64 incl = warp_inclusive_scan_u64(tsum, lane);
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  u64 wbase = 0, agg = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    u64 x = s_warp[w];
    if (w < warp) wbase += x;
    agg += x;
  }
  u64 excl = wbase + incl - tsum;

  // ---- decoupled look-back (warp 0)
  if (warp == 0) {
    volatile u64* st = s.status;
    u64 prefix = 0;
    if (tile == 0) {
      if (lane == 0) st[0] = kFlagIncl | agg;
    } else {
      if (lane == 0) st[tile] = kFlagAgg | agg;
      int64_t look = (int64_t)tile - 1;
      while (true) {
        int64_t idx = look - lane;
        u64 w = kFlagIncl;  // virtual inclusive zero before tile 0
        if (idx >= 0) {
          do { w = st[idx]; } while ((w >> 62) == 0);
        }
        uint32_t incl_mask = __ballot_sync(0xffffffffu, (w >> 62) == 2);
        int first = incl_mask ? __ffs(incl_mask) - 1 : 32;  // lanes up to the first inclusive entry contribute
        u64 c = (lane <= first) ? (w & kValMask) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += shfl_xor_u64(c, d);
        prefix += c;
        if (incl_mask) break;
        look -= 32;
      }
      if (lane == 0) st[tile] = kFlagIncl | (prefix + agg);
    }
    if (lane == 0) s_prefix = prefix;
  }
  __syncthreads();
  const u64 prefix = s_prefix;
  u64 run = prefix + excl;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t b = b0 + k;
    if (b < s.nblk) s.pb[b] = run;
    run += v[k];
  }
  if (tile == (uint32_t)(s.nscan - 1) && tid == kThreads - 1) {
    s.pb[s.nblk] = prefix + agg;
    s.plan->local_total = prefix + agg;
  }

  // ---- moment records -> per-CTA partial at the global exponent
  const int R = 3 + 2 * s.nx;
  const bool have_rec = (s.rec != nullptr) && (s.nrec > 0) && (s.delta != 0);
  if (have_rec) {
    double acc[kMaxRec];
#pragma unroll
    for (int r = 0; r < kMaxRec; ++r) acc[r] = 0.0;
    const int rpc = (s.nrec + s.nscan - 1) / s.nscan;
    for (int q = tid; q < rpc; q += kThreads) {
      int64_t J = (int64_t)tile * rpc + q;
      if (J < s.nrec) {
        const double* rc = s.rec + (size_t)J * kMaxRec;
        double et = rc[0];
        if (et > -1.0e9) {
          double f = pow2d((int)et - emax);
          for (int r = 1; r < R; ++r) acc[r] += rc[r] * (r == 2 ? f * f : f);
        }
      }
    }
    for (int r = 1; r < R; ++r) {
      double x = acc[r];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += shfl_xor_f64(x, d);
      if (lane == 0) s_red[warp][r] = x;
    }
    __syncthreads();
    if (tid >= 1 && tid < R) {
      double x = 0;
      for (int w = 0; w < kWarps; ++w) x += s_red[w][tid];
      s.part[(size_t)tile * kMaxRec + tid] = x;
    }
  }

  // ---- last CTA: plan + posterior summary
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&s.ctl->scan_done, 1u) == (uint32_t)(s.nscan - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (have_rec && tid >= 1 && tid < R) {
    double x = 0;
    for (int c = 0; c < s.nscan; ++c) x += ((volatile double*)s.part)[(size_t)c * kMaxRec + tid];
    s_tot[tid] = x;
  }
  __syncthreads();
  if (tid == 0) {
    PlanDev* pl = s.plan;
    const u64 total = ((volatile PlanDev*)pl)->local_total;
    pl->emax = emax;
    if (s.world <= 1) {
      const uint32_t t_next = s.ctl->t + (uint32_t)s.delta;
      const uint32_t U = s.explicit_U ? s.U_value
                                      : philox4x32_10(0u, 0u, t_next, kStreamResample, (uint32_t)s.seed, (uint32_t)(s.seed >> 32)).x;
      pl->base = 0;
      finalize_plan(pl, total, s.n_out_global, U);
    }
    if (have_rec) {
      pl->sw = s_tot[1]; pl->sw2 = s_tot[2];
      for (int d = 0; d < s.nx; ++d) { pl->sx[d] = s_tot[3 + d]; pl->sxx[d] = s_tot[3 + s.nx + d]; }
      if (s.world <= 1)
        write_summary(pl, s.summ ? s.summ + (s.ctl->t % kObsRing) : nullptr, s.ctl->t, s.nx, s.n_out_global, s.ll_const);
    }
    if (s.world > 1) {  // this rank's record for the all-gather
      u64* r = s.shard_rec + (size_t)s.rank * kShardRec;
      r[0] = total;
      r[1] = (u64)__double_as_longlong(have_rec ? pl->sw : 0.0);
      r[2] = (u64)__double_as_longlong(have_rec ? pl->sw2 : 0.0);
      for (int d = 0; d < PZ_MAX_NX; ++d) {
        r[3 + d] = (u64)__double_as_longlong((have_rec && d < s.nx) ? pl->sx[d] : 0.0);
        r[3 + PZ_MAX_NX + d] = (u64)__double_as_longlong((have_rec && d < s.nx) ? pl->sxx[d] : 0.0);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// pf_partition: first input block of every output tile + per-step counter maintenance
// ------------------------------------------------------------------------------------------
struct PartArgs {
  const u64* pb;      // prefix array searched (absolute on a sharded filter)
  int64_t b_lo, b_hi; // block range searched
  int32_t* tile_start;
  u64* status;
  const PlanDev* plan;
  CtlDev* ctl;
  int64_t nblk;
  int64_t slot_base;
  int64_t n_out_global;
  int32_t ntiles_out;
  int32_t tile_out;
  int32_t nscan;
  int32_t delta;
};

__global__ void __launch_bounds__(kThreads) pf_partition(const PartArgs p) {
  const int64_t J = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const PlanDev* pl = p.plan;
  if (J < p.ntiles_out) {
    SlotMap sm(pl, p.n_out_global);
    const uint32_t j0 = (uint32_t)(p.slot_base + J * p.tile_out);
    // min b in [b_lo, b_hi) with slot_end(pb[b+1]) > j0
    int64_t lo = p.b_lo, hi = p.b_hi;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (sm.slot_end(p.pb[mid + 1]) > j0) hi = mid; else lo = mid + 1;
    }
    p.tile_start[J] = (int32_t)lo;
  }
  if (J < p.nscan) p.status[J] = 0ull;
  if (J == 0) {
    p.ctl->t += (uint32_t)p.delta;
    p.ctl->emax_acc = kNZero;
    p.ctl->scan_counter = 0;
    p.ctl->scan_done = 0;
    if (p.delta) p.ctl->uniform_in = 0;
  }
}

// ------------------------------------------------------------------------------------------
// resample phase shared by the step and the ancestor kernels
// ------------------------------------------------------------------------------------------
// Sources: what one particle contributes (its weight) and what is staged for its children.
template <int NX, int NY, bool KRON>
struct StateSource {
  const float* x;      // X[cur]
  int64_t ld, n_local;
  const LgDev* m;
  float y[NY];
  bool uniform;
  float v[NX][4];
  template <bool FULL>
  __device__ __forceinline__ void load(int64_t i0, int32_t eb, uint32_t (&q)[4]) {
#pragma unroll
    for (int d = 0; d < NX; ++d) {
      const float4 f = *reinterpret_cast<const float4*>(x + (size_t)d * ld + i0);
      v[d][0] = f.x; v[d][1] = f.y; v[d][2] = f.z; v[d][3] = f.w;
    }
    if (uniform) {  // the initial point mass: equal weights (lw = 0 -> n = 0, m = 2^23)
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k] = q_at(1u << 23, 0, eb);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int32_t n; uint32_t mm;
        float xs[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) xs[d] = v[d][k];
        weight_split(model_logweight<NX, NY, KRON>(*m, xs, y), n, mm);
        q[k] = q_at(mm, n, eb);
      }
    }
    if (!FULL) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i0 + k >= n_local) q[k] = 0;
    }
  }
  __device__ __forceinline__ void put(float* s_par, int stride, uint32_t slot, int k, int64_t) const {
#pragma unroll
    for (int d = 0; d < NX; ++d) s_par[d * stride + slot] = v[d][k];
  }
};

struct LogwSource {  // generic path: stored log-weights, payload = particle index
  const float* lw;
  int64_t n_local;
  template <bool FULL>
  __device__ __forceinline__ void load(int64_t i0, int32_t eb, uint32_t (&q)[4]) {
    const float4 f = *reinterpret_cast<const float4*>(lw + i0);
    const float w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int32_t n; uint32_t mm;
      weight_split(w[k], n, mm);
      q[k] = q_at(mm, n, eb);
      if (!FULL && i0 + k >= n_local) q[k] = 0;
    }
  }
  __device__ __forceinline__ void put(float* s_par, int, uint32_t slot, int k, int64_t i0) const {
    reinterpret_cast<int32_t*>(s_par)[slot] = (int32_t)(i0 + k);
  }
};

// One half-block (128 particles, 4 per lane): fixed-point weights, warp scan, slot ranges, and
// a head mark at the first child slot of every particle that has children in [j0, jend) -- plus
// one at every 128-slot segment boundary its children span, so that each output segment can
// resolve its parents with a warp-local max-scan.
template <class Source, bool FULL>
__device__ __forceinline__ void resample_half(Source& src, float* s_par, uint16_t* s_head, int stride, int64_t i0,
                                              int32_t eb, uint32_t ksh, u64 Pb, const SlotMap& sm, uint32_t j0,
                                              uint32_t jend, int lane, uint32_t& carry, uint32_t& o_prev) {
  uint32_t q[4];
  src.template load<FULL>(i0, eb, q);
  const uint32_t c0 = q[0], c1 = c0 + q[1], c2 = c1 + q[2], c3 = c2 + q[3];
  const uint32_t incl = warp_inclusive_scan_u32(c3, lane);
  const uint32_t excl = carry + incl - c3;
  carry += __shfl_sync(0xffffffffu, incl, 31);
  uint32_t o[4];
  o[0] = sm.slot_end(Pb + (u64)((excl + c0) >> ksh));
  o[1] = sm.slot_end(Pb + (u64)((excl + c1) >> ksh));
  o[2] = sm.slot_end(Pb + (u64)((excl + c2) >> ksh));
  o[3] = sm.slot_end(Pb + (u64)((excl + c3) >> ksh));
  uint32_t o_lo = __shfl_up_sync(0xffffffffu, o[3], 1);
  if (lane == 0) o_lo = o_prev;
  o_prev = __shfl_sync(0xffffffffu, o[3], 31);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t lo = max(o_lo, j0), hi = min(o[k], jend);
    if (lo < hi) {
      uint32_t slot = lo - j0;
      const uint32_t slot_last = hi - j0 - 1u;
      s_head[slot] = (uint16_t)(slot + 1);
      src.put(s_par, stride, slot, k, i0);
      if ((slot ^ slot_last) >> 7) {  // rare: the children span a segment boundary
        for (slot = (slot + 128u) & ~127u; slot <= slot_last; slot += 128u) {
          s_head[slot] = (uint16_t)(slot + 1);
          src.put(s_par, stride, slot, k, i0);
        }
      }
    }
    o_lo = o[k];
  }
}

template <class Source>
__device__ __forceinline__ void resample_tile(Source& src, float* s_par, uint16_t* s_head, int stride,
                                              const int32_t* __restrict__ ebx, const u64* __restrict__ pbx, u64 base,
                                              int64_t b_first, int64_t b_end, int64_t n_items, const SlotMap& sm,
                                              int32_t emax, uint32_t j0, uint32_t jend, int lane, int warp) {
  for (int64_t b = b_first + warp; b < b_end; b += kWarps) {
    const u64 Pb = base + pbx[b];
    uint32_t o_prev = sm.slot_end(Pb);
    if (o_prev >= jend) break;  // this and every later block start past the tile
    const int32_t eb = ebx[b];
    const uint32_t ksh = (uint32_t)(emax - eb);
    if (eb == kNZero || ksh >= 32u) continue;  // no mass at the global scale
    uint32_t carry = 0;
    const int64_t i0 = b * kBlk + lane * 4;
    if ((b + 1) * kBlk <= n_items) {
      resample_half<Source, true>(src, s_par, s_head, stride, i0, eb, ksh, Pb, sm, j0, jend, lane, carry, o_prev);
      resample_half<Source, true>(src, s_par, s_head, stride, i0 + 128, eb, ksh, Pb, sm, j0, jend, lane, carry, o_prev);
    } else {
      resample_half<Source, false>(src, s_par, s_head, stride, i0, eb, ksh, Pb, sm, j0, jend, lane, carry, o_prev);
      resample_half<Source, false>(src, s_par, s_head, stride, i0 + 128, eb, ksh, Pb, sm, j0, jend, lane, carry, o_prev);
    }
  }
}

// Parent slot (+1) of the lane's 4 slots of one 128-slot segment.
__device__ __forceinline__ void segment_parents(const uint16_t* s_head, int sl, int lane, uint32_t (&h)[4]) {
  const uint2 hv = *reinterpret_cast<const uint2*>(&s_head[sl]);
  h[0] = hv.x & 0xffffu;
  h[1] = max(h[0], hv.x >> 16);
  h[2] = max(h[1], hv.y & 0xffffu);
  h[3] = max(h[2], hv.y >> 16);
  const uint32_t incl = warp_inclusive_max_u32(h[3], lane);
  uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = max(h[k], excl);
}

struct TileArgs {
  StepArgs a;
  const int32_t* anc;  // MODE 1: parents by stored ancestor index
};

// MODE 0: fused systematic resample (parents staged by resample_tile)
// MODE 1: parents gathered through a stored ancestor array (multinomial / debug path)
// resident CTAs per SM the register allocation is held to (NX = 1 / 4 / 6)
constexpr int min_ctas(int nx) { return nx == 1 ? 4 : (nx <= 4 ? 3 : 2); }

template <int NX, int NY, int TO, int MODE, bool KRON>
__global__ void __launch_bounds__(kThreads, min_ctas(NX)) pf_step_tile(const __grid_constant__ TileArgs ta) {
  extern __shared__ __align__(16) float s_par[];  // [NX][TO] parents, then [TO] head marks (u16)
  __shared__ int32_t s_eref[kWarps];
  __shared__ double s_red[kWarps][kMaxRec];
  uint16_t* s_head = reinterpret_cast<uint16_t*>(s_par + NX * TO);
  const StepArgs& a = ta.a;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t t = a.ctl->t;
  const int cur = t & 1, nxt = cur ^ 1;
  const float* __restrict__ xin = a.X[cur];
  float* __restrict__ xout = a.X[nxt];
  const PlanDev* pl = a.plan;
  const int64_t j0l = (int64_t)blockIdx.x * TO;  // local slot
  const int64_t jendl = (j0l + TO < a.n_local) ? j0l + TO : a.n_local;
  const uint32_t j0 = (uint32_t)(a.slot_base + j0l), jend = (uint32_t)(a.slot_base + jendl);
  const int nslots = (int)(jendl - j0l);

  float y_new[NY];
#pragma unroll
  for (int k = 0; k < NY; ++k) y_new[k] = a.obs[(size_t)(t % kObsRing) * PZ_MAX_NY + k];

  if constexpr (MODE == 0) {
    {  // clear the head marks: TO u16 = TO/8 uint4
      uint4* hz = reinterpret_cast<uint4*>(s_head);
      for (int i = tid; i < TO / 8; i += kThreads) hz[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    SlotMap sm(pl, a.n_global);
    StateSource<NX, NY, KRON> src;
    src.x = xin; src.ld = a.ld; src.n_local = a.n_local; src.m = &a.model;
    src.uniform = a.ctl->uniform_in != 0;
    const uint32_t tp = (t + kObsRing - 1) % kObsRing;
#pragma unroll
    for (int k = 0; k < NY; ++k) src.y[k] = a.obs[(size_t)tp * PZ_MAX_NY + k];
    resample_tile(src, s_par, s_head, TO, a.eb[cur], a.pbx, 0ull, (int64_t)a.tile_start[blockIdx.x], a.b_hi,
                  a.world > 1 ? INT64_MAX : a.n_local, sm, pl->emax, j0, jend, lane, warp);
  } else {
    for (int sidx = tid; sidx < nslots; sidx += kThreads) {
      const int64_t p = ta.anc[j0l + sidx];
#pragma unroll
      for (int d = 0; d < NX; ++d) s_par[d * TO + sidx] = xin[(size_t)d * a.ld + p];
    }
  }
  __syncthreads();

  // ---- output side: one warp owns one canonical output block (two segments of 128 slots)
  double acc_sw = 0, acc_sw2 = 0, acc_sx[NX], acc_sxx[NX];
#pragma unroll
  for (int d = 0; d < NX; ++d) { acc_sx[d] = 0; acc_sxx[d] = 0; }
  int32_t aref = kNZero;
  float cen[NX];
#pragma unroll
  for (int d = 0; d < NX; ++d) cen[d] = pl->center[d];
  const bool full_tile = (nslots == TO);

  for (int ob = warp; ob * kBlk < nslots; ob += kWarps) {
    int32_t nn[8]; uint32_t mm[8];
    int32_t emx = kNZero;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int sl = ob * kBlk + half * 128 + lane * 4;  // tile-relative slot of this lane's quad
      float xp[NX][4];
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        const float4 f = *reinterpret_cast<const float4*>(&s_par[d * TO + sl]);
        xp[d][0] = f.x; xp[d][1] = f.y; xp[d][2] = f.z; xp[d][3] = f.w;
      }
      if constexpr (MODE == 0) {
        uint32_t h[4];
        segment_parents(s_head, sl, lane, h);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (h[k] != (uint32_t)(sl + k + 1) && h[k] != 0u) {
#pragma unroll
            for (int d = 0; d < NX; ++d) xp[d][k] = s_par[d * TO + h[k] - 1];
          }
        }
        __syncwarp();  // every parent read of this segment precedes the write-back below
      }
      float z[NX][4];
      lg_noise4<NX>(j0 + (uint32_t)sl, t, a.seed, z);
      float xo[NX][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float xs[NX], zs[NX], xn[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) { xs[d] = xp[d][k]; zs[d] = z[d][k]; }
        model_propagate<NX, KRON>(a.model, xs, zs, xn);
#pragma unroll
        for (int d = 0; d < NX; ++d) xo[d][k] = xn[d];
        weight_split(model_logweight<NX, NY, KRON>(a.model, xn, y_new), nn[half * 4 + k], mm[half * 4 + k]);
      }
      if (!full_tile) {  // only the last tile of a population whose size is not a multiple of TO
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (sl + k >= nslots) { nn[half * 4 + k] = kNZero; mm[half * 4 + k] = 0; }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) emx = max(emx, nn[half * 4 + k]);
      if (full_tile || sl < nslots) {
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          const float4 f = make_float4(xo[d][0], xo[d][1], xo[d][2], xo[d][3]);
          *reinterpret_cast<float4*>(&xout[(size_t)d * a.ld + j0l + sl]) = f;
          *reinterpret_cast<float4*>(&s_par[d * TO + sl]) = f;
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) emx = max(emx, __shfl_xor_sync(0xffffffffu, emx, d));
    // block sum at the block exponent + centred moments (fp32 per lane, fp64 across blocks)
    uint32_t ssum = 0;
    if (emx != kNZero) {
      float p_sw = 0.f, p_sw2 = 0.f, p_sx[NX], p_sxx[NX];
#pragma unroll
      for (int d = 0; d < NX; ++d) { p_sx[d] = 0.f; p_sxx[d] = 0.f; }
      __syncwarp();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int sl = ob * kBlk + half * 128 + lane * 4;
        float xv[NX][4];
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          const float4 f = *reinterpret_cast<const float4*>(&s_par[d * TO + sl]);
          xv[d][0] = f.x; xv[d][1] = f.y; xv[d][2] = f.z; xv[d][3] = f.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t q = q_at(mm[half * 4 + k], nn[half * 4 + k], emx);
          ssum += q;
          const float w = __uint2float_rn(q);  // exact: q < 2^24
          p_sw = __fadd_rn(p_sw, w);
          p_sw2 = __fmaf_rn(w, w, p_sw2);
#pragma unroll
          for (int d = 0; d < NX; ++d) {
            // slots past the population's end hold uninitialised staging data (weight 0, but 0 * NaN = NaN)
            const float dx = (full_tile || sl + k < nslots) ? __fsub_rn(xv[d][k], cen[d]) : 0.0f;
            const float wd = __fmul_rn(w, dx);
            p_sx[d] = __fadd_rn(p_sx[d], wd);
            p_sxx[d] = __fmaf_rn(wd, dx, p_sxx[d]);
          }
        }
      }
      double f = 1.0;
      if (aref == kNZero) aref = emx;
      else if (emx > aref) {
        const double g = pow2d(aref - emx);
        acc_sw *= g; acc_sw2 *= g * g;
#pragma unroll
        for (int d = 0; d < NX; ++d) { acc_sx[d] *= g; acc_sxx[d] *= g; }
        aref = emx;
      } else f = pow2d(emx - aref);
      acc_sw += f * (double)p_sw; acc_sw2 += (f * f) * (double)p_sw2;
#pragma unroll
      for (int d = 0; d < NX; ++d) { acc_sx[d] += f * (double)p_sx[d]; acc_sxx[d] += f * (double)p_sxx[d]; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, d);
    if (lane == 0) {
      const int64_t bo = (j0l / kBlk) + ob;
      a.eb[nxt][bo] = emx;
      a.sb[bo] = (u64)ssum;
    }
  }

  // ---- CTA-level reduction of the moment sums -> one record per tile
  if (lane == 0) s_eref[warp] = aref;
  __syncthreads();
  int32_t etile = kNZero;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) etile = max(etile, s_eref[w]);
  {
    const double g = (aref == kNZero) ? 0.0 : pow2d(aref - etile);
    double vals[kMaxRec];
    vals[0] = 0; vals[1] = acc_sw * g; vals[2] = acc_sw2 * g * g;
#pragma unroll
    for (int d = 0; d < NX; ++d) { vals[3 + d] = acc_sx[d] * g; vals[3 + NX + d] = acc_sxx[d] * g; }
#pragma unroll
    for (int r = 1; r < 3 + 2 * NX; ++r) {
      double x = vals[r];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += shfl_xor_f64(x, d);
      if (lane == 0) s_red[warp][r] = x;
    }
  }
  __syncthreads();
  if (tid < 3 + 2 * NX) {
    double x = 0;
    if (tid == 0) x = (double)etile;
    else for (int w = 0; w < kWarps; ++w) x += s_red[w][tid];
    a.rec[(size_t)blockIdx.x * kMaxRec + tid] = x;
  }
  if (tid == 0 && etile != kNZero) atomicMax(&a.ctl->emax_acc, etile);
}

// ------------------------------------------------------------------------------------------
// generic path kernels (stored log-weights)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) pf_stats_from_lw(const float* __restrict__ lw, int64_t n, int64_t nblk,
                                                             int32_t* eb, u64* sb, CtlDev* ctl) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * kWarps + warp;
  if (b >= nblk) return;
  int32_t nn[8]; uint32_t mm[8];
  int32_t emx = kNZero;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int64_t i0 = b * kBlk + half * 128 + lane * 4;
    const float4 f = *reinterpret_cast<const float4*>(lw + i0);
    const float w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      weight_split(w[k], nn[half * 4 + k], mm[half * 4 + k]);
      if (i0 + k >= n) { nn[half * 4 + k] = kNZero; mm[half * 4 + k] = 0; }
      emx = max(emx, nn[half * 4 + k]);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) emx = max(emx, __shfl_xor_sync(0xffffffffu, emx, d));
  uint32_t ssum = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) ssum += q_at(mm[k], nn[k], emx);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, d);
  if (lane == 0) {
    eb[b] = emx;
    sb[b] = (emx == kNZero) ? 0ull : (u64)ssum;
    if (emx != kNZero) atomicMax(&ctl->emax_acc, emx);
  }
}

__device__ __forceinline__ const int32_t* generic_eb(const GenericArgs& g) {
  return (g.parity_ctl && (g.parity_ctl->t & 1u)) ? g.eb_alt : g.eb;
}

__global__ void pf_ctl_reset(CtlDev* ctl) {
  ctl->t = 0; ctl->emax_acc = kNZero; ctl->scan_counter = 0; ctl->scan_done = 0; ctl->uniform_in = 0;
}

template <int TO>
__global__ void __launch_bounds__(kThreads) pf_ancestors(const GenericArgs g) {
  __shared__ __align__(16) int32_t s_anc[TO];
  __shared__ __align__(16) uint16_t s_head[TO];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PlanDev* pl = g.plan;
  const int64_t j0 = (int64_t)blockIdx.x * TO;
  const int64_t jend = (j0 + TO < g.n) ? j0 + TO : g.n;
  SlotMap sm(pl, g.n);
  LogwSource src{g.lw, g.n};
  for (int sidx = tid; sidx < TO; sidx += kThreads) { s_anc[sidx] = -1; s_head[sidx] = 0; }
  __syncthreads();
  resample_tile(src, reinterpret_cast<float*>(s_anc), s_head, TO, generic_eb(g), g.pb, 0ull,
                (int64_t)g.tile_start[blockIdx.x], g.nblk, g.n, sm, pl->emax, (uint32_t)j0, (uint32_t)jend, lane, warp);
  __syncthreads();
  for (int seg = warp; seg * 128 < (int)(jend - j0); seg += kWarps) {
    const int sl = seg * 128 + lane * 4;
    uint32_t h[4];
    segment_parents(s_head, sl, lane, h);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (j0 + sl + k < jend) g.anc[j0 + sl + k] = h[k] ? s_anc[h[k] - 1] : -1;
  }
}

// N iid categorical draws on the canonical fixed-point CDF (resample2, Inference.hs:57-61)
__global__ void __launch_bounds__(kThreads) pf_multinomial(const GenericArgs g) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= g.n) return;
  const PlanDev* pl = g.plan;
  const u64 total = pl->total;
  U4 r = philox4x32_10((uint32_t)j, 0u, g.step, kStreamMulti, (uint32_t)g.seed, (uint32_t)(g.seed >> 32));
  const u64 u = ((u64)r.x << 32) | r.y;
  const u64 tau = __umul64hi(u, total);
  // largest b with pb[b] <= tau
  int64_t lo = 0, hi = g.nblk;  // invariant: pb[lo] <= tau < pb[hi]
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (g.pb[mid] <= tau) lo = mid; else hi = mid;
  }
  const int64_t b = lo;
  const int32_t eb = generic_eb(g)[b];
  const uint32_t ksh = (uint32_t)(pl->emax - eb);
  const u64 Pb = g.pb[b];
  u64 c = 0;
  int64_t i = b * kBlk, iend = (b + 1) * kBlk < g.n ? (b + 1) * kBlk : g.n;
  int32_t res = (int32_t)(iend - 1);
  for (; i < iend; ++i) {
    int32_t n; uint32_t m;
    weight_split(g.lw[i], n, m);
    c += q_at(m, n, eb);
    if (Pb + (c >> ksh) > tau) { res = (int32_t)i; break; }
  }
  g.anc[j] = res;
}

template <int NX, int NY>
__global__ void __launch_bounds__(kThreads) pf_logw(StepArgs a, float* lw_out) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n_local) return;
  const uint32_t t = a.ctl->t;
  if (a.ctl->uniform_in) { lw_out[i] = 0.0f; return; }
  const float* x = a.X[t & 1];
  const uint32_t tp = (t + kObsRing - 1) % kObsRing;
  float xs[NX], y[NY];
#pragma unroll
  for (int d = 0; d < NX; ++d) xs[d] = x[(size_t)d * a.ld + i];
#pragma unroll
  for (int k = 0; k < NY; ++k) y[k] = a.obs[(size_t)tp * PZ_MAX_NY + k];
  lw_out[i] = lg_logweight<NX, NY>(a.model, xs, y);
}

__global__ void __launch_bounds__(kThreads) pf_gather_thin(StepArgs a, const int32_t* anc, int32_t stride, int64_t n_sel,
                                                           float* out) {
  const int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (s >= n_sel) return;
  const float* x = a.X[a.ctl->t & 1];
  const int64_t p = anc[s * stride];
  for (int d = 0; d < a.model.nx; ++d) out[s * a.model.nx + d] = x[(size_t)d * a.ld + p];
}

// ------------------------------------------------------------------------------------------
// sharded filter: plan from the all-gathered rank records, absolute prefixes, boundary table
// ------------------------------------------------------------------------------------------
struct ShardPlanArgs {
  PlanDev* plan;
  CtlDev* ctl;
  const u64* rec;      // [world][kShardRec]
  const u64* pb;       // local prefix [nblk+1]
  u64* pbx;            // absolute prefix, own entries [0, nblk]
  pz_step_summary* summ;
  int64_t nblk, n_global;
  int32_t rank, world, nx, delta;
  uint64_t seed;
  double ll_const;
};

__global__ void __launch_bounds__(kThreads) pf_plan_sharded(const ShardPlanArgs p) {
  u64 base = 0, total = 0;
  for (int q = 0; q < p.world; ++q) {
    const u64 tq = p.rec[(size_t)q * kShardRec];
    if (q < p.rank) base += tq;
    total += tq;
  }
  const int64_t gid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  for (int64_t b = gid; b <= p.nblk; b += (int64_t)gridDim.x * kThreads) p.pbx[b] = base + p.pb[b];
  if (gid == 0) {
    PlanDev* pl = p.plan;
    const uint32_t t_next = p.ctl->t + (uint32_t)p.delta;
    const uint32_t U = philox4x32_10(0u, 0u, t_next, kStreamResample, (uint32_t)p.seed, (uint32_t)(p.seed >> 32)).x;
    pl->base = base;
    finalize_plan(pl, total, p.n_global, U);
    if (p.delta) {  // moment sums of all ranks, in rank order (every rank computes the same value)
      double sw = 0, sw2 = 0, sx[PZ_MAX_NX] = {}, sxx[PZ_MAX_NX] = {};
      for (int q = 0; q < p.world; ++q) {
        const u64* r = p.rec + (size_t)q * kShardRec;
        sw += __longlong_as_double((long long)r[1]);
        sw2 += __longlong_as_double((long long)r[2]);
        for (int d = 0; d < p.nx; ++d) {
          sx[d] += __longlong_as_double((long long)r[3 + d]);
          sxx[d] += __longlong_as_double((long long)r[3 + PZ_MAX_NX + d]);
        }
      }
      pl->sw = sw; pl->sw2 = sw2;
      for (int d = 0; d < p.nx; ++d) { pl->sx[d] = sx[d]; pl->sxx[d] = sxx[d]; }
      write_summary(pl, p.summ ? p.summ + (p.ctl->t % kObsRing) : nullptr, p.ctl->t, p.nx, p.n_global, p.ll_const);
    }
  }
}

// For every rank q: the global block holding the ancestor of its first slot (entry q) and of
// its last slot (entry world + q), or -1 when that ancestor is not one of this rank's particles.
__global__ void pf_boundary_table(const PlanDev* plan, const u64* pbx, int64_t nblk, int64_t n_local, int64_t n_global,
                                  int32_t rank, int32_t world, int32_t* table) {
  const int q = threadIdx.x;
  if (q >= 2 * world) return;
  SlotMap sm(plan, n_global);
  const uint32_t j = (q < world) ? (uint32_t)(q * n_local) : (uint32_t)((q - world + 1) * n_local - 1);
  int32_t res = -1;
  if (!plan->degenerate && sm.slot_end(pbx[0]) <= j && j < sm.slot_end(pbx[nblk])) {
    int64_t lo = 0, hi = nblk;  // min b with slot_end(pbx[b+1]) > j
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (sm.slot_end(pbx[mid + 1]) > j) hi = mid; else lo = mid + 1;
    }
    res = (int32_t)((int64_t)rank * nblk + lo);
  }
  table[q] = res;
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Output slots per CTA.  Larger tiles amortise the per-tile overlap of the input side (one extra
// input block per tile on average) and the CTA-level reduction; PZ_TILE_OUT=small selects the
// smaller instantiation (kept for A/B measurements, profiles/README.md).
int fused_tile_out(int nx) {
  static const bool small = [] { const char* e = getenv("PZ_TILE_OUT"); return e && !strcmp(e, "small"); }();
  if (nx == 1) return small ? 4096 : 8192;
  return small ? 2048 : 4096;
}

int launch_init(const StepArgs& a, cudaStream_t s) {
  int64_t m = a.nblk * kBlk;
  pf_init<<<cdiv(m, kThreads), kThreads, 0, s>>>(a);
  return 1;
}

static ScanArgs make_scan(const StepArgs& a, int delta, bool with_rec) {
  ScanArgs sc{};
  sc.eb0 = a.eb[0]; sc.eb1 = a.eb[1];
  sc.sb = a.sb; sc.pb = a.pb; sc.status = a.scan_status; sc.part = a.scan_part;
  sc.rec = with_rec ? a.rec : nullptr;
  sc.plan = a.plan; sc.ctl = a.ctl; sc.summ = a.summ;
  sc.nblk = a.nblk; sc.n_out_global = a.n_global; sc.n_local = a.n_local;
  sc.nscan = a.nscan; sc.nrec = a.ntiles_out; sc.nx = a.model.nx; sc.delta = delta;
  sc.explicit_U = 0; sc.U_value = 0; sc.seed = a.seed; sc.ll_const = a.ll_const; sc.world = a.world;
  sc.rank = a.rank; sc.shard_rec = a.shard_rec;
  return sc;
}

static int partition(const StepArgs& a, int tile_out, int delta, cudaStream_t s) {
  PartArgs pa{a.pbx, a.b_lo, a.b_hi, a.tile_start, a.scan_status, a.plan, a.ctl, a.nblk, a.slot_base, a.n_global,
              a.ntiles_out, tile_out, a.nscan, delta};
  int nthreads = a.ntiles_out > a.nscan ? a.ntiles_out : a.nscan;
  pf_partition<<<cdiv(nthreads, kThreads), kThreads, 0, s>>>(pa);
  return 1;
}

int launch_partition_only(const StepArgs& a, int tile_out, cudaStream_t s) { return partition(a, tile_out, 0, s); }

// Scan + partition of the population in buffer (ctl->t + delta) & 1.  delta = 1 after a step.
static int scan_partition(const StepArgs& a, int delta, int tile_out, cudaStream_t s) {
  ScanArgs sc = make_scan(a, delta, delta == 1);
  pf_scan_blocks<<<a.nscan, kThreads, 0, s>>>(sc);
  partition(a, tile_out, delta, s);
  return 2;
}

int launch_scan_partition(const StepArgs& a, int tile_out, cudaStream_t s) {
  return scan_partition(a, 0, tile_out, s);
}

template <int NX, int NY, int TO, int MODE, bool KRON>
static void launch_tile(const TileArgs& ta, cudaStream_t s) {
  auto kern = pf_step_tile<NX, NY, TO, MODE, KRON>;
  size_t smem = (size_t)NX * TO * sizeof(float) + (size_t)TO * sizeof(uint16_t);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  kern<<<ta.a.ntiles_out, kThreads, smem, s>>>(ta);
}

static int dispatch_tile(const TileArgs& ta, cudaStream_t s) {
  const int nx = ta.a.model.nx, ny = ta.a.model.ny;
  const bool fused = (ta.anc == nullptr);
  const bool kron = ta.a.model.kron != 0;
#define PZ_DISPATCH(NXv, NYv, TOv, KR)                               \
  if (nx == NXv && ny == NYv && kron == KR) {                        \
    if (fused) launch_tile<NXv, NYv, TOv, 0, KR>(ta, s);             \
    else launch_tile<NXv, NYv, TOv, 1, KR>(ta, s);                   \
    return 1;                                                        \
  }
  const bool big = fused_tile_out(nx) > (nx == 1 ? 4096 : 2048);
  if (big) {
    PZ_DISPATCH(1, 1, 8192, false)
    PZ_DISPATCH(4, 2, 4096, false)
    PZ_DISPATCH(4, 2, 4096, true)
    PZ_DISPATCH(6, 3, 4096, false)
    PZ_DISPATCH(6, 3, 4096, true)
  } else {
    PZ_DISPATCH(1, 1, 4096, false)
    PZ_DISPATCH(4, 2, 2048, false)
    PZ_DISPATCH(4, 2, 2048, true)
    PZ_DISPATCH(6, 3, 2048, false)
    PZ_DISPATCH(6, 3, 2048, true)
  }
#undef PZ_DISPATCH
  return -1;
}

// anc == nullptr: fused systematic step; else parents come from the stored ancestor array.
int launch_step(const StepArgs& a, const int32_t* anc, cudaStream_t s) {
  TileArgs ta{a, anc};
  if (dispatch_tile(ta, s) < 0) return -1;
  return 1 + scan_partition(a, 1, fused_tile_out(a.model.nx), s);
}

int launch_step_profiled(const StepArgs& a, cudaStream_t s, cudaEvent_t ev[4]) {
  TileArgs ta{a, nullptr};
  cudaEventRecord(ev[0], s);
  if (dispatch_tile(ta, s) < 0) return -1;
  cudaEventRecord(ev[1], s);
  ScanArgs sc = make_scan(a, 1, true);
  pf_scan_blocks<<<a.nscan, kThreads, 0, s>>>(sc);
  cudaEventRecord(ev[2], s);
  partition(a, fused_tile_out(a.model.nx), 1, s);
  cudaEventRecord(ev[3], s);
  return 3;
}

int launch_tile_only(const StepArgs& a, cudaStream_t s) {
  TileArgs ta{a, nullptr};
  return dispatch_tile(ta, s);
}

int launch_scan_only(const StepArgs& a, int delta, cudaStream_t s) {
  ScanArgs sc = make_scan(a, delta, delta == 1);
  pf_scan_blocks<<<a.nscan, kThreads, 0, s>>>(sc);
  return 1;
}

int launch_plan_sharded(const StepArgs& a, int delta, cudaStream_t s) {
  ShardPlanArgs p{a.plan, a.ctl, a.shard_rec, a.pb, const_cast<u64*>(a.pbx), a.summ, a.nblk, a.n_global,
                  a.rank, a.world, a.model.nx, delta, a.seed, a.ll_const};
  int grid = cdiv(a.nblk + 1, kThreads);
  pf_plan_sharded<<<grid > 1024 ? 1024 : grid, kThreads, 0, s>>>(p);
  pf_boundary_table<<<1, 64, 0, s>>>(a.plan, a.pbx, a.nblk, a.n_local, a.n_global, a.rank, a.world, a.shard_table);
  return 2;
}

int launch_partition_ext(const StepArgs& a, int tile_out, int delta, cudaStream_t s) { return partition(a, tile_out, delta, s); }

int launch_ctl_reset(CtlDev* ctl, cudaStream_t s) {
  pf_ctl_reset<<<1, 1, 0, s>>>(ctl);
  return 1;
}

int launch_generic_plan(const GenericArgs& g, bool stats_from_lw, cudaStream_t s) {
  int n = 0;
  if (stats_from_lw) {
    pf_stats_from_lw<<<cdiv(g.nblk, kWarps), kThreads, 0, s>>>(g.lw, g.n, g.nblk, g.eb, g.sb, g.ctl);
    ++n;
  }
  ScanArgs sc{};
  sc.eb0 = g.eb; sc.eb1 = g.eb; sc.sb = g.sb; sc.pb = g.pb; sc.status = g.scan_status; sc.part = g.scan_part; sc.rec = nullptr;
  sc.plan = g.plan; sc.ctl = g.ctl; sc.summ = nullptr; sc.nblk = g.nblk; sc.n_out_global = g.n; sc.n_local = g.n;
  sc.nscan = g.nscan; sc.nrec = 0; sc.nx = 0; sc.delta = 0; sc.explicit_U = 1; sc.U_value = g.U; sc.seed = g.seed;
  sc.ll_const = 0; sc.world = 1; sc.rank = 0; sc.shard_rec = nullptr;
  pf_scan_blocks<<<g.nscan, kThreads, 0, s>>>(sc);
  PartArgs pa{g.pb, 0, g.nblk, g.tile_start, g.scan_status, g.plan, g.ctl, g.nblk, 0, g.n, g.ntiles_out, kGenericTileOut, g.nscan, 0};
  int nthreads = g.ntiles_out > g.nscan ? g.ntiles_out : g.nscan;
  pf_partition<<<cdiv(nthreads, kThreads), kThreads, 0, s>>>(pa);
  return n + 2;
}

int launch_stats_from_lw(const GenericArgs& g, cudaStream_t s) {
  pf_stats_from_lw<<<cdiv(g.nblk, kWarps), kThreads, 0, s>>>(g.lw, g.n, g.nblk, g.eb, g.sb, g.ctl);
  return 1;
}

int launch_generic_ancestors(const GenericArgs& g, cudaStream_t s) {
  pf_ancestors<kGenericTileOut><<<g.ntiles_out, kThreads, 0, s>>>(g);
  return 1;
}

int launch_generic_multinomial(const GenericArgs& g, cudaStream_t s) {
  pf_multinomial<<<cdiv(g.n, kThreads), kThreads, 0, s>>>(g);
  return 1;
}

int launch_logw(const StepArgs& a, float* lw_out, cudaStream_t s) {
  const int nx = a.model.nx, ny = a.model.ny;
  int grid = cdiv(a.n_local, kThreads);
  if (nx == 1 && ny == 1) pf_logw<1, 1><<<grid, kThreads, 0, s>>>(a, lw_out);
  else if (nx == 4 && ny == 2) pf_logw<4, 2><<<grid, kThreads, 0, s>>>(a, lw_out);
  else if (nx == 6 && ny == 3) pf_logw<6, 3><<<grid, kThreads, 0, s>>>(a, lw_out);
  else return -1;
  return 1;
}

int launch_gather_thin(const StepArgs& a, const int32_t* anc, int32_t stride, int64_t n_sel, float* out,
                       cudaStream_t s) {
  pf_gather_thin<<<cdiv(n_sel, kThreads), kThreads, 0, s>>>(a, anc, stride, n_sel, out);
  return 1;
}

}  // namespace pz
