// pz_kernels.cu -- hand-written sm_100a kernels of the streaming particle-filter step.
//
// One time step of zparticles' (haskell/src/Inference.hs:101-104) is three launches:
//
//   pf_step_tile<NX,NY,TO,MODE,KRON,NT>   per output tile of TO slots:
//       input side  -- from the merge-path entry of the tile, read the 16-bit weights of the input
//                      blocks whose children reach the tile (next block prefetched), warp-scan them,
//                      map every particle's cumulative weight to its last child slot with one
//                      32x32+64 multiply-add, and store the particle's state at the shared-memory
//                      slot of its FIRST child (head mark; resample2 + `fst . (xs !!)`,
//                      Inference.hs:57-61, as a systematic resampler);
//       output side -- resolve the head of every slot (one ballot per 128 slots), then per child:
//                      Philox word -> inverse-CDF normal, x' = F x + L z
//                      (noisyIntegrateGen Demo.hs:184-187; SingleTargetTracking.hs:34;
//                      MultiTargetTracking.hs:56-67), the new log-weight (score/factor/observe,
//                      Distributions.hs:29-36,160-161; MVDistributions.hs:45-46), its block
//                      statistics (max exponent, 16-bit weights, their sum), 128-bit coalesced
//                      stores of the state and 64-bit stores of the weights, and the centred moment
//                      sums of the new population (average / weightedAverage, Util/Numeric.hs:7-22).
//   pf_scan_blocks     decoupled look-back exclusive scan of the block sums rescaled to the
//       global exponent (normalize / shiftByMax, Inference.hs:63-66, Util/Numeric.hs:12-17),
//       and, in the last CTA, the resampling plan and the posterior summary.
//   pf_partition       merge-path split (first input block of every output tile) and the slot
//       position of every block boundary.
//
// No tensor cores: the path is a streaming scan / gather; see DESIGN.md for the roofline.
#include <atomic>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "pz_kernels.cuh"
#include "pz_math.cuh"
#include "pz_model.cuh"

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
// system-scope flag traffic of the peer-memory exchange (sharded filter)
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
#ifdef PZ_DBG_STAMPS
// debug build only: %globaltimer stamps of the sharded step's exchange phases (last step wins), printed at destroy
__device__ unsigned long long g_stamp[16];
__device__ __forceinline__ void stamp(int i) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_stamp[i] = t;
}
#define PZ_STAMP(cond, i) do { if (cond) stamp(i); } while (0)
#else
#define PZ_STAMP(cond, i) do { } while (0)
#endif
// poll *p until pred(value); gives up after `limit` clock64 ticks (a dead peer must not hang the device for
// ever; the host sets the limit -- PZ_P2P_TIMEOUT_S, default 120 s, 0 = never -- because the ranks' hosts
// drive their steps independently and may be skewed by seconds: GC pauses, I/O, first-step graph upload).
// ACQUIRE = false for self-validating words (the payload is in the polled word: nothing to order)
// `code` names the wait in the sticky error word (bit mask: 1 exponent all-reduce, 2 record all-gather, 4 state margins,
// 8 boundary-position margins), reported by pz_last_error.
template <bool ACQUIRE = true, class Pred>
__device__ __forceinline__ unsigned long long wait_sys(const unsigned long long* p, Pred pred, uint32_t* err,
                                                       unsigned long long limit, uint32_t code = 1u) {
  const long long t0 = clock64();
  unsigned long long v = ld_sys(p);
  while (!pred(v)) {
    if (limit && (unsigned long long)(clock64() - t0) > limit) { atomicOr(err, code); break; }
    __nanosleep(64);
    v = ld_sys(p);
  }
  if (ACQUIRE) __threadfence_system();
  return v;
}
// 2^k as a float for k <= 0 (0 when it would underflow)
// Programmatic dependent launch (PZ_PDL=1, single-GPU fused step): a kernel launched with the programmatic-serialization
// attribute may become resident while its predecessor in the stream is still running; griddepcontrol.wait holds it until the
// predecessor has completed and its writes are visible, so nothing of the predecessor's output is touched before it.  The
// trigger follows at once: the successor's CTAs then take the SM slots that free up in this kernel's last wave and start the
// moment it ends (launch latency and ramp hidden).  Both instructions are no-ops for a plain launch.
__device__ __forceinline__ void pdl_wait_then_trigger() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// warp maximum in one instruction (CREDUX.MAX.F32, sm_100a): NaN inputs are ignored, as by the fmaxf butterfly it replaces
__device__ __forceinline__ float warp_max_f32(float v) {
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}
__device__ __forceinline__ float exp2i(int k) { return k < -126 ? 0.0f : __int_as_float((127 + k) << 23); }
// 2^k as a double (0 when it would underflow)
__device__ __forceinline__ double pow2d(int k) {
  if (k < -1022) return 0.0;
  return __hiloint2double((1023 + k) << 20, 0);
}

// Slot map of the block boundaries (DESIGN.md "Resampler"):
//   pos(C)      = floor(C * rho / 2^L)            in units of 2^-32 slots, pos(T) = N * 2^32
//   block_pos(C) = pos(C) + 2^32 - 1 - U          its high word = #{ j : j * 2^32 + U < pos(C) }
struct SlotMap {
  u64 rho;
  u64 kadd;      // 0xFFFFFFFF - U
  int lsh;       // 64 - L
  uint32_t n_out;
  __device__ __forceinline__ SlotMap(const PlanDev* pl, int64_t n_global)
      : rho(pl->rho), kadd(0xFFFFFFFFull - pl->U), lsh(64 - pl->L), n_out((uint32_t)n_global) {}
  __device__ __forceinline__ u64 block_pos(u64 C) const { return __umul64hi(C << lsh, rho) + kadd; }
  __device__ __forceinline__ uint32_t slot_end(u64 C) const { return min((uint32_t)(block_pos(C) >> 32), n_out); }
};

// Slots between the end of a particle's children and the end of its block's children, negated:
// floor((A * 2^h - r * M) / 2^(h+32)); r = block mass after the particle, A = low word of the
// block-end position, M / 2^h = 2^-32-slots per unit of block-scale weight.  Generic (rare) form
// for shifts outside [0, 31]; resample_tile inlines the fast form.
__device__ __noinline__ int32_t end_delta_generic(uint32_t r, uint32_t M, uint32_t A, int h) {
  const u64 rm = (u64)r * M;
  long long num;
  if (h >= 0) {
    const u64 t = (h >= 64) ? (rm ? 1ull : 0ull) : ((rm + ((1ull << h) - 1ull)) >> h);  // ceil(r M / 2^h)
    num = (long long)A - (long long)t;
  } else {
    num = (long long)A - (long long)(rm << (-h));
  }
  return (int32_t)(num >> 32);
}

__device__ __forceinline__ long long mad_wide_s32(int32_t a, int32_t b, long long c) {
  long long d;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
}
__device__ __forceinline__ long long mad_wide_u32(uint32_t a, uint32_t b, long long c) {
  long long d;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
}

// normals of the 4 consecutive slots j..j+3 (j % 4 == 0): z[d][k]; one Philox word per normal.  A double
// filter uses the same fp32 table normals, converted exactly.
template <int NX, class T>
__device__ __forceinline__ void lg_noise4(uint32_t j, uint32_t t, uint64_t seed, const float4* __restrict__ tab_m8,
                                          T (&z)[NX][4]) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  if constexpr (NX == 1) {
    const U4 r = philox4x32_10(j >> 2, 0u, t, kStreamProp, k0, k1);
    z[0][0] = (T)icdf_normal(r.x, tab_m8);
    z[0][1] = (T)icdf_normal(r.y, tab_m8);
    z[0][2] = (T)icdf_normal(r.z, tab_m8);
    z[0][3] = (T)icdf_normal(r.w, tab_m8);
  } else if constexpr (NX == 6) {
    // two consecutive slots share three Philox calls: even slot words 0..5, odd slot words 6..11
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      const U4 r0 = philox4x32_10((j >> 1) + pr, 0u, t, kStreamProp, k0, k1);
      const U4 r1 = philox4x32_10((j >> 1) + pr, 1u, t, kStreamProp, k0, k1);
      const U4 r2 = philox4x32_10((j >> 1) + pr, 2u, t, kStreamProp, k0, k1);
      const int s = 2 * pr;
      z[0][s] = (T)icdf_normal(r0.x, tab_m8); z[1][s] = (T)icdf_normal(r0.y, tab_m8);
      z[2][s] = (T)icdf_normal(r0.z, tab_m8); z[3][s] = (T)icdf_normal(r0.w, tab_m8);
      z[4][s] = (T)icdf_normal(r1.x, tab_m8); z[5][s] = (T)icdf_normal(r1.y, tab_m8);
      z[0][s + 1] = (T)icdf_normal(r1.z, tab_m8); z[1][s + 1] = (T)icdf_normal(r1.w, tab_m8);
      z[2][s + 1] = (T)icdf_normal(r2.x, tab_m8); z[3][s + 1] = (T)icdf_normal(r2.y, tab_m8);
      z[4][s + 1] = (T)icdf_normal(r2.z, tab_m8); z[5][s + 1] = (T)icdf_normal(r2.w, tab_m8);
    }
  } else {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
#pragma unroll
      for (int c = 0; 4 * c < NX; ++c) {
        const U4 r = philox4x32_10(j + s, (uint32_t)c, t, kStreamProp, k0, k1);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (4 * c + i < NX) z[4 * c + i][s] = (T)icdf_normal(w[i], tab_m8);
      }
    }
  }
}

// 4 consecutive state scalars: one 128-bit (float) or 256-bit (double) access
template <class T> struct Quad { T v[4]; };
__device__ __forceinline__ Quad<float> ld_quad(const float* p) {
  const float4 f = *reinterpret_cast<const float4*>(p);
  return Quad<float>{{f.x, f.y, f.z, f.w}};
}
__device__ __forceinline__ Quad<double> ld_quad(const double* p) {
  Quad<double> q;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(q.v[0]), "=d"(q.v[1]), "=d"(q.v[2]), "=d"(q.v[3]) : "l"(p));
  return q;
}
__device__ __forceinline__ void st_quad(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st_quad(double* p, const double (&v)[4]) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
// shared-memory quads
__device__ __forceinline__ Quad<float> lds_quad(const float* p) {
  const float4 f = *reinterpret_cast<const float4*>(p);
  return Quad<float>{{f.x, f.y, f.z, f.w}};
}
__device__ __forceinline__ Quad<double> lds_quad(const double* p) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  return Quad<double>{{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ void sts_quad(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void sts_quad(double* p, const double (&v)[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}
// the 32-bit word of a staged scalar that carries the head-mark sentinel (the float itself / the double's high word)
__device__ __forceinline__ uint32_t mark_word(float v) { return __float_as_uint(v); }
__device__ __forceinline__ uint32_t mark_word(double v) { return (uint32_t)__double2hiint(v); }

// ------------------------------------------------------------------------------------------
// pf_init: the point-mass population before the first observation, with equal weights
// (lw = 0 for every particle: n_b = 0, w = 1, q = 2^15)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) pf_init(StepArgs a) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < a.nblk * kBlk) {
    for (int d = 0; d < a.model.nx; ++d) {
      if (a.f64) {
        ((double*)a.X[0])[(size_t)d * a.ld + i] = a.modeld.x0[d];
        ((double*)a.X[1])[(size_t)d * a.ld + i] = a.modeld.x0[d];
      } else {
        ((float*)a.X[0])[(size_t)d * a.ld + i] = a.model.x0[d];
        ((float*)a.X[1])[(size_t)d * a.ld + i] = a.model.x0[d];
      }
    }
    a.Q[0][i] = i < a.n_local ? (uint16_t)32768 : (uint16_t)0;
    a.Q[1][i] = 0;
  }
  if (i < a.nblk) {
    int64_t cnt = a.n_local - i * kBlk;
    cnt = cnt > kBlk ? kBlk : cnt;
    a.eb[0][i] = 0;
    a.eb[1][i] = kNZero;
    a.sb[i] = (u64)cnt << 15;
  }
  if (i < a.nscan) a.scan_status[i] = 0ull;
  if (i == 0) {
    a.ctl->t = 0;
    a.ctl->emax_acc = 0;
    a.ctl->scan_counter = 0;
    a.ctl->scan_done = 0;
    a.ctl->comm_error = 0;
    a.ctl->ess_resampled = 1;
    a.ctl->pad[0] = a.ctl->pad[1] = 0;
    for (int d = 0; d < 8; ++d) {
      a.plan->center[d] = d < a.model.nx ? a.model.x0[d] : 0.0f;
      a.plan->center64[d] = d < a.model.nx ? a.modeld.x0[d] : 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// pf_scan_blocks: single-pass decoupled look-back scan over the block sums
// ------------------------------------------------------------------------------------------
constexpr u64 kFlagAgg = 1ull << 62, kFlagIncl = 2ull << 62, kValMask = (1ull << 62) - 1;

__device__ __forceinline__ void copy16(void* dst, const void* src, int64_t n16, int64_t gid, int64_t gsz) {
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint4* s = reinterpret_cast<const uint4*>(src);
  for (int64_t i = gid; i < n16; i += gsz) d[i] = s[i];
}

// Peer-memory exchange: the X / Q / eb margins of the population just produced go straight into the
// neighbours' halo regions (their arrays have our layout).  Independent of the scan, so extra CTAs of
// pf_scan_blocks do it beside the scan (no forked stream, no join: the step stays a linear chain of launches);
// the last of them to finish raises the neighbours' flags.
struct HaloPushArgs {
  const CtlDev* ctl;
  const ShardP2P* p2p;
  const void* X[2];
  int32_t esz;  // bytes per state scalar
  const uint16_t* Q[2];
  const int32_t* eb[2];
  int64_t nblk, ld, margin;
  int32_t rank, world, nx, delta;
};
__device__ __forceinline__ void halo_push_body(const HaloPushArgs& p, int cta, int ncta) {
  const int tid = threadIdx.x;
  PZ_STAMP(cta == 0 && tid == 0, 12);
  const uint32_t epoch = p.ctl->t + (uint32_t)p.delta + 1u;
  const int which = (int)((p.ctl->t + (uint32_t)p.delta) & 1u);
  const int64_t gid = (int64_t)cta * kThreads + tid, gsz = (int64_t)ncta * kThreads;
  const int64_t m = p.margin, nb = p.nblk;
  const size_t es = (size_t)p.esz;
  const char* xmine = reinterpret_cast<const char*>(p.X[which]);
  if (p.rank > 0) {  // my first m blocks: the RIGHT halo of rank-1
    char* xn = reinterpret_cast<char*>(p.p2p->Xn[0][which]);
    for (int d = 0; d < p.nx; ++d)
      copy16(xn + ((size_t)d * p.ld + nb * kBlk) * es, xmine + (size_t)d * p.ld * es, m * kBlk * p.esz / 16, gid, gsz);
    copy16(p.p2p->Qn[0][which] + nb * kBlk, p.Q[which], m * kBlk / 8, gid, gsz);
    for (int64_t i = gid; i < m; i += gsz) p.p2p->ebn[0][which][nb + i] = p.eb[which][i];
  }
  if (p.rank + 1 < p.world) {  // my last m blocks: the LEFT halo of rank+1
    char* xn = reinterpret_cast<char*>(p.p2p->Xn[1][which]);
    for (int d = 0; d < p.nx; ++d)
      copy16(xn + ((ptrdiff_t)d * p.ld - m * kBlk) * (ptrdiff_t)es, xmine + ((size_t)d * p.ld + (nb - m) * kBlk) * es, m * kBlk * p.esz / 16, gid, gsz);
    copy16(p.p2p->Qn[1][which] - m * kBlk, p.Q[which] + (nb - m) * kBlk, m * kBlk / 8, gid, gsz);
    for (int64_t i = gid; i < m; i += gsz) p.p2p->ebn[1][which][i - m] = p.eb[which][nb - m + i];
  }
  // the CTA barrier orders every thread's peer stores before thread 0's fence, and a fence is cumulative: one
  // system-scope fence per CTA (75 000 of them, one per thread, cost ~10 us of the step)
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    Mailbox* mine = p.p2p->mbox[p.rank];
    if (atomicAdd(&mine->push_done, 1ull) == (unsigned long long)ncta - 1ull) {
      __threadfence_system();
      mine->push_done = 0ull;
      PZ_STAMP(true, 13);
      if (p.rank > 0) st_sys(&p.p2p->mbox[p.rank - 1]->halo_flag[1], (unsigned long long)epoch);
      if (p.rank + 1 < p.world) st_sys(&p.p2p->mbox[p.rank + 1]->halo_flag[0], (unsigned long long)epoch);
    }
  }
}


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
  int64_t margin;  // sharded: halo margin in blocks (the record carries the prefix at both margin edges)
  const ShardP2P* p2p;  // sharded: exchange the block exponent through peer memory (nullptr: NCCL did it)
  int32_t f64;
  unsigned long long wait_ticks;
  HaloPushArgs push;   // sharded over peer memory: CTAs [nscan, nscan + npush) push the margins
  int32_t npush;
};

__device__ void finalize_plan(PlanDev* pl, u64 total, int64_t n_out_global, uint32_t U) {
  pl->total = total;
  pl->degenerate = (total == 0);
  pl->U = U;
  if (total != 0) {
    const int L = 64 - __clzll((long long)total);
    pl->L = L;
    pl->rho = (u64)((((unsigned __int128)(u64)n_out_global) << (32 + L)) / total) + 1ull;
    const int lam = (64 - __clzll((long long)pl->rho)) - 30;
    pl->Mq = (uint32_t)(pl->rho >> lam) + 1u;
    pl->h0 = L - kGuard - lam;
  } else {
    pl->L = 1;
    pl->rho = 0;
    pl->Mq = 0;
    pl->h0 = 0;
  }
}

// posterior summary from the centred moment sums in the plan (also re-centres for the next step)
__device__ void write_summary(PlanDev* pl, pz_step_summary* o, uint32_t t, int nx, int64_t n_out_global, double ll_const, int f64) {
  const double sw = pl->sw, sw2 = pl->sw2;
  for (int d = 0; d < nx; ++d) {
    const double cen = f64 ? pl->center64[d] : (double)pl->center[d];
    const double m = pl->sx[d] / sw;
    const double mean = cen + m, var = pl->sxx[d] / sw - m * m;
    pl->center[d] = (float)mean;
    pl->center64[d] = mean;
    if (o) { o->mean[d] = mean; o->var[d] = var; }
  }
  if (o) {
    for (int d = nx; d < PZ_MAX_NX; ++d) { o->mean[d] = 0.0; o->var[d] = 0.0; }
    o->step = t;
    o->log_evidence_inc = log(sw) + (double)pl->emax * 0.6931471805599453 - log((double)n_out_global) + ll_const;
    o->ess = sw * sw / sw2;
    o->n_migrated = 0;
    o->e_max = pl->emax;
    o->reserved = 0;
    o->total_weight = pl->total;
  }
}

__global__ void __launch_bounds__(kThreads) pf_scan_blocks(const ScanArgs s) {
  pdl_wait_then_trigger();
  __shared__ u64 s_warp[kWarps];
  __shared__ u64 s_prefix;
  __shared__ uint32_t s_tile, s_last;
  __shared__ double s_red[kWarps][kMaxRec];
  __shared__ double s_tot[kMaxRec];
  __shared__ double s_grp[kThreads / 16][16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((int)blockIdx.x >= s.nscan) {   // a margin-push CTA (uniform per CTA)
    halo_push_body(s.push, (int)blockIdx.x - s.nscan, s.npush);
    return;
  }
  if (tid == 0) s_tile = atomicAdd(&s.ctl->scan_counter, 1u);
  int32_t emax = s.ctl->emax_acc;            // in flight together with the tile-id atomic
  const uint32_t t_pop = s.ctl->t + (uint32_t)s.delta;
  __syncthreads();
  const uint32_t tile = s_tile;
  PZ_STAMP(tile == 0 && tid == 0, 0);
  if (s.p2p) {
    // all-reduce(max) of the block exponent over peer memory: the first CTA stores this rank's value
    // into every mailbox, every CTA polls its own mailbox for all ranks' values of this epoch
    __shared__ int32_t s_emax;
    const unsigned long long epoch = (unsigned long long)(t_pop + 1u);
    if (warp == 0) {
      if (tile == 0 && lane < s.world) st_sys(&s.p2p->mbox[lane]->emax[s.rank], (epoch << 32) | (uint32_t)emax);
      int32_t v = kNZero;
      if (lane < s.world) {
        const unsigned long long w = wait_sys<false>(&s.p2p->mbox[s.rank]->emax[lane],
                                              [epoch](unsigned long long x) { return (x >> 32) == epoch; }, &s.ctl->comm_error,
                                              s.wait_ticks, 1u);
        v = (int32_t)(uint32_t)w;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
      if (lane == 0) s_emax = v;
    }
    __syncthreads();
    emax = s_emax;
    PZ_STAMP(tile == 0 && tid == 0, 1);
  }
  const int32_t* __restrict__ ebp = (t_pop & 1u) ? s.eb1 : s.eb0;

  // ---- this CTA's slice of the moment records: loads issued now, folded while the look-back waits
  const int R = 3 + 2 * s.nx;
  const bool have_rec = (s.rec != nullptr) && (s.nrec > 0) && (s.delta != 0);
  const int rpc = have_rec ? (s.nrec + s.nscan - 1) / s.nscan : 0;
  double rv[kMaxRec];
  bool rec_live = false;
  if (have_rec && tid < rpc) {
    const int64_t J = (int64_t)tile * rpc + tid;
    if (J < s.nrec) {
      rec_live = true;
      const double* rc = s.rec + (size_t)J * kMaxRec;
#pragma unroll
      for (int r = 0; r < kMaxRec; ++r) rv[r] = (r < R) ? rc[r] : 0.0;
    }
  }

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
      if (e != kNZero) S = shift_guard(s.sb[b], emax - e);
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

  // ---- publish this tile's aggregate at once (successors can add it without waiting for our prefix)
  volatile u64* st = s.status;
  if (tid == 0) st[tile] = (tile == 0 ? kFlagIncl : kFlagAgg) | agg;

  // ---- moment records -> per-CTA partial at the global exponent (overlaps the predecessors' progress)
  if (have_rec) {
    double acc[kMaxRec];
#pragma unroll
    for (int r = 0; r < kMaxRec; ++r) acc[r] = 0.0;
    if (rec_live && rv[0] > -1.0e9) {
      const double f = pow2d((int)rv[0] - emax);
#pragma unroll
      for (int r = 1; r < kMaxRec; ++r) acc[r] = rv[r] * (r == 2 ? f * f : f);
    }
    for (int q = tid + kThreads; q < rpc; q += kThreads) {  // more than 256 records per CTA: rare
      const int64_t J = (int64_t)tile * rpc + q;
      if (J < s.nrec) {
        const double* rc = s.rec + (size_t)J * kMaxRec;
        double v2[kMaxRec];
#pragma unroll
        for (int r = 0; r < kMaxRec; ++r) v2[r] = (r < R) ? rc[r] : 0.0;
        if (v2[0] > -1.0e9) {
          const double f = pow2d((int)v2[0] - emax);
#pragma unroll
          for (int r = 1; r < kMaxRec; ++r) acc[r] += v2[r] * (r == 2 ? f * f : f);
        }
      }
    }
    for (int r = 1; r < R; ++r) {
      double x = acc[r];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += shfl_xor_f64(x, d);
      if (lane == 0) s_red[warp][r] = x;
    }
  }

  // ---- decoupled look-back (warp 0)
  if (warp == 0) {
    u64 prefix = 0;
    if (tile != 0) {
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
  if (have_rec && tid >= 1 && tid < R) {  // (s_red was completed before the barrier above)
    double x = 0;
    for (int w = 0; w < kWarps; ++w) x += s_red[w][tid];
    s.part[(size_t)tile * kMaxRec + tid] = x;
  }

  // ---- last CTA: plan + posterior summary
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&s.ctl->scan_done, 1u) == (uint32_t)(s.nscan - 1));
  __syncthreads();
  if (!s_last) return;
  PZ_STAMP(tid == 0, 2);
  __threadfence();
  if (have_rec) {  // per-CTA partials -> totals: 16 groups of 16 threads, fixed order (deterministic)
    const int r = tid & 15, grp = tid >> 4;
    double x = 0;
    if (r >= 1 && r < R) {
      // batches of 8 independent L2 loads (a load-add-load-add chain would serialise on the L2 latency)
      constexpr int G = kThreads / 16;
      int c = grp;
      for (; c + 7 * G < s.nscan; c += 8 * G) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(s.part + (size_t)(c + u * G) * kMaxRec + r);
#pragma unroll
        for (int u = 0; u < 8; ++u) x += v[u];
      }
      for (; c < s.nscan; c += G) x += __ldcg(s.part + (size_t)c * kMaxRec + r);
    }
    s_grp[grp][r] = x;
    __syncthreads();
    if (tid >= 1 && tid < R) {
      double y = 0;
      for (int q = 0; q < kThreads / 16; ++q) y += s_grp[q][tid];
      s_tot[tid] = y;
    }
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
        write_summary(pl, s.summ ? s.summ + (s.ctl->t % kObsRing) : nullptr, s.ctl->t, s.nx, s.n_out_global, s.ll_const, s.f64);
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
      // local prefix where the margin this rank sends to its left / right neighbour ends / begins
      const int64_t m = s.margin < s.nblk ? s.margin : s.nblk;
      r[16] = ((volatile u64*)s.pb)[m];
      r[17] = ((volatile u64*)s.pb)[s.nblk - m];
      r[18] = (u64)((volatile CtlDev*)s.ctl)->comm_error;  // a timeout seen by this rank: every rank reports it
    }
  }
}

// ------------------------------------------------------------------------------------------
// pf_partition: first input block of every output tile, slot position of every block boundary,
// and the per-step counter maintenance
// ------------------------------------------------------------------------------------------
struct PartArgs {
  const u64* pb;      // prefix array searched (absolute on a sharded filter)
  u64* xb;            // out: block_pos(pb[b]) for b in [b_lo, b_hi]
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
  const ShardP2P* p2p;  // sharded over peer memory: wait for the neighbours' halo stores first
  int32_t rank, world;
  unsigned long long wait_ticks;
};

__global__ void __launch_bounds__(kThreads) pf_partition(const PartArgs p) {
  pdl_wait_then_trigger();
  const int64_t J = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const PlanDev* pl = p.plan;
  PZ_STAMP(J == 0, 9);
  if (p.p2p) {
    if (threadIdx.x < 4) {  // both neighbours' margins: X / Q / eb (pf_halo_push) and prefixes (pf_plan_sharded)
      const int side = threadIdx.x & 1;
      const bool have = side == 0 ? p.rank > 0 : p.rank + 1 < p.world;
      const unsigned long long epoch = pl->shard_epoch;
      Mailbox* mine = p.p2p->mbox[p.rank];
      if (have)
        wait_sys(threadIdx.x < 2 ? &mine->halo_flag[side] : &mine->pbx_flag[side], [epoch](unsigned long long x) { return x >= epoch; },
                 &p.ctl->comm_error, p.wait_ticks, threadIdx.x < 2 ? 4u : 8u);
    }
    __syncthreads();
    PZ_STAMP(J == 0, 10);
  }
  SlotMap sm(pl, p.n_out_global);
  // one warp per output tile: 32-ary search for min b in [b_lo, b_hi) with slot_end(pb[b+1]) > j0
  // (b_hi when there is none); ~4 dependent rounds of loads instead of ~19
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t tile = J >> 5; tile < p.ntiles_out; tile += nwarps) {
    const uint32_t j0 = (uint32_t)(p.slot_base + tile * p.tile_out);
    int64_t lo = p.b_lo, hi = p.b_hi;
    while (hi > lo) {
      const int64_t n = hi - lo, step = (n + 31) >> 5;
      int64_t off = (int64_t)(lane + 1) * step;
      off = off < n ? off : n;
      const int64_t mid = lo + off - 1;
      const bool pred = sm.slot_end(p.pb[mid + 1]) > j0;
      const uint32_t m = __ballot_sync(0xffffffffu, pred);
      if (m == 0) { lo = hi; break; }
      const int f = __ffs(m) - 1;
      int64_t offf = (int64_t)(f + 1) * step;
      offf = offf < n ? offf : n;
      hi = lo + offf - 1;
      lo = lo + (int64_t)f * step;
    }
    if (lane == 0) p.tile_start[tile] = (int32_t)lo;
  }
  for (int64_t b = p.b_lo + J; b <= p.b_hi; b += (int64_t)gridDim.x * kThreads) p.xb[b] = sm.block_pos(p.pb[b]);
  if (J < p.nscan) p.status[J] = 0ull;
  if (J == 0) {
    p.ctl->t += (uint32_t)p.delta;
    p.ctl->emax_acc = kNZero;
    p.ctl->scan_counter = 0;
    p.ctl->scan_done = 0;
  }
}

// ------------------------------------------------------------------------------------------
// resample phase shared by the step and the ancestor kernels
// ------------------------------------------------------------------------------------------
// Head marks: the input side stores a particle's payload at the slot of its FIRST child only (and
// again at every 128-slot segment boundary the run of children crosses, so that every segment
// starts with a head); all other slots keep the sentinel written before the input phase.  The
// output side resolves the parent of every slot with one ballot per 128 slots (resolve_heads).
// Work is balanced whatever the weight distribution: a particle with 10^6 children costs one
// store per segment on the input side and nothing extra on the output side.
constexpr uint32_t kSentinel = 0x7FC0DEADu;  // a NaN payload no arithmetic produces; index payloads use it too

// Payload sources: what is staged for the children of a particle.
template <int NX, class T>
struct StateSource {
  typedef T Scalar;
  static constexpr int NV = NX;
  static constexpr bool kPrefetch = (NX == 1);  // one word per particle: fetched one block ahead (registers)
  const T* x;  // X[cur]
  int64_t ld;
  T v[NX][4];
  __device__ __forceinline__ Quad<T> fetch(int64_t i0) const { return ld_quad(x + i0); }
  __device__ __forceinline__ void set(const Quad<T>& f) { v[0][0] = f.v[0]; v[0][1] = f.v[1]; v[0][2] = f.v[2]; v[0][3] = f.v[3]; }
  __device__ __forceinline__ void load(int64_t i0) {
#pragma unroll
    for (int d = 0; d < NX; ++d) {
      const Quad<T> f = ld_quad(x + (size_t)d * ld + i0);
      v[d][0] = f.v[0]; v[d][1] = f.v[1]; v[d][2] = f.v[2]; v[d][3] = f.v[3];
    }
  }
  __device__ __forceinline__ void put(T* s_par, int stride, int slot, int k) const {
    PZ_ASSERT(slot >= 0 && slot < stride);
#pragma unroll
    for (int d = 0; d < NX; ++d) s_par[d * stride + slot] = v[d][k];
  }
  // dimension 0 only (the rare long-run duplicates of the NX = 1 look-back variant)
  __device__ __forceinline__ T first(int k) const { return v[0][k]; }
};

struct IndexSource {  // generic path: payload = particle index
  typedef float Scalar;
  static constexpr int NV = 1;
  static constexpr bool kPrefetch = false;
  float v[1][4];
  __device__ __forceinline__ Quad<float> fetch(int64_t) const { return Quad<float>{{0.f, 0.f, 0.f, 0.f}}; }
  __device__ __forceinline__ void set(const Quad<float>&) {}
  __device__ __forceinline__ void load(int64_t i0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[0][k] = __int_as_float((int32_t)i0 + k);
  }
  __device__ __forceinline__ void put(float* s_par, int stride, int slot, int k) const {
    PZ_ASSERT(slot >= 0 && slot < stride);
    s_par[slot] = v[0][k];
  }
  __device__ __forceinline__ float first(int k) const { return v[0][k]; }
};

// SEGDUP = true: a head at every segment boundary any run crosses (every segment starts with a head).
// SEGDUP = false: only runs longer than 128 slots are duplicated, so a segment whose first slots
// have no head finds it within the previous segment (output side look-back, NX = 1).
template <bool SEGDUP, class Source>
__device__ __forceinline__ void emit_head(const Source& src, typename Source::Scalar* s_par, int stride, int lo, int hi, int k) {
  if (hi > lo) {
    src.put(s_par, stride, lo, k);
    if constexpr (SEGDUP) {
      if (((lo ^ (hi - 1)) >> 7) != 0)
        for (int s = (lo | 127) + 1; s < hi; s += 128) src.put(s_par, stride, s, k);
    }
  }
}
// look-back variant (one word per particle): duplicates of the (rare) runs longer than 128 slots, outside the
// per-particle fast path.  The payload travels by value, so the caller's sources stay in registers.
template <class S>
__device__ __noinline__ void emit_long_runs(S payload, S* s_par, int lo, int hi) {
  for (int s = (lo | 127) + 1; s < hi; s += 128) { PZ_ASSERT(s >= 0); s_par[s] = payload; }
}

// Input side of one output tile [j0, j0 + nslots): every warp takes input blocks b_first + warp,
// + 8, ...  A lane owns 8 particles of the block -- memory offsets lane*4 + {0..3} and
// 128 + lane*4 + {0..3} -- which are 8 CONSECUTIVE positions of the block's CDF order (the spec's
// order: lane-major, so that one 32-bit warp scan of the lane totals suffices).  With r = the
// block mass after a particle, the particle's children end at
//     e = max(o1 + floor((A * 2^h - r * M) / 2^(h+32)), o0)
// (o0 / o1: first / end slot of the block's children, A: fractional slot position of the block
// end): one mad.wide.s32 and one arithmetic shift per particle.  Measuring from the block END
// makes zero-weight particles childless by construction and the last weight reach o1 exactly.
// per-block inputs of the input side (prefetched one iteration ahead)
struct BlockIn {
  u64 X0, X1;
  uint2 qa, qb;
  int32_t ebb;
  __device__ __forceinline__ void load(const uint16_t* __restrict__ qin, const int32_t* __restrict__ ebx,
                                       const u64* __restrict__ xb, int64_t b, int lane) {
    X0 = xb[b]; X1 = xb[b + 1];
    ebb = ebx[b];
    const int64_t ia = b * kBlk + lane * 4;
    qa = *reinterpret_cast<const uint2*>(qin + ia);
    qb = *reinterpret_cast<const uint2*>(qin + ia + 128);
  }
};

// One input block of the input side (see resample_tile).  Returns true when this and every later block start
// past the tile.
template <bool SEGDUP, class Source>
__device__ __forceinline__ bool resample_block(const Source& src, typename Source::Scalar* s_par, int stride, const BlockIn& in,
                                               const Quad<typename Source::Scalar>& xa, const Quad<typename Source::Scalar>& xb4,
                                               int64_t b, int32_t emax, int32_t h0, uint32_t Mq, uint32_t j0, int nslots, int lane) {
  constexpr bool kAhead = Source::kPrefetch;
  const uint32_t jend = j0 + (uint32_t)nslots;
  const uint32_t o0 = (uint32_t)(in.X0 >> 32), o1 = (uint32_t)(in.X1 >> 32);
  if (o0 >= jend) return true;                // this and every later block start past the tile
  if (o1 == o0 || o1 <= j0) return false;     // no children (every empty block), or none in this tile
  const int64_t ia = b * kBlk + lane * 4;
  Source sa = src, sb = src;
  if constexpr (kAhead) { sa.set(xa); sb.set(xb4); }
  else { sa.load(ia); sb.load(ia + 128); }  // payload of the lane's 8 particles, in flight during the scan
  const int h = h0 + (emax - in.ebb);
  uint32_t c[8];
  c[0] = in.qa.x & 0xffffu;
  c[1] = c[0] + (in.qa.x >> 16);
  c[2] = c[1] + (in.qa.y & 0xffffu);
  c[3] = c[2] + (in.qa.y >> 16);
  c[4] = c[3] + (in.qb.x & 0xffffu);
  c[5] = c[4] + (in.qb.x >> 16);
  c[6] = c[5] + (in.qb.y & 0xffffu);
  c[7] = c[6] + (in.qb.y >> 16);
  const uint32_t incl = warp_inclusive_scan_u32(c[7], lane);
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  const uint32_t rbase = total - (incl - c[7]);  // block mass from this lane's first particle on
  const int32_t o0r = (int32_t)(o0 - j0), o1r = (int32_t)(o1 - j0);
  const uint32_t A = (uint32_t)in.X1;
  int32_t e[8];
  if ((unsigned)h <= 31u) {
    // (c - rbase) * M + A * 2^h as c * M + base: one 32x32+64 multiply-add per particle
    const long long base = mad_wide_s32(-(int32_t)rbase, (int32_t)Mq, (long long)((u64)A << h));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long res = mad_wide_u32(c[k], Mq, base);
      e[k] = max(o1r + ((int32_t)(res >> 32) >> h), o0r);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = max(o1r + end_delta_generic(rbase - c[k], Mq, A, h), o0r);
  }
  int32_t eprev = __shfl_up_sync(0xffffffffu, e[7], 1);
  if (lane == 0) eprev = o0r;
  if (o0r < 0 || o1r > nslots) {  // the block's children straddle a tile edge: clip
    eprev = min(max(eprev, 0), nslots);
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = min(max(e[k], 0), nslots);
  }
  const int32_t eprev0 = eprev;
  PZ_ASSERT(eprev0 >= 0 && e[7] <= nslots && e[7] >= eprev0);
#pragma unroll
  for (int k = 0; k < 4; ++k) { emit_head<SEGDUP>(sa, s_par, stride, eprev, e[k], k); eprev = e[k]; }
#pragma unroll
  for (int k = 0; k < 4; ++k) { emit_head<SEGDUP>(sb, s_par, stride, eprev, e[4 + k], k); eprev = e[4 + k]; }
  if constexpr (!SEGDUP) {
    // a run longer than 128 slots needs a lane whose 8 particles cover more than 128 slots: one vote per block
    if (__any_sync(0xffffffffu, e[7] - eprev0 > 128)) {
      int32_t ep = eprev0;
#pragma unroll
      for (int k = 0; k < 4; ++k) { if (e[k] - ep > 128) emit_long_runs(sa.first(k), s_par, ep, e[k]); ep = e[k]; }
#pragma unroll
      for (int k = 0; k < 4; ++k) { if (e[4 + k] - ep > 128) emit_long_runs(sb.first(k), s_par, ep, e[4 + k]); ep = e[4 + k]; }
    }
  }
  return false;
}

// The warp's input blocks b_first + warp, + NW, ...: the inputs of the next block (weights, boundary positions,
// and narrow payloads) are loaded one block ahead.  (A two-register-set unrolling that avoids the ~20 register
// moves per block was measured slower: it spills at 64 registers.)
template <bool SEGDUP, int NW, class Source>
__device__ __forceinline__ void resample_tile(Source& src, typename Source::Scalar* s_par, int stride, const uint16_t* __restrict__ qin,
                                              const int32_t* __restrict__ ebx, const u64* __restrict__ xb, int64_t b_first,
                                              int64_t b_end, int32_t emax, int32_t h0, uint32_t Mq, uint32_t j0, int nslots,
                                              int lane, int warp, BlockIn cur, Quad<typename Source::Scalar> pa,
                                              Quad<typename Source::Scalar> pb) {
  // `cur` / `pa`, `pb` hold the inputs / narrow payload of block b_first + warp (loaded by the caller before its
  // prologue barrier, so that their DRAM latency overlaps the prologue)
  int64_t b = b_first + warp;
  if (b >= b_end) return;
  constexpr bool kAhead = Source::kPrefetch;
  typedef Quad<typename Source::Scalar> QuadS;
  for (; b < b_end; b += NW) {
    const BlockIn in = cur;
    const QuadS xa = pa, xb4 = pb;
    const int64_t bn = b + NW;
    if (bn < b_end) {  // next iteration's inputs: in flight during this one
      cur.load(qin, ebx, xb, bn, lane);
      if constexpr (kAhead) { pa = src.fetch(bn * kBlk + lane * 4); pb = src.fetch(bn * kBlk + lane * 4 + 128); }
    }
    if (resample_block<SEGDUP>(src, s_par, stride, in, xa, xb4, b, emax, h0, Mq, j0, nslots, lane)) break;
  }
}

// Input side of the REMOTE step kernel (a sharded filter whose pending resample reaches beyond the halo margins):
// the tile's input blocks are GLOBAL block indices; each is read from its owner's arrays over NVLink (or from this
// rank's own).  No prefetch pipeline: this is the fallback path, it only has to be right.
template <bool SEGDUP, int NW, class Source>
__device__ __forceinline__ void resample_tile_remote(Source& src, typename Source::Scalar* s_par, int stride, const ShardP2P* p2p,
                                                     int cur, int64_t nblk, int world, int64_t gb_first, int32_t emax, int32_t h0,
                                                     uint32_t Mq, uint32_t j0, int nslots, int lane, int warp) {
  typedef typename Source::Scalar S;
  const int64_t gb_end = (int64_t)world * nblk;
  for (int64_t gb = gb_first + warp; gb < gb_end; gb += NW) {
    const int q = (int)(gb / nblk);
    const int64_t lb = gb - (int64_t)q * nblk;
    BlockIn in;
    in.load(p2p->Qall[q][cur], p2p->eball[q][cur], p2p->xball[q], lb, lane);
    Quad<S> xa{}, xb4{};
    if constexpr (std::is_same<Source, IndexSource>::value) {
      // index staging: the GLOBAL particle index is staged; the output side reads the parent from its owner
      if (resample_block<SEGDUP>(src, s_par, stride, in, xa, xb4, gb, emax, h0, Mq, j0, nslots, lane)) break;
    } else {
      src.x = reinterpret_cast<const S*>(p2p->Xall[q][cur]);
      if constexpr (Source::kPrefetch) { xa = src.fetch(lb * kBlk + lane * 4); xb4 = src.fetch(lb * kBlk + lane * 4 + 128); }
      if (resample_block<SEGDUP>(src, s_par, stride, in, xa, xb4, lb, emax, h0, Mq, j0, nslots, lane)) break;
    }
  }
}

// Output side: tile-relative slot of the head of each of the lane's 4 slots sl .. sl+3 (sl = segment
// base + lane*4; the segment = the 128 slots this warp iteration covers).  bits = the payload words
// of dimension 0 at those slots.  The first slot of a segment always carries a head (emit_head).
// seg_carry: head of the slots before the segment's first head (-1 when every segment starts with one);
// returns the segment's last head.
__device__ __forceinline__ int resolve_heads(const uint4& bits, int sl, int lane, int seg_carry, int (&head)[4]) {
  const int m0 = bits.x != kSentinel ? sl : -1;
  const int m1 = bits.y != kSentinel ? sl + 1 : m0;
  const int m2 = bits.z != kSentinel ? sl + 2 : m1;
  const int m3 = bits.w != kSentinel ? sl + 3 : m2;
  const uint32_t has = __ballot_sync(0xffffffffu, m3 >= 0);
  const uint32_t before = has & ((1u << lane) - 1u);
  const int src = 31 - __clz(before);            // nearest earlier lane with a head
  const int got = __shfl_sync(0xffffffffu, m3, src & 31);
  const int carry = before ? got : seg_carry;
  head[0] = max(m0, carry); head[1] = max(m1, carry); head[2] = max(m2, carry); head[3] = max(m3, carry);
  const int last = __shfl_sync(0xffffffffu, m3, (31 - __clz(has)) & 31);
  return has ? last : seg_carry;
}

// last head of the 128-slot segment ending at slot seg_end (look-back of the NX = 1 path)
template <class T>
__device__ __forceinline__ int last_head_before(const T* s_par, int seg_end, int lane) {
  const int sl = seg_end - 128 + lane * 4;
  const Quad<T> qd = lds_quad(&s_par[sl]);
  const uint4 bits = make_uint4(mark_word(qd.v[0]), mark_word(qd.v[1]), mark_word(qd.v[2]), mark_word(qd.v[3]));
  const int m0 = bits.x != kSentinel ? sl : -1;
  const int m1 = bits.y != kSentinel ? sl + 1 : m0;
  const int m2 = bits.z != kSentinel ? sl + 2 : m1;
  const int m3 = bits.w != kSentinel ? sl + 3 : m2;
  const uint32_t has = __ballot_sync(0xffffffffu, m3 >= 0);
  return __shfl_sync(0xffffffffu, m3, (31 - __clz(has)) & 31);
}

struct TileArgs {
  StepArgs a;
  const int32_t* anc;  // MODE 1: parents by stored ancestor index
};

// MODE 0: fused systematic resample (heads staged by resample_tile, resolved on the output side)
// MODE 1: parents gathered through a stored ancestor array (multinomial / debug path)
// resident CTAs per SM the register allocation is held to (NX = 1 / 4 / 6)
#ifndef PZ_OCC1
#define PZ_OCC1 4
#endif
#ifndef PZ_OCC4
#define PZ_OCC4 3
#endif
#ifndef PZ_OCC6
#define PZ_OCC6 2
#endif
// packed fp32 model step (two slots per FFMA2 / FADD2 / FMUL2) for fp32 states: PZ_PACK_MODEL is a bit mask over
// NX = 1 / 4 / 6 (1 | 2 | 4): measured, only LG(4,2) gains (0.1452 -> 0.1432 ms at N = 10^7; LG(6,3) loses 4 % to register
// pressure, RW1D is neutral); the weights and moments are packed for every type
#ifndef PZ_PACK_MODEL
#define PZ_PACK_MODEL 2
#endif
// persistent grid (a CTA strides over the tiles; table, observation and plan constants fetched once per CTA): measured at
// N = 10^7 / 10^8 -- LG(4,2) f64 0.2293 -> 0.2242 ms, LG(6,3) f64 0.3594 -> 0.3508, LG(6,3) f32 0.1992 -> 0.1972, but
// RW1D f64 0.5176 -> 0.5611, RW1D f32 0.4142 -> 0.4613, LG(4,2) f32 0.1501 -> 0.1561: the one-word states keep one CTA per tile
// With index staging and 23-block tiles (later in the round) the wide states lose that gain too (LG(6,3) fp64 0.2518 persistent vs
// 0.2499, fp32 0.1853 vs 0.1828; LG(4,2) equal): only the remote-read fallback keeps its small persistent grid.
template <class T, int NX, bool REMOTE> constexpr bool kPersist = REMOTE;
// index staging for the one-word states (PZ_IDX1: bit 0 fp32, bit 1 fp64): the input side stages the parent's INDEX (4 bytes
// per head slot, no payload loads, no payload registers) and the output side gathers the parent's state from global memory
// (sorted ancestors: the loads of a warp fall into a few adjacent lines)
// Measured (RW1D, N = 10^8, pf_step_tile): fp64 0.4978 ms state-staged (6 CTAs per SM, 80 registers, 15-block tiles) ->
// 0.4568 ms index-staged at 8 CTAs per SM (64 registers, 23-block tiles: 4 bytes per staged slot leave room for both);
// index-staged at the old occupancy 0.5098 (the gather's extra latency without the extra warps); 9 / 10 CTAs 0.512 / 0.655
// (spills).  fp32: 0.3806 state-staged vs 0.4008 index-staged (its staged word is 4 bytes either way), so fp32 stays.
#ifndef PZ_IDX1
#define PZ_IDX1 2
#endif
// PZ_IDXN: the same for the wide states (bit 0 / 1: NX = 4 fp32 / fp64, bit 2 / 3: NX = 6 fp32 / fp64); their centred new
// states, which the moment pass reads back, then go through a per-warp scratch [NX][256] of floats instead of the
// staging buffer (every lane reads back exactly what it wrote: no synchronisation)
// Measured (N = 10^7, pf_step_tile, state-staged -> index-staged with 23-block tiles): LG(4,2) fp64 0.2236 -> 0.1768 ms,
// LG(6,3) fp64 0.3519 -> 0.2532, LG(4,2) fp32 0.1463 -> 0.1337, LG(6,3) fp32 0.1977 -> 0.1857 (the staging buffer shrinks from
// NX state words to one index word per slot, the input side loads no payload; tiles of 7 / 15 / 19 / 23 / 31 blocks were tried)
// (measured and rejected: with the bulk L2 prefetch of the tile's input range RW1D fp64 0.4558 -> 0.4682 ms, LG(4,2) fp64
// 0.1663 -> 0.1935, LG(6,3) fp64 0.2358 -> 0.2539: the 47-280 KB bursts queue ahead of the input side's own latency-critical
// weight loads; the gather's misses are better left spread over the output side)
// Measured (RW1D fp64, N = 10^8): PZ_GATHER2 = 0 0.4577 ms (144 bytes of spills at 64 registers), = 1 0.4283 ms (no spills),
// = 2 (gathered one block ahead, 16 more live registers) 0.4388 ms
#ifndef PZ_GATHER2
#define PZ_GATHER2 1
#endif
// (measured and rejected, like the bulk prefetch below: LG(4,2) fp64 0.1673 -> 0.1723 ms, LG(6,3) fp64 0.2344 -> 0.2557, fp32
// +3-4 %: on this kernel every kind of prefetch instruction costs more than the latency it hides)
#ifndef PZ_PREFETCH_NEXT
#define PZ_PREFETCH_NEXT 0
#endif
#ifndef PZ_L2_PREFETCH
#define PZ_L2_PREFETCH 0
#endif
#ifndef PZ_IDXN
#define PZ_IDXN 15
#endif
template <class T, int NX, int MODE> constexpr bool kIdxStage =
    MODE == 0 &&
    (NX == 1 ? (((PZ_IDX1) >> (sizeof(T) == 8 ? 1 : 0)) & 1) : (((PZ_IDXN) >> ((NX <= 4 ? 0 : 2) + (sizeof(T) == 8 ? 1 : 0))) & 1));
template <int NX> constexpr bool kPackModel = ((PZ_PACK_MODEL) >> (NX == 1 ? 0 : (NX <= 4 ? 1 : 2))) & 1;
constexpr int min_ctas(int nx) { return nx == 1 ? PZ_OCC1 : (nx <= 4 ? PZ_OCC4 : PZ_OCC6); }
#define PZ_NACC(nx) (2 + 2 * (nx))

// A double filter forms its block exponent from the log-weights ROUNDED TO FP32 (the fp32 definition applied
// to fl32(lw): a double maximum costs ~12 instructions per particle, the exponent only has to be within a
// fraction of a unit of the true one).  lw_key is the value that enters the block maximum.
__device__ __forceinline__ float lw_key(float lw) { return lw; }
__device__ __forceinline__ float lw_key(double lw) { return __double2float_rn(lw); }

// Output side of one canonical output block (256 slots, one warp; a lane owns slots lane*4 + {0..3}
// and 128 + lane*4 + {0..3}).  FULL: every slot of the tile is live (all tiles but the last).
// A double filter (T = double) propagates and evaluates the log-weight in double; from there on it is the fp32
// machinery: the log-weight rounded to fp32 enters the block maximum and the weight polynomial, the centred state
// x' - c (formed in double) is rounded to fp32 for the moment sums.  (Keeping those in double costs 24 registers
// and an occupancy step for nothing the 16-bit weights or the fp64 tile totals could show.)
// (The staged indices are MEMORY indices, and the CDF order inside a block is lane-major -- offsets 0-3, 128-131, 4-7, ... --
// so they are not monotone along the slots: a running maximum cannot replace the head positions.  Staging CDF positions
// instead would be monotone but needs the 5-instruction position -> offset map per slot, which is what it would save.)
// Index-staged one-word states: resolve the heads of BOTH halves of output block ob and issue the loads of their parents'
// states (PZ_GATHER2 = 1: at the top of output_block; = 2: one block ahead, while the previous block is still computing)
template <class T, int TO, bool FULL>
__device__ __forceinline__ void gather_parents(const StepArgs& a, const float* s_par, int ob, int lane, int nxt, T (&xpa)[2][4]) {
  const T* __restrict__ xin = reinterpret_cast<const T*>(a.X[nxt ^ 1]);
  int seg_carry = -1;
  if (ob > 0) seg_carry = last_head_before(s_par, ob * kBlk, lane);  // runs of <= 128 slots reach back one segment
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int sl = ob * kBlk + half * 128 + lane * 4;
    const uint4 bits = *reinterpret_cast<const uint4*>(&s_par[sl]);
    int head[4];
    seg_carry = resolve_heads(bits, sl, lane, seg_carry, head);
    const uint32_t own[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int hk = FULL ? head[k] : max(head[k], 0);
      PZ_ASSERT(hk < TO && hk <= sl + k && (hk >= 0 || a.plan->total == 0));
      const int32_t pi = (head[k] == sl + k) ? (int32_t)own[k] : __float_as_int(s_par[hk]);
      xpa[half][k] = xin[pi];
    }
  }
}

// The same for the wide states (PZ_GATHER2N: bit 0 / 1 NX = 4 fp32 / fp64, bit 2 / 3 NX = 6 fp32 / fp64): NX x 8 values in flight
template <class T, int NX, int TO, bool FULL>
__device__ __forceinline__ void gather_parents_wide(const StepArgs& a, const float* s_par, int ob, int lane, int nxt, T (&xpw)[2][NX][4]) {
  const T* __restrict__ xin = reinterpret_cast<const T*>(a.X[nxt ^ 1]);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int sl = ob * kBlk + half * 128 + lane * 4;
    const uint4 bits = *reinterpret_cast<const uint4*>(&s_par[sl]);
    int head[4];
    resolve_heads(bits, sl, lane, -1, head);   // every segment of a wide state starts with a head
    const uint32_t own[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int hk = FULL ? head[k] : max(head[k], 0);
      const int32_t pi = (head[k] == sl + k) ? (int32_t)own[k] : __float_as_int(s_par[hk]);
#pragma unroll
      for (int d = 0; d < NX; ++d) xpw[half][d][k] = xin[(int64_t)d * a.ld + pi];
    }
  }
}
// Measured (N = 10^7): LG(4,2) fp32 0.1348 -> 0.1268 ms (kept); LG(6,3) fp32 0.1854 -> 0.1885, LG(4,2) fp64 0.1680 -> 0.1683,
// LG(6,3) fp64 0.2336 -> 0.2623 (48 more live registers: spills)
#ifndef PZ_GATHER2N
#define PZ_GATHER2N 1
#endif
template <class T, int NX> constexpr bool kGather2Wide =
    NX > 1 && (((PZ_GATHER2N) >> ((NX <= 4 ? 0 : 2) + (sizeof(T) == 8 ? 1 : 0))) & 1);

// Index-staged wide states: non-binding L2 prefetches of the parents of output block ob (first and last slot of each lane's
// quad, every state dimension), issued one block ahead -- no registers are held, the later gather hits L2 instead of HBM
// (PZ_PREFETCH_NEXT)
template <class T, int NX, int TO>
__device__ __forceinline__ void prefetch_parents(const StepArgs& a, const float* s_par, int ob, int lane, int nxt) {
  const T* __restrict__ xin = reinterpret_cast<const T*>(a.X[nxt ^ 1]);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int sl = ob * kBlk + half * 128 + lane * 4;
    const uint4 bits = *reinterpret_cast<const uint4*>(&s_par[sl]);
    int head[4];
    resolve_heads(bits, sl, lane, -1, head);   // (NX > 1: every segment starts with a head)
    const int32_t p0 = (head[0] == sl) ? (int32_t)bits.x : __float_as_int(s_par[head[0]]);
    const int32_t p3 = (head[3] == sl + 3) ? (int32_t)bits.w : __float_as_int(s_par[head[3]]);
#pragma unroll
    for (int d = 0; d < NX; ++d) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(xin + (int64_t)d * a.ld + p0));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(xin + (int64_t)d * a.ld + p3));
    }
  }
}

template <class T, int NX, int NY, int TO, int MODE, bool KRON, bool FULL, bool IDX, bool REMOTE, class S>
__device__ __forceinline__ void output_block(const StepArgs& a, S* s_par, float* s_dx, const float4* __restrict__ tab_m8, int ob,
                                             int nslots, int64_t j0l, uint32_t j0, uint32_t t, int nxt, const T* y_new,
                                             const T* cen, int lane, float (&acc)[PZ_NACC(NX)], int32_t& aref,
                                             const T (*xpre)[4] = nullptr) {
  T* __restrict__ xout = reinterpret_cast<T*>(a.X[nxt]);
  uint16_t* __restrict__ qout = a.Q[nxt];
  const LgDevT<T>& model = model_of<T>(a);
  float lw[8];
  float dxk[NX == 1 ? 8 : 1];  // NX = 1 keeps the centred new state in registers; wider states go through s_par
  float mx = -INFINITY;
  int seg_carry = -1;
  constexpr bool kGather2 = IDX && NX == 1 && !REMOTE && (PZ_GATHER2 != 0);
  if constexpr (MODE == 0 && NX == 1 && !kGather2) {
    if (ob > 0) seg_carry = last_head_before(s_par, ob * kBlk, lane);  // runs of <= 128 slots reach back one segment
  }
  constexpr bool kG2W = IDX && !REMOTE && kGather2Wide<T, NX>;
  T xpw[kG2W ? 2 : 1][NX][4];
  if constexpr (kG2W) gather_parents_wide<T, NX, TO, FULL>(a, reinterpret_cast<const float*>(s_par), ob, lane, nxt, xpw);
  T xpa[kGather2 ? 2 : 1][4];
  if constexpr (kGather2) {
    if (xpre != nullptr) {   // gathered one block ahead by the caller (PZ_GATHER2 = 2, full tiles)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int k = 0; k < 4; ++k) xpa[h][k] = xpre[h][k];
      }
    } else gather_parents<T, TO, FULL>(a, reinterpret_cast<const float*>(s_par), ob, lane, nxt, xpa);
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int sl = ob * kBlk + half * 128 + lane * 4;  // tile-relative slot of this lane's quad
    T xp[NX][4];
    if constexpr (kGather2) {
#pragma unroll
      for (int k = 0; k < 4; ++k) xp[0][k] = xpa[half][k];
    } else if constexpr (kG2W) {
#pragma unroll
      for (int d = 0; d < NX; ++d) {
#pragma unroll
        for (int k = 0; k < 4; ++k) xp[d][k] = xpw[half][d][k];
      }
    } else if constexpr (IDX) {
      // staged parent indices: resolve the heads, then gather the parents' states from global memory
      const T* __restrict__ xin = reinterpret_cast<const T*>(a.X[nxt ^ 1]);
      const uint4 bits = *reinterpret_cast<const uint4*>(&s_par[sl]);
      int head[4];
      seg_carry = resolve_heads(bits, sl, lane, seg_carry, head);
      if constexpr (NX != 1) seg_carry = -1;
      const uint32_t own[4] = {bits.x, bits.y, bits.z, bits.w};
      int32_t pidx[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int hk = FULL ? head[k] : max(head[k], 0);
        PZ_ASSERT(hk < TO && hk <= sl + k && (hk >= 0 || a.plan->total == 0));  // (total == 0: a degenerate plan covers no slot)
        pidx[k] = (head[k] == sl + k) ? (int32_t)own[k] : __float_as_int(s_par[hk]);
      }
      if constexpr (REMOTE) {   // global indices: the parent's state lives on the rank that owns it (NVLink load when it is a peer)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int q = pidx[k] / (int32_t)a.n_local;
          const int32_t li = pidx[k] - q * (int32_t)a.n_local;
          PZ_ASSERT(q >= 0 && q < a.world);
          const T* __restrict__ xq = reinterpret_cast<const T*>(a.p2p->Xall[q][nxt ^ 1]);
#pragma unroll
          for (int d = 0; d < NX; ++d) xp[d][k] = xq[(int64_t)d * a.ld + li];
        }
      } else {
#pragma unroll
        for (int d = 0; d < NX; ++d) {
#pragma unroll
          for (int k = 0; k < 4; ++k) xp[d][k] = xin[(int64_t)d * a.ld + pidx[k]];
        }
      }
    } else {
#pragma unroll
    for (int d = 0; d < NX; ++d) {
      const Quad<T> f = lds_quad(&s_par[d * TO + sl]);
      xp[d][0] = f.v[0]; xp[d][1] = f.v[1]; xp[d][2] = f.v[2]; xp[d][3] = f.v[3];
    }
    }
    if constexpr (MODE == 0 && !IDX) {
      int head[4];
      const uint4 bits = make_uint4(mark_word(xp[0][0]), mark_word(xp[0][1]), mark_word(xp[0][2]), mark_word(xp[0][3]));
      seg_carry = resolve_heads(bits, sl, lane, seg_carry, head);
      if constexpr (NX != 1) seg_carry = -1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (head[k] != sl + k) {
          const int hk = FULL ? head[k] : max(head[k], 0);
          PZ_ASSERT(hk < TO && hk <= sl + k && (hk >= 0 || a.plan->total == 0));  // (total == 0: a degenerate plan covers no slot)
#pragma unroll
          for (int d = 0; d < NX; ++d) xp[d][k] = s_par[d * TO + hk];
        }
      }
      if constexpr (NX > 1) __syncwarp();  // every parent read of this segment precedes the write-back below
    }
    T z[NX][4];
    lg_noise4<NX, T>(j0 + (uint32_t)sl, t, a.seed, tab_m8, z);
    T xo[NX][4];
    if constexpr (sizeof(T) == 4 && kPackModel<NX>) {
      // fp32 state: slots k, k+1 share packed instructions (same individually rounded operations per slot)
#pragma unroll
      for (int k = 0; k < 4; k += 2) {
        float2 xs[NX], zs[NX], xn[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) { xs[d] = make_float2(xp[d][k], xp[d][k + 1]); zs[d] = make_float2(z[d][k], z[d][k + 1]); }
        model_propagate<NX, KRON>(model, xs, zs, xn);
#pragma unroll
        for (int d = 0; d < NX; ++d) { xo[d][k] = xn[d].x; xo[d][k + 1] = xn[d].y; }
        float2 l = model_logweight<NX, NY, KRON>(model, xn, y_new);
        if (!FULL && sl + k >= nslots) l.x = -INFINITY;  // past the end of the population
        if (!FULL && sl + k + 1 >= nslots) l.y = -INFINITY;
        lw[half * 4 + k] = l.x; lw[half * 4 + k + 1] = l.y;
        mx = fmaxf(mx, fmaxf(l.x, l.y));
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        T xs[NX], zs[NX], xn[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) { xs[d] = xp[d][k]; zs[d] = z[d][k]; }
        model_propagate<NX, KRON>(model, xs, zs, xn);
#pragma unroll
        for (int d = 0; d < NX; ++d) xo[d][k] = xn[d];
        float l = lw_key(model_logweight<NX, NY, KRON>(model, xn, y_new));
        if (!FULL && sl + k >= nslots) l = -INFINITY;  // past the end of the population
        lw[half * 4 + k] = l;
        mx = fmaxf(mx, l);
      }
    }
    if (FULL || sl < nslots) {
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        st_quad(&xout[(size_t)d * a.ld + j0l + sl], xo[d]);
        if constexpr (NX > 1 && !IDX) sts_quad(&s_par[d * TO + sl], xo[d]);
      }
    }
    if constexpr (NX > 1 && IDX) {   // centred new state -> this lane's corner of the warp's scratch
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        float4 dx4;
        if constexpr (sizeof(T) == 4) {
          const float2 p0 = r_sub(make_float2((float)xo[d][0], (float)xo[d][1]), (float)cen[d]), p1 = r_sub(make_float2((float)xo[d][2], (float)xo[d][3]), (float)cen[d]);
          dx4 = make_float4(p0.x, p0.y, p1.x, p1.y);
        } else {
          dx4 = make_float4((float)r_sub(xo[d][0], cen[d]), (float)r_sub(xo[d][1], cen[d]), (float)r_sub(xo[d][2], cen[d]), (float)r_sub(xo[d][3], cen[d]));
        }
        *reinterpret_cast<float4*>(&s_dx[d * kBlk + half * 128 + lane * 4]) = dx4;
      }
    }
    if constexpr (NX == 1) {
      if constexpr (sizeof(T) == 4) {
        const float2 a = r_sub(make_float2((float)xo[0][0], (float)xo[0][1]), (float)cen[0]), b = r_sub(make_float2((float)xo[0][2], (float)xo[0][3]), (float)cen[0]);
        dxk[half * 4 + 0] = a.x; dxk[half * 4 + 1] = a.y; dxk[half * 4 + 2] = b.x; dxk[half * 4 + 3] = b.y;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) dxk[half * 4 + k] = (float)r_sub(xo[0][k], cen[0]);
      }
    }
  }
  mx = warp_max_f32(mx);
  int32_t nb;
  const float nbf = block_exponent(mx, nb);
  uint32_t ssum = 0;
  if (nb != kNZero) {
    // packed fp32 (FFMA2 / FADD2 / FMUL2): slots k, k+1 of a lane ride in one instruction; the lane's sums are
    // kept as (even slots, odd slots) pairs and added once per block
    float2 pp[PZ_NACC(NX)];
#pragma unroll
    for (int r = 0; r < PZ_NACC(NX); ++r) pp[r] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int sl = ob * kBlk + half * 128 + lane * 4;
      float dxv[NX][4];
      if constexpr (NX > 1 && IDX) {
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          const float4 f = *reinterpret_cast<const float4*>(&s_dx[d * kBlk + half * 128 + lane * 4]);
          dxv[d][0] = f.x; dxv[d][1] = f.y; dxv[d][2] = f.z; dxv[d][3] = f.w;
        }
      } else if constexpr (NX > 1) {
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          const Quad<T> f = lds_quad(&s_par[d * TO + sl]);
          if constexpr (sizeof(T) == 4) {
            const float2 a = r_sub(make_float2((float)f.v[0], (float)f.v[1]), (float)cen[d]), b = r_sub(make_float2((float)f.v[2], (float)f.v[3]), (float)cen[d]);
            dxv[d][0] = a.x; dxv[d][1] = a.y; dxv[d][2] = b.x; dxv[d][3] = b.y;
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) dxv[d][k] = (float)r_sub(f.v[k], cen[d]);
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) dxv[0][k] = dxk[half * 4 + k];
      }
      uint32_t qb[4];
#pragma unroll
      for (int k = 0; k < 4; k += 2) {
        float2 w;
        weight_q2(make_float2(lw[half * 4 + k], lw[half * 4 + k + 1]), nbf, w, qb[k], qb[k + 1]);
        const bool live0 = FULL || sl + k < nslots, live1 = FULL || sl + k + 1 < nslots;
        if (!live0) w.x = 0.0f;  // (its q is already 0: lw = -inf)
        if (!live1) w.y = 0.0f;
        ssum += qb[k] + qb[k + 1];  // raw bits 0x4B000000 + q; the bias is removed after the warp sum
        pp[0] = __fadd2_rn(pp[0], w);
        pp[1] = __ffma2_rn(w, w, pp[1]);
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          const float2 dx = make_float2(live0 ? dxv[d][k] : 0.0f, live1 ? dxv[d][k + 1] : 0.0f);
          const float2 wd = __fmul2_rn(w, dx);
          pp[2 + d] = __fadd2_rn(pp[2 + d], wd);
          pp[2 + NX + d] = __ffma2_rn(wd, dx, pp[2 + NX + d]);
        }
      }
      if (FULL || sl < nslots)
        *reinterpret_cast<uint2*>(&qout[j0l + sl]) = make_uint2(__byte_perm(qb[0], qb[1], 0x5410), __byte_perm(qb[2], qb[3], 0x5410));
    }
    float p[PZ_NACC(NX)];
#pragma unroll
    for (int r = 0; r < PZ_NACC(NX); ++r) p[r] = __fadd_rn(pp[r].x, pp[r].y);
    // fold the block's sums (scale 2^nb) into the lane accumulators (scale 2^aref)
    if (nb == aref) {
#pragma unroll
      for (int r = 0; r < PZ_NACC(NX); ++r) acc[r] = __fadd_rn(acc[r], p[r]);
    } else if (aref == kNZero) {
#pragma unroll
      for (int r = 0; r < PZ_NACC(NX); ++r) acc[r] = p[r];
      aref = nb;
    } else {
      const int32_t nref = max(aref, nb);
      const float ga = exp2i(aref - nref), gp = exp2i(nb - nref);
#pragma unroll
      for (int r = 0; r < PZ_NACC(NX); ++r) {
        const float fa = (r == 1) ? __fmul_rn(ga, ga) : ga, fp = (r == 1) ? __fmul_rn(gp, gp) : gp;
        acc[r] = __fmaf_rn(acc[r], fa, __fmul_rn(p[r], fp));
      }
      aref = nref;
    }
  } else {
    // empty block: every weight is zero
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int sl = ob * kBlk + half * 128 + lane * 4;
      if (FULL || sl < nslots) *reinterpret_cast<uint2*>(&qout[j0l + sl]) = make_uint2(0u, 0u);
    }
  }
  ssum = __reduce_add_sync(0xffffffffu, ssum);
  if (lane == 0) {
    const int64_t bo = (j0l / kBlk) + ob;
    a.eb[nxt][bo] = nb;
    a.sb[bo] = (nb != kNZero) ? (u64)(ssum - 256u * 0x4B000000u) : 0ull;  // 256 biased terms (mod 2^32)
  }
}

// NT threads per CTA (NW warps): 256 by default; PZ_NT1 = 128 runs the NX = 1 kernel with half-size CTAs
// REMOTE: the fallback of a sharded filter for a resample whose ancestors reach beyond the halo margins (verdict
// plan->shard_overflow == 1, the same on every rank): tile_start holds GLOBAL block indices and the input side reads
// every block from its owner over NVLink.  A separate instantiation launched after the regular one on a small
// persistent grid: the regular kernel exits at once when the margins overflowed, the remote one in every other
// case (~3 us per step).  (As a uniform runtime branch of ONE kernel it cost the register-bound wide models 6-11 %.)
template <class T, int NX, int NY, int TO, int MODE, bool KRON, int NT, int MINB, bool REMOTE = false>
__global__ void __launch_bounds__(NT, MINB) pf_step_tile(const __grid_constant__ TileArgs ta) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char s_raw[];  // [NX][TO] staged parents (or [TO] parent indices) | [256] float4 inverse-CDF table
  constexpr bool IDX = kIdxStage<T, NX, MODE>;
  typedef typename std::conditional<IDX, float, T>::type S;                        // staged scalar
  typedef typename std::conditional<IDX, IndexSource, StateSource<NX, T>>::type Src;
  S* s_par = reinterpret_cast<S*>(s_raw);
  __shared__ int32_t s_eref[NW];
  __shared__ double s_red[NW][kMaxRec];
  // index-staged wide states: [TO] indices | per-warp scratch [NW][NX][256] floats | table
  constexpr int kStageWords = IDX ? TO : NX * TO;
  constexpr int kScratch = (IDX && NX > 1) ? NW * NX * kBlk : 0;
  float* s_dx_all = reinterpret_cast<float*>(s_par + kStageWords);
  float4* s_tab = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(s_par + kStageWords) + kScratch * sizeof(float));
  const StepArgs& a = ta.a;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait_then_trigger();
  const uint32_t t = a.ctl->t;
  const int cur = t & 1, nxt = cur ^ 1;
  const T* __restrict__ xin = reinterpret_cast<const T*>(a.X[cur]);
  const PlanDev* pl = a.plan;
  constexpr bool remote = REMOTE;
  if constexpr (MODE == 0) {
    if (a.p2p != nullptr) {
      const int32_t ov = pl->shard_overflow;
      if (REMOTE ? ov != 1 : ov != 0) return;   // (2: a peer never answered -- the error is reported, neither kernel runs)
    }
    // A degenerate plan (every weight zero: PZ_EDEGENERATE, reported by the step that produced it) covers no output slot, so
    // there is no parent to gather: steps queued behind it (pz_filter_run) leave the population as it is until pz_filter_reset.
    if (pl->degenerate) return;
  }
  if constexpr (REMOTE) {
    // every rank's boundary positions of this epoch must be complete before any block is read
    if (tid < a.world) {
      const Mailbox* mine = a.p2p->mbox[a.rank];
      const unsigned long long epoch = (unsigned long long)(t + 1u);
      wait_sys(&mine->xb_done[tid], [epoch](unsigned long long x) { return x >= epoch; }, &a.ctl->comm_error, a.wait_ticks, 0x20u);
    }
    __syncthreads();
  }
  // A persistent grid strides over the tiles (launch_tile sizes it to the resident CTA count; with one CTA per tile the
  // loop runs once): the inverse-CDF table, the observation, the plan constants and the centring constants are
  // fetched once per CTA instead of once per tile.  The walk starts in the MIDDLE of the rank's tiles, so the boundary
  // tiles of a sharded filter -- which wait for the neighbours' margins -- come up when the margins have long arrived.
  T y_new[NY];
  load_obs<T, NY>(a, t, y_new);
  for (int i = tid; i < PZ_ICDF_ROWS; i += NT) s_tab[i] = g_icdf_tab[i];   // (ordered before its first use by the tile's barriers)
  T cen[NX];
#pragma unroll
  for (int d = 0; d < NX; ++d) {
    if constexpr (sizeof(T) == 8) cen[d] = pl->center64[d]; else cen[d] = pl->center[d];
  }
  const int32_t pl_emax = pl->emax, pl_h0 = pl->h0;
  const uint32_t pl_Mq = pl->Mq;
  const int ntiles = a.ntiles_out;
  const int tile_shift = (kPersist<T, NX, REMOTE> && gridDim.x < (unsigned)ntiles) ? ntiles / 2 : 0;
  int tile_k = blockIdx.x;
  do {
  int tile = tile_k + tile_shift;
  if (tile >= ntiles) tile -= ntiles;
  const int64_t j0l = (int64_t)tile * TO;  // local slot
  const int64_t jendl = (j0l + TO < a.n_local) ? j0l + TO : a.n_local;
  const uint32_t j0 = (uint32_t)(a.slot_base + j0l);
  const int nslots = (int)(jendl - j0l);
  PZ_STAMP(blockIdx.x == 0 && tid == 0, 11);
  PZ_STAMP(blockIdx.x == gridDim.x - 1 && tid == 0, 14);

  if constexpr (MODE == 0) {
    // the first input block's weights: in flight during the prologue
    const int64_t b_first = (int64_t)a.tile_start[tile];
    PZ_ASSERT(remote || (b_first >= a.b_lo && b_first <= a.b_hi));
    if (!remote && a.p2p != nullptr) {
      // sharded over peer memory: a boundary tile (its input blocks reach into a halo) waits here for the
      // neighbour's margins of this epoch (state / weights / exponents from pf_halo_push, boundary positions from
      // pf_partition_sharded); interior tiles -- nearly all -- start at once.  The flags hold the LATEST epoch: a
      // neighbour that runs ahead may already have pushed the next step's state margins (into the other buffer
      // parity -- it cannot get further: its next exchange needs this rank's record), hence >=.
      const int64_t b_next = (tile + 1 < a.ntiles_out) ? (int64_t)a.tile_start[tile + 1] : a.b_hi;
      if (tid < 4) {
        const int side = tid & 1;
        const bool need = side == 0 ? (a.rank > 0 && b_first < 0) : (a.rank + 1 < a.world && b_next >= a.nblk);
        if (need) {
          const Mailbox* mine = a.p2p->mbox[a.rank];
          const unsigned long long epoch = (unsigned long long)(t + 1u);
          wait_sys(tid < 2 ? &mine->halo_flag[side] : &mine->pbx_flag[side], [epoch](unsigned long long x) { return x >= epoch; },
                   &a.ctl->comm_error, a.wait_ticks, tid < 2 ? 0x40u : 0x80u);
        }
      }
      __syncthreads();
    }
    if constexpr (IDX && !REMOTE) {
      // Index staging leaves the parents' states untouched until the output side gathers them: one bulk L2 prefetch per
      // state dimension (cp.async.bulk.prefetch.L2, issued by one thread) pulls the tile's input range -- the blocks
      // [b_first, first block of the next tile] -- towards the SMs while the input side runs.  (A halo block that has not
      // arrived yet may be prefetched stale: L2 is the coherence point of peer stores, a prefetch keeps no private copy.)
      // Off by default (PZ_L2_PREFETCH): measured slower, see the macro.
#if PZ_L2_PREFETCH
      if (tid == 0) {
        const int64_t b_nx = (tile + 1 < ntiles) ? (int64_t)a.tile_start[tile + 1] + 1 : a.b_hi;
        int64_t nb = (b_nx < a.b_hi ? b_nx : a.b_hi) - b_first;
        const int64_t cap = 2 * (TO / kBlk) + 2;   // a tile that draws on a long stretch of childless blocks: the head is enough
        nb = nb < 0 ? 0 : (nb > cap ? cap : nb);
        const uint32_t bytes = (uint32_t)(nb * kBlk * (int64_t)sizeof(T));
        if (bytes) {
#pragma unroll
          for (int d = 0; d < NX; ++d) {
            const T* p = xin + (int64_t)d * a.ld + b_first * kBlk;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
          }
        }
      }
#endif
    }
    BlockIn first{};
    Src src;
    if constexpr (!IDX) { src.x = xin; src.ld = a.ld; }
    Quad<S> pa{}, pb{};
    if (!remote && b_first + warp < a.b_hi) {
      first.load(a.Q[cur], a.eb[cur], a.xb, b_first + warp, lane);
      // fp32 payloads of the first block are fetched here too; for doubles the 16 extra registers across the
      // barrier cost more than the exposed latency (measured: 0.568 ms hoisted vs 0.544 ms not, N = 10^8)
      if constexpr (Src::kPrefetch && sizeof(T) == 4) {
        pa = src.fetch((b_first + warp) * kBlk + lane * 4);
        pb = src.fetch((b_first + warp) * kBlk + lane * 4 + 128);
      }
    }
    // sentinel in (the mark word of) dimension 0 of every slot: "no head here"
    uint4* sz = reinterpret_cast<uint4*>(s_par);
    for (int i = tid; i < TO * (int)sizeof(S) / 16; i += NT) sz[i] = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
    __syncthreads();
    if constexpr (REMOTE) {
      resample_tile_remote<(NX != 1), NW>(src, s_par, TO, a.p2p, cur, a.nblk, a.world, b_first, pl_emax, pl_h0, pl_Mq, j0, nslots,
                                          lane, warp);
    } else {
      if constexpr (Src::kPrefetch && sizeof(T) == 8) {
        if (b_first + warp < a.b_hi) {
          pa = src.fetch((b_first + warp) * kBlk + lane * 4);
          pb = src.fetch((b_first + warp) * kBlk + lane * 4 + 128);
        }
      }
      resample_tile<(NX != 1), NW>(src, s_par, TO, a.Q[cur], a.eb[cur], a.xb, b_first, a.b_hi, pl_emax, pl_h0, pl_Mq, j0, nslots,
                               lane, warp, first, pa, pb);
    }
  } else {
    for (int sidx = tid; sidx < nslots; sidx += NT) {
      const int64_t p = ta.anc[j0l + sidx];
      PZ_ASSERT(p >= 0 && p < a.n_local);
#pragma unroll
      for (int d = 0; d < NX; ++d) s_par[d * TO + sidx] = xin[(size_t)d * a.ld + p];
    }
  }
  __syncthreads();

  // ---- output side
  const float4* tab_m8 = s_tab - 8;
  float acc[PZ_NACC(NX)];  // sw, sw2, sx[NX], sxx[NX] at scale 2^aref (fp32 per lane)
#pragma unroll
  for (int r = 0; r < PZ_NACC(NX); ++r) acc[r] = 0.0f;
  int32_t aref = kNZero;
  if (nslots == TO) {
    if constexpr (IDX && NX == 1 && !REMOTE && PZ_GATHER2 == 2) {
      T xnext[2][4];
      if (warp < TO / kBlk) gather_parents<T, TO, true>(a, reinterpret_cast<const float*>(s_par), warp, lane, nxt, xnext);
      for (int ob = warp; ob < TO / kBlk; ob += NW) {
        T xcur[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int k = 0; k < 4; ++k) xcur[h][k] = xnext[h][k];
        }
        if (ob + NW < TO / kBlk) gather_parents<T, TO, true>(a, reinterpret_cast<const float*>(s_par), ob + NW, lane, nxt, xnext);
        output_block<T, NX, NY, TO, MODE, KRON, true, IDX, REMOTE>(a, s_par, s_dx_all + warp * NX * kBlk, tab_m8, ob, nslots, j0l, j0, t, nxt, y_new, cen, lane, acc, aref, xcur);
      }
    } else {
    for (int ob = warp; ob < TO / kBlk; ob += NW) {
      if constexpr (IDX && NX > 1 && !REMOTE && (PZ_PREFETCH_NEXT != 0)) {
        if (ob + NW < TO / kBlk) prefetch_parents<T, NX, TO>(a, reinterpret_cast<const float*>(s_par), ob + NW, lane, nxt);
      }
      output_block<T, NX, NY, TO, MODE, KRON, true, IDX, REMOTE>(a, s_par, s_dx_all + warp * NX * kBlk, tab_m8, ob, nslots, j0l, j0, t, nxt, y_new, cen, lane, acc, aref);
    }
    }
  } else {
    for (int ob = warp; ob * kBlk < nslots; ob += NW)
      output_block<T, NX, NY, TO, MODE, KRON, false, IDX, REMOTE>(a, s_par, s_dx_all + warp * NX * kBlk, tab_m8, ob, nslots, j0l, j0, t, nxt, y_new, cen, lane, acc, aref);
  }

  // ---- moment sums -> one record per tile: every warp reduces its lanes at its own scale 2^aref; after ONE barrier
  // (the one that also frees the staging buffer for the next tile) the record's threads combine the warps
  if (lane == 0) s_eref[warp] = aref;
#pragma unroll
  for (int r = 0; r < PZ_NACC(NX); ++r) {
    double x = (double)acc[r];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += shfl_xor_f64(x, d);
    if (lane == 0) s_red[warp][r + 1] = x;
  }
  __syncthreads();
  // record layout (kMaxRec doubles): [0] exponent, [1] sw, [2] sw2, [3..3+NX) sx, [3+NX..3+2NX) sxx
  if (tid < 3 + 2 * NX) {
    int32_t etile = kNZero;
#pragma unroll
    for (int w = 0; w < NW; ++w) etile = max(etile, s_eref[w]);
    double x = 0;
    if (tid == 0) x = (double)etile;
    else {
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const int32_t ew = s_eref[w];
        if (ew != kNZero) {
          const double g = pow2d(ew - etile);
          x += s_red[w][tid] * (tid == 2 ? g * g : g);
        }
      }
    }
    a.rec[(size_t)tile * kMaxRec + tid] = x;
    if (tid == 0 && etile != kNZero) atomicMax(&a.ctl->emax_acc, etile);
  }
  // (the next tile rewrites s_eref / s_red only after its own two barriers)
  } while (kPersist<T, NX, REMOTE> && (tile_k += (int)gridDim.x) < ntiles);
}

// ------------------------------------------------------------------------------------------
// generic path kernels (stored log-weights): standalone resamplers, taps, the tracker
// ------------------------------------------------------------------------------------------
// One warp per canonical block: block exponent, 16-bit weights, their sum.
__global__ void __launch_bounds__(kThreads) pf_stats_from_lw(const float* __restrict__ lw, int64_t n, int64_t nblk,
                                                             int32_t* eb, u64* sb, uint16_t* q16, CtlDev* ctl) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * kWarps + warp;
  if (b >= nblk) return;
  float l[8];
  float mx = -INFINITY;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int64_t i0 = b * kBlk + half * 128 + lane * 4;
    const float4 f = *reinterpret_cast<const float4*>(lw + i0);
    const float w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      l[half * 4 + k] = (i0 + k < n) ? sanitize_lw(w[k]) : -INFINITY;
      mx = fmaxf(mx, l[half * 4 + k]);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  int32_t nb;
  const float nbf = block_exponent(mx, nb);
  uint32_t ssum = 0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int64_t i0 = b * kBlk + half * 128 + lane * 4;
    uint32_t qb[4] = {0u, 0u, 0u, 0u};
    if (nb != kNZero) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float w;
        weight_q<true>(l[half * 4 + k], nbf, w, qb[k]);
        ssum += qb[k] & 0xffffu;
      }
    }
    *reinterpret_cast<uint2*>(&q16[i0]) = make_uint2(__byte_perm(qb[0], qb[1], 0x5410), __byte_perm(qb[2], qb[3], 0x5410));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, d);
  if (lane == 0) {
    eb[b] = nb;
    sb[b] = (u64)ssum;
    if (nb != kNZero) atomicMax(&ctl->emax_acc, nb);
  }
}

__device__ __forceinline__ bool generic_alt(const GenericArgs& g) { return g.parity_ctl && (g.parity_ctl->t & 1u); }
__device__ __forceinline__ const int32_t* generic_eb(const GenericArgs& g) { return generic_alt(g) ? g.eb_alt : g.eb; }
__device__ __forceinline__ const uint16_t* generic_q(const GenericArgs& g) { return generic_alt(g) ? g.q16_alt : g.q16; }

__global__ void pf_ctl_reset(CtlDev* ctl) {
  ctl->t = 0; ctl->emax_acc = kNZero; ctl->scan_counter = 0; ctl->scan_done = 0;
  ctl->comm_error = 0; ctl->ess_resampled = 1; ctl->pad[0] = ctl->pad[1] = 0;
}

template <int TO>
__global__ void __launch_bounds__(kThreads) pf_ancestors(const GenericArgs g) {
  __shared__ __align__(16) float s_anc[TO];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PlanDev* pl = g.plan;
  const int tile = (int)blockIdx.x + g.tile0;
  const int64_t j0 = (int64_t)tile * TO;
  const int64_t jend = (j0 + TO < g.n) ? j0 + TO : g.n;
  const int nslots = (int)(jend - j0);
  for (int sidx = tid; sidx < TO; sidx += kThreads) s_anc[sidx] = __uint_as_float(kSentinel);
  __syncthreads();
  IndexSource src;
  const int64_t b_first = (int64_t)g.tile_start[tile];
  BlockIn first{};
  if (b_first + warp < g.nblk) first.load(generic_q(g), generic_eb(g), g.xb, b_first + warp, lane);
  resample_tile<true, kWarps>(src, s_anc, TO, generic_q(g), generic_eb(g), g.xb, b_first, g.nblk, pl->emax, pl->h0, pl->Mq, (uint32_t)j0,
                      nslots, lane, warp, first, Quad<float>{}, Quad<float>{});
  __syncthreads();
  for (int seg = warp; seg * 128 < nslots; seg += kWarps) {
    const int sl = seg * 128 + lane * 4;
    const uint4 bits = *reinterpret_cast<const uint4*>(&s_anc[sl]);
    int head[4];
    resolve_heads(bits, sl, lane, -1, head);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (sl + k < nslots) {
        // (head < 0: a degenerate plan -- every weight zero -- leaves slots uncovered; the host reports PZ_EDEGENERATE)
        PZ_ASSERT(head[k] < TO && (head[k] < 0 || (__float_as_int(s_anc[head[k]]) >= 0 && __float_as_int(s_anc[head[k]]) < g.nblk * kBlk)));
        g.anc[j0 + sl + k] = head[k] >= 0 ? __float_as_int(s_anc[head[k]]) : -1;
      }
  }
}

// N iid categorical draws on the canonical fixed-point CDF (resample2, Inference.hs:57-61)
__global__ void __launch_bounds__(kThreads) pf_multinomial(const GenericArgs g) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= (g.n_draws > 0 ? g.n_draws : g.n)) return;
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
  const int32_t ksh = pl->emax - generic_eb(g)[b];
  const uint16_t* q = generic_q(g);
  const u64 Pb = g.pb[b];
  u64 c = 0;
  int32_t res = -1;
  for (int p = 0; p < kBlk; ++p) {  // the block's CDF order
    const int64_t i = b * kBlk + cdf_offset(p);
    c += q[i];
    if (res < 0 && Pb + shift_guard(c, ksh) > tau) res = (int32_t)i;
  }
  g.anc[j] = res;
}

// log-weights of the pending population (parity tap): lw_out32 (rounded to float when the filter is f64) and /
// or lw_out64 (exact in both cases)
template <class T, int NX, int NY>
__global__ void __launch_bounds__(kThreads) pf_logw(StepArgs a, float* lw_out32, double* lw_out64) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n_local) return;
  const uint32_t t = a.ctl->t;
  T l = 0;  // the initial point mass: equal weights
  if (t != 0) {
    const T* x = reinterpret_cast<const T*>(a.X[t & 1]);
    T xs[NX], y[NY];
#pragma unroll
    for (int d = 0; d < NX; ++d) xs[d] = x[(size_t)d * a.ld + i];
    load_obs<T, NY>(a, t + kObsRing - 1, y);
    l = lg_logweight<NX, NY>(model_of<T>(a), xs, y);
  }
  if (lw_out32) lw_out32[i] = (float)l;
  if (lw_out64) lw_out64[i] = (double)l;
}

// every stride-th resampled particle, slot-major, as doubles
__global__ void __launch_bounds__(kThreads) pf_gather_thin(StepArgs a, const int32_t* anc, int32_t stride, int64_t n_sel,
                                                           double* out) {
  const int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (s >= n_sel) return;
  const int64_t p = anc[s * stride];
  for (int d = 0; d < a.model.nx; ++d)
    out[s * a.model.nx + d] = a.f64 ? reinterpret_cast<const double*>(a.X[a.ctl->t & 1])[(size_t)d * a.ld + p]
                                    : (double)reinterpret_cast<const float*>(a.X[a.ctl->t & 1])[(size_t)d * a.ld + p];
}

// ------------------------------------------------------------------------------------------
// sharded filter: plan from the all-gathered rank records, absolute prefixes, margin check
// ------------------------------------------------------------------------------------------
struct ShardPlanArgs {
  PlanDev* plan;
  CtlDev* ctl;
  u64* rec;            // [world][kShardRec]: gathered by NCCL, or filled here from the mailbox
  const u64* pb;       // local prefix [nblk+1]
  u64* pbx;            // absolute prefix, own entries [0, nblk]
  pz_step_summary* summ;
  int64_t nblk, n_global, n_local;
  int32_t rank, world, nx, delta;
  uint64_t seed;
  double ll_const;
  // peer-memory exchange (p2p != nullptr): records through the mailboxes, halo margins by peer stores
  const ShardP2P* p2p;
  int64_t margin;
  int32_t f64;
  unsigned long long wait_ticks;
};

// Every rank evaluates, for EVERY rank q, whether the fixed halo margins (the last `margin` blocks of
// rank q-1 and the first `margin` blocks of rank q+1) cover the ancestors of q's slots -- from the
// gathered records alone, so that all ranks reach the same verdict without another collective.
// With the peer-memory exchange this kernel is also the all-gather (CTA 0 stores this rank's record
// into every mailbox, every CTA polls its own mailbox) and the halo exchange (all CTAs store the two
// margins straight into the neighbours' halo regions; the last CTA to finish raises their flags).
__global__ void __launch_bounds__(kThreads) pf_plan_sharded(const ShardPlanArgs p) {
  const int tid = threadIdx.x;
  const uint32_t epoch = p.ctl->t + (uint32_t)p.delta + 1u;
  const u64* rec = p.rec;
  __shared__ u64 s_rec[kMaxRanks * kShardRec];
  PZ_STAMP(blockIdx.x == 0 && tid == 0, 3);
  if (p.p2p) {
    // all-gather of the records over peer memory: CTA 0 stores this rank's record into every mailbox as
    // self-validating half-words, every CTA polls its own mailbox until all of this epoch are there
    Mailbox* mine = p.p2p->mbox[p.rank];
    const int nw = p.world * 2 * kShardRec;
    if (blockIdx.x == 0) {
      for (int i = tid; i < nw; i += kThreads) {
        const int q = i / (2 * kShardRec), w = i % (2 * kShardRec);
        const u64 v = p.rec[(size_t)p.rank * kShardRec + (w >> 1)];
        st_sys(&p.p2p->mbox[q]->rec2[p.rank][w], ((u64)epoch << 32) | (uint32_t)((w & 1) ? (v >> 32) : v));
      }
    }
    uint32_t* s_half = reinterpret_cast<uint32_t*>(s_rec);  // little endian: half w of word k of rank q = s_half[(q * kShardRec + k) * 2 + (w & 1)]
    for (int i = tid; i < nw; i += kThreads) {
      const u64 x = wait_sys<false>(&mine->rec2[i / (2 * kShardRec)][i % (2 * kShardRec)],
                                    [epoch](unsigned long long x) { return (uint32_t)(x >> 32) == epoch; }, &p.ctl->comm_error,
                                    p.wait_ticks, 2u);
      s_half[i] = (uint32_t)x;
    }
    __syncthreads();
    rec = s_rec;
    PZ_STAMP(blockIdx.x == 0 && tid == 0, 4);
  }
  u64 base = 0, total = 0;
  for (int q = 0; q < p.world; ++q) {
    const u64 tq = ((const volatile u64*)rec)[(size_t)q * kShardRec];
    if (q < p.rank) base += tq;
    total += tq;
  }
  const int64_t gid = (int64_t)blockIdx.x * kThreads + tid, gsz = (int64_t)gridDim.x * kThreads;
  for (int64_t b = gid; b <= p.nblk; b += gsz) p.pbx[b] = base + p.pb[b];
  PZ_STAMP(blockIdx.x == 0 && tid == 0, 5);
  if (p.p2p) {
    // prefix margins -> the neighbours' halo regions (my first m blocks are the RIGHT halo of rank-1, my last
    // m blocks the LEFT halo of rank+1); the X / Q / eb margins went out in pf_halo_push, beside the scan
    const int64_t m = p.margin, nb = p.nblk;
    const int npush = (int)((m + kThreads - 1) / kThreads);  // CTAs that have prefix entries to store
    if ((int)blockIdx.x < npush) {
      for (int64_t i = gid; i < m; i += gsz) {
        if (p.rank > 0) p.p2p->pbxn[0][nb + 1 + i] = base + p.pb[1 + i];
        if (p.rank + 1 < p.world) p.p2p->pbxn[1][i - m] = base + p.pb[nb - m + i];
      }
      PZ_STAMP(blockIdx.x == 0 && tid == 0, 6);
      __threadfence_system();
      __syncthreads();
      PZ_STAMP(blockIdx.x == 0 && tid == 0, 7);
      if (tid == 0) {
        Mailbox* mine = p.p2p->mbox[p.rank];
        if (atomicAdd(&mine->pbx_done, 1ull) == (unsigned long long)npush - 1ull) {
          PZ_STAMP(true, 8);
          __threadfence_system();
          mine->pbx_done = 0ull;
          if (p.rank > 0) st_sys(&p.p2p->mbox[p.rank - 1]->pbx_flag[1], (unsigned long long)epoch);
          if (p.rank + 1 < p.world) st_sys(&p.p2p->mbox[p.rank + 1]->pbx_flag[0], (unsigned long long)epoch);
        }
      }
    }
  }
  if (gid == 0) {
    PlanDev* pl = p.plan;
    const uint32_t t_next = p.ctl->t + (uint32_t)p.delta;
    const uint32_t U = philox4x32_10(0u, 0u, t_next, kStreamResample, (uint32_t)p.seed, (uint32_t)(p.seed >> 32)).x;
    pl->base = base;
    pl->shard_epoch = epoch;
    finalize_plan(pl, total, p.n_global, U);
    int32_t overflow = 0;
    int64_t migrated = 0;
    if (total != 0) {
      SlotMap sm(pl, p.n_global);
      u64 bq = 0;  // CDF base of rank q
      for (int q = 0; q < p.world; ++q) {
        const volatile u64* r = (const volatile u64*)rec + (size_t)q * kShardRec;
        const u64 tq = r[0];
        // rank q's margins serve its neighbours: its first blocks rank q-1's last slots, its last blocks rank q+1's first slots
        if (q > 0 && sm.slot_end(bq + r[16]) < (uint32_t)((int64_t)q * p.n_local)) overflow = 1;
        if (q + 1 < p.world && sm.slot_end(bq + r[17]) > (uint32_t)((int64_t)(q + 1) * p.n_local)) overflow = 1;
        if (q == p.rank) {  // my slots whose ancestor is not one of my particles
          const int64_t j_lo = (int64_t)q * p.n_local, j_hi = j_lo + p.n_local;
          const int64_t s0 = sm.slot_end(bq), s1 = sm.slot_end(bq + tq);
          const int64_t a0 = (s0 < j_hi ? s0 : j_hi) - j_lo, a1 = j_hi - (s1 > j_lo ? s1 : j_lo);
          migrated = (a0 > 0 ? a0 : 0) + (a1 > 0 ? a1 : 0);
        }
        bq += tq;
      }
    }
    bool comm_err = p.ctl->comm_error != 0;
    for (int q = 0; q < p.world; ++q) comm_err = comm_err || ((const volatile u64*)rec)[(size_t)q * kShardRec + 18] != 0;
    if (comm_err) overflow = 2;
    pl->shard_overflow = overflow;
    if (p.delta) {  // moment sums of all ranks, in rank order (every rank computes the same value)
      double sw = 0, sw2 = 0, sx[PZ_MAX_NX] = {}, sxx[PZ_MAX_NX] = {};
      for (int q = 0; q < p.world; ++q) {
        const volatile u64* r = (const volatile u64*)rec + (size_t)q * kShardRec;
        sw += __longlong_as_double((long long)r[1]);
        sw2 += __longlong_as_double((long long)r[2]);
        for (int d = 0; d < p.nx; ++d) {
          sx[d] += __longlong_as_double((long long)r[3 + d]);
          sxx[d] += __longlong_as_double((long long)r[3 + PZ_MAX_NX + d]);
        }
      }
      pl->sw = sw; pl->sw2 = sw2;
      for (int d = 0; d < p.nx; ++d) { pl->sx[d] = sx[d]; pl->sxx[d] = sxx[d]; }
      pz_step_summary* o = p.summ ? p.summ + (p.ctl->t % kObsRing) : nullptr;
      write_summary(pl, o, p.ctl->t, p.nx, p.n_global, p.ll_const, p.f64);
      if (o) o->n_migrated = overflow ? -(int64_t)overflow : migrated;
    }
  }
}


// ------------------------------------------------------------------------------------------
// pf_partition_sharded (peer-memory exchange): everything between the scan and the next step kernel in ONE
// launch -- the all-gather of the rank records through the mailboxes, the plan (every CTA derives it from the
// gathered records; CTA 0 stores it, the posterior summary and the margin verdict), the boundary positions of
// this rank's blocks (xb = block_pos(base + local prefix)) with the two margins stored straight into the
// neighbours' halos, and the merge-path split.  Nothing here waits for a neighbour's margins except the (few)
// boundary tiles whose first input block lies in a halo; the step kernel's boundary CTAs wait for their halos
// themselves, so the halo latency hides behind the interior tiles.
// ------------------------------------------------------------------------------------------
struct PartShardArgs {
  PlanDev* plan;
  CtlDev* ctl;
  u64* rec;            // [world][kShardRec]: this rank's record (written by pf_scan_blocks)
  const u64* pb;       // local prefix [nblk+1]
  u64* xb;             // boundary positions, own entries [0, nblk], halo entries [b_lo, 0) and (nblk, b_hi]
  int32_t* tile_start;
  u64* status;
  pz_step_summary* summ;
  int64_t nblk, n_global, n_local, slot_base, b_lo, b_hi, margin;
  int32_t rank, world, nx, delta, ntiles_out, tile_out, nscan, f64;
  uint64_t seed;
  double ll_const;
  const ShardP2P* p2p;
  unsigned long long wait_ticks;
};

__global__ void __launch_bounds__(kThreads) pf_partition_sharded(const PartShardArgs p) {
  __shared__ u64 s_rec[kMaxRanks * kShardRec];
  __shared__ u64 s_base[kMaxRanks + 1];   // CDF base of every rank
  __shared__ PlanDev s_plan;
  __shared__ int32_t s_overflow;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t t_cur = ((volatile CtlDev*)p.ctl)->t;
  const uint32_t epoch = t_cur + (uint32_t)p.delta + 1u;
  Mailbox* mine = p.p2p->mbox[p.rank];
  PZ_STAMP(blockIdx.x == 0 && tid == 0, 3);
  // ---- all-gather of the records: CTA 0 stores this rank's record into every mailbox as self-validating
  // half-words, every CTA polls its own mailbox until all of this epoch are there
  {
    const int nw = p.world * 2 * kShardRec;
    if (blockIdx.x == 0) {
      for (int i = tid; i < nw; i += kThreads) {
        const int q = i / (2 * kShardRec), w = i % (2 * kShardRec);
        const u64 v = p.rec[(size_t)p.rank * kShardRec + (w >> 1)];
        st_sys(&p.p2p->mbox[q]->rec2[p.rank][w], ((u64)epoch << 32) | (uint32_t)((w & 1) ? (v >> 32) : v));
      }
    }
    uint32_t* s_half = reinterpret_cast<uint32_t*>(s_rec);
    for (int i = tid; i < nw; i += kThreads) {
      const u64 x = wait_sys<false>(&mine->rec2[i / (2 * kShardRec)][i % (2 * kShardRec)],
                                    [epoch](unsigned long long x) { return (uint32_t)(x >> 32) == epoch; }, &p.ctl->comm_error,
                                    p.wait_ticks, 2u);
      s_half[i] = (uint32_t)x;
    }
    __syncthreads();
  }
  PZ_STAMP(blockIdx.x == 0 && tid == 0, 4);
  u64 base = 0, total = 0;
  for (int q = 0; q < p.world; ++q) {
    const u64 tq = s_rec[(size_t)q * kShardRec];
    if (q < p.rank) base += tq;
    total += tq;
  }
  // ---- the plan: every CTA derives it (one 128-bit division), CTA 0 publishes it
  if (tid == 0) {
    const uint32_t t_next = t_cur + (uint32_t)p.delta;
    const uint32_t U = philox4x32_10(0u, 0u, t_next, kStreamResample, (uint32_t)p.seed, (uint32_t)(p.seed >> 32)).x;
    s_plan.base = base;
    s_plan.emax = ((volatile PlanDev*)p.plan)->emax;
    finalize_plan(&s_plan, total, p.n_global, U);
    // margin verdict from the gathered records alone (the same on every rank, in every CTA): do the fixed margins
    // (the last `margin` blocks of rank q-1, the first of rank q+1) cover the ancestors of every rank's slots?
    int32_t overflow = 0;
    u64 bq = 0;
    SlotMap sm0(&s_plan, p.n_global);
    for (int q = 0; q < p.world; ++q) {
      const u64* r = s_rec + (size_t)q * kShardRec;
      s_base[q] = bq;
      if (total != 0) {
        if (q > 0 && sm0.slot_end(bq + r[16]) < (uint32_t)((int64_t)q * p.n_local)) overflow = 1;
        if (q + 1 < p.world && sm0.slot_end(bq + r[17]) > (uint32_t)((int64_t)(q + 1) * p.n_local)) overflow = 1;
      }
      bq += r[0];
    }
    s_base[p.world] = bq;
    s_overflow = overflow;
  }
  __syncthreads();
  SlotMap sm(&s_plan, p.n_global);
  const bool remote = s_overflow != 0;   // the next resample reads remote blocks directly
  const int64_t gid = (int64_t)blockIdx.x * kThreads + tid, gsz = (int64_t)gridDim.x * kThreads;
  const int64_t nb = p.nblk, m = p.margin;
  // ---- boundary positions of my blocks; the margins also go into the neighbours' halos (my boundaries 1..m are
  // the RIGHT halo of rank-1 at its indices nb+1..nb+m; my boundaries nb-m..nb-1 the LEFT halo of rank+1 at -m..-1)
  // (the first npush CTAs store the margins -- and only they pay a system-scope fence, one per CTA)
  const int npush = (int)min((int64_t)gridDim.x, (m + kThreads - 1) / kThreads);
  if ((int)blockIdx.x < npush) {
    for (int64_t i = gid; i < m; i += (int64_t)npush * kThreads) {
      if (p.rank > 0) p.p2p->xbn[0][nb + 1 + i] = total != 0 ? sm.block_pos(base + p.pb[1 + i]) : 0ull;
      if (p.rank + 1 < p.world) p.p2p->xbn[1][i - m] = total != 0 ? sm.block_pos(base + p.pb[nb - m + i]) : 0ull;
    }
  }
  for (int64_t b = gid; b <= nb; b += gsz) p.xb[b] = total != 0 ? sm.block_pos(base + p.pb[b]) : 0ull;
  PZ_STAMP(blockIdx.x == 0 && tid == 0, 6);
  __syncthreads();
  if (tid == 0) {
    // "my xb is complete" needs every CTA; the margin flags only the pushing ones.  One counter: every CTA arrives
    // (device-scope fence: its xb stores are in L2, where the peers' NVLink reads go), the pushing ones after a
    // system-scope fence; the last arrival raises all flags.
    if ((int)blockIdx.x < npush) __threadfence_system(); else __threadfence();
    if (atomicAdd(&mine->pbx_done, 1ull) == (unsigned long long)gridDim.x - 1ull) {
      __threadfence_system();
      mine->pbx_done = 0ull;
      if (p.rank > 0) st_sys(&p.p2p->mbox[p.rank - 1]->pbx_flag[1], (unsigned long long)epoch);
      if (p.rank + 1 < p.world) st_sys(&p.p2p->mbox[p.rank + 1]->pbx_flag[0], (unsigned long long)epoch);
      for (int q = 0; q < p.world; ++q) st_sys(&p.p2p->mbox[q]->xb_done[p.rank], (unsigned long long)epoch);  // my xb is complete
      PZ_STAMP(true, 8);
    }
  }
  // ---- CTA 0: plan, posterior summary, margin verdict (from the gathered records alone: the same on every rank)
  if (gid == 0) {
    PlanDev* pl = p.plan;
    pl->base = base;
    pl->shard_epoch = epoch;
    pl->total = s_plan.total; pl->degenerate = s_plan.degenerate; pl->U = s_plan.U; pl->L = s_plan.L; pl->rho = s_plan.rho;
    pl->Mq = s_plan.Mq; pl->h0 = s_plan.h0;
    int32_t overflow = s_overflow;
    int64_t migrated = 0;
    if (total != 0) {  // my slots whose ancestor is not one of my particles
      const int64_t j_lo = (int64_t)p.rank * p.n_local, j_hi = j_lo + p.n_local;
      const int64_t s0 = sm.slot_end(base), s1 = sm.slot_end(base + s_rec[(size_t)p.rank * kShardRec]);
      const int64_t a0 = (s0 < j_hi ? s0 : j_hi) - j_lo, a1 = j_hi - (s1 > j_lo ? s1 : j_lo);
      migrated = (a0 > 0 ? a0 : 0) + (a1 > 0 ? a1 : 0);
    }
    bool comm_err = ((volatile CtlDev*)p.ctl)->comm_error != 0;
    for (int q = 0; q < p.world; ++q) comm_err = comm_err || s_rec[(size_t)q * kShardRec + 18] != 0;
    if (comm_err) overflow = 2;
    pl->shard_overflow = overflow;
    if (p.delta) {  // moment sums of all ranks, in rank order (every rank computes the same value)
      double sw = 0, sw2 = 0, sx[PZ_MAX_NX] = {}, sxx[PZ_MAX_NX] = {};
      for (int q = 0; q < p.world; ++q) {
        const u64* r = s_rec + (size_t)q * kShardRec;
        sw += __longlong_as_double((long long)r[1]);
        sw2 += __longlong_as_double((long long)r[2]);
        for (int d = 0; d < p.nx; ++d) {
          sx[d] += __longlong_as_double((long long)r[3 + d]);
          sxx[d] += __longlong_as_double((long long)r[3 + PZ_MAX_NX + d]);
        }
      }
      pl->sw = sw; pl->sw2 = sw2;
      for (int d = 0; d < p.nx; ++d) { pl->sx[d] = sx[d]; pl->sxx[d] = sxx[d]; }
      pz_step_summary* o = p.summ ? p.summ + (t_cur % kObsRing) : nullptr;
      write_summary(pl, o, t_cur, p.nx, p.n_global, p.ll_const, p.f64);
      // (margins too small -- overflow == 1 -- is not an error: the next step takes the remote-read path)
      if (o) o->n_migrated = overflow == 2 ? -2 : migrated;
    }
  }
  // ---- merge-path split: one warp per output tile.  The first input block of a tile lies among my own blocks
  // unless the tile starts before my first block's children end (left halo) or after my last block's (right halo)
  const u64 P0 = base, PN = base + p.pb[nb];
  const uint32_t end0 = total != 0 ? sm.slot_end(P0) : 0u, endN = total != 0 ? sm.slot_end(PN) : 0u;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t tile = gid >> 5; tile < p.ntiles_out; tile += nwarps) {
    const uint32_t j0 = (uint32_t)(p.slot_base + tile * p.tile_out);
    int64_t lo, hi;
    if (remote) {
      // the margins do not cover this resample: GLOBAL split over the blocks of all ranks, reading the owners' local
      // prefixes over NVLink (they are final: every rank's record has been gathered)
      lo = 0; hi = (int64_t)p.world * nb;
      while (hi > lo) {
        const int64_t n = hi - lo, step = (n + 31) >> 5;
        int64_t off = (int64_t)(lane + 1) * step;
        off = off < n ? off : n;
        const int64_t mid = lo + off - 1, gb1 = mid + 1;          // boundary gb1 = end of global block mid
        const int q = (int)(gb1 / nb);
        const u64 C = q >= p.world ? total : s_base[q] + ((volatile const u64*)p.p2p->pball[q])[gb1 - (int64_t)q * nb];
        const bool pred = total != 0 && sm.slot_end(C) > j0;
        const uint32_t mk = __ballot_sync(0xffffffffu, pred);
        if (mk == 0) { lo = hi; break; }
        const int f = __ffs(mk) - 1;
        int64_t offf = (int64_t)(f + 1) * step;
        offf = offf < n ? offf : n;
        hi = lo + offf - 1;
        lo = lo + (int64_t)f * step;
      }
      if (lane == 0) p.tile_start[tile] = (int32_t)lo;   // a GLOBAL block index
      continue;
    }
    int side = -1;   // -1: own blocks; 0 / 1: the left / right halo
    if (p.rank > 0 && end0 > j0) { lo = p.b_lo; hi = 0; side = 0; }
    else if (p.rank + 1 < p.world && endN <= j0) { lo = nb; hi = p.b_hi; side = 1; }
    else { lo = 0; hi = nb; }
    if (side >= 0) {  // this epoch's boundary positions from the neighbour
      if (lane == 0)
        wait_sys(&mine->pbx_flag[side], [epoch](unsigned long long x) { return x >= (unsigned long long)epoch; }, &p.ctl->comm_error, p.wait_ticks, 0x10u);
      __syncwarp();
      __threadfence_system();
    }
    // 32-ary search for min b in [lo, hi) with slot_end(boundary b + 1) > j0 (hi when there is none)
    while (hi > lo) {
      const int64_t n = hi - lo, step = (n + 31) >> 5;
      int64_t off = (int64_t)(lane + 1) * step;
      off = off < n ? off : n;
      const int64_t mid = lo + off - 1;
      uint32_t se;
      if (side < 0 || mid + 1 == 0 || mid + 1 == nb) se = sm.slot_end(base + p.pb[mid + 1]);    // my own boundaries (0 and nb included)
      else se = min((uint32_t)(((volatile u64*)p.xb)[mid + 1] >> 32), (uint32_t)p.n_global);  // a halo entry
      const bool pred = total != 0 && se > j0;
      const uint32_t mk = __ballot_sync(0xffffffffu, pred);
      if (mk == 0) { lo = hi; break; }
      const int f = __ffs(mk) - 1;
      int64_t offf = (int64_t)(f + 1) * step;
      offf = offf < n ? offf : n;
      hi = lo + offf - 1;
      lo = lo + (int64_t)f * step;
    }
    // (a tile of the left-halo case whose search found nothing starts at my block 0; of the right-halo case: b_hi)
    if (lane == 0) p.tile_start[tile] = (int32_t)lo;
  }
  if (gid < p.nscan) p.status[gid] = 0ull;
  PZ_STAMP(blockIdx.x == 0 && tid == 0, 10);
  // ---- the last CTA to finish advances the step counter (every CTA read it above)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&mine->part_done, 1ull) == (unsigned long long)gridDim.x - 1ull) {
      mine->part_done = 0ull;
      PZ_STAMP(true, 9);
      p.ctl->t = t_cur + (uint32_t)p.delta;
      p.ctl->emax_acc = kNZero;
      p.ctl->scan_counter = 0;
      p.ctl->scan_done = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Output slots per CTA: (a multiple of the CTA's warp count) - 1 canonical blocks, because a tile of m
// output blocks draws on m + 1 input blocks in the typical case and the warps of the CTA take input
// blocks round-robin (a remainder of one block would leave the other warps waiting at the barrier).  Larger tiles amortise the
// extra input block and the CTA-level reduction.
// CTA size and tile size of the step kernel (measured, profiles/README.md): half-size CTAs put twice as
// many independent, phase-shifted tiles on an SM and halve the barrier domain.  A double filter stages
// 8 bytes per slot and dimension, so its tiles are about half as long for the same shared memory.
#ifndef PZ_NT1
#define PZ_NT1 128   // threads per CTA, NX = 1
#endif
#ifndef PZ_TB1
#define PZ_TB1 23    // output blocks per tile, NX = 1, float (a multiple of the warp count, minus 1)
#endif
#ifndef PZ_TB1D
#define PZ_TB1D 23   // output blocks per tile, NX = 1, double (index-staged: 4 bytes per slot)
#endif
#ifndef PZ_NT4D
#define PZ_NT4D 128  // threads per CTA, NX = 4, double
#endif
#ifndef PZ_NTN
#define PZ_NTN 256   // threads per CTA, NX > 1
#endif
#ifndef PZ_TBN
#define PZ_TBN 23    // output blocks per tile, NX > 1, float
#endif
#ifndef PZ_TBND
#define PZ_TBND 23   // output blocks per tile, NX > 1, double
#endif
#ifndef PZ_OCC4D
#define PZ_OCC4D 2   // resident CTAs per SM of the double NX = 4 kernel
#endif
#ifndef PZ_OCC6D
#define PZ_OCC6D 2   // resident CTAs per SM of the double NX = 6 kernel
#endif
#ifndef PZ_OCC1D
#define PZ_OCC1D 4   // resident 256-thread-equivalents per SM the double NX = 1 kernel's registers are held to (8 CTAs of 128)
#endif
// small = true: the second tile size (PZ_TILE_OUT=small), kept so that the tests can check that results do
// not depend on the tiling
constexpr int tile_blocks(int nx, bool f64, bool small) {
  return nx == 1 ? (f64 ? (small ? 7 : PZ_TB1D) : (small ? 15 : PZ_TB1)) : (f64 ? (small ? 3 : PZ_TBND) : (small ? 7 : PZ_TBN));
}
static bool tile_small() {
  static const bool small = [] { const char* e = getenv("PZ_TILE_OUT"); return e && !strcmp(e, "small"); }();
  return small;
}
int fused_tile_out(int nx, int f64) { return tile_blocks(nx, f64 != 0, tile_small()) * kBlk; }

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
  sc.rank = a.rank; sc.shard_rec = a.shard_rec; sc.margin = a.margin; sc.p2p = a.p2p;
  sc.f64 = a.f64; sc.wait_ticks = a.wait_ticks;
  sc.npush = 0;
  if (a.p2p != nullptr) {
    const int esz = a.f64 ? 8 : 4;
    sc.push = HaloPushArgs{a.ctl, a.p2p, {a.X[0], a.X[1]}, esz, {a.Q[0], a.Q[1]}, {a.eb[0], a.eb[1]}, a.nblk, a.ld, a.margin,
                           a.rank, a.world, a.model.nx, delta};
    // 16-byte stores per margin and side: nx * m * 16 * esz (X) + m * 32 (Q); two per thread, at most two CTAs per SM beside the scan
    int grid = cdiv((int64_t)(a.model.nx * 16 * esz + 32) * a.margin, 2 * kThreads);
    sc.npush = grid < 1 ? 1 : (grid > 296 ? 296 : grid);
  }
  return sc;
}

static int partition_grid(int64_t ntiles, int64_t nscan, int64_t nbound) {
  int64_t nthreads = ntiles * 32 > nscan ? ntiles * 32 : nscan;  // one warp per output tile
  if (nbound > nthreads) nthreads = nbound;  // one block boundary per thread (grid-stride beyond 4096 CTAs)
  const int g = cdiv(nthreads, kThreads);
  return g > 4096 ? 4096 : g;
}

// PZ_PDL=1: the kernels of a single-GPU fused step are launched with programmatic stream serialization
static bool use_pdl(const StepArgs& a) {
  static const int on = [] { const char* e = getenv("PZ_PDL"); return e ? atoi(e) : 0; }();
  return on != 0 && a.p2p == nullptr && a.world == 1;
}
template <class Kern, class Arg>
static void launch_maybe_pdl(Kern kern, int grid, int block, size_t smem, cudaStream_t s, bool pdl, const Arg& arg) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, arg);
}

static int partition(const StepArgs& a, int tile_out, int delta, cudaStream_t s) {
  PartArgs pa{a.pbx, a.xb, a.b_lo, a.b_hi, a.tile_start, a.scan_status, a.plan, a.ctl, a.nblk, a.slot_base, a.n_global,
              a.ntiles_out, tile_out, a.nscan, delta, a.p2p, a.rank, a.world, a.wait_ticks};
  int grid = partition_grid(a.ntiles_out, a.nscan, a.b_hi - a.b_lo + 1);
  if (a.p2p && grid > 592) grid = 592;  // every CTA polls a flag: keep the whole grid resident
  if (use_pdl(a)) launch_maybe_pdl(pf_partition, grid, kThreads, 0, s, true, pa);
  else pf_partition<<<grid, kThreads, 0, s>>>(pa);
  return 1;
}

int launch_partition_only(const StepArgs& a, int tile_out, cudaStream_t s) { return partition(a, tile_out, 0, s); }

// Scan + partition of the population in buffer (ctl->t + delta) & 1.  delta = 1 after a step.
static int scan_partition(const StepArgs& a, int delta, int tile_out, cudaStream_t s) {
  ScanArgs sc = make_scan(a, delta, delta == 1);
  if (use_pdl(a)) launch_maybe_pdl(pf_scan_blocks, a.nscan + sc.npush, kThreads, 0, s, true, sc);
  else pf_scan_blocks<<<a.nscan + sc.npush, kThreads, 0, s>>>(sc);
  partition(a, tile_out, delta, s);
  return 2;
}

int launch_scan_partition(const StepArgs& a, int tile_out, cudaStream_t s) {
  return scan_partition(a, 0, tile_out, s);
}

template <class T, int NX, int NY, int MODE, bool KRON, bool SMALL>
static int launch_tile(const TileArgs& ta, cudaStream_t s) {
  constexpr bool F64 = sizeof(T) == 8;
  // (LG(4,2) fp64 runs 128-thread CTAs, 4 per SM: 0.1730 -> 0.1676 ms at N = 10^7; the same costs LG(6,3) fp64 30 %)
  constexpr int NT = (NX == 1) ? PZ_NT1 : ((NX == 4 && F64) ? PZ_NT4D : PZ_NTN);
  constexpr int TO = tile_blocks(NX, F64, SMALL) * kBlk;
  constexpr int OCC = F64 ? (NX == 1 ? PZ_OCC1D : (NX <= 4 ? PZ_OCC4D : PZ_OCC6D)) : min_ctas(NX);
  // resident CTAs per SM the registers are held to, in units of the kernel's own CTA size (PZ_MINB1 / PZ_MINB1D: A/B hooks
  // for the one-word states, whose CTAs have 128 threads)
#if defined(PZ_MINB1D) && defined(PZ_MINB1)
  constexpr int MINB = NX == 1 ? (F64 ? PZ_MINB1D : PZ_MINB1) : OCC * (kThreads / NT);
#elif defined(PZ_MINB1D)
  constexpr int MINB = (F64 && NX == 1) ? PZ_MINB1D : OCC * (kThreads / NT);
#elif defined(PZ_MINB1)
  constexpr int MINB = (!F64 && NX == 1) ? PZ_MINB1 : OCC * (kThreads / NT);
#else
  constexpr int MINB = OCC * (kThreads / NT);
#endif
  auto kern = pf_step_tile<T, NX, NY, TO, MODE, KRON, NT, MINB>;
  // staged parents [NX][TO] (or parent indices [TO]) + inverse-CDF table
  const size_t smem_state = (size_t)NX * TO * sizeof(T) + PZ_ICDF_ROWS * sizeof(float4);
  const size_t smem = kIdxStage<T, NX, MODE>
                          ? (size_t)TO * sizeof(float) + (NX > 1 ? (size_t)(NT / 32) * NX * kBlk * sizeof(float) : 0) + PZ_ICDF_ROWS * sizeof(float4)
                          : smem_state;
  // function attributes are per context (device); concurrent first launches set the same value
  static std::atomic<unsigned long long> attr_done{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(attr_done.load(std::memory_order_acquire) & bit)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  // persistent grid: as many CTAs as are resident at once (PZ_PERSIST=0: one CTA per tile; PZ_PERSIST=k: k CTAs per SM)
  static std::atomic<int> resident[64];
  int grid = resident[dev & 63].load(std::memory_order_acquire);
  if (grid == 0) {
    int nsm = 148, per_sm = 1;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    const char* e = getenv("PZ_PERSIST");
    if (e && atoi(e) > 0) per_sm = atoi(e);
    grid = (!kPersist<T, NX, false> || (e && atoi(e) == 0)) ? -1 : nsm * (per_sm > 0 ? per_sm : 1);
    resident[dev & 63].store(grid, std::memory_order_release);
  }
  if (grid < 0 || grid > ta.a.ntiles_out) grid = ta.a.ntiles_out;
  if (MODE == 0 && use_pdl(ta.a)) launch_maybe_pdl(kern, grid, NT, smem, s, true, ta);
  else kern<<<grid, NT, smem, s>>>(ta);
  if constexpr (MODE == 0) {
    if (ta.a.p2p != nullptr) {   // sharded over peer memory: the remote-read fallback follows (it exits at once unless the margins overflowed)
      auto rkern = pf_step_tile<T, NX, NY, TO, MODE, KRON, NT, MINB, true>;
      static std::atomic<unsigned long long> rattr_done{0};
      if (!(rattr_done.load(std::memory_order_acquire) & bit)) {
        cudaFuncSetAttribute(rkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        rattr_done.fetch_or(bit, std::memory_order_release);
      }
      const int rgrid = ta.a.ntiles_out < 148 ? ta.a.ntiles_out : 148;
      rkern<<<rgrid, NT, smem, s>>>(ta);
      return 2;
    }
  }
  return 1;
}

static int dispatch_tile(const TileArgs& ta, cudaStream_t s) {
  const int nx = ta.a.model.nx, ny = ta.a.model.ny;
  const bool fused = (ta.anc == nullptr);
  const bool kron = ta.a.model.kron != 0;
  const bool f64 = ta.a.f64 != 0;
  const bool small = tile_small();
#define PZ_DISPATCH(Tv, NXv, NYv, KR)                                     \
  if (f64 == (sizeof(Tv) == 8) && nx == NXv && ny == NYv && kron == KR) { \
    if (small) {                                                          \
      if (fused) return launch_tile<Tv, NXv, NYv, 0, KR, true>(ta, s);    \
      return launch_tile<Tv, NXv, NYv, 1, KR, true>(ta, s);               \
    }                                                                     \
    if (fused) return launch_tile<Tv, NXv, NYv, 0, KR, false>(ta, s);     \
    return launch_tile<Tv, NXv, NYv, 1, KR, false>(ta, s);                \
  }
  PZ_DISPATCH(float, 1, 1, false)
  PZ_DISPATCH(float, 1, 1, true)
  PZ_DISPATCH(float, 4, 2, false)
  PZ_DISPATCH(float, 4, 2, true)
  PZ_DISPATCH(float, 6, 3, false)
  PZ_DISPATCH(float, 6, 3, true)
  PZ_DISPATCH(double, 1, 1, false)
  PZ_DISPATCH(double, 1, 1, true)
  PZ_DISPATCH(double, 4, 2, false)
  PZ_DISPATCH(double, 4, 2, true)
  PZ_DISPATCH(double, 6, 3, false)
  PZ_DISPATCH(double, 6, 3, true)
#undef PZ_DISPATCH
  return -1;
}

// anc == nullptr: fused systematic step; else parents come from the stored ancestor array.
int launch_step(const StepArgs& a, const int32_t* anc, cudaStream_t s) {
  TileArgs ta{a, anc};
  const int nt = dispatch_tile(ta, s);
  if (nt < 0) return -1;
  return nt + scan_partition(a, 1, fused_tile_out(a.model.nx, a.f64), s);
}

int launch_step_profiled(const StepArgs& a, cudaStream_t s, cudaEvent_t ev[4]) {
  TileArgs ta{a, nullptr};
  cudaEventRecord(ev[0], s);
  if (dispatch_tile(ta, s) < 0) return -1;
  cudaEventRecord(ev[1], s);
  ScanArgs sc = make_scan(a, 1, true);
  pf_scan_blocks<<<a.nscan + sc.npush, kThreads, 0, s>>>(sc);
  cudaEventRecord(ev[2], s);
  partition(a, fused_tile_out(a.model.nx, a.f64), 1, s);
  cudaEventRecord(ev[3], s);
  return 3;
}

int launch_tile_only(const StepArgs& a, cudaStream_t s) {
  TileArgs ta{a, nullptr};
  return dispatch_tile(ta, s);
}

int launch_scan_only(const StepArgs& a, int delta, cudaStream_t s) {
  ScanArgs sc = make_scan(a, delta, delta == 1);
  pf_scan_blocks<<<a.nscan + sc.npush, kThreads, 0, s>>>(sc);
  return 1;
}

#ifdef PZ_DBG_STAMPS
void dbg_read_stamps(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_stamp, sizeof(g_stamp)); }
#endif

int launch_plan_sharded(const StepArgs& a, int delta, cudaStream_t s) {
  ShardPlanArgs p{a.plan, a.ctl, a.shard_rec, a.pb, const_cast<u64*>(a.pbx), a.summ, a.nblk, a.n_global, a.n_local,
                  a.rank, a.world, a.model.nx, delta, a.seed, a.ll_const, a.p2p, a.margin, a.f64, a.wait_ticks};
  int grid = cdiv(a.nblk + 1, kThreads);
  grid = grid > 592 ? 592 : grid;  // every CTA polls a flag: keep the whole grid resident (148 SMs x 4)
  pf_plan_sharded<<<grid, kThreads, 0, s>>>(p);
  return 1;
}

int launch_partition_ext(const StepArgs& a, int tile_out, int delta, cudaStream_t s) { return partition(a, tile_out, delta, s); }

int launch_partition_sharded(const StepArgs& a, int tile_out, int delta, cudaStream_t s) {
  PartShardArgs p{};
  p.plan = a.plan; p.ctl = a.ctl; p.rec = a.shard_rec; p.pb = a.pb; p.xb = a.xb; p.tile_start = a.tile_start; p.status = a.scan_status;
  p.summ = a.summ; p.nblk = a.nblk; p.n_global = a.n_global; p.n_local = a.n_local; p.slot_base = a.slot_base; p.b_lo = a.b_lo; p.b_hi = a.b_hi;
  p.margin = a.margin; p.rank = a.rank; p.world = a.world; p.nx = a.model.nx; p.delta = delta; p.ntiles_out = a.ntiles_out; p.tile_out = tile_out;
  p.nscan = a.nscan; p.f64 = a.f64; p.seed = a.seed; p.ll_const = a.ll_const; p.p2p = a.p2p; p.wait_ticks = a.wait_ticks;
  int grid = partition_grid(a.ntiles_out, a.nscan, a.nblk + 1);
  if (grid > 592) grid = 592;  // every CTA polls its mailbox: keep the whole grid resident (148 SMs x 4)
  pf_partition_sharded<<<grid, kThreads, 0, s>>>(p);
  return 1;
}


int launch_ctl_reset(CtlDev* ctl, cudaStream_t s) {
  pf_ctl_reset<<<1, 1, 0, s>>>(ctl);
  return 1;
}

int launch_generic_plan(const GenericArgs& g, bool stats_from_lw, cudaStream_t s) {
  int n = 0;
  if (stats_from_lw) {
    pf_stats_from_lw<<<cdiv(g.nblk, kWarps), kThreads, 0, s>>>(g.lw, g.n, g.nblk, g.eb, g.sb, g.q16, g.ctl);
    ++n;
  }
  ScanArgs sc{};
  sc.eb0 = g.eb; sc.eb1 = g.eb; sc.sb = g.sb; sc.pb = g.pb; sc.status = g.scan_status; sc.part = g.scan_part; sc.rec = nullptr;
  sc.plan = g.plan; sc.ctl = g.ctl; sc.summ = nullptr; sc.nblk = g.nblk; sc.n_out_global = g.n; sc.n_local = g.n;
  sc.nscan = g.nscan; sc.nrec = 0; sc.nx = 0; sc.delta = 0; sc.explicit_U = 1; sc.U_value = g.U; sc.seed = g.seed;
  sc.ll_const = 0; sc.world = 1; sc.rank = 0; sc.shard_rec = nullptr; sc.margin = 0; sc.p2p = nullptr;
  sc.f64 = 0; sc.wait_ticks = 0;
  pf_scan_blocks<<<g.nscan, kThreads, 0, s>>>(sc);
  PartArgs pa{g.pb, g.xb, 0, g.nblk, g.tile_start, g.scan_status, g.plan, g.ctl, g.nblk, 0, g.n, g.ntiles_out, kGenericTileOut, g.nscan, 0,
              nullptr, 0, 1, 0ull};
  pf_partition<<<partition_grid(g.ntiles_out, g.nscan, g.nblk + 1), kThreads, 0, s>>>(pa);
  return n + 2;
}

int launch_stats_from_lw(const GenericArgs& g, cudaStream_t s) {
  pf_stats_from_lw<<<cdiv(g.nblk, kWarps), kThreads, 0, s>>>(g.lw, g.n, g.nblk, g.eb, g.sb, g.q16, g.ctl);
  return 1;
}

int launch_generic_ancestors(const GenericArgs& g, cudaStream_t s) {
  pf_ancestors<kGenericTileOut><<<g.ntiles_out, kThreads, 0, s>>>(g);   // tiles [g.tile0, g.tile0 + g.ntiles_out)
  return 1;
}

// slices of a device's arrays -> the same offsets of its peers' arrays
__global__ void __launch_bounds__(kThreads) pf_push_slices(const __grid_constant__ PushArgs p) {
  const int64_t gid = (int64_t)blockIdx.x * kThreads + threadIdx.x, gsz = (int64_t)gridDim.x * kThreads;
  for (int a = 0; a < p.n_arr; ++a) {
    const int64_t n16 = p.bytes[a] / 16, n4 = (p.bytes[a] - n16 * 16) / 4;
    const uint4* src = reinterpret_cast<const uint4*>(p.src[a]);
    for (int64_t i = gid; i < n16; i += gsz) {
      const uint4 v = src[i];
      for (int q = 0; q < p.n_peers; ++q) reinterpret_cast<uint4*>(p.dst[q][a])[i] = v;
    }
    if (gid < n4) {
      const uint32_t v = reinterpret_cast<const uint32_t*>(src + n16)[gid];
      for (int q = 0; q < p.n_peers; ++q) reinterpret_cast<uint32_t*>(reinterpret_cast<uint4*>(p.dst[q][a]) + n16)[gid] = v;
    }
  }
}
int launch_push_slices(const PushArgs& p, cudaStream_t s) {
  int64_t mx = 0;
  for (int a = 0; a < p.n_arr; ++a) mx = p.bytes[a] > mx ? p.bytes[a] : mx;
  int grid = (int)((mx / 16 + kThreads - 1) / kThreads);
  grid = grid < 1 ? 1 : (grid > 592 ? 592 : grid);
  pf_push_slices<<<grid, kThreads, 0, s>>>(p);
  return 1;
}

// global block exponent of a gathered eb array (every device of a multi-device tracker holds all of them)
__global__ void __launch_bounds__(kThreads) pf_emax_from_eb(const int32_t* __restrict__ eb, int64_t nblk, CtlDev* ctl) {
  int32_t m = kNZero;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nblk; i += (int64_t)gridDim.x * kThreads) m = max(m, eb[i]);
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m != kNZero) atomicMax(&ctl->emax_acc, m);
}
int launch_emax_from_eb(const int32_t* eb, int64_t nblk, CtlDev* ctl, cudaStream_t s) {
  int grid = cdiv(nblk, kThreads);
  grid = grid > 148 ? 148 : grid;
  pf_emax_from_eb<<<grid, kThreads, 0, s>>>(eb, nblk, ctl);
  return 1;
}

int launch_generic_multinomial(const GenericArgs& g, cudaStream_t s) {
  pf_multinomial<<<cdiv(g.n_draws > 0 ? g.n_draws : g.n, kThreads), kThreads, 0, s>>>(g);
  return 1;
}

int launch_logw(const StepArgs& a, float* lw_out32, double* lw_out64, cudaStream_t s) {
  const int nx = a.model.nx, ny = a.model.ny;
  int grid = cdiv(a.n_local, kThreads);
#define PZ_LOGW(Tv, NXv, NYv) pf_logw<Tv, NXv, NYv><<<grid, kThreads, 0, s>>>(a, lw_out32, lw_out64)
  if (a.f64) {
    if (nx == 1 && ny == 1) PZ_LOGW(double, 1, 1);
    else if (nx == 4 && ny == 2) PZ_LOGW(double, 4, 2);
    else if (nx == 6 && ny == 3) PZ_LOGW(double, 6, 3);
    else return -1;
  } else {
    if (nx == 1 && ny == 1) PZ_LOGW(float, 1, 1);
    else if (nx == 4 && ny == 2) PZ_LOGW(float, 4, 2);
    else if (nx == 6 && ny == 3) PZ_LOGW(float, 6, 3);
    else return -1;
  }
#undef PZ_LOGW
  return 1;
}

int launch_gather_thin(const StepArgs& a, const int32_t* anc, int32_t stride, int64_t n_sel, double* out,
                       cudaStream_t s) {
  pf_gather_thin<<<cdiv(n_sel, kThreads), kThreads, 0, s>>>(a, anc, stride, n_sel, out);
  return 1;
}

}  // namespace pz
