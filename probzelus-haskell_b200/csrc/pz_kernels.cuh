// pz_kernels.cuh -- device data structures and kernel entry points of the particle-filter
// step (one time step of zparticles', haskell/src/Inference.hs:101-104).
//
// HBM layout (per device, structure-of-arrays, all arrays padded to kPad particles):
//   X[2][NX][ld]    fp32 particle state, double-buffered: step t reads X[t&1], writes X[(t+1)&1]
//   Q[2][ld]        u16 fixed-point weight of every particle at its block's scale 2^(n_b - 15)
//   eb[2][nblk]     per canonical block (256 particles): block exponent n_b
//   sb[nblk]        per block: sum of its q
//   pb[nblk+1]      exclusive prefix of the block sums rescaled to the global exponent
//   xb[nblk+1]      threshold-shifted slot position of every block boundary (2^-32 slots)
//   tile_start[]    first input block of every output tile (merge-path partition)
//   rec[]           per output tile: centred moment partial sums for the posterior summary
// Log-weights and ancestors are never stored on the fused path: a particle carries its 16-bit
// weight (4 bytes of traffic per update instead of ~19 instructions to recompute it), and the
// ancestor gather is staged through shared memory inside the kernel that propagates the child.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/pzgpu.h"

// -DPZ_BOUNDS (make -C csrc bounds): device-side assertions on every staged / gathered index of the resampling and tracker
// kernels.  compute-sanitizer is closed on the GPU pool this build is measured on, so the parity suite is run once against
// this build instead (tools/gpu_bounds.sh): a violated bound traps the kernel and fails the test with a CUDA error.
#ifdef PZ_BOUNDS
#include <cassert>
#define PZ_ASSERT(x) assert(x)
#else
#define PZ_ASSERT(x) ((void)0)
#endif

namespace pz {

constexpr int kBlk = 256;        // canonical resampling block (particles)
constexpr int kThreads = 256;    // CTA size of every kernel
constexpr int kWarps = kThreads / 32;
constexpr int kScanItems = 8;    // block sums per thread in pf_scan_blocks
constexpr int kScanTile = kThreads * kScanItems;
constexpr int kPad = 2048;       // allocation granularity in particles
constexpr int kObsRing = 4096;   // rows of the device observation / summary rings
constexpr int kMaxRec = 3 + 2 * PZ_MAX_NX;  // doubles per moment record

// Derived parameters of the linear-Gaussian family (RW1D = LG(1,1)) in the arithmetic type of the
// filter: T = float (state_f64 = 0) or double (state_f64 = 1, the reference's own precision,
// Inference.hs:28 `type Weight = Log Double`).
template <class T>
struct LgDevT {
  T F[PZ_MAX_NX][PZ_MAX_NX];
  T Lq[PZ_MAX_NX][PZ_MAX_NX];
  T H[PZ_MAX_NY][PZ_MAX_NX];
  T Rinv[PZ_MAX_NY][PZ_MAX_NY];
  T x0[PZ_MAX_NX];
  T wscale;
  int32_t nx, ny;
  // Position/velocity ("Kronecker") structure, detected on the host: state = (pos[p], vel[p]),
  // F = [[a, b], [c, d]] (x) I_p, Q = diag(qp I, qv I), H = [I_p 0], R = r I_p.  The specialised
  // code path performs exactly the operations of the dense spec that have non-zero coefficients,
  // in the same order, so the results are bit-identical.
  int32_t kron;
  T ka, kb, kc, kd, klp, klv, krinv;
};
typedef LgDevT<float> LgDev;
typedef LgDevT<double> LgDevD;

// Resampling plan + posterior moment sums of one weighted population (device resident).
struct PlanDev {
  uint64_t total;       // T: global fixed-point total
  uint64_t base;        // CDF offset of this rank's first particle (0 on one GPU)
  uint64_t local_total; // this rank's share of T
  uint64_t rho;         // floor(N * 2^(32+L) / T) + 1: slot position of C is floor(C * rho / 2^L) * 2^-32
  uint32_t Mq;          // 30-bit slope multiplier floor(rho / 2^lam) + 1, lam = bitlen(rho) - 30
  int32_t h0;           // L - kGuard - lam: a block at exponent distance ksh uses the shift h0 + ksh
  int32_t L;            // bit length of T
  int32_t emax;         // global block exponent
  int32_t degenerate;   // 1 when T == 0
  int32_t shard_overflow;  // sharded: 1 when some rank's ancestors reach beyond the halo margins, 2: a peer never answered
  uint32_t shard_epoch;    // sharded: epoch of the exchange that produced this plan
  uint32_t U;           // the uniform of the resample that will consume this plan
  float center[8];      // fp32 posterior mean: centring constants of the next moment sums
  double center64[8];   // the same in double (f64 filters centre on the unrounded mean)
  double sw, sw2;       // sum w, sum w^2 at scale 2^emax
  double sx[PZ_MAX_NX], sxx[PZ_MAX_NX];  // centred first / second moment sums
};

// Step counter and accumulators living in device memory so that one captured CUDA graph
// serves every step.
struct CtlDev {
  uint32_t t;             // index of the observation the NEXT fused step assimilates
  int32_t emax_acc;       // atomicMax target of the kernel producing block statistics
  uint32_t scan_counter;  // dynamic tile ids of pf_scan_blocks
  uint32_t scan_done;     // "last CTA" counter of pf_scan_blocks
  uint32_t comm_error;    // sharded: a peer-memory wait timed out
  uint32_t ess_resampled; // adaptive resampling (generic path): 1 when the last step resampled
  uint32_t pad[2];
};

struct ShardP2P;  // peer-memory exchange of a sharded filter (below)

struct StepArgs {
  void* X[2];              // particle state [NX][ld] of float (f64 = 0) or double (f64 = 1), halo layout as Q
  uint16_t* Q[2];          // 16-bit weights of the two populations (same halo layout as X)
  int32_t* eb[2];
  uint64_t* sb;
  uint64_t* pb;            // local exclusive prefix of the rescaled block sums [nblk+1]
  const uint64_t* pbx;     // CDF prefix the step kernel reads: == pb on one GPU; on a sharded filter the
                           // ABSOLUTE prefix (rank base added), with halo entries at [-hb, 0) and (nblk, nblk+hb]
  uint64_t* xb;            // slot positions of the block boundaries [b_lo, b_hi], written by pf_partition
  int64_t b_lo, b_hi;      // valid (halo-extended) block range of pbx / xb / eb / X / Q for the pending resample
  int32_t* tile_start;
  double* rec;             // [ntiles_out][kMaxRec]
  uint64_t* scan_status;   // [nscan]
  double* scan_part;       // [nscan][kMaxRec]
  PlanDev* plan;
  CtlDev* ctl;
  const float* obs;        // ring [kObsRing][PZ_MAX_NY]
  const double* obs64;     // the same observations in double (f64 filters)
  pz_step_summary* summ;   // ring [kObsRing]
  int64_t ld;              // leading dimension of X (particles)
  int64_t n_local;         // particles on this rank (input and output)
  int64_t n_global;
  int64_t slot_base;       // global index of this rank's first slot / particle
  int64_t nblk;            // local canonical blocks
  int32_t ntiles_out;      // output tiles of the fused kernel
  int32_t nscan;           // CTAs of pf_scan_blocks
  uint64_t seed;
  double ll_const;
  int32_t resampler;
  int32_t world;           // > 1: the plan is finalised by pf_plan_sharded after the all-gather
  int32_t rank;
  uint64_t* shard_rec;     // sharded: [world][kShardRec] gathered {local_total, sw, sw2, sx[], sxx[], prefix at the margin edges}
  int64_t margin;          // sharded: halo margin in blocks exchanged with each neighbour every step
  const ShardP2P* p2p;     // sharded: peer-memory exchange (nullptr: NCCL collectives between the kernels)
  int32_t f64;             // 1: state storage and model arithmetic in double (descriptor flag state_f64)
  unsigned long long wait_ticks;  // sharded: clock64 ticks a peer-memory poll may take before the step is failed (0: no limit)
  LgDev model;
  LgDevD modeld;           // the same parameters unrounded (used when f64)
};
constexpr int kShardRec = 19;  // word 18: comm_error of the sending rank (so that every rank reports a timeout)
constexpr int kMaxRanks = 32;

// Peer-memory exchange of a sharded filter (one process per GPU, buffers mapped with CUDA IPC over
// NVLink): every rank owns a mailbox its peers store into, and maps the halo regions of its two
// neighbours.  Small messages (block exponent, records) travel as self-validating 64-bit words,
// (epoch << 32) | 32 payload bits, so neither side needs a fence or a flag; the halo margins are
// plain 128-bit peer stores + a system-scope fence + an epoch flag.  Every rank polls its OWN
// memory; no collective library call is on the path.
struct Mailbox {
  unsigned long long emax[kMaxRanks];                // (epoch << 32) | block exponent of rank q
  unsigned long long rec2[kMaxRanks][2 * kShardRec]; // rank q's record (pf_scan_blocks' last CTA): (epoch << 32) | low / high half of each word
  unsigned long long halo_flag[2];                   // epoch of the X / Q / eb margins in my left / right halo (written by that neighbour's pf_halo_push)
  unsigned long long pbx_flag[2];                    // epoch of the boundary-position (xb) margins in my left / right halo (that neighbour's pf_partition_sharded)
  unsigned long long push_done;                      // CTAs of pf_halo_push that finished their stores
  unsigned long long pbx_done;                       // CTAs of pf_partition_sharded that finished their margin stores
  unsigned long long part_done;                      // CTAs of pf_partition_sharded that finished (the last one advances the step counter)
  unsigned long long xb_done[kMaxRanks];             // epoch of rank q's complete boundary-position array (read by the remote step kernel)
};
struct ShardP2P {
  Mailbox* mbox[kMaxRanks];   // mbox[q]: rank q's mailbox (mbox[rank] is local memory)
  void* Xn[2][2];             // [side 0 = left neighbour, 1 = right][buffer]: its X (same halo layout as ours)
  uint16_t* Qn[2][2];
  int32_t* ebn[2][2];
  uint64_t* pbxn[2];
  uint64_t* xbn[2];           // the neighbours' boundary-position arrays (halo layout as ours)
  // every rank's arrays (own regions), for the step whose ancestors reach beyond the halo margins: the remote
  // step kernel and the global merge-path split read any block of any rank directly over NVLink
  void* Xall[kMaxRanks][2];
  uint16_t* Qall[kMaxRanks][2];
  int32_t* eball[kMaxRanks][2];
  uint64_t* xball[kMaxRanks];
  uint64_t* pball[kMaxRanks];  // local prefixes (final once the rank's record is gathered)
};  // 64-bit words per rank in the all-gathered record

// Host-callable launchers (pz_kernels.cu).  Each returns the number of kernels it enqueued.
int launch_init(const StepArgs& a, cudaStream_t s);
int launch_scan_partition(const StepArgs& a, int tile_out, cudaStream_t s);
int launch_partition_only(const StepArgs& a, int tile_out, cudaStream_t s);  // a.tile_start / a.ntiles_out for another tile size
int launch_step(const StepArgs& a, const int32_t* anc, cudaStream_t s);  // step tile kernel + scan + partition
int fused_tile_out(int nx, int f64);
// same as launch_step(a, nullptr, s) with an event recorded after each of the three kernels
int launch_step_profiled(const StepArgs& a, cudaStream_t s, cudaEvent_t ev[4]);
// sharded filter (one process per GPU): the step is split around the NCCL collectives
int launch_tile_only(const StepArgs& a, cudaStream_t s);                 // fused step kernel
int launch_scan_only(const StepArgs& a, int delta, cudaStream_t s);      // after the emax all-reduce
int launch_plan_sharded(const StepArgs& a, int delta, cudaStream_t s);   // after the all-gather: plan, absolute prefix, summary, margin check
// peer memory: record all-gather + plan + summary + margin check + boundary positions (own and pushed margins) + merge-path split, one kernel
int launch_partition_sharded(const StepArgs& a, int tile_out, int delta, cudaStream_t s);
int launch_partition_ext(const StepArgs& a, int tile_out, int delta, cudaStream_t s);  // after the halo exchange

// Generic (stored log-weight) path: standalone resampler, ancestor taps, visualiser read.
struct GenericArgs {
  const float* lw;       // [n] stored log-weights (pf_stats_from_lw input) or nullptr
  uint16_t* q16;         // [n] 16-bit weights: written by pf_stats_from_lw ...
  const uint16_t* q16_alt;  // ... or, with parity_ctl set, q16 / q16_alt = the filter's buffers 0 / 1
  uint64_t* xb;          // [nblk+1]
  int32_t* anc;          // [n_out]
  int32_t* eb;           // block exponents (standalone) ...
  int32_t* eb_alt;       // ... or, with parity_ctl set, eb / eb_alt = the filter's buffers 0 / 1
  const CtlDev* parity_ctl;
  uint64_t* sb;
  uint64_t* pb;
  int32_t* tile_start;
  uint64_t* scan_status;
  double* scan_part;
  PlanDev* plan;
  CtlDev* ctl;
  int64_t n;
  int64_t nblk;
  int32_t ntiles_out;
  int32_t nscan;
  uint32_t U;            // explicit uniform (standalone API)
  uint64_t seed;         // multinomial
  uint32_t step;
  int64_t n_draws;       // multinomial: number of iid draws (0: n)
  int32_t tile0;         // pf_ancestors: first output tile of this launch (a device of a multi-device tracker takes a range)
};
int launch_ctl_reset(CtlDev* ctl, cudaStream_t s);
int launch_generic_plan(const GenericArgs& g, bool stats_from_lw, cudaStream_t s);
int launch_generic_ancestors(const GenericArgs& g, cudaStream_t s);
int launch_generic_multinomial(const GenericArgs& g, cudaStream_t s);
int launch_logw(const StepArgs& a, float* lw_out32, double* lw_out64, cudaStream_t s);  // log-weights of the pending population (parity tap)
int launch_gather_thin(const StepArgs& a, const int32_t* anc, int32_t stride, int64_t n_sel, double* out,
                       cudaStream_t s);
constexpr int kGenericTileOut = 2048;
int launch_stats_from_lw(const GenericArgs& g, cudaStream_t s);

// multi-target tracker (pz_mtt.cu)
int mtt_words_per_particle();
size_t mtt_params_size();
void mtt_fill_params(const pz_model_desc* d, void* out_mttdev);
#ifdef PZ_DBG_STAMPS
void dbg_read_stamps(unsigned long long* out);  // debug build: %globaltimer stamps of the sharded exchange phases
#endif
int mtt_step_ctas(int64_t n);  // CTAs of mtt_step = rows of its per-CTA summary partials (5 doubles each)
// A multi-device tracker (one process, D devices): device r owns the particles and output slots [j0, j_end) of the
// global population; parents are read from their owners through peer pointers (sin_all, n_per particles per device).
struct MttShard {
  int64_t j0 = 0, j_end = 0;         // global range of this launch (j_end = n for a single device)
  int64_t n_per = 0;                 // particles per device (0: single device, parents in `sin`)
  const uint32_t* sin_all[8] = {};   // parent records of every device (current parity)
};
constexpr int kMttMaxDevices = 8;
int launch_mtt_step(const void* params, const uint32_t* sin, uint32_t* sout, const int32_t* anc, float* lw, const float* obs,
                    const CtlDev* ctl, unsigned long long* overflow, double* part, int64_t n, int32_t n_obs, uint64_t seed,
                    cudaStream_t s, const MttShard* shard = nullptr);
// n_ovf > 1: `overflow` holds one dropped-birth counter per device
int launch_mtt_summary(const double* part, int64_t n, PlanDev* plan, const CtlDev* ctl, pz_step_summary* summ, double ll_const,
                       const unsigned long long* overflow, cudaStream_t s, int n_ovf = 1);
// Slices of up to 6 arrays stored into the same offsets of up to 8 devices' arrays (128-bit peer stores over NVLink):
// the all-gather of a multi-device tracker's block statistics, 16-bit weights and summary partials.
struct PushArgs {
  int n_arr = 0, n_peers = 0;
  const void* src[6] = {};
  void* dst[8][6] = {};
  int64_t bytes[6] = {};   // multiples of 4; src / dst 16-byte aligned
};
int launch_push_slices(const PushArgs& p, cudaStream_t s);
int launch_emax_from_eb(const int32_t* eb, int64_t nblk, CtlDev* ctl, cudaStream_t s);
int launch_mtt_gather_multi(const MttShard& sh, const int32_t* anc, int32_t stride, int64_t n_sel, uint32_t* out, cudaStream_t s, int words);
// MultiCamera tracker (pz_mtt.cu): same calling convention as the MTT step
int mcam_words_per_particle();
size_t mcam_params_size();
void mcam_fill_params(const pz_model_desc* d, void* out_mcamdev);
int launch_mcam_step(const void* params, const uint32_t* sin, uint32_t* sout, const int32_t* anc, float* lw, const float* obs,
                     const CtlDev* ctl, unsigned long long* overflow, double* part, int64_t n, int32_t n_obs, uint64_t seed,
                     cudaStream_t s, const MttShard* shard = nullptr);
int launch_mtt_gather(const uint32_t* sdata, const int32_t* anc, int32_t stride, int64_t n_sel, uint32_t* out, cudaStream_t s, int words);

}  // namespace pz
