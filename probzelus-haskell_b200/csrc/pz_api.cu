// pz_api.cu -- the C ABI of include/pzgpu.h: filter handle, device memory, stream, CUDA
// graph and the launch sequence of one particle-filter step.
//
// Reference interface replaced (per call):
//   pz_filter_create   zparticles n model   (Inference.hs:107-108) / zdsparticles (113-114)
//   pz_filter_step     ZStream.step         (Util/ZStream.hs:16) applied to zparticles'
//   pz_filter_run      ZS.runStream         (Util/ZStream.hs:65-69)
//   pz_filter_read_particles   the `[b]` output list of zparticles' (Inference.hs:104)
//
// There is no CPU fallback: every entry point that computes returns PZ_ECUDA when no CUDA
// device is usable.
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>  // types only: the library is loaded with dlopen when a sharded filter is created

#include "pz_generic.cuh"
#include "pz_kernels.cuh"

using namespace pz;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define PZ_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) return fail(PZ_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// Position/velocity ("Kronecker") structure of a derived parameter set: F = [[a,b],[c,d]] (x) I_p,
// Q = diag(qp I, qv I), H = [I_p 0], R = r I_p; and the scalar random walk with unit F and H (multiplying by
// exactly 1 is exact, so the specialised path skips it).
template <class T>
void detect_structure(LgDevT<T>* p) {
  const int nx = p->nx, ny = p->ny;
  p->kron = 0;
  if (getenv("PZ_NO_KRON")) return;
  if (nx == 1 && ny == 1 && p->F[0][0] == (T)1 && p->H[0][0] == (T)1) {
    p->kron = 1; p->klp = p->Lq[0][0]; p->krinv = p->Rinv[0][0];
  }
  if (nx % 2 == 0 && ny == nx / 2) {
    const int P = nx / 2;
    bool ok = true;
    const T a = p->F[0][0], b = p->F[0][P], c = p->F[P][0], dd = p->F[P][P];
    const T lp = p->Lq[0][0], lv = p->Lq[P][P], ri = p->Rinv[0][0];
    for (int i = 0; i < nx && ok; ++i)
      for (int j = 0; j < nx; ++j) {
        const int bi = i / P, bj = j / P;
        const T fexp = (i % P == j % P) ? (bi == 0 ? (bj == 0 ? a : b) : (bj == 0 ? c : dd)) : (T)0;
        const T lexp = (i == j) ? (bi == 0 ? lp : lv) : (T)0;
        if (p->F[i][j] != fexp || p->Lq[i][j] != lexp) { ok = false; break; }
      }
    for (int i = 0; i < ny && ok; ++i) {
      for (int j = 0; j < nx; ++j) if (p->H[i][j] != ((i == j) ? (T)1 : (T)0)) ok = false;
      for (int j = 0; j < ny; ++j) if (p->Rinv[i][j] != ((i == j) ? ri : (T)0)) ok = false;
    }
    if (ok) { p->kron = 1; p->ka = a; p->kb = b; p->kc = c; p->kd = dd; p->klp = lp; p->klv = lv; p->krinv = ri; }
  }
}

// Derived model parameters: Cholesky of Q, inverse and log-determinant of R in double; the fp32 set is
// rounded once (DESIGN.md "Model descriptor"), the f64 set is kept as computed.
int derive_lg(const pz_model_desc* d, LgDev* p, LgDevD* pd, double* ll_const) {
  const int nx = d->nx, ny = d->ny;
  if (nx < 1 || nx > PZ_MAX_NX || ny < 1 || ny > PZ_MAX_NY) return fail(PZ_EINVAL, "unsupported dims nx=%d ny=%d", nx, ny);
  memset(p, 0, sizeof(*p));
  memset(pd, 0, sizeof(*pd));
  p->nx = nx; p->ny = ny;
  pd->nx = nx; pd->ny = ny;
  double L[PZ_MAX_NX][PZ_MAX_NX] = {};
  for (int j = 0; j < nx; ++j) {
    double s = d->Q[j * nx + j];
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    if (s < 0) return fail(PZ_EINVAL, "Q is not positive semi-definite");
    const double ljj = s > 0 ? sqrt(s) : 0.0;
    L[j][j] = ljj;
    for (int i = j + 1; i < nx; ++i) {
      double t = d->Q[i * nx + j];
      for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
      L[i][j] = ljj > 0 ? t / ljj : 0.0;
    }
  }
  double A[PZ_MAX_NY][2 * PZ_MAX_NY] = {};
  for (int i = 0; i < ny; ++i) {
    for (int j = 0; j < ny; ++j) A[i][j] = d->R[i * ny + j];
    A[i][ny + i] = 1.0;
  }
  double logdet = 0;
  for (int c = 0; c < ny; ++c) {
    const double piv = A[c][c];
    if (!(piv > 0)) return fail(PZ_EINVAL, "R is not positive definite");
    logdet += log(piv);
    for (int j = 0; j < 2 * ny; ++j) A[c][j] /= piv;
    for (int r = 0; r < ny; ++r) {
      if (r == c) continue;
      const double f = A[r][c];
      for (int j = 0; j < 2 * ny; ++j) A[r][j] -= f * A[c][j];
    }
  }
  for (int i = 0; i < nx; ++i) {
    for (int j = 0; j < nx; ++j) {
      p->F[i][j] = (float)d->F[i * nx + j]; p->Lq[i][j] = (float)L[i][j];
      pd->F[i][j] = d->F[i * nx + j]; pd->Lq[i][j] = L[i][j];
    }
    p->x0[i] = (float)d->x0[i];
    pd->x0[i] = d->x0[i];
  }
  for (int i = 0; i < ny; ++i) {
    for (int j = 0; j < nx; ++j) { p->H[i][j] = (float)d->H[i * nx + j]; pd->H[i][j] = d->H[i * nx + j]; }
    for (int j = 0; j < ny; ++j) { p->Rinv[i][j] = (float)A[i][ny + j]; pd->Rinv[i][j] = A[i][ny + j]; }
  }
  const bool dbl = (ny == 1 && d->scalar_ll_2x);
  p->wscale = dbl ? -1.0f : -0.5f;
  pd->wscale = dbl ? -1.0 : -0.5;
  detect_structure(p);
  detect_structure(pd);
  *ll_const = (dbl ? -1.0 : -0.5) * (logdet + ny * log(2 * M_PI));
  return PZ_OK;
}

// scratch of the generic (stored log-weight) path
struct GenericScratch {
  float* lw = nullptr;
  double* lw64 = nullptr;       // f64 filters: exact log-weight tap
  double* thin = nullptr;       // read_particles scratch (grown on demand)
  int64_t thin_cap = 0;
  uint16_t* q16 = nullptr;
  uint64_t* xb = nullptr;
  int32_t* anc = nullptr;
  int32_t* eb = nullptr;
  uint64_t* sb = nullptr;
  uint64_t* pb = nullptr;
  int32_t* tile_start = nullptr;
  uint64_t* scan_status = nullptr;
  double* scan_part = nullptr;
  PlanDev* plan = nullptr;
  CtlDev* ctl = nullptr;
  int64_t cap = 0;  // particles
  void release() {
    cudaFree(lw); cudaFree(lw64); cudaFree(thin); cudaFree(q16); cudaFree(xb); cudaFree(anc); cudaFree(eb); cudaFree(sb); cudaFree(pb); cudaFree(tile_start);
    cudaFree(scan_status); cudaFree(scan_part); cudaFree(plan); cudaFree(ctl);
    *this = GenericScratch();
  }
};

int generic_alloc(GenericScratch* g, int64_t n, bool full, cudaStream_t st) {
  if (g->cap >= n && (!full || g->eb)) return PZ_OK;
  g->release();
  const int64_t ld = round_up(n, kPad), nblk = (n + kBlk - 1) / kBlk;
  const int64_t nscan = (nblk + kScanTile - 1) / kScanTile, ntiles = (n + kGenericTileOut - 1) / kGenericTileOut;
  PZ_CUDA(cudaMalloc(&g->lw, ld * sizeof(float)));
  PZ_CUDA(cudaMemsetAsync(g->lw, 0, ld * sizeof(float), st));
  PZ_CUDA(cudaMalloc(&g->q16, ld * sizeof(uint16_t)));
  PZ_CUDA(cudaMemsetAsync(g->q16, 0, ld * sizeof(uint16_t), st));
  PZ_CUDA(cudaMalloc(&g->xb, (nblk + 2) * sizeof(uint64_t)));
  PZ_CUDA(cudaMalloc(&g->anc, ld * sizeof(int32_t)));
  PZ_CUDA(cudaMalloc(&g->tile_start, (ntiles + 1) * sizeof(int32_t)));
  if (full) {
    PZ_CUDA(cudaMalloc(&g->eb, (nblk + 1) * sizeof(int32_t)));
    PZ_CUDA(cudaMalloc(&g->sb, (nblk + 1) * sizeof(uint64_t)));
    PZ_CUDA(cudaMalloc(&g->pb, (nblk + 2) * sizeof(uint64_t)));
    PZ_CUDA(cudaMalloc(&g->scan_status, (nscan + 1) * sizeof(uint64_t)));
    PZ_CUDA(cudaMemsetAsync(g->scan_status, 0, (nscan + 1) * sizeof(uint64_t), st));
    PZ_CUDA(cudaMalloc(&g->scan_part, (nscan + 1) * kMaxRec * sizeof(double)));
    PZ_CUDA(cudaMalloc(&g->plan, sizeof(PlanDev)));
    PZ_CUDA(cudaMemsetAsync(g->plan, 0, sizeof(PlanDev), st));
    PZ_CUDA(cudaMalloc(&g->ctl, sizeof(CtlDev)));
    PZ_CUDA(cudaMemsetAsync(g->ctl, 0, sizeof(CtlDev), st));
  }
  g->cap = n;
  return PZ_OK;
}

}  // namespace

// Delayed-sampling (Rao-Blackwellised) linear-Gaussian filter: the one Gaussian marginal all particles share.
struct RbDev {
  double mean[PZ_MAX_NX];
  double cov[PZ_MAX_NX * PZ_MAX_NX];
  double F[PZ_MAX_NX * PZ_MAX_NX], Q[PZ_MAX_NX * PZ_MAX_NX], H[PZ_MAX_NY * PZ_MAX_NX], R[PZ_MAX_NY * PZ_MAX_NY];
  double y[PZ_MAX_NY];
  double ll_scale;   // 2 for the reference's doubled scalar score (Distributions.hs:160-161), else 1
  int32_t nx, ny;
  uint32_t t;
};

// One step of the delayed-sampling stream: makeMarginal = kalmanPredict (MVDistributions.hs:75-76,
// DelayedSampling.hs:135-139), score' of the marginal predictive (DelayedSampling.hs:124-128), makeConditional
// = kalmanUpdate (MVDistributions.hs:83-100, DelayedSampling.hs:143-152).  Every particle is identical, so the
// whole population is this one recursion in double on one thread.
__global__ void rb_kalman_step(RbDev* st, pz_step_summary* out, double n_particles) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int nx = st->nx, ny = st->ny;
  double m[PZ_MAX_NX], P[PZ_MAX_NX][PZ_MAX_NX], FP[PZ_MAX_NX][PZ_MAX_NX];
  for (int i = 0; i < nx; ++i) {
    double s = 0;
    for (int k = 0; k < nx; ++k) s += st->F[i * nx + k] * st->mean[k];
    m[i] = s;
    for (int j = 0; j < nx; ++j) {
      double a = 0;
      for (int k = 0; k < nx; ++k) a += st->F[i * nx + k] * st->cov[k * nx + j];
      FP[i][j] = a;
    }
  }
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < nx; ++j) {
      double a = 0;
      for (int k = 0; k < nx; ++k) a += FP[i][k] * st->F[j * nx + k];
      P[i][j] = a + st->Q[i * nx + j];
    }
  // S = R + H P H^T, innovation, Gauss-Jordan inverse (ny <= 3, SPD)
  double PHt[PZ_MAX_NX][PZ_MAX_NY], S[PZ_MAX_NY][2 * PZ_MAX_NY], v[PZ_MAX_NY];
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      double a = 0;
      for (int k = 0; k < nx; ++k) a += P[i][k] * st->H[j * nx + k];
      PHt[i][j] = a;
    }
  for (int i = 0; i < ny; ++i) {
    for (int j = 0; j < ny; ++j) {
      double a = 0;
      for (int k = 0; k < nx; ++k) a += st->H[i * nx + k] * PHt[k][j];
      S[i][j] = a + st->R[i * ny + j];
      S[i][ny + j] = (i == j) ? 1.0 : 0.0;
    }
    double a = 0;
    for (int k = 0; k < nx; ++k) a += st->H[i * nx + k] * m[k];
    v[i] = st->y[i] - a;
  }
  double logdet = 0;
  for (int c = 0; c < ny; ++c) {
    const double piv = S[c][c];
    logdet += log(piv);
    for (int j = 0; j < 2 * ny; ++j) S[c][j] /= piv;
    for (int r = 0; r < ny; ++r) {
      if (r == c) continue;
      const double f = S[r][c];
      for (int j = 0; j < 2 * ny; ++j) S[r][j] -= f * S[c][j];
    }
  }
  double quad = 0, Siv[PZ_MAX_NY];
  for (int i = 0; i < ny; ++i) {
    double a = 0;
    for (int j = 0; j < ny; ++j) a += S[i][ny + j] * v[j];
    Siv[i] = a;
    quad += v[i] * a;
  }
  const double ll = -0.5 * (quad + logdet + ny * 1.8378770664093453);
  for (int i = 0; i < nx; ++i) {
    double a = 0;
    for (int j = 0; j < ny; ++j) a += PHt[i][j] * Siv[j];
    st->mean[i] = m[i] + a;
  }
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < nx; ++j) {
      double a = 0;
      for (int k = 0; k < ny; ++k)
        for (int l = 0; l < ny; ++l) a += PHt[i][k] * S[k][ny + l] * PHt[j][l];
      st->cov[i * nx + j] = P[i][j] - a;
    }
  out->step = st->t;
  out->log_evidence_inc = st->ll_scale * ll;
  out->ess = n_particles;          // equal weights
  for (int d = 0; d < PZ_MAX_NX; ++d) {
    out->mean[d] = d < nx ? st->mean[d] : 0.0;
    out->var[d] = d < nx ? st->cov[d * nx + d] : 0.0;
  }
  out->n_migrated = 0;
  out->e_max = 0;
  out->reserved = 0;
  out->total_weight = (uint64_t)n_particles << 15;  // every particle at q = 2^15
  st->t += 1;
}

struct pz_filter {
  pz_model_desc desc;
  StepArgs a;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t push_stream = nullptr;                 // (unused since the margin push moved into the scan kernel)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaGraphExec_t graph = nullptr;
  cudaGraphExec_t sgraph[2] = {nullptr, nullptr};  // sharded: one captured step (kernels + NCCL) per buffer parity
  // PZ_SHARD_TIMING=1 (implies no graph): events between the kernels of a sharded step, averages printed at destroy
  bool shard_timing = false;
  cudaEvent_t tev[5] = {};
  double tacc[4] = {0, 0, 0, 0};
  int64_t tcount = 0;
  bool tpending = false;
  float* h_obs = nullptr;            // pinned [kObsRing][PZ_MAX_NY]
  pz_step_summary* h_summ = nullptr; // pinned [kObsRing]
  float* d_obs = nullptr;
  uint32_t t = 0;                    // host mirror of ctl->t
  double last_ms = 0;
  int64_t last_launches = 0;
  GenericScratch g;                  // taps / multinomial path
  int launches_per_step = 0;
  // sharded filter (one process per GPU)
  ncclComm_t comm = nullptr;
  int64_t hb = 0;                    // halo capacity in blocks on each side
  void* xbuf[2] = {nullptr, nullptr};   // float or double [nx][ld]
  double* d_obs64 = nullptr;
  double* h_obs64 = nullptr;            // pinned [kObsRing][PZ_MAX_NY]
  uint16_t* qbuf[2] = {nullptr, nullptr};
  int32_t* ebbuf[2] = {nullptr, nullptr};
  uint64_t* pbxbuf = nullptr;
  uint64_t* xbbuf = nullptr;
  int32_t* warm = nullptr;           // sharded: scratch of the NCCL peer warm-up
  Mailbox* mbox = nullptr;           // sharded: this rank's mailbox (peer-memory exchange)
  ShardP2P* d_p2p = nullptr;
  std::vector<void*> ipc_opened;     // peer mappings to close
  // delayed-sampling linear-Gaussian filter
  bool is_rb = false;
  RbDev* rb = nullptr;
  bool is_gen = false;               // generic-path filter (stored log-weights): carried weights, adaptive resampling, Walker
  GenArgs ga;
  float* gen_lw[2] = {nullptr, nullptr};
  double* gen_obs = nullptr;         // device [PZ_MAX_NY]
  double* h_gen_obs = nullptr;       // pinned
  double* gen_part = nullptr;
  GenState* gen_state = nullptr;
  bool is_beta = false;              // Beta-Bernoulli under delayed sampling
  BetaDev* beta = nullptr;
  bool is_multi = false;             // single-process multi-device filter: the shards (one per device) do the work
  std::vector<pz_filter*> shards;
  // multi-target tracker
  bool is_mtt = false;
  uint32_t* mtt_s[2] = {nullptr, nullptr};   // particle records, double-buffered
  float* mtt_obs = nullptr;                  // device [m_max][2]
  float* h_mtt_obs = nullptr;                // pinned
  unsigned long long* mtt_overflow = nullptr;
  double* mtt_part = nullptr;                // per-CTA summary partials of mtt_step
  std::vector<char> mtt_params;
  int32_t mtt_mmax = 0;
  bool is_mcam = false;                      // the MultiCamera tracker: same host path, its own record and step kernel
  int tracker_words = 0;                     // 32-bit words per particle record
  int tracker_obs_w = 2;                     // doubles per observation row (2: MTT detections; 16: MultiCamera rows)
  uint32_t* tracker_thin = nullptr;          // read_tracks scratch: the thinned records (grown on demand, kept across calls)
  int64_t tracker_thin_cap = 0;              // in 32-bit words
  // multi-device tracker (a shard of a pz_filter_create_multi tracker): this device owns particles [mtt_j0, mtt_jend) of
  // the global population; the gathered arrays (16-bit weights, block exponents / sums, summary partials, dropped-birth
  // counters) exist once per step parity, so that a peer's push of step t+1 never lands in what step t still reads
  int mtt_world = 1, mtt_rank = 0;
  int64_t mtt_j0 = 0, mtt_jend = 0, mtt_per = 0;
  uint16_t* mtt_q16b = nullptr;
  uint64_t* mtt_sbb = nullptr;
  double* mtt_partb = nullptr;
  unsigned long long* mtt_ovf[2] = {nullptr, nullptr};   // [kMttMaxDevices] counters each
};

namespace {

// ---- NCCL, loaded on demand (a single-GPU host never needs libnccl) ----------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  static std::mutex mu;  // two threads may create their first sharded filters at the same time
  std::lock_guard<std::mutex> lock(mu);
  if (g_nccl.handle) return PZ_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(PZ_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
#define PZ_SYM(field, name)                                                      \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                     \
  if (!g_nccl.field) return fail(PZ_ENCCL, "libnccl lacks %s", name)
  PZ_SYM(GetUniqueId, "ncclGetUniqueId");
  PZ_SYM(CommInitRank, "ncclCommInitRank");
  PZ_SYM(CommDestroy, "ncclCommDestroy");
  PZ_SYM(AllReduce, "ncclAllReduce");
  PZ_SYM(AllGather, "ncclAllGather");
  PZ_SYM(Send, "ncclSend");
  PZ_SYM(Recv, "ncclRecv");
  PZ_SYM(GroupStart, "ncclGroupStart");
  PZ_SYM(GroupEnd, "ncclGroupEnd");
  PZ_SYM(GetErrorString, "ncclGetErrorString");
#undef PZ_SYM
  g_nccl.handle = h;
  return PZ_OK;
}

#define PZ_NCCL(call)                                                                         \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess) return fail(PZ_ENCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
  } while (0)

void shard_timing_collect(pz_filter* f) {
  if (!f->tpending) return;
  f->tpending = false;
  if (cudaEventSynchronize(f->tev[4]) != cudaSuccess) { cudaGetLastError(); return; }
  for (int i = 0; i < 4; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, f->tev[i], f->tev[i + 1]) == cudaSuccess) f->tacc[i] += ms;
  }
  f->tcount += 1;
}

void free_filter(pz_filter* f) {
  if (!f) return;
  if (f->is_multi) {
    for (pz_filter* sh : f->shards) if (sh && sh->stream) { cudaSetDevice(sh->device); cudaStreamSynchronize(sh->stream); }
    for (pz_filter* sh : f->shards) free_filter(sh);
    delete f;
    return;
  }
  cudaSetDevice(f->device);
  if (f->stream) cudaStreamSynchronize(f->stream);
  if (f->graph) cudaGraphExecDestroy(f->graph);
  for (int i = 0; i < 2; ++i) if (f->sgraph[i]) cudaGraphExecDestroy(f->sgraph[i]);
#ifdef PZ_DBG_STAMPS
  if (f->a.world > 1) {
    unsigned long long st[16] = {};
    cudaStreamSynchronize(f->stream);
    dbg_read_stamps(st);
    char line[512];
    int o = snprintf(line, sizeof(line), "[pz stamps] rank %d (ns after scan start):", f->a.rank);
    for (int i = 0; i < 15; ++i) o += snprintf(line + o, sizeof(line) - o, " [%d]=%lld", i, (long long)(st[i] - st[0]));
    fprintf(stderr, "%s\n", line);
  }
#endif
  if (f->shard_timing) {
    shard_timing_collect(f);
    if (f->tcount)
      fprintf(stderr, "[pz shard timing] rank %d of %d, %lld steps: step kernels %.1f us, scan + margin push %.1f us, (unused) %.1f us, plan + partition %.1f us\n",
              f->a.rank, f->a.world, (long long)f->tcount, 1e3 * f->tacc[0] / f->tcount, 1e3 * f->tacc[1] / f->tcount,
              1e3 * f->tacc[2] / f->tcount, 1e3 * f->tacc[3] / f->tcount);
    for (auto& e : f->tev) if (e) cudaEventDestroy(e);
  }
  if (f->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(f->comm);
  StepArgs& a = f->a;
  cudaFree(f->xbuf[0]); cudaFree(f->xbuf[1]); cudaFree(f->qbuf[0]); cudaFree(f->qbuf[1]); cudaFree(f->xbbuf); cudaFree(f->ebbuf[0]); cudaFree(f->ebbuf[1]); cudaFree(a.sb); cudaFree(a.pb);
  cudaFree(f->pbxbuf); cudaFree(a.tile_start); cudaFree(a.rec); cudaFree(a.scan_status); cudaFree(a.scan_part);
  cudaFree(a.plan); cudaFree(a.ctl); cudaFree(f->d_obs); cudaFree(f->d_obs64); cudaFree(a.summ); cudaFree(a.shard_rec);
  cudaFree(f->rb);
  cudaFree(f->warm);
  for (void* p : f->ipc_opened) cudaIpcCloseMemHandle(p);
  cudaFree(f->mbox); cudaFree(f->d_p2p);
  cudaFree(f->mtt_s[0]); cudaFree(f->mtt_s[1]); cudaFree(f->mtt_obs); cudaFree(f->mtt_overflow); cudaFree(f->mtt_part);
  cudaFree(f->tracker_thin);
  cudaFree(f->mtt_q16b); cudaFree(f->mtt_sbb); cudaFree(f->mtt_partb); cudaFree(f->mtt_ovf[0]); cudaFree(f->mtt_ovf[1]);
  cudaFree(f->gen_lw[0]); cudaFree(f->gen_lw[1]); cudaFree(f->gen_obs); cudaFree(f->gen_part); cudaFree(f->gen_state); cudaFree(f->beta);
  if (f->h_gen_obs) cudaFreeHost(f->h_gen_obs);
  if (f->h_mtt_obs) cudaFreeHost(f->h_mtt_obs);
  f->g.release();
  if (f->h_obs) cudaFreeHost(f->h_obs);
  if (f->h_obs64) cudaFreeHost(f->h_obs64);
  if (f->h_summ) cudaFreeHost(f->h_summ);
  if (f->ev0) cudaEventDestroy(f->ev0);
  if (f->ev1) cudaEventDestroy(f->ev1);
  if (f->push_stream) cudaStreamDestroy(f->push_stream);
  if (f->ev_fork) cudaEventDestroy(f->ev_fork);
  if (f->ev_join) cudaEventDestroy(f->ev_join);
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
}

int stage_obs(pz_filter* f, const double* obs, int64_t k, int32_t obs_dim) {
  // rows for steps [t, t+k) -> ring rows (t+i) % kObsRing
  const int64_t r0 = f->t % kObsRing;
  const int64_t first = (r0 + k <= kObsRing) ? k : kObsRing - r0;
  if (f->a.f64) {  // a double filter reads the observations unrounded
    for (int64_t i = 0; i < k; ++i) {
      double* row = f->h_obs64 + ((f->t + i) % kObsRing) * PZ_MAX_NY;
      for (int c = 0; c < PZ_MAX_NY; ++c) row[c] = c < obs_dim ? obs[i * obs_dim + c] : 0.0;
    }
    const size_t rowb = PZ_MAX_NY * sizeof(double);
    PZ_CUDA(cudaMemcpyAsync(f->d_obs64 + r0 * PZ_MAX_NY, f->h_obs64 + r0 * PZ_MAX_NY, first * rowb, cudaMemcpyHostToDevice, f->stream));
    if (first < k)
      PZ_CUDA(cudaMemcpyAsync(f->d_obs64, f->h_obs64, (k - first) * rowb, cudaMemcpyHostToDevice, f->stream));
    return PZ_OK;
  }
  for (int64_t i = 0; i < k; ++i) {
    float* row = f->h_obs + ((f->t + i) % kObsRing) * PZ_MAX_NY;
    for (int c = 0; c < PZ_MAX_NY; ++c) row[c] = c < obs_dim ? (float)obs[i * obs_dim + c] : 0.0f;
  }
  const size_t rowb = PZ_MAX_NY * sizeof(float);
  PZ_CUDA(cudaMemcpyAsync(f->d_obs + r0 * PZ_MAX_NY, f->h_obs + r0 * PZ_MAX_NY, first * rowb, cudaMemcpyHostToDevice, f->stream));
  if (first < k)
    PZ_CUDA(cudaMemcpyAsync(f->d_obs, f->h_obs, (k - first) * rowb, cudaMemcpyHostToDevice, f->stream));
  return PZ_OK;
}

int fetch_summaries(pz_filter* f, uint32_t t0, int64_t k) {
  const int64_t r0 = t0 % kObsRing;
  const int64_t first = (r0 + k <= kObsRing) ? k : kObsRing - r0;
  PZ_CUDA(cudaMemcpyAsync(f->h_summ + r0, f->a.summ + r0, first * sizeof(pz_step_summary), cudaMemcpyDeviceToHost, f->stream));
  if (first < k)
    PZ_CUDA(cudaMemcpyAsync(f->h_summ, f->a.summ, (k - first) * sizeof(pz_step_summary), cudaMemcpyDeviceToHost, f->stream));
  return PZ_OK;
}

// ancestors of the pending resample into f->g.anc (systematic: canonical; multinomial: step t)
int sharded_advance(pz_filter* f, int delta, int* launches);

int pending_ancestors(pz_filter* f, int* launches) {
  StepArgs& a = f->a;
  if (a.world > 1) return fail(PZ_ESTATE, "ancestor / particle taps are not available on a sharded filter");
  int rc = generic_alloc(&f->g, a.n_local, false, f->stream);
  if (rc) return rc;
  int n = 0;
  GenericArgs g{};
  g.lw = f->g.lw; g.anc = f->g.anc; g.eb = a.eb[0]; g.eb_alt = a.eb[1]; g.parity_ctl = a.ctl;
  if (f->is_mtt || f->is_gen) {   // stored log-weights (and their q16); one eb buffer
    g.q16 = f->g.q16; g.q16_alt = f->g.q16; g.eb = a.eb[0]; g.eb_alt = a.eb[0]; g.parity_ctl = nullptr;
  } else { g.q16 = a.Q[0]; g.q16_alt = a.Q[1]; }
  g.xb = f->g.xb;
  g.sb = a.sb; g.pb = a.pb; g.tile_start = f->g.tile_start; g.scan_status = a.scan_status; g.scan_part = a.scan_part;
  g.plan = a.plan; g.ctl = a.ctl; g.n = a.n_local; g.nblk = a.nblk;
  g.ntiles_out = (int32_t)((a.n_local + kGenericTileOut - 1) / kGenericTileOut);
  g.nscan = a.nscan; g.U = 0; g.seed = a.seed; g.step = f->t;
  if (a.resampler == PZ_RESAMPLE_MULTINOMIAL_REF) {
    n += launch_generic_multinomial(g, f->stream);
  } else {
    // partition for the generic tile size, then the ancestor kernel
    StepArgs b = a;
    b.tile_start = f->g.tile_start;
    b.ntiles_out = g.ntiles_out;
    b.xb = f->g.xb;  // same values as a.xb; kept separate so that the filter's own plan is untouched
    n += launch_partition_only(b, kGenericTileOut, f->stream);
    n += launch_generic_ancestors(g, f->stream);
  }
  if (f->is_gen && f->ga.ess_threshold > 0) n += launch_gen_anc_identity(f->ga, f->stream);  // no resample decided: the population itself
  if (launches) *launches += n;
  PZ_CUDA(cudaGetLastError());
  return PZ_OK;
}

// ---- sharded filter: everything between two step kernels ---------------------------------------
// emax all-reduce -> scan -> all-gather -> plan (+ margin check) -> fixed-margin halo exchange with the
// two neighbours -> partition.  No host synchronisation: the sequence is stream-ordered (and captured
// in a CUDA graph per buffer parity).  `delta` = 1 after a step kernel, 0 for the initial population.
// A rank whose ancestors reach beyond the margins is reported by every rank in the step summary
// (n_migrated = -1 -> PZ_ECAPACITY); migration volume is the slots whose ancestor lives on a neighbour.
int sharded_advance(pz_filter* f, int delta, int* launches) {
  StepArgs& a = f->a;
  const int R = a.world, me = a.rank;
  const int64_t nb = a.nblk, m = f->hb;
  cudaStream_t s = f->stream;
  const int which = (int)((f->t + (uint32_t)delta) & 1u);  // buffer of the population just produced
  if (a.p2p) {  // peer-memory exchange inside the kernels: no collective call at all
    const bool tm = f->shard_timing && delta;
    // the X / Q / eb margins travel to the neighbours while this rank scans: extra CTAs of the scan kernel push them
    *launches += launch_scan_only(a, delta, s);
    if (tm) { cudaEventRecord(f->tev[2], s); cudaEventRecord(f->tev[3], s); }
    *launches += launch_partition_sharded(a, fused_tile_out(a.model.nx, a.f64), delta, s);
    if (tm) { cudaEventRecord(f->tev[4], s); f->tpending = true; }
    PZ_CUDA(cudaGetLastError());
    return PZ_OK;
  }
  PZ_NCCL(g_nccl.AllReduce(&a.ctl->emax_acc, &a.ctl->emax_acc, 1, ncclInt32, ncclMax, f->comm, s));
  *launches += launch_scan_only(a, delta, s);
  PZ_NCCL(g_nccl.AllGather(a.shard_rec + (size_t)me * kShardRec, a.shard_rec, kShardRec, ncclUint64, f->comm, s));
  *launches += launch_plan_sharded(a, delta, s);
  char* x = reinterpret_cast<char*>(a.X[which]);
  const size_t es = a.f64 ? 8 : 4;
  uint16_t* q = a.Q[which];
  int32_t* eb = a.eb[which];
  uint64_t* pbx = const_cast<uint64_t*>(a.pbx);
  PZ_NCCL(g_nccl.GroupStart());
  if (me > 0) {  // my first m blocks -> the right halo of rank me-1; its last m blocks -> my left halo
    for (int d = 0; d < a.model.nx; ++d) {
      PZ_NCCL(g_nccl.Send(x + (size_t)d * a.ld * es, m * kBlk * es, ncclUint8, me - 1, f->comm, s));
      PZ_NCCL(g_nccl.Recv(x + ((size_t)d * a.ld - m * kBlk) * es, m * kBlk * es, ncclUint8, me - 1, f->comm, s));
    }
    PZ_NCCL(g_nccl.Send(q, m * kBlk * 2, ncclUint8, me - 1, f->comm, s));
    PZ_NCCL(g_nccl.Recv(q - m * kBlk, m * kBlk * 2, ncclUint8, me - 1, f->comm, s));
    PZ_NCCL(g_nccl.Send(eb, m, ncclInt32, me - 1, f->comm, s));
    PZ_NCCL(g_nccl.Recv(eb - m, m, ncclInt32, me - 1, f->comm, s));
    PZ_NCCL(g_nccl.Send(pbx + 1, m, ncclUint64, me - 1, f->comm, s));       // pbx[1 .. m]; pbx[0] equals its pbx[nblk]
    PZ_NCCL(g_nccl.Recv(pbx - m, m, ncclUint64, me - 1, f->comm, s));       // its pbx[nblk-m .. nblk-1]
  }
  if (me + 1 < R) {
    for (int d = 0; d < a.model.nx; ++d) {
      PZ_NCCL(g_nccl.Send(x + ((size_t)d * a.ld + (nb - m) * kBlk) * es, m * kBlk * es, ncclUint8, me + 1, f->comm, s));
      PZ_NCCL(g_nccl.Recv(x + ((size_t)d * a.ld + nb * kBlk) * es, m * kBlk * es, ncclUint8, me + 1, f->comm, s));
    }
    PZ_NCCL(g_nccl.Send(q + (nb - m) * kBlk, m * kBlk * 2, ncclUint8, me + 1, f->comm, s));
    PZ_NCCL(g_nccl.Recv(q + nb * kBlk, m * kBlk * 2, ncclUint8, me + 1, f->comm, s));
    PZ_NCCL(g_nccl.Send(eb + (nb - m), m, ncclInt32, me + 1, f->comm, s));
    PZ_NCCL(g_nccl.Recv(eb + nb, m, ncclInt32, me + 1, f->comm, s));
    PZ_NCCL(g_nccl.Send(pbx + (nb - m), m, ncclUint64, me + 1, f->comm, s));
    PZ_NCCL(g_nccl.Recv(pbx + nb + 1, m, ncclUint64, me + 1, f->comm, s));
  }
  PZ_NCCL(g_nccl.GroupEnd());
  *launches += launch_partition_ext(a, fused_tile_out(a.model.nx, a.f64), delta, s);
  PZ_CUDA(cudaGetLastError());
  return PZ_OK;
}

// the error a step's summary implies (the same verdict on every rank)
int report_step(pz_filter* f, const pz_step_summary& s, uint32_t t) {
  if (s.total_weight == 0) return fail(PZ_EDEGENERATE, "step %u: every particle has zero weight (all -inf or NaN log-weights)", t);
  if (s.n_migrated == -2) {
    CtlDev c{};
    cudaMemcpy(&c, f->a.ctl, sizeof(c), cudaMemcpyDeviceToHost);
    cudaGetLastError();
    return fail(PZ_ENCCL, "step %u: a peer rank never answered the peer-memory exchange (rank %d, wait mask 0x%x: 1 exponent, 2 records, "
                "4/0x40 state margins, 8/0x10/0x80 boundary positions; 0 = reported by a peer)", t, f->a.rank, c.comm_error);
  }
  if (s.n_migrated < 0)
    return fail(PZ_ECAPACITY, "step %u: a rank's ancestors reach beyond the halo margin of %lld blocks (NCCL exchange path: set PZ_HALO_BLOCKS; "
                "the peer-memory path reads remote blocks instead)", t, (long long)f->hb);
  return PZ_OK;
}

int enqueue_step(pz_filter* f) {
  StepArgs& a = f->a;
  if (a.world > 1) {
    if (f->sgraph[f->t & 1u]) {
      PZ_CUDA(cudaGraphLaunch(f->sgraph[f->t & 1u], f->stream));
      f->last_launches += f->launches_per_step;
    } else {
      if (f->shard_timing) { shard_timing_collect(f); cudaEventRecord(f->tev[0], f->stream); }
      int n = launch_tile_only(a, f->stream);
      if (f->shard_timing) cudaEventRecord(f->tev[1], f->stream);
      int rc = sharded_advance(f, 1, &n);
      f->last_launches += n;
      if (rc) return rc;
    }
  } else if (a.resampler == PZ_RESAMPLE_MULTINOMIAL_REF) {
    int n = 0;
    int rc = pending_ancestors(f, &n);
    if (rc) return rc;
    n += launch_step(a, f->g.anc, f->stream);
    f->last_launches += n;
  } else if (f->graph) {
    PZ_CUDA(cudaGraphLaunch(f->graph, f->stream));
    f->last_launches += f->launches_per_step;
  } else {
    int n = launch_step(a, nullptr, f->stream);
    if (n < 0) return fail(PZ_EINVAL, "no kernel instantiated for nx=%d ny=%d", a.model.nx, a.model.ny);
    f->last_launches += n;
  }
  f->t += 1;
  return PZ_OK;
}

// the gathered arrays of step parity `par` (a single-device tracker has one set)
struct MttBufs { uint16_t* q16; int32_t* eb; uint64_t* sb; double* part; unsigned long long* ovf; };
MttBufs mtt_bufs(pz_filter* f, int par) {
  if (f->mtt_world <= 1) return MttBufs{f->g.q16, f->ebbuf[0], f->a.sb, f->mtt_part, f->mtt_overflow};
  return MttBufs{par ? f->mtt_q16b : f->g.q16, f->ebbuf[par], par ? f->mtt_sbb : f->a.sb, par ? f->mtt_partb : f->mtt_part, f->mtt_ovf[par]};
}

GenericArgs mtt_generic(pz_filter* f, int par) {
  StepArgs& a = f->a;
  const MttBufs b = mtt_bufs(f, par);
  GenericArgs g{};
  g.lw = f->g.lw; g.q16 = b.q16; g.q16_alt = b.q16; g.xb = a.xb; g.anc = f->g.anc; g.eb = b.eb; g.eb_alt = b.eb;
  g.parity_ctl = nullptr;
  g.sb = b.sb; g.pb = a.pb; g.tile_start = a.tile_start; g.scan_status = a.scan_status; g.scan_part = a.scan_part;
  g.plan = a.plan; g.ctl = a.ctl; g.n = a.n_local; g.nblk = a.nblk; g.ntiles_out = a.ntiles_out; g.nscan = a.nscan;
  g.U = 0; g.seed = a.seed; g.step = f->t;
  return g;
}

// block statistics of this device's stored log-weights (its slice of the global arrays of parity `par`)
int mtt_stats(pz_filter* f, int par) {
  GenericArgs g = mtt_generic(f, par);
  if (f->mtt_world > 1) {
    const int64_t b0 = f->mtt_j0 / kBlk;
    g.lw += f->mtt_j0; g.q16 += f->mtt_j0; g.eb += b0; g.sb += b0;
    g.n = f->mtt_jend - f->mtt_j0; g.nblk = (g.n + kBlk - 1) / kBlk;
  }
  return launch_stats_from_lw(g, f->stream);
}

// scan -> (summary) -> partition over the (gathered) arrays of parity `par`
int mtt_plan_global(pz_filter* f, int par, int delta, double ll_const) {
  StepArgs b = f->a;
  const MttBufs mb = mtt_bufs(f, par);
  b.eb[0] = b.eb[1] = mb.eb; b.sb = mb.sb;
  int n = 0;
  if (f->mtt_world > 1) n += launch_emax_from_eb(mb.eb, b.nblk, b.ctl, f->stream);   // the other devices' exponents
  n += launch_scan_only(b, delta, f->stream);
  if (delta)
    n += launch_mtt_summary(mb.part, b.n_local, b.plan, b.ctl, b.summ, ll_const, mb.ovf, f->stream, f->mtt_world > 1 ? f->mtt_world : 1);
  n += launch_partition_ext(b, kGenericTileOut, delta, f->stream);
  return n;
}

// stats of the stored log-weights -> scan -> (summary) -> partition
int mtt_plan(pz_filter* f, int delta, double ll_const) {
  return mtt_stats(f, 0) + mtt_plan_global(f, 0, delta, ll_const);
}

// rank / world / per: a shard of a multi-device tracker owns particles [rank * per, min((rank + 1) * per, n)); its plan
// arrays cover the whole population (every device derives the global plan from the gathered block statistics)
int create_mtt(const pz_model_desc* model, uint64_t n, uint64_t seed, int device, pz_filter** out, int rank = 0, int world = 1,
               int64_t per = 0) {
  const bool mcam = model->kind == PZ_MODEL_MULTICAM;
  if (!mcam && (model->nx != 4 || model->ny != 2)) return fail(PZ_EINVAL, "the MTT descriptor needs the (4,2) track model");
  if (mcam) {
    if (model->nx != 4 || model->ny != 0) return fail(PZ_EINVAL, "the MultiCamera descriptor has a 4-dimensional box state (ny = 0: H = I and R = meas_var I are implied)");
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
        if (model->F[i * 4 + j] != (i == j ? 1.0 : 0.0) || (i != j && model->Q[i * 4 + j] != 0.0))
          return fail(PZ_EINVAL, "the MultiCamera kernel takes F = I and a diagonal Q (MultiCamera.hs:110-116)");
    if (model->mtt.pd != 1.0) return fail(PZ_EINVAL, "MultiCamera: every live track is observed (pd = 1)");
  }
  if (model->mtt.k_max != 16) return fail(PZ_EINVAL, "k_max must be 16 (one lane per track slot)");
  if (model->mtt.m_max < 1 || model->mtt.m_max > 32) return fail(PZ_EINVAL, "m_max must be in [1, 32]");
  if (n < 1 || n > (1ull << 28)) return fail(PZ_EINVAL, "n_particles must be in [1, 2^28] for the tracker");
  if (model->resampler != PZ_RESAMPLE_SYSTEMATIC) return fail(PZ_EINVAL, "the tracker uses the systematic resampler");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PZ_ECUDA, "no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(PZ_EINVAL, "device %d out of range (%d devices)", device, ndev);
  PZ_CUDA(cudaSetDevice(device));
  pz_filter* f = new (std::nothrow) pz_filter();
  if (!f) return fail(PZ_ENOMEM, "out of host memory");
  f->desc = *model; f->device = device; f->is_mtt = true; f->mtt_mmax = model->mtt.m_max;
  f->is_mcam = mcam; f->tracker_words = mcam ? mcam_words_per_particle() : mtt_words_per_particle(); f->tracker_obs_w = mcam ? 16 : 2;
  f->mtt_world = world; f->mtt_rank = rank; f->mtt_per = world > 1 ? per : 0;
  f->mtt_j0 = world > 1 ? rank * per : 0;
  f->mtt_jend = world > 1 ? (((int64_t)(rank + 1) * per < (int64_t)n) ? (rank + 1) * per : (int64_t)n) : (int64_t)n;
  StepArgs& a = f->a;
  memset(&a, 0, sizeof(a));
  a.model.nx = 4; a.model.ny = 2;
  a.n_global = a.n_local = (int64_t)n; a.world = 1; a.seed = seed; a.resampler = model->resampler;
  a.nblk = (a.n_local + kBlk - 1) / kBlk;
  a.nscan = (int32_t)((a.nblk + kScanTile - 1) / kScanTile);
  a.ntiles_out = (int32_t)((a.n_local + kGenericTileOut - 1) / kGenericTileOut);
  a.ld = round_up(a.n_local, kPad);
  if (mcam) { f->mtt_params.resize(mcam_params_size()); mcam_fill_params(model, f->mtt_params.data()); }
  else { f->mtt_params.resize(mtt_params_size()); mtt_fill_params(model, f->mtt_params.data()); }
  const size_t sbytes = (size_t)(f->mtt_jend - f->mtt_j0) * f->tracker_words * sizeof(uint32_t);
#define PZ_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int c_ = fail(PZ_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); free_filter(f); return c_; } } while (0)
  PZ_TRY(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  PZ_TRY(cudaEventCreate(&f->ev0));
  PZ_TRY(cudaEventCreate(&f->ev1));
  for (int i = 0; i < 2; ++i) { PZ_TRY(cudaMalloc(&f->mtt_s[i], sbytes)); PZ_TRY(cudaMemsetAsync(f->mtt_s[i], 0, sbytes, f->stream)); }
  PZ_TRY(cudaMalloc(&f->ebbuf[0], (a.nblk + 8) * sizeof(int32_t)));
  a.eb[0] = a.eb[1] = f->ebbuf[0];
  PZ_TRY(cudaMalloc(&a.sb, (a.nblk + 8) * sizeof(uint64_t)));
  PZ_TRY(cudaMalloc(&a.pb, (a.nblk + 8) * sizeof(uint64_t)));
  a.pbx = a.pb; a.b_lo = 0; a.b_hi = a.nblk;
  PZ_TRY(cudaMalloc(&f->xbbuf, (a.nblk + 8) * sizeof(uint64_t)));
  a.xb = f->xbbuf;
  PZ_TRY(cudaMalloc(&a.tile_start, (a.ntiles_out + 1) * sizeof(int32_t)));
  PZ_TRY(cudaMalloc(&a.scan_status, (a.nscan + 1) * sizeof(uint64_t)));
  PZ_TRY(cudaMemsetAsync(a.scan_status, 0, (a.nscan + 1) * sizeof(uint64_t), f->stream));
  PZ_TRY(cudaMalloc(&a.scan_part, (size_t)(a.nscan + 1) * kMaxRec * sizeof(double)));
  PZ_TRY(cudaMalloc(&a.plan, sizeof(PlanDev)));
  PZ_TRY(cudaMemsetAsync(a.plan, 0, sizeof(PlanDev), f->stream));
  PZ_TRY(cudaMalloc(&a.ctl, sizeof(CtlDev)));
  PZ_TRY(cudaMemsetAsync(a.ctl, 0, sizeof(CtlDev), f->stream));
  PZ_TRY(cudaMalloc(&a.summ, (size_t)kObsRing * sizeof(pz_step_summary)));
  PZ_TRY(cudaMemsetAsync(a.summ, 0, (size_t)kObsRing * sizeof(pz_step_summary), f->stream));
  PZ_TRY(cudaMallocHost(&f->h_summ, (size_t)kObsRing * sizeof(pz_step_summary)));
  PZ_TRY(cudaMalloc(&f->mtt_obs, 32 * 16 * sizeof(float)));
  PZ_TRY(cudaMallocHost(&f->h_mtt_obs, 32 * 16 * sizeof(float)));
  PZ_TRY(cudaMalloc(&f->mtt_overflow, sizeof(unsigned long long)));
  PZ_TRY(cudaMemsetAsync(f->mtt_overflow, 0, sizeof(unsigned long long), f->stream));
  PZ_TRY(cudaMalloc(&f->mtt_part, (size_t)mtt_step_ctas(a.n_local) * 5 * sizeof(double)));
  {
    int rc = generic_alloc(&f->g, a.n_local, false, f->stream);  // stored log-weights (all 0: equal weights) + ancestors
    if (rc) { free_filter(f); return rc; }
  }
  launch_ctl_reset(a.ctl, f->stream);
  if (world > 1) {   // second parity of the gathered arrays; the first plan follows in create_mtt_multi (it needs every shard)
    PZ_TRY(cudaMalloc(&f->ebbuf[1], (a.nblk + 8) * sizeof(int32_t)));
    PZ_TRY(cudaMalloc(&f->mtt_sbb, (a.nblk + 8) * sizeof(uint64_t)));
    PZ_TRY(cudaMalloc(&f->mtt_q16b, (size_t)a.ld * sizeof(uint16_t)));
    PZ_TRY(cudaMemsetAsync(f->mtt_q16b, 0, (size_t)a.ld * sizeof(uint16_t), f->stream));
    PZ_TRY(cudaMalloc(&f->mtt_partb, (size_t)mtt_step_ctas(a.n_local) * 5 * sizeof(double)));
    for (int i = 0; i < 2; ++i) {
      PZ_TRY(cudaMalloc(&f->mtt_ovf[i], kMttMaxDevices * sizeof(unsigned long long)));
      PZ_TRY(cudaMemsetAsync(f->mtt_ovf[i], 0, kMttMaxDevices * sizeof(unsigned long long), f->stream));
    }
    PZ_TRY(cudaEventCreateWithFlags(&f->ev_fork, cudaEventDisableTiming));
    PZ_TRY(cudaStreamSynchronize(f->stream));
    *out = f;
    return PZ_OK;
  }
  mtt_plan(f, 0, 0.0);
  PZ_TRY(cudaGetLastError());
  PZ_TRY(cudaStreamSynchronize(f->stream));
#undef PZ_TRY
  *out = f;
  return PZ_OK;
}

// observation rows -> device (fp32); MTT: (z0, z1); MultiCamera: (camera, pose[10], box[4], appearance reading)
int tracker_stage_obs(pz_filter* f, const double* obs, int32_t n_obs) {
  const int w = f->tracker_obs_w;
  for (int i = 0; i < w * n_obs; ++i) f->h_mtt_obs[i] = (float)obs[i];
  if (n_obs) PZ_CUDA(cudaMemcpyAsync(f->mtt_obs, f->h_mtt_obs, (size_t)w * n_obs * sizeof(float), cudaMemcpyHostToDevice, f->stream));
  return PZ_OK;
}
int tracker_launch_step(pz_filter* f, const uint32_t* sin, uint32_t* sout, double* part, int32_t n_obs, const MttShard* shard) {
  if (f->is_mcam)
    return launch_mcam_step(f->mtt_params.data(), sin, sout, f->g.anc, f->g.lw, f->mtt_obs, f->a.ctl, f->mtt_overflow, part, f->a.n_local,
                            n_obs, f->a.seed, f->stream, shard);
  return launch_mtt_step(f->mtt_params.data(), sin, sout, f->g.anc, f->g.lw, f->mtt_obs, f->a.ctl, f->mtt_overflow, part, f->a.n_local,
                         n_obs, f->a.seed, f->stream, shard);
}
// per-step constants of the score added on the host.  MTT: -(lambda_c + lambda_b) - log M!  (poisson_ll,
// Distributions.hs:168-169; :256-260); MultiCamera: every term is accumulated by the kernel.
double tracker_ll_const(const pz_filter* f, int32_t n_obs) {
  if (f->is_mcam) return 0.0;
  return -(f->desc.mtt.clutter_lambda + f->desc.mtt.new_track_lambda) - lgamma(n_obs + 1.0);
}
int tracker_check_obs(const pz_filter* f, const double* obs, int32_t n_obs) {
  if (n_obs < 0 || n_obs > f->mtt_mmax) return fail(PZ_ECAPACITY, "%d observations exceed m_max = %d", n_obs, f->mtt_mmax);
  if (f->is_mcam)
    for (int i = 0; i < n_obs; ++i)
      if (!(obs[i * 16] == 0.0 || obs[i * 16] == 1.0)) return fail(PZ_EINVAL, "observation %d: camera id %g (expected 0 or 1)", i, obs[i * 16]);
  return PZ_OK;
}

int mtt_step_host(pz_filter* f, const double* obs, int32_t n_obs, pz_step_summary* out) {
  StepArgs& a = f->a;
  int rc0 = tracker_check_obs(f, obs, n_obs);
  if (rc0) return rc0;
  const uint32_t t0 = f->t;
  rc0 = tracker_stage_obs(f, obs, n_obs);
  if (rc0) return rc0;
  PZ_CUDA(cudaEventRecord(f->ev0, f->stream));
  GenericArgs g = mtt_generic(f, 0);
  int n = launch_generic_ancestors(g, f->stream);
  n += tracker_launch_step(f, f->mtt_s[t0 & 1], f->mtt_s[(t0 + 1) & 1], f->mtt_part, n_obs, nullptr);
  const double ll_const = tracker_ll_const(f, n_obs);
  n += mtt_plan(f, 1, ll_const);
  f->t += 1;
  PZ_CUDA(cudaEventRecord(f->ev1, f->stream));
  PZ_CUDA(cudaMemcpyAsync(f->h_summ + (t0 % kObsRing), a.summ + (t0 % kObsRing), sizeof(pz_step_summary), cudaMemcpyDeviceToHost, f->stream));
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  PZ_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, f->ev0, f->ev1);
  f->last_ms = ms; f->last_launches = n;
  const pz_step_summary& s = f->h_summ[t0 % kObsRing];
  if (out) *out = s;
  if (s.total_weight == 0) return fail(PZ_EDEGENERATE, "step %u: every particle has zero weight", t0);
  return PZ_OK;
}

// ---- multi-device tracker (one process, D devices; MultiTargetTracking.hs:133-145 under zdsparticles) ------------
// Device r owns particles and output slots [r * per, (r + 1) * per).  A step on every device:
//   pf_ancestors over its own output tiles (global ancestor indices from the global plan) -> mtt_step over its own
//   particles, parent records read from their owners over NVLink -> block statistics of its own log-weights ->
//   pf_push_slices: its slices of the 16-bit weights, block exponents / sums, summary partials and dropped-birth counter
//   stored into every peer (the all-gather) -> [wait for every peer's push] -> global exponent, scan, summary and
//   partition over the gathered arrays, redundantly on every device (the scan of 4 * 10^5 blocks takes 25 us).
// Particles, Philox counters and summary partials are keyed on GLOBAL indices: the run equals the one-device run bit
// for bit (tests/test_mtt.py::test_multi_device_tracker_equals_single_device).
int mtt_push(pz_filter* parent, int r, int par) {
  pz_filter* sh = parent->shards[r];
  const int D = (int)parent->shards.size();
  const MttBufs mine = mtt_bufs(sh, par);
  const int64_t j0 = sh->mtt_j0, cnt = sh->mtt_jend - sh->mtt_j0;
  const int64_t b0 = j0 / kBlk, nb = (cnt + kBlk - 1) / kBlk;
  const int64_t c0 = j0 / 128, nc = mtt_step_ctas(cnt);
  PushArgs p{};
  p.n_arr = 4;
  p.src[0] = mine.q16 + j0;  p.bytes[0] = nb * kBlk * (int64_t)sizeof(uint16_t);
  p.src[1] = mine.eb + b0;   p.bytes[1] = nb * (int64_t)sizeof(int32_t);
  p.src[2] = mine.sb + b0;   p.bytes[2] = nb * (int64_t)sizeof(uint64_t);
  p.src[3] = mine.part + c0 * 5; p.bytes[3] = nc * 5 * (int64_t)sizeof(double);
  PushArgs o{};   // the dropped-birth counter goes to every device, this one included
  o.n_arr = 1; o.src[0] = sh->mtt_overflow; o.bytes[0] = sizeof(unsigned long long);
  for (int q = 0; q < D; ++q) {
    const MttBufs theirs = mtt_bufs(parent->shards[q], par);
    o.dst[o.n_peers++][0] = theirs.ovf + r;
    if (q == r) continue;
    void** d = p.dst[p.n_peers++];
    d[0] = theirs.q16 + j0; d[1] = theirs.eb + b0; d[2] = theirs.sb + b0; d[3] = theirs.part + c0 * 5;
  }
  int n = launch_push_slices(p, sh->stream);
  n += launch_push_slices(o, sh->stream);
  PZ_CUDA(cudaEventRecord(sh->ev_fork, sh->stream));
  return n;
}

int enable_peer_access(int n_devices, const int* device_ids) {
  for (int r = 0; r < n_devices; ++r) {
    cudaSetDevice(device_ids[r]);
    for (int q = 0; q < n_devices; ++q) {
      if (q == r || device_ids[q] == device_ids[r]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, device_ids[r], device_ids[q]);
      if (!can) return fail(PZ_ECUDA, "device %d cannot access device %d (no peer path)", device_ids[r], device_ids[q]);
      cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[q], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(PZ_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", device_ids[r], device_ids[q], cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
  return PZ_OK;
}

int create_mtt_multi(const pz_model_desc* model, uint64_t n, uint64_t seed, int n_devices, const int* device_ids, pz_filter** out) {
  const int D = n_devices;
  if (D > kMttMaxDevices) return fail(PZ_EINVAL, "the multi-device tracker takes up to %d devices", kMttMaxDevices);
  // whole resampling tiles (2048 slots = 8 blocks = 16 mtt_step CTAs) per device
  const int64_t per = round_up(((int64_t)n + D - 1) / D, (int64_t)kGenericTileOut);
  if ((int64_t)(D - 1) * per >= (int64_t)n) return fail(PZ_EINVAL, "%llu particles are too few for %d devices (%d per tile)", (unsigned long long)n, D, kGenericTileOut);
  pz_filter* parent = new (std::nothrow) pz_filter();
  if (!parent) return fail(PZ_ENOMEM, "out of host memory");
  parent->desc = *model; parent->is_multi = true; parent->is_mtt = true; parent->device = device_ids[0];
  parent->mtt_mmax = model->mtt.m_max; parent->mtt_world = D; parent->mtt_per = per;
  parent->is_mcam = model->kind == PZ_MODEL_MULTICAM;
  parent->tracker_words = parent->is_mcam ? mcam_words_per_particle() : mtt_words_per_particle();
  parent->tracker_obs_w = parent->is_mcam ? 16 : 2;
  memset(&parent->a, 0, sizeof(parent->a));
  auto bail = [&](int code) { free_filter(parent); return code; };
  for (int r = 0; r < D; ++r) {
    pz_filter* sh = nullptr;
    int rc = create_mtt(model, n, seed, device_ids[r], &sh, r, D, per);
    if (rc) return bail(rc);
    parent->shards.push_back(sh);
  }
  parent->a = parent->shards[0]->a;
  parent->a.world = 1;   // (the plan is global on every device: no halo exchange, no rank-relative indices)
  int rc = enable_peer_access(D, device_ids);
  if (rc) return bail(rc);
  // the plan of the prior population (all log-weights 0)
  for (int r = 0; r < D; ++r) {
    pz_filter* sh = parent->shards[r];
    if (cudaSetDevice(sh->device) != cudaSuccess) return bail(fail(PZ_ECUDA, "cudaSetDevice(%d) failed", sh->device));
    mtt_stats(sh, 0);
    rc = mtt_push(parent, r, 0);
    if (rc < 0) return bail(rc);
  }
  for (int q = 0; q < D; ++q) {
    pz_filter* sh = parent->shards[q];
    cudaSetDevice(sh->device);
    for (int r = 0; r < D; ++r) if (r != q) cudaStreamWaitEvent(sh->stream, parent->shards[r]->ev_fork, 0);
    mtt_plan_global(sh, 0, 0, 0.0);
  }
  for (int q = 0; q < D; ++q) {
    pz_filter* sh = parent->shards[q];
    cudaSetDevice(sh->device);
    cudaError_t e = cudaStreamSynchronize(sh->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return bail(fail(PZ_ECUDA, "multi-device tracker: first plan on device %d: %s", sh->device, cudaGetErrorString(e)));
  }
  *out = parent;
  return PZ_OK;
}

MttShard mtt_shard_view(pz_filter* parent, pz_filter* sh, int rec_parity) {
  MttShard ms;
  ms.j0 = sh->mtt_j0; ms.j_end = sh->mtt_jend; ms.n_per = parent->mtt_per;
  for (size_t q = 0; q < parent->shards.size(); ++q) ms.sin_all[q] = parent->shards[q]->mtt_s[rec_parity];
  return ms;
}

int mtt_multi_step_host(pz_filter* p, const double* obs, int32_t n_obs, pz_step_summary* out) {
  int rc0 = tracker_check_obs(p->shards[0], obs, n_obs);
  if (rc0) return rc0;
  const int D = (int)p->shards.size();
  const uint32_t t0 = p->t;
  const int pin = (int)(t0 & 1u), pout = pin ^ 1;
  const double ll_const = tracker_ll_const(p->shards[0], n_obs);
  int64_t launches = 0;
  for (int r = 0; r < D; ++r) {   // own slots: ancestors -> tracker step -> statistics -> push
    pz_filter* f = p->shards[r];
    PZ_CUDA(cudaSetDevice(f->device));
    rc0 = tracker_stage_obs(f, obs, n_obs);
    if (rc0) return rc0;
    PZ_CUDA(cudaEventRecord(f->ev0, f->stream));
    GenericArgs g = mtt_generic(f, pin);
    g.tile0 = (int32_t)(f->mtt_j0 / kGenericTileOut);
    g.ntiles_out = (int32_t)((f->mtt_jend - f->mtt_j0 + kGenericTileOut - 1) / kGenericTileOut);
    launches += launch_generic_ancestors(g, f->stream);
    const MttShard ms = mtt_shard_view(p, f, (int)(t0 & 1u));
    launches += tracker_launch_step(f, nullptr, f->mtt_s[(t0 + 1) & 1], mtt_bufs(f, pout).part + (f->mtt_j0 / 128) * 5, n_obs, &ms);
    launches += mtt_stats(f, pout);
    int rc = mtt_push(p, r, pout);
    if (rc < 0) return rc;
    launches += rc;
  }
  for (int q = 0; q < D; ++q) {   // the global plan, on every device, once every push has landed
    pz_filter* f = p->shards[q];
    PZ_CUDA(cudaSetDevice(f->device));
    for (int r = 0; r < D; ++r) if (r != q) PZ_CUDA(cudaStreamWaitEvent(f->stream, p->shards[r]->ev_fork, 0));
    launches += mtt_plan_global(f, pout, 1, ll_const);
    f->t += 1;
    PZ_CUDA(cudaEventRecord(f->ev1, f->stream));
    PZ_CUDA(cudaMemcpyAsync(f->h_summ + (t0 % kObsRing), f->a.summ + (t0 % kObsRing), sizeof(pz_step_summary), cudaMemcpyDeviceToHost, f->stream));
  }
  double ms_max = 0;
  for (int q = 0; q < D; ++q) {
    pz_filter* f = p->shards[q];
    PZ_CUDA(cudaSetDevice(f->device));
    PZ_CUDA(cudaStreamSynchronize(f->stream));
    PZ_CUDA(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, f->ev0, f->ev1);
    ms_max = ms > ms_max ? ms : ms_max;
  }
  p->t += 1;
  p->last_ms = ms_max; p->last_launches = launches;
  const pz_step_summary& s = p->shards[0]->h_summ[t0 % kObsRing];   // every device derives the same summary
  if (out) *out = s;
  if (s.total_weight == 0) return fail(PZ_EDEGENERATE, "step %u: every particle has zero weight", t0);
  return PZ_OK;
}

// Peer-memory exchange of a sharded filter: map every rank's mailbox and the two neighbours' particle
// arrays with CUDA IPC (handles travel through one NCCL all-gather at creation).  All ranks must agree on
// the outcome, so the verdict is all-reduced; on any failure the filter keeps the NCCL exchange.
int setup_p2p(pz_filter* f) {
  StepArgs& a = f->a;
  const int R = a.world, me = a.rank;
  if (getenv("PZ_SHARD_NCCL") || R > kMaxRanks) return PZ_OK;
  // buffers of every rank are mapped by every rank: the mailbox, and -- for the remote-read fallback and the
  // neighbours' halos -- state, weights, exponents, boundary positions and local prefixes
  constexpr int NH = 10;
  struct Pack { cudaIpcMemHandle_t h[NH]; };
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  Pack mine;
  memset(&mine, 0, sizeof(mine));
  bool ok = cudaMalloc(&f->mbox, sizeof(Mailbox)) == cudaSuccess && cudaMemset(f->mbox, 0, sizeof(Mailbox)) == cudaSuccess;
  void* bases[NH] = {f->mbox, f->xbuf[0], f->xbuf[1], f->qbuf[0], f->qbuf[1], f->ebbuf[0], f->ebbuf[1], f->pbxbuf, f->xbbuf, a.pb};
  for (int i = 0; i < NH && ok; ++i) ok = cudaIpcGetMemHandle(&mine.h[i], bases[i]) == cudaSuccess;
  // all-gather of the handles.  Every rank must enter both collectives below whatever failed locally: a local failure
  // only lowers this rank's verdict.
  unsigned char* d_all = nullptr;
  std::vector<Pack> all((size_t)R);
  int fatal = PZ_OK;
  if (cudaMalloc(&d_all, (size_t)R * sizeof(Pack)) != cudaSuccess) { ok = false; d_all = nullptr; cudaGetLastError(); }
  {
    // (without the staging buffer this rank still joins the collective, in place on its handle-free scratch)
    unsigned char* buf = d_all;
    if (buf && cudaMemcpy(buf + (size_t)me * sizeof(Pack), &mine, sizeof(Pack), cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
    if (buf) {
      ncclResult_t r = g_nccl.AllGather(buf + (size_t)me * sizeof(Pack), buf, sizeof(Pack), ncclUint8, f->comm, f->stream);
      if (r != ncclSuccess) { ok = false; fatal = fail(PZ_ENCCL, "all-gather of the IPC handles failed: %s", g_nccl.GetErrorString(r)); }
      if (cudaStreamSynchronize(f->stream) != cudaSuccess) ok = false;
      if (cudaMemcpy(all.data(), buf, (size_t)R * sizeof(Pack), cudaMemcpyDeviceToHost) != cudaSuccess) ok = false;
      cudaFree(buf);
    } else {
      fatal = fail(PZ_ECUDA, "cudaMalloc of the IPC staging buffer failed");   // the peers' all-gather cannot complete without this rank
    }
    cudaGetLastError();
  }
  if (fatal != PZ_OK) return fatal;
  ShardP2P h{};
  const int64_t H = f->hb * kBlk;
  const size_t esz = a.f64 ? 8 : 4;
  auto open = [&](const cudaIpcMemHandle_t& hd) -> void* {
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    f->ipc_opened.push_back(p);
    return p;
  };
  std::vector<std::vector<void*>> peer((size_t)R, std::vector<void*>(NH, nullptr));
  for (int q = 0; q < R && ok; ++q) {
    if (q == me) { for (int i = 0; i < NH; ++i) peer[q][i] = bases[i]; continue; }
    for (int i = 0; i < NH && ok; ++i) { peer[q][i] = open(all[q].h[i]); ok = peer[q][i] != nullptr; }
  }
  for (int q = 0; q < R && ok; ++q) {
    void** p = peer[q].data();
    h.mbox[q] = (Mailbox*)p[0];
    for (int b = 0; b < 2; ++b) {
      h.Xall[q][b] = (char*)p[1 + b] + (size_t)H * esz;
      h.Qall[q][b] = (uint16_t*)p[3 + b] + H;
      h.eball[q][b] = (int32_t*)p[5 + b] + f->hb;
    }
    h.xball[q] = (uint64_t*)p[8] + f->hb;
    h.pball[q] = (uint64_t*)p[9];
  }
  for (int side = 0; side < 2 && ok; ++side) {
    const int q = side == 0 ? me - 1 : me + 1;
    if (q < 0 || q >= R) continue;
    void** p = peer[q].data();
    for (int b = 0; b < 2; ++b) { h.Xn[side][b] = h.Xall[q][b]; h.Qn[side][b] = h.Qall[q][b]; h.ebn[side][b] = h.eball[q][b]; }
    h.pbxn[side] = (uint64_t*)p[7] + f->hb;
    h.xbn[side] = h.xball[q];
  }
  if (ok) ok = cudaMalloc(&f->d_p2p, sizeof(ShardP2P)) == cudaSuccess &&
               cudaMemcpy(f->d_p2p, &h, sizeof(ShardP2P), cudaMemcpyHostToDevice) == cudaSuccess;
  if (ok) ok = cudaStreamCreateWithFlags(&f->push_stream, cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreateWithFlags(&f->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&f->ev_join, cudaEventDisableTiming) == cudaSuccess;
  cudaGetLastError();
  // agreement: min over ranks of the local verdict (every rank enters the all-reduce)
  const int32_t verdict = ok ? 1 : 0;
  int32_t agreed = 0;
  const bool staged = cudaMemcpy(f->warm, &verdict, sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess;
  ncclResult_t r = g_nccl.AllReduce(f->warm, f->warm, 1, ncclInt32, ncclMin, f->comm, f->stream);
  const bool synced = cudaStreamSynchronize(f->stream) == cudaSuccess;
  if (r != ncclSuccess) return fail(PZ_ENCCL, "all-reduce of the peer-memory verdict failed: %s", g_nccl.GetErrorString(r));
  if (!staged || !synced || cudaMemcpy(&agreed, f->warm, sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    return fail(PZ_ECUDA, "peer-memory set-up: the verdict all-reduce could not be staged on this rank");
  }
  if (agreed) a.p2p = f->d_p2p;
  return PZ_OK;
}

int create_rb(const pz_model_desc* model, uint64_t n, int device, pz_filter** out) {
  const int nx = model->nx, ny = model->ny;
  if (nx < 1 || nx > PZ_MAX_NX || ny < 1 || ny > PZ_MAX_NY) return fail(PZ_EINVAL, "unsupported dims nx=%d ny=%d", nx, ny);
  if (n < 1 || n > (1ull << 30)) return fail(PZ_EINVAL, "n_particles must be in [1, 2^30]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PZ_ECUDA, "no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(PZ_EINVAL, "device %d out of range (%d devices)", device, ndev);
  PZ_CUDA(cudaSetDevice(device));
  pz_filter* f = new (std::nothrow) pz_filter();
  if (!f) return fail(PZ_ENOMEM, "out of host memory");
  f->desc = *model; f->device = device; f->is_rb = true;
  memset(&f->a, 0, sizeof(f->a));
  f->a.model.nx = nx; f->a.model.ny = ny; f->a.n_local = f->a.n_global = (int64_t)n; f->a.world = 1;
  RbDev h{};
  h.nx = nx; h.ny = ny; h.t = 0;
  for (int i = 0; i < nx * nx; ++i) { h.F[i] = model->F[i]; h.Q[i] = model->Q[i]; }
  for (int i = 0; i < ny * nx; ++i) h.H[i] = model->H[i];
  for (int i = 0; i < ny * ny; ++i) h.R[i] = model->R[i];
  for (int i = 0; i < nx; ++i) h.mean[i] = model->x0[i];  // point mass: zero covariance
  h.ll_scale = (ny == 1 && model->scalar_ll_2x) ? 2.0 : 1.0;
#define PZ_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int c_ = fail(PZ_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); free_filter(f); return c_; } } while (0)
  PZ_TRY(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  PZ_TRY(cudaEventCreate(&f->ev0));
  PZ_TRY(cudaEventCreate(&f->ev1));
  PZ_TRY(cudaMalloc(&f->rb, sizeof(RbDev)));
  PZ_TRY(cudaMemcpy(f->rb, &h, sizeof(RbDev), cudaMemcpyHostToDevice));
  PZ_TRY(cudaMalloc(&f->a.summ, sizeof(pz_step_summary)));
  PZ_TRY(cudaMallocHost(&f->h_summ, sizeof(pz_step_summary)));
#undef PZ_TRY
  *out = f;
  return PZ_OK;
}

int rb_step_host(pz_filter* f, const double* obs, pz_step_summary* out) {
  PZ_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(f->rb) + offsetof(RbDev, y), obs, f->a.model.ny * sizeof(double),
                          cudaMemcpyHostToDevice, f->stream));
  PZ_CUDA(cudaEventRecord(f->ev0, f->stream));
  rb_kalman_step<<<1, 32, 0, f->stream>>>(f->rb, f->a.summ, (double)f->a.n_local);
  PZ_CUDA(cudaEventRecord(f->ev1, f->stream));
  PZ_CUDA(cudaMemcpyAsync(f->h_summ, f->a.summ, sizeof(pz_step_summary), cudaMemcpyDeviceToHost, f->stream));
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  PZ_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, f->ev0, f->ev1);
  f->last_ms += ms; f->last_launches += 1;
  f->t += 1;
  if (out) *out = *f->h_summ;
  if (!(f->h_summ->log_evidence_inc == f->h_summ->log_evidence_inc))
    return fail(PZ_EDEGENERATE, "step %u: the innovation covariance is not positive definite", f->t - 1);
  return PZ_OK;
}


// ---- generic-path filters (stored log-weights): carried weights, adaptive resampling, the walker ----------
GenericArgs gen_generic(pz_filter* f) {
  StepArgs& a = f->a;
  GenericArgs g{};
  g.lw = f->gen_lw[f->t & 1]; g.q16 = f->g.q16; g.q16_alt = f->g.q16; g.xb = a.xb; g.anc = f->g.anc; g.eb = a.eb[0]; g.eb_alt = a.eb[0];
  g.parity_ctl = nullptr;
  g.sb = a.sb; g.pb = a.pb; g.tile_start = a.tile_start; g.scan_status = a.scan_status; g.scan_part = a.scan_part;
  g.plan = a.plan; g.ctl = a.ctl; g.n = a.n_local; g.nblk = a.nblk; g.ntiles_out = a.ntiles_out; g.nscan = a.nscan;
  g.U = 0; g.seed = a.seed; g.step = f->t;
  return g;
}

// stats of the stored log-weights of buffer `which` -> scan -> (summary) -> partition
int gen_plan(pz_filter* f, int which, int delta) {
  StepArgs& a = f->a;
  GenericArgs g = gen_generic(f);
  g.lw = f->gen_lw[which];
  int n = launch_stats_from_lw(g, f->stream);
  n += launch_scan_only(a, delta, f->stream);
  if (delta) n += launch_gen_summary(f->ga, f->stream);
  n += launch_partition_ext(a, kGenericTileOut, delta, f->stream);
  return n;
}

int gen_start(pz_filter* f) {  // the prior population and its (trivial) plan
  launch_ctl_reset(f->a.ctl, f->stream);
  launch_gen_init(f->ga, f->stream);
  gen_plan(f, 0, 0);
  f->t = 0;
  PZ_CUDA(cudaGetLastError());
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  return PZ_OK;
}

int create_gen(const pz_model_desc* model, uint64_t n, uint64_t seed, int device, pz_filter** out) {
  const bool walker = model->kind == PZ_MODEL_WALKER;
  if (!walker) {
    if (!((model->nx == 1 && model->ny == 1) || (model->nx == 4 && model->ny == 2) || (model->nx == 6 && model->ny == 3)))
      return fail(PZ_EINVAL, "no kernel instantiated for nx=%d ny=%d (supported: (1,1), (4,2), (6,3))", model->nx, model->ny);
    if (model->state_f64) return fail(PZ_EINVAL, "carry_weights / ess_threshold filters store fp32 state (state_f64 = 0)");
  }
  if (n < 1 || n > (1ull << 28)) return fail(PZ_EINVAL, "n_particles must be in [1, 2^28] on the stored-log-weight path");
  if (!(model->ess_threshold >= 0.0 && model->ess_threshold <= 1.0)) return fail(PZ_EINVAL, "ess_threshold must be in [0, 1]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PZ_ECUDA, "no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(PZ_EINVAL, "device %d out of range (%d devices)", device, ndev);
  PZ_CUDA(cudaSetDevice(device));
  pz_filter* f = new (std::nothrow) pz_filter();
  if (!f) return fail(PZ_ENOMEM, "out of host memory");
  f->desc = *model; f->device = device; f->is_gen = true;
  StepArgs& a = f->a;
  memset(&a, 0, sizeof(a));
  double ll_const = 0;
  if (walker) {
    a.model.nx = 5; a.model.ny = 2; a.f64 = 1;
    // two doubled scalar scores per step: each -log(2 pi) - log(sigma^2) (Distributions.hs:160-161)
    const double v = model->walker.obs_std * model->walker.obs_std;
    if (!(v > 0) || !(model->walker.dt > 0)) { delete f; return fail(PZ_EINVAL, "bad walker parameters"); }
    ll_const = 0.0;  // the constant is inside score_normal_2x already
  } else {
    int rc = derive_lg(model, &a.model, &a.modeld, &ll_const);
    if (rc) { delete f; return rc; }
  }
  a.ll_const = ll_const;
  a.n_global = a.n_local = (int64_t)n; a.world = 1; a.seed = seed; a.resampler = model->resampler;
  a.nblk = (a.n_local + kBlk - 1) / kBlk;
  a.nscan = (int32_t)((a.nblk + kScanTile - 1) / kScanTile);
  a.ntiles_out = (int32_t)((a.n_local + kGenericTileOut - 1) / kGenericTileOut);
  a.ld = round_up(a.n_local, kPad);
  const size_t esz = a.f64 ? 8 : 4;
  const size_t xbytes = (size_t)a.model.nx * a.ld * esz;
#define PZ_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int c_ = fail(PZ_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); free_filter(f); return c_; } } while (0)
  PZ_TRY(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  PZ_TRY(cudaEventCreate(&f->ev0));
  PZ_TRY(cudaEventCreate(&f->ev1));
  for (int i = 0; i < 2; ++i) {
    PZ_TRY(cudaMalloc(&f->xbuf[i], xbytes));
    PZ_TRY(cudaMemsetAsync(f->xbuf[i], 0, xbytes, f->stream));
    a.X[i] = f->xbuf[i];
    PZ_TRY(cudaMalloc(&f->gen_lw[i], (size_t)a.ld * sizeof(float)));
    PZ_TRY(cudaMemsetAsync(f->gen_lw[i], 0, (size_t)a.ld * sizeof(float), f->stream));
  }
  PZ_TRY(cudaMalloc(&f->ebbuf[0], (a.nblk + 8) * sizeof(int32_t)));
  a.eb[0] = a.eb[1] = f->ebbuf[0];
  PZ_TRY(cudaMalloc(&a.sb, (a.nblk + 8) * sizeof(uint64_t)));
  PZ_TRY(cudaMalloc(&a.pb, (a.nblk + 8) * sizeof(uint64_t)));
  a.pbx = a.pb; a.b_lo = 0; a.b_hi = a.nblk;
  PZ_TRY(cudaMalloc(&f->xbbuf, (a.nblk + 8) * sizeof(uint64_t)));
  a.xb = f->xbbuf;
  PZ_TRY(cudaMalloc(&a.tile_start, (a.ntiles_out + 1) * sizeof(int32_t)));
  PZ_TRY(cudaMalloc(&a.scan_status, (a.nscan + 1) * sizeof(uint64_t)));
  PZ_TRY(cudaMemsetAsync(a.scan_status, 0, (a.nscan + 1) * sizeof(uint64_t), f->stream));
  PZ_TRY(cudaMalloc(&a.scan_part, (size_t)(a.nscan + 1) * kMaxRec * sizeof(double)));
  PZ_TRY(cudaMalloc(&a.plan, sizeof(PlanDev)));
  PZ_TRY(cudaMemsetAsync(a.plan, 0, sizeof(PlanDev), f->stream));
  PZ_TRY(cudaMalloc(&a.ctl, sizeof(CtlDev)));
  PZ_TRY(cudaMemsetAsync(a.ctl, 0, sizeof(CtlDev), f->stream));
  PZ_TRY(cudaMalloc(&a.summ, sizeof(pz_step_summary)));
  PZ_TRY(cudaMemsetAsync(a.summ, 0, sizeof(pz_step_summary), f->stream));
  PZ_TRY(cudaMallocHost(&f->h_summ, sizeof(pz_step_summary)));
  PZ_TRY(cudaMalloc(&f->gen_obs, PZ_MAX_NY * sizeof(double)));
  PZ_TRY(cudaMallocHost(&f->h_gen_obs, PZ_MAX_NY * sizeof(double)));
  PZ_TRY(cudaMalloc(&f->gen_part, (size_t)gen_summary_parts(a.n_local) * gen_summary_rec() * sizeof(double)));
  PZ_TRY(cudaMalloc(&f->gen_state, sizeof(GenState)));
  {
    int rc = generic_alloc(&f->g, a.n_local, false, f->stream);  // q16 + ancestors (+ an unused lw buffer)
    if (rc) { free_filter(f); return rc; }
  }
  GenArgs& g = f->ga;
  memset(&g, 0, sizeof(g));
  g.kind = model->kind; g.carry_weights = model->carry_weights ? 1 : 0;
  g.n_summary = walker ? 6 : a.model.nx;
  g.ess_threshold = model->ess_threshold;
  g.X[0] = a.X[0]; g.X[1] = a.X[1]; g.lw[0] = f->gen_lw[0]; g.lw[1] = f->gen_lw[1];
  g.q16 = f->g.q16; g.eb = a.eb[0]; g.anc = f->g.anc; g.obs = f->gen_obs; g.plan = a.plan; g.ctl = a.ctl; g.gen = f->gen_state;
  g.summ = a.summ; g.part = f->gen_part; g.n = a.n_local; g.ld = a.ld; g.seed = seed; g.ll_const = ll_const;
  g.model = a.model; g.walker = model->walker;
  {
    int rc = gen_start(f);
    if (rc) { free_filter(f); return rc; }
  }
#undef PZ_TRY
  *out = f;
  return PZ_OK;
}

int gen_step_host(pz_filter* f, const double* obs, int32_t obs_dim, pz_step_summary* out) {
  StepArgs& a = f->a;
  const uint32_t t0 = f->t;
  for (int i = 0; i < PZ_MAX_NY; ++i) f->h_gen_obs[i] = i < obs_dim ? obs[i] : 0.0;
  PZ_CUDA(cudaMemcpyAsync(f->gen_obs, f->h_gen_obs, PZ_MAX_NY * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  PZ_CUDA(cudaEventRecord(f->ev0, f->stream));
  int n = 0;
  int rc = pending_ancestors(f, &n);   // canonical systematic (or multinomial) ancestors; identity when the adaptive rule says so
  if (rc) return rc;
  n += launch_gen_step(f->ga, f->stream);
  n += gen_plan(f, (int)((t0 + 1) & 1u), 1);
  f->t += 1;
  PZ_CUDA(cudaEventRecord(f->ev1, f->stream));
  PZ_CUDA(cudaMemcpyAsync(f->h_summ, a.summ, sizeof(pz_step_summary), cudaMemcpyDeviceToHost, f->stream));
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  PZ_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, f->ev0, f->ev1);
  f->last_ms = ms; f->last_launches = n;
  if (out) *out = *f->h_summ;
  if (f->h_summ->total_weight == 0) return fail(PZ_EDEGENERATE, "step %u: every particle has zero weight", t0);
  return PZ_OK;
}

int create_beta(const pz_model_desc* model, uint64_t n, int device, pz_filter** out) {
  if (!(model->beta_a > 0) || !(model->beta_b > 0)) return fail(PZ_EINVAL, "the Beta prior needs a, b > 0");
  if (!model->delayed) return fail(PZ_EINVAL, "the Beta-Bernoulli descriptor runs under delayed sampling (delayed = 1)");
  if (n < 1 || n > (1ull << 30)) return fail(PZ_EINVAL, "n_particles must be in [1, 2^30]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PZ_ECUDA, "no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(PZ_EINVAL, "device %d out of range (%d devices)", device, ndev);
  PZ_CUDA(cudaSetDevice(device));
  pz_filter* f = new (std::nothrow) pz_filter();
  if (!f) return fail(PZ_ENOMEM, "out of host memory");
  f->desc = *model; f->device = device; f->is_rb = true; f->is_beta = true;
  memset(&f->a, 0, sizeof(f->a));
  f->a.model.nx = 1; f->a.model.ny = 1; f->a.n_local = f->a.n_global = (int64_t)n; f->a.world = 1;
  BetaDev h{};
  h.a = model->beta_a; h.b = model->beta_b; h.t = 0;
#define PZ_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int c_ = fail(PZ_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); free_filter(f); return c_; } } while (0)
  PZ_TRY(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  PZ_TRY(cudaEventCreate(&f->ev0));
  PZ_TRY(cudaEventCreate(&f->ev1));
  PZ_TRY(cudaMalloc(&f->beta, sizeof(BetaDev)));
  PZ_TRY(cudaMemcpy(f->beta, &h, sizeof(BetaDev), cudaMemcpyHostToDevice));
  PZ_TRY(cudaMalloc(&f->a.summ, sizeof(pz_step_summary)));
  PZ_TRY(cudaMallocHost(&f->h_summ, sizeof(pz_step_summary)));
#undef PZ_TRY
  *out = f;
  return PZ_OK;
}

int beta_step_host(pz_filter* f, const double* obs, pz_step_summary* out) {
  if (!(obs[0] == 0.0 || obs[0] == 1.0)) return fail(PZ_EINVAL, "a Bernoulli observation is 0 or 1");
  PZ_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(f->beta) + offsetof(BetaDev, y), obs, sizeof(double), cudaMemcpyHostToDevice, f->stream));
  PZ_CUDA(cudaEventRecord(f->ev0, f->stream));
  launch_beta_bernoulli(f->beta, f->a.summ, (double)f->a.n_local, f->stream);
  PZ_CUDA(cudaEventRecord(f->ev1, f->stream));
  PZ_CUDA(cudaMemcpyAsync(f->h_summ, f->a.summ, sizeof(pz_step_summary), cudaMemcpyDeviceToHost, f->stream));
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  PZ_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, f->ev0, f->ev1);
  f->last_ms += ms; f->last_launches += 1;
  f->t += 1;
  if (out) *out = *f->h_summ;
  return PZ_OK;
}

int capture_sharded_graphs(pz_filter* f);

// shard_only: one shard of a single-process multi-device filter (pz_filter_create_multi): allocate and initialise,
// leave the peer wiring, the first plan and the graph capture to the caller
int create_common(const pz_model_desc* model, uint64_t n_global, uint64_t seed, int device, int rank, int world,
                  const void* nccl_id, pz_filter** out, bool shard_only = false) {
  if (!model || !out) return fail(PZ_EINVAL, "null argument");
  *out = nullptr;
  if (model->kind == PZ_MODEL_MTT || model->kind == PZ_MODEL_MULTICAM) {
    if (world != 1) return fail(PZ_EINVAL, "one tracker per process: several GPUs through pz_filter_create_multi");
    return create_mtt(model, n_global, seed, device, out);
  }
  if (model->kind == PZ_MODEL_BETA_BERNOULLI) {
    if (world != 1) return fail(PZ_EINVAL, "a delayed-sampling filter runs on one GPU (all particles are identical)");
    return create_beta(model, n_global, device, out);
  }
  if (model->kind == PZ_MODEL_WALKER) {
    if (world != 1) return fail(PZ_EINVAL, "the walker runs on one GPU");
    return create_gen(model, n_global, seed, device, out);
  }
  if (model->kind != PZ_MODEL_RW1D && model->kind != PZ_MODEL_LINGAUSS) return fail(PZ_EINVAL, "unknown model kind %d", model->kind);
  if (!model->delayed && (model->carry_weights || model->ess_threshold > 0)) {
    if (world != 1) return fail(PZ_EINVAL, "carry_weights / ess_threshold filters run on one GPU (stored-log-weight path)");
    return create_gen(model, n_global, seed, device, out);
  }
  if (model->delayed) {
    if (world != 1) return fail(PZ_EINVAL, "a delayed-sampling filter runs on one GPU (all particles are identical)");
    return create_rb(model, n_global, device, out);
  }
  if (!((model->nx == 1 && model->ny == 1) || (model->nx == 4 && model->ny == 2) || (model->nx == 6 && model->ny == 3)))
    return fail(PZ_EINVAL, "no kernel instantiated for nx=%d ny=%d (supported: (1,1), (4,2), (6,3))", model->nx, model->ny);
  if (n_global < 1 || n_global > (1ull << 30)) return fail(PZ_EINVAL, "n_particles must be in [1, 2^30]");
  if (world < 1 || rank < 0 || rank >= world) return fail(PZ_EINVAL, "bad rank %d / world %d", rank, world);
  if (world > 1) {
    if (world > 32) return fail(PZ_EINVAL, "at most 32 ranks");
    if (n_global % ((uint64_t)world * kBlk) != 0)
      return fail(PZ_EINVAL, "a sharded filter needs n_particles divisible by world * 256 (whole resampling blocks per rank)");
    if (model->resampler != PZ_RESAMPLE_SYSTEMATIC) return fail(PZ_EINVAL, "sharded filters use the systematic resampler");
    if (!shard_only) {
      if (!nccl_id) return fail(PZ_EINVAL, "null nccl_unique_id");
      int rc = load_nccl();
      if (rc) return rc;
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PZ_ECUDA, "no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(PZ_EINVAL, "device %d out of range (%d devices)", device, ndev);
  PZ_CUDA(cudaSetDevice(device));
  pz_filter* f = new (std::nothrow) pz_filter();
  if (!f) return fail(PZ_ENOMEM, "out of host memory");
  f->desc = *model;
  f->device = device;
  StepArgs& a = f->a;
  memset(&a, 0, sizeof(a));
  int rc = derive_lg(model, &a.model, &a.modeld, &a.ll_const);
  if (rc) { delete f; return rc; }
  a.f64 = model->state_f64 ? 1 : 0;
  {
    // peer-memory polls give up after PZ_P2P_TIMEOUT_S seconds (default 120; 0 = never): the hosts of the ranks
    // drive their steps independently and may be skewed by seconds
    double secs = 120.0;
    if (const char* e = getenv("PZ_P2P_TIMEOUT_S")) secs = atof(e);
    a.wait_ticks = secs > 0 ? (unsigned long long)(secs * 1.9e9) : 0ull;
  }
  a.n_global = (int64_t)n_global;
  a.n_local = (int64_t)(n_global / (uint64_t)world);
  a.slot_base = (int64_t)rank * a.n_local;
  a.world = world;
  a.rank = rank;
  a.seed = seed;
  a.resampler = model->resampler;
  a.nblk = (a.n_local + kBlk - 1) / kBlk;
  a.nscan = (int32_t)((a.nblk + kScanTile - 1) / kScanTile);
  const int to = fused_tile_out(model->nx, a.f64);
  a.ntiles_out = (int32_t)((a.n_local + to - 1) / to);
  if (world > 1) {
    // halo margin exchanged with each neighbour every step: nblk/32 clamped to [64, 2048] blocks
    int64_t hb = a.nblk / 32;
    hb = hb < 64 ? 64 : (hb > 2048 ? 2048 : hb);
    if (const char* e = getenv("PZ_HALO_BLOCKS")) hb = atoll(e) > 0 ? atoll(e) : hb;
    f->hb = hb < a.nblk ? hb : a.nblk;
    a.margin = f->hb;
  }
  const int64_t H = f->hb * kBlk;
  a.ld = round_up(a.n_local, kPad) + 2 * H;
  const size_t esz = a.f64 ? sizeof(double) : sizeof(float);
  const size_t xbytes = (size_t)model->nx * a.ld * esz;
#define PZ_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int c_ = fail(PZ_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); free_filter(f); return c_; } } while (0)
  PZ_TRY(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  PZ_TRY(cudaEventCreate(&f->ev0));
  PZ_TRY(cudaEventCreate(&f->ev1));
  for (int i = 0; i < 2; ++i) {
    // every initialising memset is ordered on the filter's own (non-blocking) stream, ahead of pf_init
    PZ_TRY(cudaMalloc(&f->xbuf[i], xbytes));
    PZ_TRY(cudaMemsetAsync(f->xbuf[i], 0, xbytes, f->stream));
    a.X[i] = (char*)f->xbuf[i] + (size_t)H * esz;
    PZ_TRY(cudaMalloc(&f->qbuf[i], (size_t)a.ld * sizeof(uint16_t)));
    PZ_TRY(cudaMemsetAsync(f->qbuf[i], 0, (size_t)a.ld * sizeof(uint16_t), f->stream));
    a.Q[i] = f->qbuf[i] + H;
    PZ_TRY(cudaMalloc(&f->ebbuf[i], (a.nblk + 2 * f->hb + 8) * sizeof(int32_t)));
    a.eb[i] = f->ebbuf[i] + f->hb;
  }
  PZ_TRY(cudaMalloc(&a.sb, (a.nblk + 8) * sizeof(uint64_t)));
  PZ_TRY(cudaMalloc(&a.pb, (a.nblk + 8) * sizeof(uint64_t)));
  if (world > 1) {
    PZ_TRY(cudaMalloc(&f->pbxbuf, (a.nblk + 2 * f->hb + 8) * sizeof(uint64_t)));
    a.pbx = f->pbxbuf + f->hb;
    PZ_TRY(cudaMalloc(&a.shard_rec, (size_t)world * kShardRec * sizeof(uint64_t)));
    PZ_TRY(cudaMalloc(&f->warm, 2 * sizeof(int32_t)));
  } else {
    a.pbx = a.pb;
  }
  a.b_lo = rank > 0 ? -f->hb : 0;                       // fixed halo-extended block range of a sharded filter
  a.b_hi = a.nblk + (rank + 1 < world ? f->hb : 0);
  PZ_TRY(cudaMalloc(&f->xbbuf, (a.nblk + 2 * f->hb + 8) * sizeof(uint64_t)));
  a.xb = f->xbbuf + f->hb;
  PZ_TRY(cudaMalloc(&a.tile_start, (a.ntiles_out + 1) * sizeof(int32_t)));
  PZ_TRY(cudaMalloc(&a.rec, (size_t)(a.ntiles_out + 1) * kMaxRec * sizeof(double)));
  PZ_TRY(cudaMalloc(&a.scan_status, (a.nscan + 1) * sizeof(uint64_t)));
  PZ_TRY(cudaMemsetAsync(a.scan_status, 0, (a.nscan + 1) * sizeof(uint64_t), f->stream));
  PZ_TRY(cudaMalloc(&a.scan_part, (size_t)(a.nscan + 1) * kMaxRec * sizeof(double)));
  PZ_TRY(cudaMalloc(&a.plan, sizeof(PlanDev)));
  PZ_TRY(cudaMemsetAsync(a.plan, 0, sizeof(PlanDev), f->stream));
  PZ_TRY(cudaMalloc(&a.ctl, sizeof(CtlDev)));
  PZ_TRY(cudaMemsetAsync(a.ctl, 0, sizeof(CtlDev), f->stream));
  PZ_TRY(cudaMalloc(&f->d_obs, (size_t)kObsRing * PZ_MAX_NY * sizeof(float)));
  PZ_TRY(cudaMemsetAsync(f->d_obs, 0, (size_t)kObsRing * PZ_MAX_NY * sizeof(float), f->stream));
  a.obs = f->d_obs;
  PZ_TRY(cudaMalloc(&f->d_obs64, (size_t)kObsRing * PZ_MAX_NY * sizeof(double)));
  PZ_TRY(cudaMemsetAsync(f->d_obs64, 0, (size_t)kObsRing * PZ_MAX_NY * sizeof(double), f->stream));
  a.obs64 = f->d_obs64;
  PZ_TRY(cudaMalloc(&a.summ, (size_t)kObsRing * sizeof(pz_step_summary)));
  PZ_TRY(cudaMemsetAsync(a.summ, 0, (size_t)kObsRing * sizeof(pz_step_summary), f->stream));
  PZ_TRY(cudaMallocHost(&f->h_obs, (size_t)kObsRing * PZ_MAX_NY * sizeof(float)));
  PZ_TRY(cudaMallocHost(&f->h_obs64, (size_t)kObsRing * PZ_MAX_NY * sizeof(double)));
  PZ_TRY(cudaMallocHost(&f->h_summ, (size_t)kObsRing * sizeof(pz_step_summary)));
  launch_init(a, f->stream);
  if (shard_only) {
    PZ_TRY(cudaGetLastError());
    PZ_TRY(cudaStreamSynchronize(f->stream));
    *out = f;
    return PZ_OK;
  }
  if (world > 1) {
    ncclUniqueId id;
    memcpy(&id, nccl_id, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&f->comm, world, id, rank);
    if (r != ncclSuccess) { int c = fail(PZ_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); free_filter(f); return c; }
    // establish every peer-to-peer connection now (NCCL sets channels up lazily on first use, which
    // would otherwise land in the first step that happens to migrate particles to a new peer)
    {
      bool ok = g_nccl.GroupStart() == ncclSuccess;
      for (int peer = 0; peer < world && ok; ++peer) {
        if (peer == rank) continue;
        if (peer != rank - 1 && peer != rank + 1) continue;  // only neighbours ever exchange halos
        ok = g_nccl.Send(f->warm, 1, ncclInt32, peer, f->comm, f->stream) == ncclSuccess &&
             g_nccl.Recv(f->warm + 1, 1, ncclInt32, peer, f->comm, f->stream) == ncclSuccess;
      }
      ok = (g_nccl.GroupEnd() == ncclSuccess) && ok;
      if (!ok) { int c = fail(PZ_ENCCL, "NCCL peer warm-up failed"); free_filter(f); return c; }
    }
    rc = setup_p2p(f);
    if (rc) { free_filter(f); return rc; }
    int n = 0;
    rc = sharded_advance(f, 0, &n);
    if (rc) { free_filter(f); return rc; }
  } else {
    launch_scan_partition(a, to, f->stream);
  }
  PZ_TRY(cudaGetLastError());
  PZ_TRY(cudaStreamSynchronize(f->stream));
  if (world > 1 && a.p2p && getenv("PZ_SHARD_TIMING")) {
    f->shard_timing = true;
    for (auto& e : f->tev) PZ_TRY(cudaEventCreate(&e));
  }
  if (world > 1 && !f->shard_timing) capture_sharded_graphs(f);
  // one captured graph serves every systematic step (the step index lives in ctl->t)
  if (world == 1 && a.resampler == PZ_RESAMPLE_SYSTEMATIC) {
    cudaGraph_t graph = nullptr;
    PZ_TRY(cudaStreamBeginCapture(f->stream, cudaStreamCaptureModeThreadLocal));
    f->launches_per_step = launch_step(a, nullptr, f->stream);
    PZ_TRY(cudaStreamEndCapture(f->stream, &graph));
    PZ_TRY(cudaGraphInstantiate(&f->graph, graph, 0));
    cudaGraphDestroy(graph);
  }
#undef PZ_TRY
  *out = f;
  return PZ_OK;
}


// One captured graph per buffer parity: step kernel, exchange kernels (or NCCL calls, which are capturable); a failed
// capture leaves the eager stream-ordered path in place.
int capture_sharded_graphs(pz_filter* f) {
  if (getenv("PZ_NO_SHARD_GRAPH")) return PZ_OK;
  StepArgs& a = f->a;
  const uint32_t t_save = f->t;
  for (int par = 0; par < 2; ++par) {
    cudaGraph_t graph = nullptr;
    f->t = (uint32_t)par;  // only the parity matters for the buffers named in the captured calls
    bool ok = cudaStreamBeginCapture(f->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    int n = 0;
    if (ok) {
      n = launch_tile_only(a, f->stream);
      ok = sharded_advance(f, 1, &n) == PZ_OK;
      ok = (cudaStreamEndCapture(f->stream, &graph) == cudaSuccess) && ok && graph;
    }
    if (ok) ok = cudaGraphInstantiate(&f->sgraph[par], graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      for (int i = 0; i < 2; ++i) if (f->sgraph[i]) { cudaGraphExecDestroy(f->sgraph[i]); f->sgraph[i] = nullptr; }
      break;
    }
    f->launches_per_step = n;
  }
  f->t = t_save;
  return PZ_OK;
}

// ---- single-process multi-device filter (one host thread, the filter owns every device it is given) ----
// The shards are the same sharded filters a multi-process host creates, wired with direct peer pointers
// (cudaDeviceEnablePeerAccess) instead of CUDA IPC, and driven from one thread: every shard's step is enqueued
// on its own device's stream (the kernels of different devices wait for each other through the mailboxes),
// then all are awaited.
int create_multi(const pz_model_desc* model, uint64_t n_global, uint64_t seed, int n_devices, const int* device_ids, pz_filter** out) {
  if (!model || !out || !device_ids) return fail(PZ_EINVAL, "null argument");
  *out = nullptr;
  if (n_devices < 1 || n_devices > kMaxRanks) return fail(PZ_EINVAL, "n_devices must be in [1, %d]", kMaxRanks);
  if (n_devices == 1) return create_common(model, n_global, seed, device_ids[0], 0, 1, nullptr, out);
  // (test hook: PZ_MULTI_SAME_DEVICE=1 lets the shards of a multi-device TRACKER share a device -- its exchange is ordered by
  // events, not by spinning kernels -- so that the sharded path runs on a one-GPU box)
  const bool tracker = model->kind == PZ_MODEL_MTT || model->kind == PZ_MODEL_MULTICAM;
  const bool same_ok = tracker && getenv("PZ_MULTI_SAME_DEVICE") != nullptr;
  for (int i = 0; i < n_devices; ++i)
    for (int j = 0; j < i; ++j)
      if (device_ids[i] == device_ids[j] && !same_ok) return fail(PZ_EINVAL, "device %d listed twice", device_ids[i]);
  if (tracker) return create_mtt_multi(model, n_global, seed, n_devices, device_ids, out);
  if (model->kind != PZ_MODEL_RW1D && model->kind != PZ_MODEL_LINGAUSS)
    return fail(PZ_EINVAL, "multi-device filters: the linear-Gaussian families (RW1D / LINGAUSS) and the tracker (MTT)");
  if (model->delayed || model->carry_weights || model->ess_threshold > 0)
    return fail(PZ_EINVAL, "delayed-sampling / stored-log-weight filters run on one GPU");
  pz_filter* parent = new (std::nothrow) pz_filter();
  if (!parent) return fail(PZ_ENOMEM, "out of host memory");
  parent->desc = *model; parent->is_multi = true; parent->device = device_ids[0];
  memset(&parent->a, 0, sizeof(parent->a));
  const int R = n_devices;
  auto bail = [&](int code) { free_filter(parent); return code; };
  for (int r = 0; r < R; ++r) {
    pz_filter* sh = nullptr;
    int rc = create_common(model, n_global, seed, device_ids[r], r, R, nullptr, &sh, true);
    if (rc) return bail(rc);
    parent->shards.push_back(sh);
  }
  parent->a = parent->shards[0]->a;   // model, sizes (a.n_local = particles per shard)
  parent->a.world = R;
  // peer access between every pair (mailboxes are all-to-all)
  for (int r = 0; r < R; ++r) {
    cudaSetDevice(device_ids[r]);
    for (int q = 0; q < R; ++q) {
      if (q == r) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, device_ids[r], device_ids[q]);
      if (!can) return bail(fail(PZ_ECUDA, "device %d cannot access device %d (no peer path)", device_ids[r], device_ids[q]));
      cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[q], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return bail(fail(PZ_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", device_ids[r], device_ids[q], cudaGetErrorString(e)));
      cudaGetLastError();
    }
  }
#define PZ_TRYM(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(fail(PZ_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_))); } while (0)
  for (int r = 0; r < R; ++r) {
    pz_filter* sh = parent->shards[r];
    PZ_TRYM(cudaSetDevice(sh->device));
    PZ_TRYM(cudaMalloc(&sh->mbox, sizeof(Mailbox)));
    PZ_TRYM(cudaMemset(sh->mbox, 0, sizeof(Mailbox)));
    PZ_TRYM(cudaStreamCreateWithFlags(&sh->push_stream, cudaStreamNonBlocking));
    PZ_TRYM(cudaEventCreateWithFlags(&sh->ev_fork, cudaEventDisableTiming));
    PZ_TRYM(cudaEventCreateWithFlags(&sh->ev_join, cudaEventDisableTiming));
    PZ_TRYM(cudaDeviceSynchronize());
  }
  for (int r = 0; r < R; ++r) {
    pz_filter* sh = parent->shards[r];
    PZ_TRYM(cudaSetDevice(sh->device));
    ShardP2P h{};
    const size_t esz = sh->a.f64 ? 8 : 4;
    const int64_t H = sh->hb * kBlk;
    for (int q = 0; q < R; ++q) {
      pz_filter* o = parent->shards[q];
      h.mbox[q] = o->mbox;
      for (int b = 0; b < 2; ++b) { h.Xall[q][b] = o->a.X[b]; h.Qall[q][b] = o->a.Q[b]; h.eball[q][b] = o->a.eb[b]; }
      h.xball[q] = o->a.xb;
      h.pball[q] = o->a.pb;
    }
    for (int side = 0; side < 2; ++side) {
      const int q = side == 0 ? r - 1 : r + 1;
      if (q < 0 || q >= R) continue;
      pz_filter* nb = parent->shards[q];
      for (int b = 0; b < 2; ++b) {
        h.Xn[side][b] = (char*)nb->xbuf[b] + (size_t)H * esz;
        h.Qn[side][b] = nb->qbuf[b] + H;
        h.ebn[side][b] = nb->ebbuf[b] + nb->hb;
      }
      h.pbxn[side] = nb->pbxbuf + nb->hb;
      h.xbn[side] = nb->xbbuf + nb->hb;
    }
    PZ_TRYM(cudaMalloc(&sh->d_p2p, sizeof(ShardP2P)));
    PZ_TRYM(cudaMemcpy(sh->d_p2p, &h, sizeof(ShardP2P), cudaMemcpyHostToDevice));
    sh->a.p2p = sh->d_p2p;
  }
  // the plan of the prior population: every shard's exchange kernels are enqueued before any is awaited
  for (int r = 0; r < R; ++r) {
    pz_filter* sh = parent->shards[r];
    PZ_TRYM(cudaSetDevice(sh->device));
    int n = 0;
    int rc = sharded_advance(sh, 0, &n);
    if (rc) return bail(rc);
  }
  for (int r = 0; r < R; ++r) {
    pz_filter* sh = parent->shards[r];
    PZ_TRYM(cudaSetDevice(sh->device));
    PZ_TRYM(cudaStreamSynchronize(sh->stream));
  }
  for (int r = 0; r < R; ++r) {
    pz_filter* sh = parent->shards[r];
    PZ_TRYM(cudaSetDevice(sh->device));
    capture_sharded_graphs(sh);
  }
#undef PZ_TRYM
  *out = parent;
  return PZ_OK;
}

int report_step(pz_filter* f, const pz_step_summary& s, uint32_t t);

// one or k steps on every shard: enqueue everywhere, then await everywhere
int multi_advance(pz_filter* p, const double* obs, int64_t k, int32_t obs_dim, pz_step_summary* out, bool run) {
  const uint32_t t0 = p->t;
  for (pz_filter* sh : p->shards) {
    PZ_CUDA(cudaSetDevice(sh->device));
    sh->last_launches = 0;
    int rc = stage_obs(sh, obs, k, obs_dim);
    if (rc) return rc;
    PZ_CUDA(cudaEventRecord(sh->ev0, sh->stream));
    for (int64_t i = 0; i < k; ++i) {
      rc = enqueue_step(sh);
      if (rc) return rc;
    }
    PZ_CUDA(cudaEventRecord(sh->ev1, sh->stream));
    rc = fetch_summaries(sh, t0, k);
    if (rc) return rc;
  }
  double ms_max = 0;
  int64_t launches = 0;
  for (pz_filter* sh : p->shards) {
    PZ_CUDA(cudaSetDevice(sh->device));
    PZ_CUDA(cudaStreamSynchronize(sh->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, sh->ev0, sh->ev1);
    ms_max = ms > ms_max ? ms : ms_max;
    launches += sh->last_launches;
  }
  p->t += (uint32_t)k;
  p->last_ms = run ? p->last_ms + ms_max : ms_max;
  p->last_launches = run ? p->last_launches + launches : launches;
  int status = PZ_OK;
  for (int64_t i = 0; i < k; ++i) {
    pz_step_summary s = p->shards[0]->h_summ[(t0 + i) % kObsRing];   // every shard derives the same posterior summary
    int64_t mig = 0;
    for (pz_filter* sh : p->shards) {
      const int64_t m = sh->h_summ[(t0 + i) % kObsRing].n_migrated;
      if (m < 0) { mig = m; break; }   // an error verdict (the same on every shard)
      mig += m;
    }
    s.n_migrated = mig;
    if (out) out[i] = s;
    if (status == PZ_OK) status = report_step(p->shards[0], s, t0 + (uint32_t)i);
  }
  return status;
}

}  // namespace

extern "C" {

int pz_abi_version(void) { return PZ_ABI_VERSION; }
const char* pz_last_error(void) { return g_err; }

static void fill_common(pz_model_desc* d, int kind, int nx, int ny) {
  memset(d, 0, sizeof(*d));
  d->kind = kind; d->nx = nx; d->ny = ny;
  d->scalar_ll_2x = 1;  // reference behaviour (Distributions.hs:160-161)
  d->resampler = PZ_RESAMPLE_SYSTEMATIC;
}

int pz_model_rw1d(pz_model_desc* d, double q_var, double r_var) {
  if (!d || !(q_var >= 0) || !(r_var > 0)) return fail(PZ_EINVAL, "bad rw1d arguments");
  fill_common(d, PZ_MODEL_RW1D, 1, 1);
  d->F[0] = 1.0; d->Q[0] = q_var; d->H[0] = 1.0; d->R[0] = r_var;
  return PZ_OK;
}

int pz_model_stt(pz_model_desc* d) {
  if (!d) return fail(PZ_EINVAL, "null descriptor");
  fill_common(d, PZ_MODEL_LINGAUSS, 6, 3);
  const double tdiff = 1.0;  // SingleTargetTracking.hs:42
  for (int i = 0; i < 6; ++i) d->F[i * 6 + i] = 1.0;
  for (int i = 0; i < 3; ++i) {
    d->F[i * 6 + 3 + i] = tdiff;             // motionMatrix :48-52
    d->Q[i * 6 + i] = tdiff * 0.01;          // posCov :40
    d->Q[(3 + i) * 6 + 3 + i] = tdiff * 0.1; // velCov :41
    d->H[i * 6 + i] = 1.0;                   // posFromPosVel :53-54
    d->R[i * 3 + i] = 10.0;                  // :35
  }
  return PZ_OK;
}

int pz_model_mtt_track(pz_model_desc* d) {
  if (!d) return fail(PZ_EINVAL, "null descriptor");
  fill_common(d, PZ_MODEL_LINGAUSS, 4, 2);
  const double tdiff = 1.0;  // MultiTargetTracking.hs:37
  for (int i = 0; i < 4; ++i) d->F[i * 4 + i] = 1.0;
  for (int i = 0; i < 2; ++i) {
    d->F[i * 4 + 2 + i] += tdiff * 1.0;            // motionMatrix :59-63
    d->F[(2 + i) * 4 + i] += tdiff * (-1.0 / 100);
    d->F[(2 + i) * 4 + 2 + i] += tdiff * (-0.1);
    d->Q[i * 4 + i] = tdiff * 0.01;                // motionCov :64-67
    d->Q[(2 + i) * 4 + 2 + i] = tdiff * 0.1;
    d->H[i * 4 + i] = 1.0;                         // posFromPosVel :72-73
    d->R[i * 2 + i] = 1.0;                         // :70
  }
  return PZ_OK;
}

int pz_model_mtt(pz_model_desc* d) {
  int rc = pz_model_mtt_track(d);
  if (rc) return rc;
  d->kind = PZ_MODEL_MTT;
  d->mtt.clutter_lambda = 3.0;       // MultiTargetTracking.hs:40
  d->mtt.new_track_lambda = 0.1;     // :41
  d->mtt.pd = 0.8;                   // :42
  d->mtt.survival_prob = exp(-0.02); // :43
  d->mtt.clutter_var = 10.0;         // :46
  d->mtt.birth_pos_var = 5.0;        // :52-54
  d->mtt.birth_vel_var = 0.1;
  d->mtt.meas_var = 1.0;
  d->mtt.k_max = 16;
  d->mtt.m_max = 32;
  return PZ_OK;
}

int pz_model_multicam(pz_model_desc* d) {
  if (!d) return fail(PZ_EINVAL, "null descriptor");
  fill_common(d, PZ_MODEL_MULTICAM, 4, 0);           // ny = 0: the box is observed directly, box ~ N(posWH, meas_var I) (:130)
  const double tdiff = 1.0;                          // MultiCamera.hs:77-78
  for (int i = 0; i < 4; ++i) {
    d->F[i * 4 + i] = 1.0;                           // motionMatrix = eye (:113)
    d->Q[i * 4 + i] = tdiff * (i < 2 ? 1.0 : 0.1);   // motionCov (:114-116)
  }
  d->x0[0] = d->x0[1] = 0.0; d->x0[2] = d->x0[3] = 1.0;   // mean of the box prior and of the clutter boxes (:91, :187)
  d->mtt.clutter_lambda = 1.0;        // amtClutter, per camera (:184-185)
  d->mtt.new_track_lambda = 0.1;      // birthRate * tdiff (:195, :206)
  d->mtt.pd = 1.0;                    // every live track is measured (:140-146)
  d->mtt.survival_prob = exp(-tdiff * 0.02);   // :120, :126
  d->mtt.clutter_var = 1.0;           // clutter box covariance I (:187)
  d->mtt.birth_pos_var = 1.0;         // box prior covariance diag(1, 1, 0.2, 0.2) (:93)
  d->mtt.birth_vel_var = 0.2;
  d->mtt.meas_var = 1.0;
  d->mtt.k_max = 16;
  d->mtt.m_max = 32;
  return PZ_OK;
}

int pz_model_walker(pz_model_desc* d) {
  if (!d) return fail(PZ_EINVAL, "null descriptor");
  fill_common(d, PZ_MODEL_WALKER, 5, 2);
  d->carry_weights = 1;                 // runOurFilter = particles 1000 (Walker.hs:129-131; Inference.hs:85-86)
  pz_walker_params& w = d->walker;
  w.pos_noise = 0.01; w.obs_std = 10.0; w.dt = 10.0;
  w.p_init[0] = 0.7; w.p_init[1] = 0.25; w.p_init[2] = 0.05;
  // motionTypeTransition (Walker.hs:45-49): half lives in seconds, [from][to]
  w.half_life[0][1] = 30.0; w.half_life[0][2] = 300.0;
  w.half_life[1][1] = 1.0; w.half_life[1][0] = 10.0; w.half_life[1][2] = 60.0;
  w.half_life[2][2] = 0.5; w.half_life[2][1] = 2.0;
  w.speed_lo[1] = 0.0; w.speed_hi[1] = 2.0; w.speed_lo[2] = 2.0; w.speed_hi[2] = 7.0;
  w.exp_mean_as_written = 1;
  return PZ_OK;
}

int pz_model_beta_bernoulli(pz_model_desc* d, double a, double b) {
  if (!d || !(a > 0) || !(b > 0)) return fail(PZ_EINVAL, "bad Beta parameters");
  fill_common(d, PZ_MODEL_BETA_BERNOULLI, 1, 1);
  d->delayed = 1; d->beta_a = a; d->beta_b = b;
  return PZ_OK;
}

int pz_filter_create(const pz_model_desc* model, uint64_t n_particles, uint64_t seed, int device, pz_filter** out) {
  return create_common(model, n_particles, seed, device, 0, 1, nullptr, out);
}

int pz_filter_create_multi(const pz_model_desc* model, uint64_t n_particles, uint64_t seed, int n_devices, const int* device_ids,
                           pz_filter** out) {
  return create_multi(model, n_particles, seed, n_devices, device_ids, out);
}

int pz_comm_unique_id(void* out128) {
  if (!out128) return fail(PZ_EINVAL, "null argument");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  PZ_NCCL(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, sizeof(id));
  return PZ_OK;
}

int pz_filter_create_sharded(const pz_model_desc* model, uint64_t n_global, uint64_t seed, int device, int rank,
                             int world, const void* nccl_unique_id, pz_filter** out) {
  return create_common(model, n_global, seed, device, rank, world, nccl_unique_id, out);
}

int64_t pz_filter_n_local(const pz_filter* f) {
  if (!f) return 0;
  if (f->is_multi && f->is_mtt) return f->a.n_global;
  return f->is_multi ? f->a.n_local * (int64_t)f->shards.size() : f->a.n_local;   // a multi-device handle holds the whole population
}

int pz_filter_step(pz_filter* f, const double* obs, int32_t n_obs, int32_t obs_dim, pz_step_summary* out) {
  if (!f || (!obs && n_obs > 0)) return fail(PZ_EINVAL, "null argument");
  if (f->is_mtt) {
    const int ow = f->tracker_obs_w;
    if (obs_dim != ow) return fail(PZ_EINVAL, "tracker observations are rows of %d doubles", ow);
    if (f->is_multi) return mtt_multi_step_host(f, obs, n_obs, out);
    PZ_CUDA(cudaSetDevice(f->device));
    return mtt_step_host(f, obs, n_obs, out);
  }
  if (n_obs != 1 || obs_dim != f->a.model.ny) return fail(PZ_EINVAL, "expected 1 observation of dim %d, got %d x %d", f->a.model.ny, n_obs, obs_dim);
  if (f->is_multi) return multi_advance(f, obs, 1, obs_dim, out, false);
  PZ_CUDA(cudaSetDevice(f->device));
  if (f->is_beta) { f->last_ms = 0; f->last_launches = 0; return beta_step_host(f, obs, out); }
  if (f->is_rb) { f->last_ms = 0; f->last_launches = 0; return rb_step_host(f, obs, out); }
  if (f->is_gen) return gen_step_host(f, obs, obs_dim, out);
  const uint32_t t0 = f->t;
  f->last_launches = 0;
  int rc = stage_obs(f, obs, 1, obs_dim);
  if (rc) return rc;
  PZ_CUDA(cudaEventRecord(f->ev0, f->stream));
  rc = enqueue_step(f);
  if (rc) return rc;
  PZ_CUDA(cudaEventRecord(f->ev1, f->stream));
  rc = fetch_summaries(f, t0, 1);
  if (rc) return rc;
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, f->ev0, f->ev1);
  f->last_ms = ms;
  pz_step_summary& s = f->h_summ[t0 % kObsRing];
  if (out) *out = s;
  return report_step(f, s, t0);
}

int pz_filter_step_profile(pz_filter* f, const double* obs, int32_t n_obs, int32_t obs_dim, double ms[3]) {
  if (!f || !obs || !ms) return fail(PZ_EINVAL, "null argument");
  if (n_obs != 1 || obs_dim != f->a.model.ny) return fail(PZ_EINVAL, "expected 1 observation of dim %d", f->a.model.ny);
  if (f->a.resampler != PZ_RESAMPLE_SYSTEMATIC || f->a.world > 1 || f->is_mtt || f->is_rb || f->is_gen || f->is_multi)
    return fail(PZ_ESTATE, "profiled step: single-GPU systematic filters only");
  PZ_CUDA(cudaSetDevice(f->device));
  cudaEvent_t ev[4] = {};
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ev[i]);
  int rc = e == cudaSuccess ? stage_obs(f, obs, 1, obs_dim) : fail(PZ_ECUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e));
  if (rc == PZ_OK) {
    f->last_launches = launch_step_profiled(f->a, f->stream, ev);
    f->t += 1;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(f->stream);
    if (e != cudaSuccess) rc = fail(PZ_ECUDA, "profiled step failed: %s", cudaGetErrorString(e));
  }
  for (int i = 0; i < 3 && rc == PZ_OK; ++i) {
    float x = 0;
    cudaEventElapsedTime(&x, ev[i], ev[i + 1]);
    ms[i] = x;
  }
  for (int i = 0; i < 4; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
  return rc;
}

int pz_filter_run(pz_filter* f, const double* obs, int64_t n_steps, int32_t obs_dim, pz_step_summary* out) {
  if (!f || !obs || n_steps < 0) return fail(PZ_EINVAL, "bad argument");
  if (f->is_mtt) return fail(PZ_ESTATE, "pz_filter_run: the tracker takes a variable number of observations per step; use pz_filter_step");
  if (obs_dim != f->a.model.ny) return fail(PZ_EINVAL, "expected observations of dim %d", f->a.model.ny);
  f->last_launches = 0;
  f->last_ms = 0;
  if (f->is_multi) {
    int64_t done = 0;
    int status = PZ_OK;
    while (done < n_steps) {
      const int64_t k = (n_steps - done) < (kObsRing - 1) ? (n_steps - done) : (kObsRing - 1);
      int rc = multi_advance(f, obs + done * obs_dim, k, obs_dim, out ? out + done : nullptr, true);
      if (rc && status == PZ_OK) status = rc;
      if (rc == PZ_ECUDA || rc == PZ_EINVAL) return rc;
      done += k;
    }
    return status;
  }
  PZ_CUDA(cudaSetDevice(f->device));
  if (f->is_rb || f->is_gen) {
    double ms = 0; int64_t nl = 0;
    for (int64_t i = 0; i < n_steps; ++i) {
      int rc = f->is_beta ? beta_step_host(f, obs + i * obs_dim, out ? out + i : nullptr)
               : f->is_rb ? rb_step_host(f, obs + i * obs_dim, out ? out + i : nullptr)
                          : gen_step_host(f, obs + i * obs_dim, obs_dim, out ? out + i : nullptr);
      if (rc) return rc;
      if (f->is_gen) { ms += f->last_ms; nl += f->last_launches; }
    }
    if (f->is_gen) { f->last_ms = ms; f->last_launches = nl; }
    return PZ_OK;
  }
  int64_t done = 0;
  int status = PZ_OK;
  while (done < n_steps) {
    const int64_t k = (n_steps - done) < (kObsRing - 1) ? (n_steps - done) : (kObsRing - 1);
    const uint32_t t0 = f->t;
    int rc = stage_obs(f, obs + done * obs_dim, k, obs_dim);
    if (rc) return rc;
    PZ_CUDA(cudaEventRecord(f->ev0, f->stream));
    for (int64_t i = 0; i < k; ++i) {
      rc = enqueue_step(f);
      if (rc) return rc;
    }
    PZ_CUDA(cudaEventRecord(f->ev1, f->stream));
    rc = fetch_summaries(f, t0, k);
    if (rc) return rc;
    PZ_CUDA(cudaStreamSynchronize(f->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, f->ev0, f->ev1);
    f->last_ms += ms;
    for (int64_t i = 0; i < k; ++i) {
      const pz_step_summary& s = f->h_summ[(t0 + i) % kObsRing];
      if (out) out[done + i] = s;
      if (status == PZ_OK) status = report_step(f, s, t0 + (uint32_t)i);
    }
    done += k;
  }
  return status;
}

int pz_filter_last_timing(pz_filter* f, double* ms, int64_t* launches) {
  if (!f) return fail(PZ_EINVAL, "null filter");
  if (ms) *ms = f->last_ms;
  if (launches) *launches = f->last_launches;
  return PZ_OK;
}

static int read_state_common(pz_filter* f, float* out32, double* out64, int64_t out_len) {
  if (!f || (!out32 && !out64)) return fail(PZ_EINVAL, "null argument");
  if (f->is_mtt) return fail(PZ_ESTATE, "tracker state is read with pz_filter_read_tracks / pz_filter_read_mcam_tracks");
  if (f->is_multi) {   // shard after shard: out[d][r * n_shard + i]
    const int R = (int)f->shards.size(), nx = f->a.model.nx;
    const int64_t ns = f->a.n_local, n = ns * R;
    if (out_len < nx * n) return fail(PZ_EINVAL, "buffer too small: need %lld scalars", (long long)(nx * n));
    std::vector<float> t32(out32 ? (size_t)nx * ns : 0);
    std::vector<double> t64(out64 ? (size_t)nx * ns : 0);
    for (int r = 0; r < R; ++r) {
      int rc = read_state_common(f->shards[r], out32 ? t32.data() : nullptr, out64 ? t64.data() : nullptr, nx * ns);
      if (rc) return rc;
      for (int d = 0; d < nx; ++d) {
        if (out32) memcpy(out32 + (size_t)d * n + (size_t)r * ns, t32.data() + (size_t)d * ns, ns * sizeof(float));
        if (out64) memcpy(out64 + (size_t)d * n + (size_t)r * ns, t64.data() + (size_t)d * ns, ns * sizeof(double));
      }
    }
    return PZ_OK;
  }
  const StepArgs& a = f->a;
  if (f->is_mtt) return fail(PZ_ESTATE, "tracker state is read with pz_filter_read_tracks");
  if (f->is_rb) return fail(PZ_ESTATE, "a delayed-sampling filter holds no sampled state; use pz_filter_read_posterior");
  if (out_len < a.model.nx * a.n_local) return fail(PZ_EINVAL, "buffer too small: need %lld scalars", (long long)(a.model.nx * a.n_local));
  if (out32 && a.f64) return fail(PZ_ESTATE, "the filter stores its state in double (state_f64 = 1): use pz_filter_read_state_f64");
  PZ_CUDA(cudaSetDevice(f->device));
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  const size_t es = a.f64 ? 8 : 4;
  const char* x = reinterpret_cast<const char*>(a.X[f->t & 1]);
  if (out64 && !a.f64) {  // widen on the host
    std::vector<float> tmp((size_t)a.n_local);
    for (int d = 0; d < a.model.nx; ++d) {
      PZ_CUDA(cudaMemcpy(tmp.data(), x + (size_t)d * a.ld * es, a.n_local * es, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < a.n_local; ++i) out64[(size_t)d * a.n_local + i] = (double)tmp[i];
    }
    return PZ_OK;
  }
  char* out = out64 ? reinterpret_cast<char*>(out64) : reinterpret_cast<char*>(out32);
  for (int d = 0; d < a.model.nx; ++d)
    PZ_CUDA(cudaMemcpy(out + (size_t)d * a.n_local * es, x + (size_t)d * a.ld * es, a.n_local * es, cudaMemcpyDeviceToHost));
  return PZ_OK;
}

int pz_filter_read_state(pz_filter* f, float* out, int64_t out_len) { return read_state_common(f, out, nullptr, out_len); }
int pz_filter_read_state_f64(pz_filter* f, double* out, int64_t out_len) { return read_state_common(f, nullptr, out, out_len); }

int pz_filter_read_posterior(pz_filter* f, double* mean_out, double* cov_out) {
  if (!f || !mean_out || !cov_out) return fail(PZ_EINVAL, "null argument");
  if (!f->is_rb) return fail(PZ_ESTATE, "not a delayed-sampling filter (model.delayed = 0)");
  PZ_CUDA(cudaSetDevice(f->device));
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  if (f->is_beta) {   // MBeta alpha beta (DelayedSampling.hs:63,72)
    BetaDev h;
    PZ_CUDA(cudaMemcpy(&h, f->beta, sizeof(BetaDev), cudaMemcpyDeviceToHost));
    const double s = h.a + h.b;
    mean_out[0] = h.a; mean_out[1] = h.b; cov_out[0] = h.a * h.b / (s * s * (s + 1.0));
    return PZ_OK;
  }
  RbDev h;
  PZ_CUDA(cudaMemcpy(&h, f->rb, sizeof(RbDev), cudaMemcpyDeviceToHost));
  const int nx = h.nx;
  for (int i = 0; i < nx; ++i) mean_out[i] = h.mean[i];
  for (int i = 0; i < nx * nx; ++i) cov_out[i] = h.cov[i];
  return PZ_OK;
}

static int read_logw_common(pz_filter* f, float* out32, double* out64, int64_t n) {
  if (!f || (!out32 && !out64)) return fail(PZ_EINVAL, "null argument");
  if (f->is_multi) return fail(PZ_ESTATE, "log-weight taps are per shard; not available on a multi-device handle");
  if (f->is_rb) return fail(PZ_ESTATE, "a delayed-sampling filter has equal weights and no sampled state");
  if (n < f->a.n_local) return fail(PZ_EINVAL, "buffer too small");
  PZ_CUDA(cudaSetDevice(f->device));
  int rc = generic_alloc(&f->g, f->a.n_local, false, f->stream);
  if (rc) return rc;
  if (out64) {
    if (!f->g.lw64) PZ_CUDA(cudaMalloc(&f->g.lw64, (size_t)round_up(f->a.n_local, kPad) * sizeof(double)));
    if (f->is_mtt || f->is_gen) {  // stored fp32 log-weights: widen on the host
      std::vector<float> tmp((size_t)f->a.n_local);
      PZ_CUDA(cudaMemcpyAsync(tmp.data(), f->is_gen ? f->gen_lw[f->t & 1] : f->g.lw, f->a.n_local * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
      PZ_CUDA(cudaStreamSynchronize(f->stream));
      for (int64_t i = 0; i < f->a.n_local; ++i) out64[i] = (double)tmp[i];
      return PZ_OK;
    }
    launch_logw(f->a, nullptr, f->g.lw64, f->stream);
    PZ_CUDA(cudaMemcpyAsync(out64, f->g.lw64, f->a.n_local * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  } else {
    if (!f->is_mtt && !f->is_gen) launch_logw(f->a, f->g.lw, nullptr, f->stream);
    PZ_CUDA(cudaMemcpyAsync(out32, f->is_gen ? f->gen_lw[f->t & 1] : f->g.lw, f->a.n_local * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
  }
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  return PZ_OK;
}

int pz_filter_read_logw(pz_filter* f, float* out, int64_t n) { return read_logw_common(f, out, nullptr, n); }
int pz_filter_read_logw_f64(pz_filter* f, double* out, int64_t n) { return read_logw_common(f, nullptr, out, n); }

int pz_filter_read_ancestors(pz_filter* f, int32_t* out, int64_t n) {
  if (!f || !out) return fail(PZ_EINVAL, "null argument");
  if (f->is_multi) return fail(PZ_ESTATE, "ancestor taps are per device; not available on a multi-device handle");
  if (f->is_rb) return fail(PZ_ESTATE, "a delayed-sampling filter has equal weights and no sampled state");
  if (n < f->a.n_local) return fail(PZ_EINVAL, "buffer too small");
  PZ_CUDA(cudaSetDevice(f->device));
  int rc = pending_ancestors(f, nullptr);
  if (rc) return rc;
  PZ_CUDA(cudaMemcpyAsync(out, f->g.anc, f->a.n_local * sizeof(int32_t), cudaMemcpyDeviceToHost, f->stream));
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  return PZ_OK;
}

int pz_filter_read_particles(pz_filter* f, int32_t stride, double* out, int64_t out_len) {
  if (!f || !out || stride < 1) return fail(PZ_EINVAL, "bad argument");
  if (f->is_mtt) return fail(PZ_ESTATE, "tracker particles are read with pz_filter_read_tracks");
  const StepArgs& a = f->a;
  const int64_t n_sel = (a.n_local + stride - 1) / stride;
  if (out_len < n_sel * a.model.nx) return fail(PZ_EINVAL, "buffer too small: need %lld doubles", (long long)(n_sel * a.model.nx));
  PZ_CUDA(cudaSetDevice(f->device));
  if (f->is_rb) {  // every particle is the marginal's mean (the reference emits the marginal itself per particle)
    double mean[PZ_MAX_NX], cov[PZ_MAX_NX * PZ_MAX_NX];
    int rc = pz_filter_read_posterior(f, mean, cov);
    if (rc) return rc;
    if (f->is_beta) mean[0] = mean[0] / (mean[0] + mean[1]);
    for (int64_t i = 0; i < n_sel; ++i)
      for (int d = 0; d < a.model.nx; ++d) out[i * a.model.nx + d] = mean[d];
    return PZ_OK;
  }
  int rc = pending_ancestors(f, nullptr);
  if (rc) return rc;
  // device scratch kept with the filter (the visualiser reads every step)
  if (f->g.thin_cap < n_sel * a.model.nx) {
    cudaFree(f->g.thin); f->g.thin = nullptr; f->g.thin_cap = 0;
    PZ_CUDA(cudaMalloc(&f->g.thin, (size_t)n_sel * a.model.nx * sizeof(double)));
    f->g.thin_cap = n_sel * a.model.nx;
  }
  launch_gather_thin(a, f->g.anc, stride, n_sel, f->g.thin, f->stream);
  PZ_CUDA(cudaMemcpyAsync(out, f->g.thin, (size_t)n_sel * a.model.nx * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  return PZ_OK;
}

// thinned resampled population of a tracker: the raw records (W words each) of every stride-th output slot's parent
static int tracker_gather(pz_filter* f, int32_t stride, int64_t n_sel_cap, std::vector<uint32_t>* h, int64_t* n_sel_out, int* W_out) {
  const StepArgs& a = f->a;
  const int64_t n_sel = (a.n_local + stride - 1) / stride;
  if (n_sel_cap < n_sel) return fail(PZ_EINVAL, "buffers too small: need %lld particles", (long long)n_sel);
  pz_filter* parent = f;
  if (f->is_multi) f = f->shards[0];   // every device holds the global plan: device 0 derives all ancestors and gathers from the owners
  PZ_CUDA(cudaSetDevice(f->device));
  int rc = PZ_OK;
  if (parent->is_multi) {
    GenericArgs g = mtt_generic(f, (int)(parent->t & 1u));
    launch_generic_ancestors(g, f->stream);
  } else rc = pending_ancestors(f, nullptr);
  if (rc) return rc;
  const int W = f->tracker_words;
  if (f->tracker_thin_cap < n_sel * W) {   // the visualiser calls this every frame: no allocation per call
    cudaFree(f->tracker_thin); f->tracker_thin = nullptr; f->tracker_thin_cap = 0;
    PZ_CUDA(cudaMalloc(&f->tracker_thin, (size_t)n_sel * W * sizeof(uint32_t)));
    f->tracker_thin_cap = n_sel * W;
  }
  uint32_t* d_out = f->tracker_thin;
  if (parent->is_multi) launch_mtt_gather_multi(mtt_shard_view(parent, f, (int)(parent->t & 1u)), f->g.anc, stride, n_sel, d_out, f->stream, W);
  else launch_mtt_gather(f->mtt_s[f->t & 1], f->g.anc, stride, n_sel, d_out, f->stream, W);
  h->resize((size_t)n_sel * W);
  cudaError_t e = cudaMemcpyAsync(h->data(), d_out, h->size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(f->stream);
  if (e != cudaSuccess) return fail(PZ_ECUDA, "read_tracks: %s", cudaGetErrorString(e));
  *n_sel_out = n_sel; *W_out = W;
  return PZ_OK;
}

int pz_filter_read_tracks(pz_filter* f, int32_t stride, int32_t* count_out, int32_t* id_out, double* mean_out, double* cov_out,
                          int64_t n_sel_cap) {
  if (!f || !count_out || !id_out || !mean_out || !cov_out || stride < 1) return fail(PZ_EINVAL, "bad argument");
  if (!f->is_mtt) return fail(PZ_ESTATE, "not a tracker filter");
  if (f->is_mcam) return fail(PZ_ESTATE, "MultiCamera tracks are read with pz_filter_read_mcam_tracks");
  std::vector<uint32_t> h;
  int64_t n_sel = 0; int W = 0;
  int rc = tracker_gather(f, stride, n_sel_cap, &h, &n_sel, &W);
  if (rc) return rc;
  static const int tri[4][4] = {{0, 1, 2, 3}, {1, 4, 5, 6}, {2, 5, 7, 8}, {3, 6, 8, 9}};
  for (int64_t i = 0; i < n_sel; ++i) {
    const uint32_t* p = h.data() + (size_t)i * W;
    const int cnt = (int)p[0];
    count_out[i] = cnt;
    for (int k = 0; k < 16; ++k) {
      const uint32_t* r = p + 4 + 16 * k;
      float fl[16];
      memcpy(fl, r, sizeof(fl));
      id_out[i * 16 + k] = k < cnt ? (int32_t)r[0] : -1;
      for (int c = 0; c < 4; ++c) mean_out[(i * 16 + k) * 4 + c] = k < cnt ? (double)fl[1 + c] : 0.0;
      for (int x = 0; x < 4; ++x) for (int y = 0; y < 4; ++y) cov_out[(i * 16 + k) * 16 + x * 4 + y] = k < cnt ? (double)fl[5 + tri[x][y]] : 0.0;
    }
  }
  return PZ_OK;
}

int pz_filter_read_mcam_tracks(pz_filter* f, int32_t stride, int32_t* count_out, int32_t* id_out, int32_t* identity_out,
                               int32_t* camera_out, double* start_out, double* mean_out, double* var_out, int32_t* n_app_out,
                               double* app_mean_out, double* app_cov_out, int64_t n_sel_cap) {
  if (!f || !count_out || !id_out || !identity_out || !camera_out || !mean_out || !var_out || stride < 1) return fail(PZ_EINVAL, "bad argument");
  if (!f->is_mtt || !f->is_mcam) return fail(PZ_ESTATE, "not a MultiCamera filter");
  std::vector<uint32_t> h;
  int64_t n_sel = 0; int W = 0;
  int rc = tracker_gather(f, stride, n_sel_cap, &h, &n_sel, &W);
  if (rc) return rc;
  const int K = 16, A = 16, D = 10, apps0 = 4 + K * 12;
  for (int64_t i = 0; i < n_sel; ++i) {
    const uint32_t* p = h.data() + (size_t)i * W;
    const int cnt = (int)p[0], na = (int)p[2];
    count_out[i] = cnt;
    if (n_app_out) n_app_out[i] = na;
    for (int k = 0; k < K; ++k) {
      float fl[12];
      memcpy(fl, p + 4 + 12 * k, sizeof(fl));
      const uint32_t* r = p + 4 + 12 * k;
      id_out[i * K + k] = k < cnt ? (int32_t)r[0] : -1;
      identity_out[i * K + k] = k < cnt ? (int32_t)r[1] : -1;
      camera_out[i * K + k] = k < cnt ? (int32_t)r[2] : -1;
      if (start_out) start_out[i * K + k] = k < cnt ? (double)fl[3] : 0.0;
      for (int c = 0; c < 4; ++c) { mean_out[(i * K + k) * 4 + c] = k < cnt ? (double)fl[4 + c] : 0.0; var_out[(i * K + k) * 4 + c] = k < cnt ? (double)fl[8 + c] : 0.0; }
    }
    if (app_mean_out && app_cov_out)
      for (int ap = 0; ap < A; ++ap) {
        float fl[66];
        memcpy(fl, p + apps0 + 66 * ap, sizeof(fl));
        for (int x = 0; x < D; ++x) {
          app_mean_out[(i * A + ap) * D + x] = ap < na ? (double)fl[x] : 0.0;
          for (int y = 0; y < D; ++y) {
            const int lo = x <= y ? x : y, hi = x <= y ? y : x;
            app_cov_out[((i * A + ap) * D + x) * D + y] = ap < na ? (double)fl[D + lo * D - (lo * (lo - 1)) / 2 + (hi - lo)] : 0.0;
          }
        }
      }
  }
  return PZ_OK;
}

void pz_filter_destroy(pz_filter* f) { free_filter(f); }

// ---- standalone resamplers on stored log-weights --------------------------------------------
static int standalone(const double* logw, int64_t n, bool multinomial, uint32_t U, uint64_t seed, uint32_t step,
                      int32_t* anc_out, int64_t n_draws = 0) {
  if (!logw || !anc_out || n < 1 || n > (1ll << 30) || n_draws < 0 || n_draws > n) return fail(PZ_EINVAL, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PZ_ECUDA, "no CUDA device (this library has no CPU fallback)");
  GenericScratch gs;
  int rc = generic_alloc(&gs, n, true, nullptr);
  if (rc) { gs.release(); return rc; }
  std::vector<float> lw32((size_t)n);
  for (int64_t i = 0; i < n; ++i) lw32[i] = (float)logw[i];
  cudaStream_t s = nullptr;
  GenericArgs g{};
  g.lw = gs.lw; g.q16 = gs.q16; g.q16_alt = gs.q16; g.xb = gs.xb; g.anc = gs.anc; g.eb = gs.eb; g.eb_alt = gs.eb;
  g.parity_ctl = nullptr; g.sb = gs.sb; g.pb = gs.pb;
  g.tile_start = gs.tile_start; g.scan_status = gs.scan_status; g.scan_part = gs.scan_part; g.plan = gs.plan; g.ctl = gs.ctl;
  g.n = n; g.nblk = (n + kBlk - 1) / kBlk;
  g.ntiles_out = (int32_t)((n + kGenericTileOut - 1) / kGenericTileOut);
  g.nscan = (int32_t)((g.nblk + kScanTile - 1) / kScanTile);
  g.U = U; g.seed = seed; g.step = step; g.n_draws = n_draws;
  PlanDev plan;
  cudaError_t e = cudaMemcpy(gs.lw, lw32.data(), n * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    launch_ctl_reset(gs.ctl, s);
    launch_generic_plan(g, true, s);
    if (multinomial) launch_generic_multinomial(g, s);
    else launch_generic_ancestors(g, s);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(&plan, gs.plan, sizeof(plan), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && !plan.degenerate) e = cudaMemcpy(anc_out, gs.anc, (n_draws > 0 ? n_draws : n) * sizeof(int32_t), cudaMemcpyDeviceToHost);
  gs.release();
  if (e != cudaSuccess) return fail(PZ_ECUDA, "resample: %s", cudaGetErrorString(e));
  if (plan.degenerate) return fail(PZ_EDEGENERATE, "every log-weight is -inf or NaN");
  return PZ_OK;
}

int pz_resample_systematic(const double* logw, int64_t n, double u, int32_t* anc_out) {
  if (!(u >= 0.0 && u < 1.0)) return fail(PZ_EINVAL, "u must be in [0, 1)");
  return standalone(logw, n, false, (uint32_t)(u * 4294967296.0), 0, 0, anc_out);
}

int pz_resample_multinomial(const double* logw, int64_t n, uint64_t seed, uint32_t step, int32_t* anc_out) {
  return standalone(logw, n, true, 0, seed, step, anc_out);
}

int pz_importance_resampling(const double* logw, int64_t n, uint64_t seed, uint32_t step, int64_t k, int32_t* idx_out) {
  if (k < 1) return fail(PZ_EINVAL, "k must be at least 1");
  return standalone(logw, n, true, 0, seed, step, idx_out, k);
}

static int dist_tap(bool sample, int32_t kind, const double* params, int32_t np, uint64_t seed, const double* x, int64_t n, double* out) {
  static const int need[7] = {0, 1, 1, 1, 2, 1, 2};
  if (!params || !out || n < 1 || kind < 1 || kind > 6 || np < need[kind] || np > 8 || (!sample && !x)) return fail(PZ_EINVAL, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PZ_ECUDA, "no CUDA device (this library has no CPU fallback)");
  double *d_x = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_out, n * sizeof(double));
  if (e == cudaSuccess && !sample) e = cudaMalloc(&d_x, n * sizeof(double));
  if (e == cudaSuccess && !sample) e = cudaMemcpy(d_x, x, n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    launch_dist(sample, kind, params, np, seed, n, d_x, d_out, nullptr);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_x); cudaFree(d_out);
  if (e != cudaSuccess) return fail(PZ_ECUDA, "pz_dist: %s", cudaGetErrorString(e));
  return PZ_OK;
}

int pz_dist_sample(int32_t kind, const double* params, int32_t n_params, uint64_t seed, int64_t n, double* out) {
  return dist_tap(true, kind, params, n_params, seed, nullptr, n, out);
}

int pz_dist_score(int32_t kind, const double* params, int32_t n_params, const double* x, int64_t n, double* out) {
  return dist_tap(false, kind, params, n_params, 0, x, n, out);
}

int pz_filter_reset(pz_filter* f, uint64_t seed) {
  if (!f) return fail(PZ_EINVAL, "null filter");
  if (f->a.world > 1 || f->is_multi) return fail(PZ_ESTATE, "pz_filter_reset: single-GPU filters only (destroy and re-create the shards together)");
  PZ_CUDA(cudaSetDevice(f->device));
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  cudaGetLastError();
  if (f->is_beta) {
    BetaDev h{};
    h.a = f->desc.beta_a; h.b = f->desc.beta_b;
    PZ_CUDA(cudaMemcpy(f->beta, &h, sizeof(h), cudaMemcpyHostToDevice));
    f->t = 0;
    return PZ_OK;
  }
  if (f->is_rb) {
    RbDev h;
    PZ_CUDA(cudaMemcpy(&h, f->rb, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < PZ_MAX_NX; ++i) h.mean[i] = i < h.nx ? f->desc.x0[i] : 0.0;
    memset(h.cov, 0, sizeof(h.cov));
    h.t = 0;
    PZ_CUDA(cudaMemcpy(f->rb, &h, sizeof(h), cudaMemcpyHostToDevice));
    f->t = 0;
    return PZ_OK;
  }
  if (f->is_mtt) return fail(PZ_ESTATE, "pz_filter_reset: not available for the tracker (destroy and re-create)");
  if (f->is_gen) {
    f->a.seed = seed; f->ga.seed = seed;
    return gen_start(f);
  }
  StepArgs& a = f->a;
  a.seed = seed;
  launch_init(a, f->stream);
  launch_scan_partition(a, fused_tile_out(a.model.nx, a.f64), f->stream);
  PZ_CUDA(cudaGetLastError());
  PZ_CUDA(cudaStreamSynchronize(f->stream));
  f->t = 0;
  if (f->graph) {   // the captured step carries the seed by value: capture it again
    cudaGraphExecDestroy(f->graph);
    f->graph = nullptr;
    cudaGraph_t graph = nullptr;
    PZ_CUDA(cudaStreamBeginCapture(f->stream, cudaStreamCaptureModeThreadLocal));
    f->launches_per_step = launch_step(a, nullptr, f->stream);
    PZ_CUDA(cudaStreamEndCapture(f->stream, &graph));
    PZ_CUDA(cudaGraphInstantiate(&f->graph, graph, 0));
    cudaGraphDestroy(graph);
  }
  return PZ_OK;
}

}  // extern "C"
