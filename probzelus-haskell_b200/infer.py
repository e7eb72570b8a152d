"""ctypes binding of libpzgpu.so and the `infer`-style stream API.

Reference interface mirrored here (names and argument meaning kept):

    zparticles   :: Int -> ZStream (Weighted m) a b -> ZStream m a [b]   Inference.hs:107-108
    zdsparticles :: Int -> ZStream (StateT Heap (Weighted m)) a b -> ..  Inference.hs:113-114
    ZStream { step :: a -> f (ZStream f a b, b) }                        Util/ZStream.hs:16
    runStream    :: (a -> m ()) -> ZStream m () a -> m ()                Util/ZStream.hs:65-69

The model argument is a kernel-side descriptor instead of a Haskell closure.
"""
import ctypes as C
import os

import numpy as np

PZ_MAX_NX, PZ_MAX_NY = 6, 3
PZ_RESAMPLE_SYSTEMATIC, PZ_RESAMPLE_MULTINOMIAL_REF = 0, 1
PZ_MODEL_RW1D, PZ_MODEL_LINGAUSS, PZ_MODEL_MTT, PZ_MODEL_WALKER, PZ_MODEL_BETA_BERNOULLI, PZ_MODEL_MULTICAM = 1, 2, 3, 4, 5, 6
PZ_DIST = {"bernoulli": 1, "categorical": 2, "poisson": 3, "uniform": 4, "exponential": 5, "normal": 6}
_ERRORS = {-1: "EINVAL", -2: "ECUDA", -3: "ENCCL", -4: "EDEGENERATE", -5: "ECAPACITY", -6: "ENOMEM", -7: "ESTATE"}


class PzError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"PZ_{_ERRORS.get(code, code)}: {msg}")
        self.code = code


class MttParams(C.Structure):
    _fields_ = [("clutter_lambda", C.c_double), ("new_track_lambda", C.c_double), ("pd", C.c_double),
                ("survival_prob", C.c_double), ("clutter_var", C.c_double), ("birth_pos_var", C.c_double),
                ("birth_vel_var", C.c_double), ("meas_var", C.c_double), ("k_max", C.c_int32),
                ("m_max", C.c_int32)]


class WalkerParams(C.Structure):
    """pz_walker_params (include/pzgpu.h): Examples/Walker.hs constants."""
    _fields_ = [("pos_noise", C.c_double), ("obs_std", C.c_double), ("dt", C.c_double), ("p_init", C.c_double * 3),
                ("half_life", (C.c_double * 3) * 3), ("speed_lo", C.c_double * 3), ("speed_hi", C.c_double * 3),
                ("exp_mean_as_written", C.c_int32), ("reserved", C.c_int32)]


class ModelDesc(C.Structure):
    """pz_model_desc (include/pzgpu.h)."""
    _fields_ = [("kind", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("scalar_ll_2x", C.c_int32),
                ("resampler", C.c_int32), ("delayed", C.c_int32), ("state_f64", C.c_int32), ("reserved", C.c_int32),
                ("F", C.c_double * (PZ_MAX_NX * PZ_MAX_NX)), ("Q", C.c_double * (PZ_MAX_NX * PZ_MAX_NX)),
                ("H", C.c_double * (PZ_MAX_NY * PZ_MAX_NX)), ("R", C.c_double * (PZ_MAX_NY * PZ_MAX_NY)),
                ("x0", C.c_double * PZ_MAX_NX), ("mtt", MttParams),
                ("ess_threshold", C.c_double), ("carry_weights", C.c_int32), ("reserved2", C.c_int32),
                ("walker", WalkerParams), ("beta_a", C.c_double), ("beta_b", C.c_double)]


class StepSummary(C.Structure):
    """pz_step_summary (include/pzgpu.h)."""
    _fields_ = [("step", C.c_int64), ("log_evidence_inc", C.c_double), ("ess", C.c_double),
                ("mean", C.c_double * PZ_MAX_NX), ("var", C.c_double * PZ_MAX_NX), ("n_migrated", C.c_int64),
                ("e_max", C.c_int32), ("reserved", C.c_int32), ("total_weight", C.c_uint64)]


_SYMBOLS = [
    "pz_abi_version", "pz_last_error", "pz_model_rw1d", "pz_model_stt", "pz_model_mtt_track", "pz_model_mtt",
    "pz_filter_create", "pz_comm_unique_id", "pz_filter_create_sharded", "pz_filter_step", "pz_filter_run",
    "pz_filter_last_timing", "pz_filter_step_profile", "pz_filter_read_particles", "pz_filter_read_tracks", "pz_filter_read_ancestors", "pz_filter_read_state",
    "pz_filter_read_logw", "pz_filter_read_state_f64", "pz_filter_read_logw_f64", "pz_filter_read_posterior", "pz_filter_n_local", "pz_resample_systematic", "pz_resample_multinomial",
    "pz_filter_destroy", "pz_model_walker", "pz_model_beta_bernoulli", "pz_importance_resampling",
    "pz_dist_sample", "pz_dist_score", "pz_filter_reset", "pz_filter_create_multi",
    "pz_model_multicam", "pz_filter_read_mcam_tracks",
]


def exported_symbols():
    """Every entry point include/pzgpu.h declares."""
    return list(_SYMBOLS)


def lib_path():
    override = os.environ.get("PZ_LIB_PATH")  # A/B builds of the same ABI (tools/, profiles/README.md)
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpzgpu.so")


_lib = None


def lib():
    """Load libpzgpu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise PzError(-2, f"{path} is missing: run __graft_entry__.build() (make -C probzelus-haskell_b200/csrc)")
    L = C.CDLL(path)
    f64p, i32p, f32p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_float)
    L.pz_last_error.restype = C.c_char_p
    L.pz_model_rw1d.argtypes = [C.POINTER(ModelDesc), C.c_double, C.c_double]
    for nm in ("pz_model_stt", "pz_model_mtt_track", "pz_model_mtt", "pz_model_multicam"):
        getattr(L, nm).argtypes = [C.POINTER(ModelDesc)]
    L.pz_filter_create.argtypes = [C.POINTER(ModelDesc), C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
    L.pz_comm_unique_id.argtypes = [C.c_void_p]
    L.pz_filter_create_sharded.argtypes = [C.POINTER(ModelDesc), C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.POINTER(C.c_void_p)]
    L.pz_filter_step.argtypes = [C.c_void_p, f64p, C.c_int32, C.c_int32, C.POINTER(StepSummary)]
    L.pz_filter_run.argtypes = [C.c_void_p, f64p, C.c_int64, C.c_int32, C.POINTER(StepSummary)]
    L.pz_filter_last_timing.argtypes = [C.c_void_p, f64p, C.POINTER(C.c_int64)]
    L.pz_filter_step_profile.argtypes = [C.c_void_p, f64p, C.c_int32, C.c_int32, f64p]
    L.pz_filter_read_particles.argtypes = [C.c_void_p, C.c_int32, f64p, C.c_int64]
    L.pz_filter_read_tracks.argtypes = [C.c_void_p, C.c_int32, i32p, i32p, f64p, f64p, C.c_int64]
    L.pz_filter_read_mcam_tracks.argtypes = [C.c_void_p, C.c_int32, i32p, i32p, i32p, i32p, f64p, f64p, f64p, i32p, f64p, f64p, C.c_int64]
    L.pz_filter_read_ancestors.argtypes = [C.c_void_p, i32p, C.c_int64]
    L.pz_filter_read_state.argtypes = [C.c_void_p, f32p, C.c_int64]
    L.pz_filter_read_logw.argtypes = [C.c_void_p, f32p, C.c_int64]
    L.pz_filter_read_state_f64.argtypes = [C.c_void_p, f64p, C.c_int64]
    L.pz_filter_read_logw_f64.argtypes = [C.c_void_p, f64p, C.c_int64]
    L.pz_filter_read_posterior.argtypes = [C.c_void_p, f64p, f64p]
    L.pz_filter_n_local.argtypes = [C.c_void_p]
    L.pz_filter_n_local.restype = C.c_int64
    L.pz_resample_systematic.argtypes = [f64p, C.c_int64, C.c_double, i32p]
    L.pz_resample_multinomial.argtypes = [f64p, C.c_int64, C.c_uint64, C.c_uint32, i32p]
    L.pz_filter_destroy.argtypes = [C.c_void_p]
    L.pz_filter_destroy.restype = None
    L.pz_model_walker.argtypes = [C.POINTER(ModelDesc)]
    L.pz_model_beta_bernoulli.argtypes = [C.POINTER(ModelDesc), C.c_double, C.c_double]
    L.pz_importance_resampling.argtypes = [f64p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int64, i32p]
    L.pz_dist_sample.argtypes = [C.c_int32, f64p, C.c_int32, C.c_uint64, C.c_int64, f64p]
    L.pz_dist_score.argtypes = [C.c_int32, f64p, C.c_int32, f64p, C.c_int64, f64p]
    L.pz_filter_reset.argtypes = [C.c_void_p, C.c_uint64]
    L.pz_filter_create_multi.argtypes = [C.POINTER(ModelDesc), C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise PzError(rc, lib().pz_last_error().decode())


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# ---- model descriptors (kernel-side plugin surface) -------------------------------------------
def _variants(d, carry_weights, ess_threshold):
    d.carry_weights, d.ess_threshold = int(carry_weights), float(ess_threshold)
    return d


def walker(resampler=PZ_RESAMPLE_SYSTEMATIC, carry_weights=True, ess_threshold=0.0, exp_mean_as_written=True):
    """Examples/Walker.hs:20-148 under `particles N` (cumulative weights, Inference.hs:73-86)."""
    d = ModelDesc()
    _check(lib().pz_model_walker(C.byref(d)))
    d.resampler = resampler
    d.walker.exp_mean_as_written = int(exp_mean_as_written)
    return _variants(d, carry_weights, ess_threshold)


def beta_bernoulli(a=1.0, b=1.0):
    """betaBernoulli a b (Examples/DelayedSamplingAbstract.hs:57-69) under delayed sampling."""
    d = ModelDesc()
    _check(lib().pz_model_beta_bernoulli(C.byref(d), a, b))
    return d


def rw1d(q_var=1.0, r_var=3.0, scalar_ll_2x=True, resampler=PZ_RESAMPLE_SYSTEMATIC, delay=False, f64=False,
         carry_weights=False, ess_threshold=0.0):
    """1-D Gaussian random walk + noisy observation: kalman1DObserved, Examples/Demo.hs:184-198
    (delay=True: kalman1DObservedDelayed, Demo.hs:215-242).  f64=True: state and arithmetic in double, the
    reference's own precision (Inference.hs:28)."""
    d = ModelDesc()
    _check(lib().pz_model_rw1d(C.byref(d), q_var, r_var))
    d.scalar_ll_2x, d.resampler, d.delayed, d.state_f64 = int(scalar_ll_2x), resampler, int(delay), int(f64)
    return _variants(d, carry_weights, ess_threshold)


def stt(resampler=PZ_RESAMPLE_SYSTEMATIC, delay=False, f64=False, carry_weights=False, ess_threshold=0.0):
    """Single-target tracker: Examples/SingleTargetTracking.hs:28-54.  delay=False is the bootstrap filter,
    delay=True the delayed-sampling stream in which every particle carries the exact Kalman marginal."""
    d = ModelDesc()
    _check(lib().pz_model_stt(C.byref(d)))
    d.resampler, d.delayed, d.state_f64 = resampler, int(delay), int(f64)
    return _variants(d, carry_weights, ess_threshold)


def mtt_track(resampler=PZ_RESAMPLE_SYSTEMATIC, delay=False, f64=False, carry_weights=False, ess_threshold=0.0):
    """2-D position/velocity track model of the multi-target tracker: MultiTargetTracking.hs:56-73."""
    d = ModelDesc()
    _check(lib().pz_model_mtt_track(C.byref(d)))
    d.resampler, d.delayed, d.state_f64 = resampler, int(delay), int(f64)
    return _variants(d, carry_weights, ess_threshold)


def multicam():
    """MultiCamera tracker constants: Examples/MultiCamera.hs:40-227.  Observation rows: (camera, pose[10], box[4], reading)."""
    d = ModelDesc()
    _check(lib().pz_model_multicam(C.byref(d)))
    return d


def mtt():
    """Multi-target tracker constants: Examples/MultiTargetTracking.hs:36-145."""
    d = ModelDesc()
    _check(lib().pz_model_mtt(C.byref(d)))
    return d


# ---- the stream ----------------------------------------------------------------------------------
class Posterior:
    """One element of the output stream: the particle approximation of p(x_t | y_1..t)."""

    def __init__(self, stream, summary):
        self._stream = stream
        nx = 6 if stream.model.kind == PZ_MODEL_WALKER else stream.nx   # the walker summarises x, y, vx, vy, P(walking), P(running)
        self.step = summary.step
        self.mean = np.array(summary.mean[:nx])
        self.var = np.array(summary.var[:nx])
        self.ess = summary.ess
        self.log_evidence_inc = summary.log_evidence_inc
        self.e_max = summary.e_max
        self.total_weight = summary.total_weight
        self.n_migrated = summary.n_migrated
        self.resampled = bool(summary.reserved)   # stored-log-weight filters: this step's population came from a resample

    def particles(self, stride=1):
        """The reference's `[b]`: equally weighted particles after resampling (every stride-th)."""
        if self._stream.t != self.step + 1:
            raise PzError(-7, "the stream has advanced past this posterior")
        return self._stream.read_particles(stride)


class ZStream:
    """ZStream IO Obs Posterior built around pz_filter_step (Util/ZStream.hs:16-21)."""

    def __init__(self, model, n_particles, seed=0, device=0, rank=0, world=1, nccl_id=None, devices=None):
        self.model, self.nx, self.ny = model, model.nx, model.ny
        self.t = 0
        self.rank, self.world = rank, world
        self._h = C.c_void_p()
        if devices is not None:   # one process, several devices (pz_filter_create_multi)
            ids = (C.c_int * len(devices))(*devices)
            _check(lib().pz_filter_create_multi(C.byref(model), int(n_particles), seed, len(devices), ids, C.byref(self._h)))
        elif world > 1:
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            _check(lib().pz_filter_create_sharded(C.byref(model), int(n_particles), seed, device, rank, world, buf,
                                                  C.byref(self._h)))
        else:
            _check(lib().pz_filter_create(C.byref(model), int(n_particles), seed, device, C.byref(self._h)))
        self.n = int(lib().pz_filter_n_local(self._h))  # particles held by this process
        self.n_global = int(n_particles)

    def step(self, obs):
        y = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1)
        s = StepSummary()
        if self.model.kind == PZ_MODEL_MTT:   # a (possibly empty) list of 2-D detections
            rc = lib().pz_filter_step(self._h, _ptr(y, C.c_double) if y.size else None, y.size // 2, 2, C.byref(s))
        elif self.model.kind == PZ_MODEL_MULTICAM:   # a (possibly empty) list of rows (camera, pose[10], box[4], reading)
            rc = lib().pz_filter_step(self._h, _ptr(y, C.c_double) if y.size else None, y.size // 16, 16, C.byref(s))
        else:
            rc = lib().pz_filter_step(self._h, _ptr(y, C.c_double), 1, y.size, C.byref(s))
        if rc in (0, -3, -4, -5):   # the library advanced its step for every one of these codes
            self.t += 1
        _check(rc)
        return Posterior(self, s)

    def read_tracks(self, stride=1):
        """Tracker output: list (one per selected resampled particle) of lists of (id, mean[4], cov[4,4])."""
        n_sel = (self.n + stride - 1) // stride
        cnt = np.zeros(n_sel, dtype=np.int32); ids = np.zeros((n_sel, 16), dtype=np.int32)
        mean = np.zeros((n_sel, 16, 4)); cov = np.zeros((n_sel, 16, 4, 4))
        _check(lib().pz_filter_read_tracks(self._h, stride, _ptr(cnt, C.c_int32), _ptr(ids, C.c_int32), _ptr(mean, C.c_double),
                                           _ptr(cov, C.c_double), n_sel))
        return [[(int(ids[i, k]), mean[i, k].copy(), cov[i, k].copy()) for k in range(cnt[i])] for i in range(n_sel)]

    def read_mcam_tracks(self, stride=1, appearances=False):
        """MultiCamera output (processObservationsStream, MultiCamera.hs:208-216): per selected resampled particle a dict
        {"tracks": [(trackID, identity, camera, startTime, mean[4], var[4])], "n_app": A[, "app": [(mean[10], cov[10,10])]]}."""
        n_sel = (self.n + stride - 1) // stride
        cnt = np.zeros(n_sel, np.int32); ids = np.zeros((n_sel, 16), np.int32); idn = np.zeros((n_sel, 16), np.int32)
        cam = np.zeros((n_sel, 16), np.int32); st = np.zeros((n_sel, 16)); mean = np.zeros((n_sel, 16, 4)); var = np.zeros((n_sel, 16, 4))
        na = np.zeros(n_sel, np.int32)
        am = np.zeros((n_sel, 16, 10)) if appearances else None
        ac = np.zeros((n_sel, 16, 10, 10)) if appearances else None
        _check(lib().pz_filter_read_mcam_tracks(self._h, stride, _ptr(cnt, C.c_int32), _ptr(ids, C.c_int32), _ptr(idn, C.c_int32),
                                                _ptr(cam, C.c_int32), _ptr(st, C.c_double), _ptr(mean, C.c_double), _ptr(var, C.c_double),
                                                _ptr(na, C.c_int32), _ptr(am, C.c_double) if appearances else None,
                                                _ptr(ac, C.c_double) if appearances else None, n_sel))
        out = []
        for i in range(n_sel):
            p = {"tracks": [(int(ids[i, k]), int(idn[i, k]), int(cam[i, k]), float(st[i, k]), mean[i, k].copy(), var[i, k].copy())
                            for k in range(cnt[i])], "n_app": int(na[i])}
            if appearances:
                p["app"] = [(am[i, a].copy(), ac[i, a].copy()) for a in range(na[i])]
            out.append(p)
        return out

    def run(self, observations, want_summaries=True):
        """runStream over a whole observation array [T, ny]; returns the T summaries."""
        obs = np.ascontiguousarray(observations, dtype=np.float64).reshape(-1, self.ny)
        T = obs.shape[0]
        out = (StepSummary * T)() if want_summaries else None
        rc = lib().pz_filter_run(self._h, _ptr(obs, C.c_double), T, self.ny, out)
        self.t += T
        _check(rc)
        return [Posterior(self, s) for s in out] if want_summaries else None

    def step_profile(self, obs):
        """One step outside the CUDA graph; returns ms of (fused kernel, scan, partition)."""
        y = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1)
        ms = np.zeros(3)
        _check(lib().pz_filter_step_profile(self._h, _ptr(y, C.c_double), 1, y.size, _ptr(ms, C.c_double)))
        self.t += 1
        return ms

    def last_timing(self):
        ms, n = C.c_double(), C.c_int64()
        _check(lib().pz_filter_last_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def read_particles(self, stride=1):
        n_sel = (self.n + stride - 1) // stride
        out = np.zeros((n_sel, self.nx))
        _check(lib().pz_filter_read_particles(self._h, stride, _ptr(out, C.c_double), out.size))
        return out

    def read_ancestors(self):
        out = np.zeros(self.n, dtype=np.int32)
        _check(lib().pz_filter_read_ancestors(self._h, _ptr(out, C.c_int32), out.size))
        return out

    def read_state(self):
        """Pre-resample population, SoA [nx, n], in the filter's storage type."""
        if self.model.state_f64 or self.model.kind == PZ_MODEL_WALKER:
            out = np.zeros((self.nx, self.n), dtype=np.float64)
            _check(lib().pz_filter_read_state_f64(self._h, _ptr(out, C.c_double), out.size))
            return out
        out = np.zeros((self.nx, self.n), dtype=np.float32)
        _check(lib().pz_filter_read_state(self._h, _ptr(out, C.c_float), out.size))
        return out

    def read_posterior(self):
        """Delayed-sampling filters: the Gaussian marginal (mean, cov) every particle carries (MDistr)."""
        mean = np.zeros(self.nx); cov = np.zeros((self.nx, self.nx))
        _check(lib().pz_filter_read_posterior(self._h, _ptr(mean, C.c_double), _ptr(cov, C.c_double)))
        return mean, cov

    def read_logw(self):
        if self.model.state_f64:
            out = np.zeros(self.n, dtype=np.float64)
            _check(lib().pz_filter_read_logw_f64(self._h, _ptr(out, C.c_double), out.size))
            return out
        out = np.zeros(self.n, dtype=np.float32)
        _check(lib().pz_filter_read_logw(self._h, _ptr(out, C.c_float), out.size))
        return out

    def reset(self, seed=0):
        """Back to the prior at step 0 (pz_filter_reset): the way on after a degenerate step."""
        _check(lib().pz_filter_reset(self._h, seed))
        self.t = 0

    def close(self):
        if self._h:
            lib().pz_filter_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def zparticles(num_particles, model, seed=0, device=0):
    """zparticles numParticles model (Inference.hs:107-108)."""
    return ZStream(model, num_particles, seed, device)


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the host distributes it to every rank)."""
    buf = (C.c_char * 128)()
    _check(lib().pz_comm_unique_id(buf))
    return bytes(buf)


def zparticles_sharded(num_particles, model, nccl_id, rank, world, seed=0, device=0):
    """zparticles over `world` GPUs, one process per GPU: this rank holds global slots
    [rank * N / world, (rank + 1) * N / world); resampling is global."""
    return ZStream(model, num_particles, seed, device, rank, world, nccl_id)


def zparticles_multi(num_particles, model, devices, seed=0):
    """zparticles over several GPUs driven by THIS process (one host thread): global resampling, the exchange
    over NVLink peer memory; step / run / read_state as on one GPU."""
    return ZStream(model, num_particles, seed, devices=list(devices))


def zdsparticles(num_particles, model, seed=0, device=0):
    """zdsparticles (Inference.hs:113-114): the per-particle heap is the fixed-capacity state."""
    return ZStream(model, num_particles, seed, device)


def infer(model, num_particles, seed=0, device=0):
    """ProbZelus `infer`: model x particle count -> stream of posteriors."""
    return ZStream(model, num_particles, seed, device)


def run_stream(act, stream, observations):
    """ZS.runStream act stream (Util/ZStream.hs:65-69), driven by a finite observation list."""
    for y in observations:
        act(stream.step(y))


def resample_systematic(logw, u):
    lw = np.ascontiguousarray(logw, dtype=np.float64)
    anc = np.zeros(lw.size, dtype=np.int32)
    _check(lib().pz_resample_systematic(_ptr(lw, C.c_double), lw.size, float(u), _ptr(anc, C.c_int32)))
    return anc


def importance_resampling(logw, seed, step, k=1):
    """importanceResampling (Inference.hs:68-71): k draws from Categorical(exp(logw - max))."""
    lw = np.ascontiguousarray(logw, dtype=np.float64)
    idx = np.zeros(k, dtype=np.int32)
    _check(lib().pz_importance_resampling(_ptr(lw, C.c_double), lw.size, seed, step, k, _ptr(idx, C.c_int32)))
    return idx


def dist_sample(kind, params, seed, n):
    """n draws of a Distributions.hs primitive on the GPU (pz_dist_sample)."""
    p = np.ascontiguousarray(params, dtype=np.float64)
    out = np.zeros(n)
    _check(lib().pz_dist_sample(PZ_DIST[kind], _ptr(p, C.c_double), p.size, seed, n, _ptr(out, C.c_double)))
    return out


def dist_score(kind, params, x):
    """log-densities of a Distributions.hs primitive on the GPU (pz_dist_score)."""
    p = np.ascontiguousarray(params, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(x.size)
    _check(lib().pz_dist_score(PZ_DIST[kind], _ptr(p, C.c_double), p.size, _ptr(x, C.c_double), x.size, _ptr(out, C.c_double)))
    return out


def resample_multinomial(logw, seed, step):
    lw = np.ascontiguousarray(logw, dtype=np.float64)
    anc = np.zeros(lw.size, dtype=np.int32)
    _check(lib().pz_resample_multinomial(_ptr(lw, C.c_double), lw.size, seed, step, _ptr(anc, C.c_int32)))
    return anc


# ---- output encoding (the JSON lines the Java glimpsetracks visualiser reads) -------------------
def _num(x):
    x = float(x)
    return repr(x) if x == x and abs(x) != float("inf") else "null"   # aeson encodes NaN / Infinity as null


def json_line_tracks(ground_truth, particles, observations):
    """One line of `examples mtt N` / `examples stt delay N` (test/Examples.hs:23-24,
    SingleTargetTracking.hs:67-69): [[ [id,[x..]].. ], [ [ [id,{"mean":[..],"cov":[[..]..]}].. ] per particle ], [[ox,oy(,oz)]..]].
    ground_truth: [(id, state)], particles: [[(id, mean, cov-or-None)]], observations: [[..]].
    A track with cov None is an RConst and serialises as a bare array (DSProg.hs:52-54)."""
    def vec(v):
        return "[" + ",".join(_num(a) for a in v) + "]"

    def track(t):
        tid, mean, cov = t
        if cov is None:
            return "[%d,%s]" % (tid, vec(mean))
        return '[%d,{"mean":%s,"cov":[%s]}]' % (tid, vec(mean), ",".join(vec(r) for r in cov))

    gt = "[" + ",".join("[%d,%s]" % (i, vec(x)) for i, x in ground_truth) + "]"
    ps = "[" + ",".join("[" + ",".join(track(t) for t in p) + "]" for p in particles) + "]"
    ob = "[" + ",".join(vec(o) for o in observations) + "]"
    return "[" + gt + "," + ps + "," + ob + "]"


def json_line_mcam(ground_truth, particles, observations):
    """One line of `MultiCamera.runExample` (Examples/MultiCamera.hs:221-227): the aeson encoding of
    (groundTruth :: [Track], particles :: [[MarginalTrack]], obs :: [(CameraID, R 10, R 4, R 1)]).  A track is the generic
    record encoding of TrackG (:48-57) in declaration order -- posWH, startTime, trackID, identity, camera -- with posWH a bare
    array for the ground truth (R 4, MVDistributions.hs:24-25) and {"mean":..,"cov":..} for a marginal (DSProg.hs:52-54,
    DelayedSampling.hs:69-73).  ground_truth: [(trackID, identity, camera, box[4])] (mcam_scenario's rows; startTime unknown: 0);
    particles: ZStream.read_mcam_tracks()'s list; observations: rows (camera, pose[10], box[4], reading)."""
    def vec(v):
        return "[" + ",".join(_num(a) for a in v) + "]"

    def obj(pos, start, tid, ident, cam):
        return '{"posWH":%s,"startTime":%s,"trackID":%d,"identity":%d,"camera":%d}' % (pos, _num(start), tid, ident, cam)

    gt = "[" + ",".join(obj(vec(t[-1]), t[3] if len(t) > 4 else 0.0, t[0], t[1], t[2]) for t in ground_truth) + "]"

    def marginal(mean, var):
        cov = ",".join(vec([var[i] if i == j else 0.0 for j in range(4)]) for i in range(4))
        return '{"mean":%s,"cov":[%s]}' % (vec(mean), cov)
    ps = "[" + ",".join("[" + ",".join(obj(marginal(m, v), st, tid, idn, cam) for tid, idn, cam, st, m, v in p["tracks"]) + "]"
                        for p in particles) + "]"
    ob = "[" + ",".join("[%d,%s,%s,%s]" % (int(o[0]), vec(o[1:11]), vec(o[11:15]), vec(o[15:16])) for o in observations) + "]"
    return "[" + gt + "," + ps + "," + ob + "]"


def json_line_generic1d(values):
    """One line for Generic1D.java (:333-348): a flat array, one double per plotted series."""
    return "[" + ",".join(_num(v) for v in values) + "]"
