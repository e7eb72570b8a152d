#!/usr/bin/env python
"""bench.py -- particle-updates/sec of the streaming particle-filter step on B200.

    python bench.py --gpus N --steps K --warmup W            (this build)
    python bench.py --impl reference --gpus N --steps K ...  (CPU port of the reference path)

A "step" is one pass of the hot path (resample + gather + propagate + weight + normalise,
zparticles', haskell/src/Inference.hs:101-104) over all N_p particles.  Headline workload at 1 GPU:
BASELINE.json configs[1], the 1-D Gaussian random-walk model at N_p = 10^8 on one B200, state and
arithmetic in double (the reference's precision, Inference.hs:28; `--dtype f32` selects fp32 state).

value : N_p * K / (device time of K back-to-back steps), observations resident in HBM.
e2e   : the same through the public API with HOST observation buffers, one ZStream.step
        per observation (pinned H2D of the observation + D2H of the posterior summary + sync
        inside the timed region).
roofline / cpu_baseline: see DESIGN.md "Measurement".
legs  : (1 GPU, default run) the other single-GPU configurations in the same JSON line: the fp32 variant
        of the headline, configs[2] (LG(4,2) and LG(6,3), N = 10^7, both storage types), configs[3]
        (multi-target tracker, N = 10^6) and a sustained leg (10^3 steps with clocks and power).
sharded_parity : (N > 1) before timing, the live ranks run a small sharded filter next to a single-GPU
        filter of the same population on rank 0 and compare state bits, totals and exponents.

The CPU oracle (oracle/) is executed only by the cpu_baseline leg and by --impl reference.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODELS = {"rw1d": (1, 1), "lg42": (4, 2), "lg63": (6, 3), "mtt": (4, 2), "mcam": (4, 0)}
HEADLINE_PARTICLES = 100_000_000


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons sampled DURING the timed region: NVML (about 1 ms per sample) when
    nvidia_ml_py is importable, else nvidia-smi (about 50 ms per sample)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.sm, self.pw, self.sm_max, self.reasons, self.source = [], [], None, set(), "nvidia-smi"
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
        except Exception:
            self.h = None

    def _sample_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            self.pw.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        r = [c.strip() for c in out.split(",")]
        if r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
        self.sm_max = float(r[1])
        try:
            self.pw.append(float(r[2]))
        except ValueError:
            pass
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self._sample_nvml() if self.h is not None else self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.002 if self.h is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["clock sampling unavailable"], "samples": 0}
        sm = sorted(self.sm)
        out = {"sm_mhz": sm[len(sm) // 2], "sm_mhz_min": sm[0], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
               "samples": len(sm), "source": self.source}
        if self.pw:
            out["power_w_median"] = sorted(self.pw)[len(self.pw) // 2]
            out["power_w_max"] = max(self.pw)
        return out


def ncu_traffic(name, dtype, n_p):
    """dram__bytes_read + dram__bytes_write of one pf_step_tile launch from the committed ncu capture of this
    model / dtype (profiles/traffic.json, written by tools/ncu_buckets.py) when it was taken at this N; else None.
    Not measured in this run: the source is named in roofline.traffic_source."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh).get(f"{name}_{dtype}")
        return (float(t["dram_bytes"]), t.get("report")) if t and int(t["particles"]) == int(n_p) else (None, None)
    except Exception:
        return None, None


def synth_obs(d, T, seed=0x0B5E2026):
    """Synthetic observation stream from the generative model of the descriptor (Demo.hs:203-207,
    SingleTargetTracking.hs:61-65): x_t ~ N(F x_{t-1}, Q), y_t ~ N(H x_t, R), numpy, fixed seed.  The GPU
    arm must not execute anything under oracle/, so it does not use the oracle's generator."""
    nx, ny = d.nx, d.ny
    F = np.array(d.F[: nx * nx]).reshape(nx, nx); Q = np.array(d.Q[: nx * nx]).reshape(nx, nx)
    Hm = np.array(d.H[: ny * nx]).reshape(ny, nx); R = np.array(d.R[: ny * ny]).reshape(ny, ny)
    Lq = np.linalg.cholesky(Q + 1e-300 * np.eye(nx)); Lr = np.linalg.cholesky(R)
    rng = np.random.default_rng(seed)
    x = np.array(d.x0[:nx], dtype=np.float64)
    obs = np.zeros((T, ny))
    for t in range(T):
        x = F @ x + Lq @ rng.standard_normal(nx)
        obs[t] = Hm @ x + Lr @ rng.standard_normal(ny)
    return obs


def model_desc(pz, name, f64=False):
    return {"rw1d": pz.rw1d, "lg42": pz.mtt_track, "lg63": pz.stt}[name](f64=f64)


def helper_desc(H, name):
    return {"rw1d": H.model_rw1d, "lg42": H.model_mtt_track, "lg63": H.model_stt}[name]()


def make_config(args):
    """The workload description shared verbatim by both arms (the driver compares them)."""
    nx, ny = MODELS[args.model]
    n_p = args.total_particles // max(args.gpus, 1) if args.total_particles else args.particles
    esz = 8 if args.dtype == "f64" else 4
    return {
        "workload": f"{args.model} bootstrap particle filter, N={n_p} particles per GPU x {args.gpus} GPU(s), "
                    f"one observation per step (BASELINE.json configs[1]; configs[4] when sharded)",
        "model": args.model, "particles_per_gpu": n_p, "state_dim": nx, "obs_dim": ny,
        "state_storage": "fp64" if args.dtype == "f64" else "fp32",
        "weights": "u16 fixed point per particle, u64 block prefix", "resampler": "systematic (canonical PZ-SYS-256/q16)",
        "l2": "inputs exceed L2 (state %.0f MB per buffer); no flush needed" % (nx * esz * n_p / 1e6),
        "reference_arm": "`--impl reference` times the CPU port of the reference algorithm (double, multinomial resample2) "
                         "on a BOUNDED SAMPLE of this workload: the same model and stream at the N stated in its "
                         "cpu_baseline.sample, sized so that its K + W steps take about a minute on the host cores",
    }


def cpu_port_rate(H, name, n_sample, steps, threads, faithful=False, optimised=False):
    """particle-updates/s of the CPU port.  faithful=False: the reference's algorithm (double,
    log-domain normalise, multinomial draws) with the O(N) per-draw CDF walk replaced by a binary
    search and every loop (fold, CDF build, draws, gather) chunked over the cores with OpenMP;
    faithful=True: the literal O(N^2) walk, one thread; optimised=True: SURVEY.md 8(d)(ii), not the
    reference's algorithm but the best a host does with the same step (max-shifted weights, one O(N)
    systematic resample over a parallel prefix sum)."""
    lib = H.oracle()
    d = helper_desc(H, name)
    obs, _ = H.gen_obs(d, steps)
    t0 = time.perf_counter()
    rc = lib.pzo_refpf_run(C.byref(d), n_sample, 1, H.ptr(obs, C.c_double), steps, 2 if optimised else (0 if faithful else 1), None, None,
                           1 if faithful else threads)
    dt = time.perf_counter() - t0
    assert rc == 0
    return n_sample * steps / dt, dt


def run_reference(args):
    """--impl reference: the CPU port of the reference's path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pzhelpers as H
    threads = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    n_full = make_config(args)["particles_per_gpu"]
    # size the sample so that K + W steps take about 40 s: calibrate on two steps at 10^6 particles
    n_cal = min(n_full, 1_000_000)
    cpu_port_rate(H, args.model, n_cal, 1, threads)
    rate_cal, _ = cpu_port_rate(H, args.model, n_cal, 2, threads)
    n_sample = int(min(n_full, max(100_000, rate_cal * 40.0 / (K + W))))
    if args.cpu_particles:
        n_sample = min(n_full, args.cpu_particles)
    cpu_port_rate(H, args.model, n_sample, max(1, W), threads)  # W warm-up steps
    rate, dt = cpu_port_rate(H, args.model, n_sample, K, threads)
    orate, odt = cpu_port_rate(H, args.model, n_sample, K, threads, optimised=True)
    line = {
        "impl": "reference", "metric": "particle-updates/sec", "value": rate, "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.total_particles else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args),
        "sample": {"particles": n_sample, "steps": K, "of_particles_per_gpu": n_full},
        "cpu_baseline": {"value": rate, "unit": "particle-updates/s", "cores": threads, "kind": "port",
                         "sample": f"N={n_sample} particles (of the workload's {n_full}) x {K} steps of the same model and stream; reference algorithm "
                                   "(Inference.hs:57-66,101-104: double, logaddexp normalise, multinomial resample2) "
                                   "with the O(N) per-draw CDF walk replaced by a binary search; fold, CDF build, draws and gather "
                                   "chunked over all cores with OpenMP; the literal O(N^2) reference cannot finish one step at this N",
                         "optimised_systematic": {"value": orate, "cores": threads,
                                                  "sample": f"same sample; not the reference's algorithm: max-shifted weights, O(N) systematic "
                                                            f"resampling over a parallel prefix sum, all loops over all cores ({odt:.1f} s)"}},
        "e2e": {"value": rate, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def synth_mtt_scenario(d, T, seed=5, m_max=32, warm_tracks=0):
    """Detections of the tracker's generative model (sampleStep under MP.sim, MultiTargetTracking.hs:75-99,
    126-131) in numpy: births Poisson(lambda_b) from the birth prior, motion x' ~ N(F x, Q), detection with
    probability ps*pd (a missed track survives with (1-pd)/(1-ps*pd)), clutter Poisson(lambda_c) ~ N(0, cv I),
    detections shuffled.  warm_tracks seeds that many tracks at t = 0 so that short benches see a populated
    scene (configs[3]: "up to 16 targets").  The GPU arm does not execute oracle/."""
    rng = np.random.default_rng(seed)
    F = np.array(d.F[:16]).reshape(4, 4); Lq = np.sqrt(np.diag(np.array(d.Q[:16]).reshape(4, 4)))
    m = d.mtt
    pspd = m.survival_prob * m.pd
    p_coast = (1.0 - m.pd) / (1.0 - pspd)
    birth_sd = np.sqrt([m.birth_pos_var, m.birth_pos_var, m.birth_vel_var, m.birth_vel_var])
    tracks = [rng.standard_normal(4) * birth_sd for _ in range(warm_tracks)]
    out = []
    for _ in range(T):
        det, keep = [], []
        for x in tracks:
            x = F @ x + Lq * rng.standard_normal(4)
            if rng.random() < pspd:
                det.append(x[:2] + np.sqrt(m.meas_var) * rng.standard_normal(2)); keep.append(x)
            elif rng.random() < p_coast:
                keep.append(x)
        for _ in range(rng.poisson(m.new_track_lambda)):
            if len(keep) < 16:
                x = rng.standard_normal(4) * birth_sd
                det.append(x[:2] + np.sqrt(m.meas_var) * rng.standard_normal(2)); keep.append(x)
        for _ in range(rng.poisson(m.clutter_lambda)):
            det.append(np.sqrt(m.clutter_var) * rng.standard_normal(2))
        tracks = keep
        det = det[:m_max]
        rng.shuffle(det)
        out.append(np.array(det, dtype=np.float64).reshape(-1, 2))
    return out


def measure_mtt(pz, n_p, K, W, tracks, with_cpu, devices=1):
    """configs[3]: the multi-target tracker (N = 10^6, up to 16 tracks), detections streamed step by step.
    devices > 1: configs[4]'s sharded tracker -- ONE process drives `devices` GPUs (pz_filter_create_multi), n_p particles in total."""
    import torch
    obs = synth_mtt_scenario(pz.mtt(), W + K, seed=5, warm_tracks=tracks)
    zs = pz.zparticles(n_p, pz.mtt(), seed=1) if devices <= 1 else pz.zparticles_multi(n_p, pz.mtt(), list(range(devices)), seed=1)
    for o in obs[:W]:
        zs.step(o)
    torch.cuda.synchronize()
    dev_ms, launches, bytes_alg = 0.0, 0, 0.0
    t0 = time.perf_counter()
    for o in obs[W:]:
        post = zs.step(o)
        ms, nl = zs.last_timing()
        dev_ms += ms; launches += nl
        bytes_alg += 2.0 * n_p * (16 + 64 * post.mean[0])  # header + live track records, read + written
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    zs.close()
    peak, peak_src = measured_peak()
    achieved = bytes_alg / (dev_ms * 1e-3) / 1e9
    out = {
        "value": n_p * K / (dev_ms * 1e-3), "unit": "particle-updates/s", "ms_per_step": dev_ms / K, "dtype": "f32",
        "workload": f"multi-target tracker (MultiTargetTracking.hs), N={n_p} particles, up to 16 Rao-Blackwellised tracks (BASELINE.json configs[3])",
        "mean_tracks_last_step": float(post.mean[0]), "steps": K, "n_devices": devices,
        "e2e": {"value": n_p * K / wall, "unit": "particle-updates/s", "h2d_bytes_per_step": 8 * int(np.mean([len(o) for o in obs[W:]])),
                "d2h_bytes_per_step": 144, "ms_per_step": wall / K * 1e3},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "mtt_step (+ resample kernels)", "peak_source": peak_src, "alg_bytes_per_launch": bytes_alg / K,
                     "kernel_ms": dev_ms / K, "kernel_ms_source": "whole step, device events",
                     "alg_bytes_definition": "live track records (64 B each) + 16 B header per particle, read once and written once"},
    }
    if with_cpu:
        import pzhelpers as H
        n_cpu = 2000
        orc = H.OracleMtt(H.model_mtt(), n_cpu, 1)
        t0 = time.perf_counter()
        for o in obs[: W + K]:
            orc.step(o)
        cpu_dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu * (W + K) / cpu_dt, "unit": "particle-updates/s", "cores": 1, "kind": "port",
                               "sample": f"N={n_cpu} x {W + K} steps of the same detection stream, oracle restatement in double, 1 thread ({cpu_dt:.1f} s)"}
    return out


def synth_mcam_scenario(d, T, seed=7, clutter=0.05, warm_tracks=4):
    """Observation rows of the two-camera model (MultiCamera.hs:183-206) in numpy: survival, Poisson births with a camera and
    an identity, one row per live track and camera (box + unit pose + appearance reading), Poisson clutter.  Rows are
    (camera, pose[10], box[4], reading).  The GPU arm does not execute oracle/."""
    rng = np.random.default_rng(seed)
    m = d.mtt
    q = np.sqrt(np.array([d.Q[i * 4 + i] for i in range(4)])); x0 = np.array(list(d.x0)[:4])
    pv = np.sqrt([m.birth_pos_var, m.birth_pos_var, m.birth_vel_var, m.birth_vel_var])
    apps, tracks, out = [], [], []

    def birth():
        if not apps or rng.random() < 1.0 / (len(apps) + 1):
            apps.append(rng.standard_normal(10)); ident = len(apps) - 1
        else:
            ident = int(rng.integers(len(apps)))
        return [int(rng.integers(2)), ident, x0 + pv * rng.standard_normal(4)]
    tracks = [birth() for _ in range(warm_tracks)]
    for _ in range(T):
        tracks = [[c, i, x + q * rng.standard_normal(4)] for c, i, x in tracks if rng.random() < m.survival_prob]
        for _ in range(rng.poisson(m.new_track_lambda)):
            if len(tracks) < 16:
                tracks.append(birth())
        rows = []
        for cam in (0, 1):
            for c, i, x in tracks:
                if c != cam:
                    continue
                pose = rng.standard_normal(10); pose /= np.linalg.norm(pose)
                rows.append(np.concatenate([[cam], pose, x + rng.standard_normal(4), [pose @ apps[i] + rng.standard_normal()]]))
            for _ in range(rng.poisson(clutter)):
                rows.append(np.concatenate([[cam], rng.standard_normal(10), x0 + rng.standard_normal(4), [rng.standard_normal()]]))
        out.append(np.array(rows[:32], dtype=np.float64).reshape(-1, 16))
    return out


def measure_mcam(pz, n_p, K, W):
    """the MultiCamera tracker (N particles, 4 targets in the scene at t = 0, clutter 0.05 per camera: at the
    reference's amtClutter = 1 the model itself starves the population, tests/test_mcam.py)."""
    import torch
    d = pz.multicam(); d.mtt.clutter_lambda = 0.05
    obs = synth_mcam_scenario(d, W + K)
    zs = pz.zparticles(n_p, d, seed=1)
    for o in obs[:W]:
        zs.step(o)
    torch.cuda.synchronize()
    dev_ms, launches = 0.0, 0
    t0 = time.perf_counter()
    for o in obs[W:]:
        post = zs.step(o)
        ms, nl = zs.last_timing()
        dev_ms += ms; launches += nl
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    zs.close()
    return {"value": n_p * K / (dev_ms * 1e-3), "unit": "particle-updates/s", "ms_per_step": dev_ms / K, "dtype": "f32", "steps": K,
            "workload": f"MultiCamera tracker (MultiCamera.hs), N={n_p} particles, 2 cameras, appearance marginals over R^10",
            "mean_tracks_last_step": float(post.mean[0]),
            "e2e": {"value": n_p * K / wall, "unit": "particle-updates/s", "ms_per_step": wall / K * 1e3},
            "gpu_launches": int(launches)}


def measure_lg(pz, name, dtype, n_p, K, W, device=0, profile=True, sustained=0):
    """One single-GPU leg: device-resident K steps, end-to-end K steps, per-kernel events."""
    import torch
    nx, ny = MODELS[name]
    esz = 8 if dtype == "f64" else 4
    d = model_desc(pz, name, dtype == "f64")
    obs = synth_obs(d, W + 2 * K + 16 + sustained)
    zs = pz.zparticles(n_p, d, seed=1, device=device)
    zs.run(obs[:W], want_summaries=False)
    torch.cuda.synchronize()
    zs.run(obs[W:W + K], want_summaries=False)
    torch.cuda.synchronize()
    dev_ms, launches = zs.last_timing()
    off = W + K
    for i in range(3):
        zs.step(obs[off + i])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        zs.step(obs[off + 3 + i])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    out = {"value": n_p * K / (dev_ms * 1e-3), "unit": "particle-updates/s", "ms_per_step": dev_ms / K, "dtype": dtype, "steps": K,
           "workload": f"{name} bootstrap particle filter, N={n_p}, state {'fp64' if esz == 8 else 'fp32'}",
           "e2e": {"value": n_p * K / e2e_s, "unit": "particle-updates/s", "h2d_bytes_per_step": 8 * ny, "d2h_bytes_per_step": 144,
                   "ms_per_step": e2e_s / K * 1e3},
           "gpu_launches": int(launches)}
    if profile:
        prof = np.array([zs.step_profile(obs[off + 3 + K + i]) for i in range(10)])
        peak, peak_src = measured_peak()
        t_tile = float(np.mean(prof[:, 0]))
        alg = 2.0 * nx * esz * n_p
        tr, rep = ncu_traffic(name, dtype, n_p)
        out["roofline"] = {"bound": "hbm", "achieved": alg / (t_tile * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": alg / (t_tile * 1e-3) / 1e9 / peak, "traffic": tr,
                           "traffic_source": f"ncu capture of one launch at this N, committed under profiles/ ({rep}); not measured in this run" if tr else None,
                           "kernel": "pf_step_tile", "alg_bytes_per_launch": alg, "kernel_ms": t_tile,
                           "step_kernels_ms": {"pf_step_tile": t_tile, "pf_scan_blocks": float(np.mean(prof[:, 1])),
                                               "pf_partition": float(np.mean(prof[:, 2]))}}
    if sustained:
        # the configs[1] stream length: 10^3 back-to-back steps, clocks and power sampled during the run
        so = obs[-sustained:]
        sampler = ClockSampler(device); sampler.start()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        zs.run(so, want_summaries=False)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms, _ = zs.last_timing()
        out["sustained"] = {"steps": sustained, "value": n_p * sustained / (ms * 1e-3), "ms_per_step": ms / sustained,
                            "wall_s": wall, "clocks": sampler.summary()}
    zs.close()
    return out


def sharded_parity(pz, dist, torch, rank, world, local_rank, nccl_id_fn):
    """Bit-equality of the sharded filter with the single-GPU filter on the LIVE ranks: world * 256 * 101 particles,
    6 steps, rw1d and lg42 (fp32) plus rw1d in double; rank 0 runs the single-GPU filter, the shards are all-gathered."""
    out = {}
    n = world * 256 * 101
    for name, dtype in (("rw1d", "f32"), ("lg42", "f32"), ("rw1d", "f64")):
        d = model_desc(pz, name, dtype == "f64")
        obs = synth_obs(d, 6, seed=77)
        zs = pz.zparticles_sharded(n, d, nccl_id_fn(), rank, world, seed=4242, device=local_rank)
        ref = pz.zparticles(n, d, seed=4242, device=local_rank) if rank == 0 else None
        verdict = "bit-exact"
        bt = np.uint64 if dtype == "f64" else np.uint32
        for t in range(6):
            post = zs.step(obs[t])
            mine = torch.from_numpy(zs.read_state().copy()).to(f"cuda:{local_rank}")
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            if rank == 0 and verdict == "bit-exact":
                pr = ref.step(obs[t])
                full = torch.cat(parts, dim=1).cpu().numpy()
                same = np.array_equal(np.ascontiguousarray(full).view(bt), ref.read_state().view(bt))
                same = same and post.total_weight == pr.total_weight and post.e_max == pr.e_max
                if not same:
                    verdict = f"MISMATCH at step {t}"
        flag = torch.tensor([1 if verdict == "bit-exact" else 0], device=f"cuda:{local_rank}")
        dist.broadcast(flag, 0)
        zs.close()
        if ref is not None:
            ref.close()
        dist.barrier()
        out[f"{name}_{dtype}"] = verdict if rank == 0 else ("bit-exact" if int(flag.item()) else "MISMATCH")
    out["particles"] = n
    out["steps"] = 6
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this build has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pz = g.load_package()
    name = args.model
    nx, ny = MODELS[name]
    cfg = make_config(args)
    n_p = cfg["particles_per_gpu"]
    scaling = "strong" if args.total_particles else "weak"
    if world > 1:
        n_p -= n_p % 256  # a sharded filter holds whole resampling blocks (256 particles) per rank
    K, W = args.steps, args.warmup
    f64 = args.dtype == "f64"
    esz = 8 if f64 else 4
    d = model_desc(pz, name, f64)
    obs = synth_obs(d, W + K + 2 * K + 16)

    def nccl_id():
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(pz.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    parity = None
    if world > 1:
        if not args.no_parity:
            parity = sharded_parity(pz, dist, torch, rank, world, local_rank, nccl_id)
        zs = pz.zparticles_sharded(n_p * world, d, nccl_id(), rank, world, seed=1, device=local_rank)
    else:
        zs = pz.zparticles(n_p, d, seed=1, device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput: K back-to-back steps, observations staged in HBM
    zs.run(obs[:W], want_summaries=False)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    zs.run(obs[W:W + K], want_summaries=False)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms, launches = zs.last_timing()
    clocks = sampler.summary() if sampler else None
    if world > 1:
        tt = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms = float(tt.item())
    value = n_p * world * K / (dev_ms * 1e-3)

    # ---- end to end: host observation -> ZStream.step -> posterior summary on the host
    off = W + K
    for i in range(min(W, 3)):
        zs.step(obs[off + i])
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        post = zs.step(obs[off + 3 + i])
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e = n_p * world * K / e2e_s

    # ---- roofline of the dominant kernel (fused resample+propagate+weight), CUDA events on its stream
    if world == 1:
        prof = np.array([zs.step_profile(obs[off + 3 + K + i]) for i in range(10)])
        kernel_src = "CUDA events around pf_step_tile on its stream (10 steps outside the graph)"
    else:  # sharded: the step kernel is interleaved with the exchange kernels; bound the kernel by the whole step
        prof = np.array([[dev_ms / K, float("nan"), float("nan")]])
        kernel_src = "whole step (kernel + exchange), device events"
    t_tile_ms = float(np.mean(prof[:, 0]))
    alg_bytes = 2.0 * nx * esz * n_p  # read the state once, write it once; the kernel also moves 4 B of u16 weights
    achieved = alg_bytes / (t_tile_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    # every rank tears its shard down together, then rank 0 reports
    n_migrated = int(post.n_migrated)
    if world > 1:
        mt = torch.tensor([n_migrated], device="cuda", dtype=torch.int64)
        dist.all_reduce(mt)
        n_migrated = int(mt.item())
    post_step, post_mean0, post_ess = int(post.step), float(post.mean[0]), float(post.ess)
    zs.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- the other single-GPU configurations (default run only): fp32 twin, configs[2], configs[3], sustained
    legs = None
    if world == 1 and not args.no_legs and name == "rw1d" and n_p == HEADLINE_PARTICLES:
        legs = {}
        kl = min(K, 20)
        other = "f32" if f64 else "f64"

        def leg(key, fn):   # an extra leg must never take the headline down: its failure is reported in its own slot
            try:
                legs[key] = fn()
            except Exception as exc:
                legs[key] = {"error": str(exc)[:200]}
        leg(f"rw1d_{other}", lambda: measure_lg(pz, "rw1d", other, n_p, kl, 3))
        for m in ("lg42", "lg63"):
            for dt in ("f32", "f64"):
                leg(f"{m}_{dt}", lambda m=m, dt=dt: measure_lg(pz, m, dt, 10_000_000, kl, 3))
        leg("mtt_f32", lambda: measure_mtt(pz, 1_000_000, 10, 3, args.mtt_tracks, with_cpu=False))
        # (at the reference's clutter rate the MultiCamera model starves a population -- passert False; the leg runs at 0.05)
        leg("multicam_f32", lambda: measure_mcam(pz, 100_000, 10, 3))
        import torch
        if torch.cuda.device_count() >= 2:   # configs[4]'s sharded tracker: one process, every GPU of the box, 10^6 particles per GPU
            nd = min(torch.cuda.device_count(), 8)
            leg(f"mtt_f32_{nd}_devices", lambda: measure_mtt(pz, 1_000_000 * nd, 10, 3, args.mtt_tracks, with_cpu=False, devices=nd))
        leg("sustained_rw1d_" + args.dtype, lambda: measure_lg(pz, "rw1d", args.dtype, n_p, 5, 3, profile=False, sustained=1000)["sustained"])

    # ---- CPU baseline on this box's host cores (bounded sample)
    threads = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu:
        import pzhelpers as H  # the CPU oracle: executed by this leg only
        n_sample = min(n_p, args.cpu_particles or 2_000_000)
        steps_cpu = 5
        rate, dt = cpu_port_rate(H, name, n_sample, steps_cpu, threads)
        frate, fdt = cpu_port_rate(H, name, 1000, 100, 1, faithful=True)
        orate, odt = cpu_port_rate(H, name, n_sample, steps_cpu, threads, optimised=True)
        cpu = {"value": rate, "unit": "particle-updates/s", "cores": threads, "kind": "port",
               "sample": f"N={n_sample} x {steps_cpu} steps, reference algorithm in double, binary-search multinomial, every loop over all cores with OpenMP ({dt:.1f} s)",
               "faithful_o_n2": {"value": frate, "cores": 1,
                                 "sample": f"configs[0]: N=1000 x 100 steps, literal O(N^2) resample2 ({fdt:.1f} s)"},
               "optimised_systematic": {"value": orate, "cores": threads,
                                        "sample": f"same sample; not the reference's algorithm: max-shifted weights, O(N) systematic "
                                                  f"resampling over a parallel prefix sum, all loops over all cores ({odt:.1f} s)"}}
    tr, rep = ncu_traffic(name, args.dtype, n_p) if world == 1 else (None, None)
    line = {
        "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e, "unit": "particle-updates/s", "h2d_bytes_per_step": 8 * ny, "d2h_bytes_per_step": 144,
                "ms_per_step": e2e_s / K * 1e3},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "wall_s_timed": wall,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # also against the nominal 8 TB/s, and for the whole step (value x B_alg per GPU)
                     "frac_nominal_8TBs": achieved / 8000.0,
                     "whole_step": {"achieved": value / world * 2 * esz * nx / 1e9, "frac": value / world * 2 * esz * nx / 1e9 / peak},
                     "traffic": tr,
                     "traffic_source": f"ncu capture of one launch at this N, committed under profiles/ ({rep}); not measured in this run" if tr else None,
                     "kernel": "pf_step_tile", "peak_source": peak_src,
                     "alg_bytes_per_launch": alg_bytes, "alg_bytes_per_update": 2 * esz * nx, "kernel_ms": t_tile_ms, "kernel_ms_source": kernel_src,
                     "step_kernels_ms": {"pf_step_tile": t_tile_ms, "pf_scan_blocks": float(np.mean(prof[:, 1])),
                                         "pf_partition": float(np.mean(prof[:, 2]))}},
        "cpu_baseline": cpu,
        "posterior_check": {"step": post_step, "mean0": post_mean0, "ess": post_ess, "n_migrated_last_step_all_ranks": n_migrated},
        "parity": "partial: bit-exact against this repository's CPU oracle; unpinned against the Haskell binary (no GHC, no reference-held vectors)",
    }
    if parity is not None:
        line["sharded_parity"] = parity
    if legs is not None:
        line["legs"] = legs
    print(json.dumps(line), flush=True)


def run_mtt(args):
    import torch
    import __graft_entry__ as g
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this build has no CPU fallback")
    pz = g.load_package()
    sampler = ClockSampler(0); sampler.start()
    leg = measure_mtt(pz, args.particles, args.steps, args.warmup, args.mtt_tracks, with_cpu=not args.no_cpu, devices=args.devices)
    clocks = sampler.summary()
    line = {"metric": "particle-updates/sec", "value": leg["value"], "unit": leg["unit"], "n_gpus": args.devices, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": leg["workload"], "model": "mtt", "particles_per_gpu": args.particles,
                       "mean_tracks_last_step": leg["mean_tracks_last_step"], "l2": "particle records exceed L2 (1040 B per particle)"},
            "e2e": leg["e2e"], "gpu_launches": leg["gpu_launches"], "clocks": clocks, "roofline": leg["roofline"],
            "cpu_baseline": leg.get("cpu_baseline")}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="rw1d", choices=list(MODELS))
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"], help="state storage and arithmetic (descriptor flag state_f64)")
    ap.add_argument("--particles", type=int, default=HEADLINE_PARTICLES, help="particles per GPU (weak scaling, the default)")
    ap.add_argument("--total-particles", type=int, default=0,
                    help="strong scaling: the whole population, split evenly over the GPUs (overrides --particles)")
    ap.add_argument("--cpu-particles", type=int, default=0, help="CPU sample size (0: sized automatically)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the extra single-GPU legs of the default run")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded bit-equality check before the timed region")
    ap.add_argument("--mtt-tracks", type=int, default=8, help="tracker bench: targets present at t = 0")
    ap.add_argument("--devices", type=int, default=1, help="tracker bench (--model mtt): GPUs driven by this ONE process (pz_filter_create_multi); --particles is the total")
    args = ap.parse_args()
    args.gpus = int(os.environ.get("WORLD_SIZE", args.gpus))
    if args.impl == "reference":
        run_reference(args)
    elif args.model == "mtt":
        if args.particles == HEADLINE_PARTICLES:
            args.particles = 1_000_000
        run_mtt(args)
    elif args.model == "mcam":   # the MultiCamera tracker alone: a plain line, no roofline claim
        import torch
        import __graft_entry__ as g
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; this build has no CPU fallback")
        n_p = 100_000 if args.particles == HEADLINE_PARTICLES else args.particles
        leg = measure_mcam(g.load_package(), n_p, args.steps, args.warmup)
        print(json.dumps({"metric": "particle-updates/sec", "value": leg["value"], "unit": leg["unit"], "n_gpus": 1, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": leg["workload"], "model": "mcam", "particles_per_gpu": n_p},
                          "e2e": leg["e2e"], "gpu_launches": leg["gpu_launches"]}))
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
