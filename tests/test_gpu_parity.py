"""GPU parity tests (run with -m gpu on the B200 box).  Every call goes through the C ABI
(libpzgpu.so via ctypes); the CPU oracle is the checker.

Bars: bit-exact for ancestors, fixed-point block statistics and the fp32 particle state
(integer / index work and the canonical fp32 spec); 1e-6 relative for the posterior summaries
(floating reductions whose order and partial precision differ between kernels and oracle); 3 standard errors for
posterior moments against the exact Kalman filter."""
import ctypes as C

import numpy as np
import pytest

import pzhelpers as H

pytestmark = pytest.mark.gpu


def weight_vectors(n, rng):
    yield "uniform", np.zeros(n)
    yield "gaussian", rng.standard_normal(n) * 3
    hot = np.full(n, -np.inf); hot[(n * 7) // 11] = -2.5
    yield "one_hot", hot
    yield "geometric", -0.01 * np.arange(n, dtype=np.float64)
    mix = rng.standard_normal(n) * 2
    mix[rng.random(n) < 0.3] = -np.inf
    if n > 3:
        mix[1] = np.nan
    yield "with_inf_nan", mix
    yield "wide_range", -np.abs(rng.standard_normal(n)) * 200


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 1000, 2048, 2049, 4097, 100003, 1 << 20])
def test_resample_systematic_bit_exact(pz, n):
    rng = np.random.default_rng(n)
    for name, lw in weight_vectors(n, rng):
        for u in (0.0, 0.37, 1 - 2.0 ** -32):
            U = int(u * 2 ** 32)
            rc, ref = H.oracle_systematic(lw.astype(np.float32), U)
            if rc == H.PZ_EDEGENERATE:
                with pytest.raises(pz.PzError) as e:
                    pz.resample_systematic(lw, u)
                assert e.value.code == H.PZ_EDEGENERATE
                continue
            got = pz.resample_systematic(lw, u)
            assert np.array_equal(got, ref), (name, n, u, np.flatnonzero(got != ref)[:5])


def test_resample_golden_fixture(pz):
    g = np.load(H.ROOT + "/tests/golden/philox_normal_kat.npz")
    got = pz.resample_systematic(g["lw_r"].astype(np.float64), int(g["U_r"]) / 2.0 ** 32)
    assert np.array_equal(got, g["anc_r"])


def test_resample_degenerate_reports_error(pz):
    for lw in (np.full(1000, -np.inf), np.full(1000, np.nan)):
        with pytest.raises(pz.PzError) as e:
            pz.resample_systematic(lw, 0.5)
        assert e.value.code == H.PZ_EDEGENERATE


@pytest.mark.parametrize("n", [1000, 100003])
def test_resample_multinomial_bit_exact(pz, n):
    rng = np.random.default_rng(n + 1)
    lw = rng.standard_normal(n) * 2
    for step in (0, 5):
        rc, ref = H.oracle_multinomial(lw.astype(np.float32), 1234, step)
        assert rc == 0
        assert np.array_equal(pz.resample_multinomial(lw, 1234, step), ref)


MODELS = {"rw1d": (H.model_rw1d, "rw1d"), "lg42": (H.model_mtt_track, "mtt_track"), "lg63": (H.model_stt, "stt")}


def bits(a):
    """Bit pattern of a float32 / float64 array (bit-exact comparisons)."""
    return a.view(np.uint64 if a.dtype == np.float64 else np.uint32)


@pytest.mark.parametrize("f64", [False, True], ids=["f32", "f64"])
@pytest.mark.parametrize("model", ["rw1d", "lg42", "lg63"])
@pytest.mark.parametrize("n", [1, 300, 4096, 10007, 200000])
def test_filter_trajectory_bit_exact(pz, model, n, f64):
    """Whole-filter parity: after every step the population (fp32, or double for state_f64 filters), the
    fixed-point total and the global exponent equal the oracle's bit for bit; summaries agree to 1e-9."""
    hd = MODELS[model][0](f64=f64)
    d = getattr(pz, MODELS[model][1])(f64=f64)
    T = 8
    obs, _ = H.gen_obs(hd, T)
    zs = pz.zparticles(n, d, seed=4242)
    orc = H.OracleFilter(hd, n, 4242)
    for t in range(T):
        # ancestors of the resample this step performs
        if t in (0, 3):
            anc_gpu = zs.read_ancestors()
        post = zs.step(obs[t])
        rc, s = orc.step(obs[t])
        assert rc == 0
        x_ref, lw_ref, anc_ref = orc.state()
        if t in (0, 3):
            assert np.array_equal(anc_gpu, anc_ref), (model, n, t)
        x = zs.read_state()
        assert x.dtype == x_ref.dtype == (np.float64 if f64 else np.float32)
        assert np.array_equal(bits(x), bits(x_ref)), (model, n, t)
        assert np.array_equal(bits(zs.read_logw()), bits(lw_ref))
        assert post.total_weight == s.total_weight and post.e_max == s.e_max and post.step == t
        nx = hd.nx
        # summaries are floating reductions (fp32 per-lane partials, fp64 above): tolerance, not bits
        assert np.allclose(post.mean, s.mean[:nx], rtol=1e-6, atol=2e-6)
        assert np.allclose(post.var, s.var[:nx], rtol=1e-4, atol=1e-6)
        assert abs(post.ess - s.ess) < 1e-5 * s.ess
        assert abs(post.log_evidence_inc - s.log_evidence_inc) < 1e-5
    zs.close(); orc.close()


@pytest.mark.parametrize("f64", [False, True], ids=["f32", "f64"])
def test_filter_bit_exact_at_16m_particles(pz, f64):
    """Bit parity at N = 2^24 (the oracle's step is an OpenMP loop: a few seconds): state, log-weights, ancestors,
    totals -- the same bar as the small sizes, two thousand tiles instead of a handful."""
    n = 1 << 24
    hd = H.model_rw1d(f64=f64)
    obs, _ = H.gen_obs(hd, 3)
    zs = pz.zparticles(n, pz.rw1d(f64=f64), seed=2024)
    orc = H.OracleFilter(hd, n, 2024)
    for t in range(3):
        anc_gpu = zs.read_ancestors() if t == 2 else None
        post = zs.step(obs[t])
        rc, s = orc.step(obs[t])
        assert rc == 0
        x_ref, lw_ref, anc_ref = orc.state()
        if anc_gpu is not None:
            assert np.array_equal(anc_gpu, anc_ref)
        assert np.array_equal(bits(zs.read_state()), bits(x_ref)), t
        assert post.total_weight == s.total_weight and post.e_max == s.e_max
        assert abs(post.mean[0] - s.mean[0]) < 2e-6 and abs(post.ess - s.ess) < 1e-5 * s.ess
    assert np.array_equal(bits(zs.read_logw()), bits(lw_ref))
    zs.close(); orc.close()


@pytest.mark.parametrize("f64", [False, True], ids=["f32", "f64"])
@pytest.mark.parametrize("model", ["rw1d", "lg42", "lg63"])
def test_filter_outlier_observations_bit_exact(pz, model, f64):
    """Observations far in the tails collapse the weights onto a handful of particles: runs of thousands
    of children (the long-run duplicates and the segment look-back of the step kernel), empty blocks,
    shifts outside the fast range of the slot map.  Still bit-exact against the oracle."""
    hd = MODELS[model][0](f64=f64)
    d = getattr(pz, MODELS[model][1])(f64=f64)
    n = 60000 + 123
    obs, _ = H.gen_obs(hd, 10)
    obs = obs.copy()
    obs[2] += 14.0; obs[3] -= 25.0; obs[6] += 40.0; obs[7] += 40.0
    zs = pz.zparticles(n, d, seed=77)
    orc = H.OracleFilter(hd, n, 77)
    max_run = 0
    for t in range(len(obs)):
        post = zs.step(obs[t])
        rc, s = orc.step(obs[t])
        assert rc == 0
        x_ref, _, anc_ref = orc.state()
        max_run = max(max_run, int(np.bincount(anc_ref).max()))
        assert np.array_equal(bits(zs.read_state()), bits(x_ref)), (model, t)
        assert post.total_weight == s.total_weight and post.e_max == s.e_max
        assert abs(post.ess - s.ess) < 1e-4 * s.ess + 1e-6
    assert max_run > 300   # the scenario really produced long runs
    zs.close(); orc.close()


def test_filter_golden_fixture(pz):
    """The committed fixture (made by tools/make_golden.py) without re-running the oracle."""
    g = np.load(H.ROOT + "/tests/golden/philox_normal_kat.npz")
    for name, mk in (("rw1d", pz.rw1d), ("lg42", pz.mtt_track), ("lg63", pz.stt)):
        zs = pz.zparticles(3000, mk(), seed=42)
        for y in g[name + "_obs"]:
            post = zs.step(y)
        assert np.array_equal(zs.read_state().view(np.uint32), g[name + "_x"].view(np.uint32))
        assert post.total_weight == int(g[name + "_total"])
        assert np.allclose(post.mean, g[name + "_mean"], rtol=1e-6, atol=2e-6)
        zs.close()


@pytest.mark.parametrize("f64", [False, True], ids=["f32", "f64"])
def test_run_equals_step(pz, f64):
    """runStream over a staged observation array == repeated ZStream.step (bitwise)."""
    d = pz.rw1d(f64=f64)
    obs, _ = H.gen_obs(H.model_rw1d(), 40)
    a = pz.zparticles(50000, d, seed=9)
    b = pz.zparticles(50000, d, seed=9)
    pa = [a.step(y) for y in obs]
    pb = b.run(obs)
    assert np.array_equal(bits(a.read_state()), bits(b.read_state()))
    for x, y in zip(pa, pb):
        assert x.total_weight == y.total_weight and x.mean[0] == y.mean[0] and x.step == y.step  # same kernels: bitwise
    ms, launches = b.last_timing()
    assert launches == 3 * 40 and ms > 0


def test_reproducible_and_seed_sensitive(pz):
    d = pz.mtt_track()
    obs, _ = H.gen_obs(H.model_mtt_track(), 5)
    outs = []
    for seed in (1, 1, 2):
        zs = pz.zparticles(20000, d, seed=seed)
        zs.run(obs)
        outs.append(zs.read_state().copy())
        zs.close()
    assert np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("model,f64", [("rw1d", False), ("lg42", False), ("lg63", False), ("rw1d", True), ("lg63", True)])
def test_posterior_matches_kalman(pz, model, f64):
    """Posterior mean / variance within 3 SE of the exact Kalman filter over 20 seeds at
    N = 10^5 (BASELINE.md section 5 gate 2); also the log-evidence for the properly normalised scores."""
    hd = MODELS[model][0](f64=f64)
    d = getattr(pz, MODELS[model][1])(f64=f64)
    T, n, seeds = 30, 100000, 20
    obs, _ = H.gen_obs(hd, T)
    km, kc, kl = H.kalman(hd, obs)
    nx = hd.nx
    means = np.zeros((seeds, T, nx)); vars_ = np.zeros((seeds, T, nx)); ll = np.zeros(seeds)
    for s in range(seeds):
        zs = pz.zparticles(n, d, seed=1000 + s)
        posts = zs.run(obs)
        means[s] = [p.mean for p in posts]; vars_[s] = [p.var for p in posts]
        ll[s] = sum(p.log_evidence_inc for p in posts)
        zs.close()
    se_m = means.std(axis=0, ddof=1) / np.sqrt(seeds)
    zm = (means.mean(axis=0) - km) / np.maximum(se_m, 1e-12)
    kvar = np.array([np.diag(c) for c in kc])
    se_v = vars_.std(axis=0, ddof=1) / np.sqrt(seeds)
    zv = (vars_.mean(axis=0) - kvar) / np.maximum(se_v, 1e-12)
    H.assert_within_mc_tolerance(means, km, "mean")
    H.assert_within_mc_tolerance(vars_, kvar, "var")
    if model != "rw1d":
        assert abs(ll.mean() - kl.sum()) < 4 * ll.std(ddof=1) / np.sqrt(seeds) + 0.02


def test_quantised_weights_carry_no_visible_bias_at_large_n(pz):
    """N = 2^28: the Monte Carlo error of the posterior mean is ~10^-4 standard deviations per step and
    ~10^-5 for the average over 200 steps, small enough to expose a systematic effect of the 16-bit
    block-relative weights or of the fp32 model step -- none is visible against the exact Kalman filter."""
    hd = H.model_rw1d()
    T, n = 200, 1 << 28
    obs, _ = H.gen_obs(hd, T)
    km, kc, _ = H.kalman(hd, obs)           # exact target of the doubled score (R_eff = R / 2)
    zs = pz.zparticles(n, pz.rw1d(), seed=11)
    posts = zs.run(obs)
    err = np.array([p.mean[0] for p in posts]) - km[:, 0]
    sd = np.sqrt(kc[:, 0, 0])
    assert np.sqrt(np.mean((err[20:] / sd[20:]) ** 2)) < 4e-4    # per step: ~ sqrt(2 / N) with resampling noise
    assert abs(np.mean(err[20:] / sd[20:])) < 6e-5               # time average: no bias at the 10^-5 level
    verr = np.array([p.var[0] for p in posts])[20:] / kc[20:, 0, 0] - 1
    assert abs(verr.mean()) < 2e-4
    zs.close()


def test_long_stream_stays_on_the_kalman_posterior(pz):
    """BASELINE configs[2] at full size: the 2-D position/velocity tracker, N = 10^7 particles over 10^4 steps.  The
    filter neither degenerates nor drifts: over the last 2000 steps the posterior mean tracks the exact
    Kalman mean to Monte Carlo accuracy and the log-evidence per step matches."""
    hd = H.model_mtt_track()
    T, n = 10000, 10000000
    obs, _ = H.gen_obs(hd, T)
    km, kc, kl = H.kalman(hd, obs)
    zs = pz.zparticles(n, pz.mtt_track(), seed=2026)
    posts = zs.run(obs)
    assert len(posts) == T and posts[-1].step == T - 1
    mean = np.array([p.mean for p in posts[-2000:]])
    var = np.array([p.var for p in posts[-2000:]])
    ess = np.array([p.ess for p in posts[-2000:]])
    kvar = np.array([np.diag(c) for c in kc[-2000:]])
    err = (mean - km[-2000:]) / np.sqrt(kvar)
    assert np.sqrt(np.mean(err ** 2)) < 0.01 and np.max(np.abs(err)) < 0.08   # ~ 1/sqrt(ESS) with path degeneracy
    assert np.all(np.abs(var.mean(axis=0) / kvar.mean(axis=0) - 1) < 0.02)
    assert ess.min() > 100 and np.median(ess) > 0.2 * n   # a 4-sigma innovation costs most of the ESS for one step
    ll = np.array([p.log_evidence_inc for p in posts])
    assert abs(ll.sum() - kl.sum()) < 1e-3 * abs(kl.sum())
    zs.close()


@pytest.mark.parametrize("model", ["rw1d", "mtt_track", "stt"])
def test_delayed_sampling_is_the_exact_kalman_recursion(pz, model):
    """model.delayed = 1 (`delay=True`, SingleTargetTracking.hs:36-38; Demo.hs:215-242): every particle carries
    the Gaussian marginal of makeMarginal / makeConditional (DelayedSampling.hs:135-152), i.e. the exact Kalman
    posterior, with equal weights."""
    hd = {"rw1d": H.model_rw1d, "mtt_track": H.model_mtt_track, "stt": H.model_stt}[model]()
    d = getattr(pz, model)(delay=True)
    T, n = 60, 5000
    obs, _ = H.gen_obs(hd, T)
    hk = {"rw1d": lambda: H.model_rw1d(scalar_ll_2x=0), "mtt_track": H.model_mtt_track, "stt": H.model_stt}[model]()
    km, kc, kl = H.kalman(hk, obs)     # the delayed posterior conditions on the true R (only the SCORE is doubled)
    zs = pz.zparticles(n, d, seed=1)
    for t in range(T):
        post = zs.step(obs[t])
        mean, cov = zs.read_posterior()
        assert np.allclose(mean, km[t], rtol=1e-10, atol=1e-10) and np.allclose(cov, kc[t], rtol=1e-9, atol=1e-12)
        assert np.allclose(post.mean, km[t], atol=1e-10) and np.allclose(post.var, np.diag(kc[t]), atol=1e-12)
        assert post.ess == n and post.step == t
        scale = 2.0 if model == "rw1d" else 1.0   # gaussian_ll' is twice the log-density (Distributions.hs:160-161)
        assert abs(post.log_evidence_inc - scale * kl[t]) < 1e-9 * max(1.0, abs(kl[t]))
    parts = post.particles(stride=1000)
    assert parts.shape == (5, hd.nx) and np.allclose(parts, km[-1][None, :])
    with pytest.raises(pz.PzError):
        zs.read_state()
    zs.close()


def test_filters_on_two_devices_of_one_process(pz):
    """One process, filters on device 0 and device 1 (a C host may do this): same seed, same bits."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    obs, _ = H.gen_obs(H.model_mtt_track(), 5)
    a = pz.zparticles(30000, pz.mtt_track(), seed=5, device=0)
    b = pz.zparticles(30000, pz.mtt_track(), seed=5, device=1)
    for y in obs:
        pa, pb = a.step(y), b.step(y)
        assert pa.total_weight == pb.total_weight
    assert np.array_equal(a.read_state().view(np.uint32), b.read_state().view(np.uint32))
    a.close(); b.close()


def test_multinomial_reference_mode(pz):
    """resampler = MULTINOMIAL_REF (resample2, Inference.hs:57-61): bit-exact vs the oracle and
    statistically equal to the systematic filter."""
    hd = H.model_rw1d(resampler=H.PZ_RESAMPLE_MULTINOMIAL_REF)
    d = pz.rw1d(resampler=pz.PZ_RESAMPLE_MULTINOMIAL_REF)
    n, T = 20000, 6
    obs, _ = H.gen_obs(hd, T)
    zs = pz.zparticles(n, d, seed=77)
    orc = H.OracleFilter(hd, n, 77)
    for t in range(T):
        post = zs.step(obs[t])
        rc, s = orc.step(obs[t])
        x_ref, _, _ = orc.state()
        assert np.array_equal(zs.read_state().view(np.uint32), x_ref.view(np.uint32)), t
        assert post.total_weight == s.total_weight


def test_read_particles_is_resampled_population(pz):
    d = pz.stt()
    hd = H.model_stt()
    n = 30000
    obs, _ = H.gen_obs(hd, 4)
    zs = pz.zparticles(n, d, seed=3)
    for y in obs:
        post = zs.step(y)
    x = zs.read_state()
    anc = zs.read_ancestors()
    p = post.particles(stride=7)
    assert p.shape == ((n + 6) // 7, 6)
    assert np.array_equal(p.astype(np.float32), x[:, anc[::7]].T)
    # equally weighted resampled particles reproduce the weighted posterior mean
    full = zs.read_particles(1)
    assert np.allclose(full.mean(axis=0), post.mean, atol=5 * np.sqrt(post.var / n) + 1e-3)


def test_degenerate_weights_are_reported(pz):
    d = pz.rw1d()
    zs = pz.zparticles(5000, d, seed=1)
    zs.step([0.3])
    with pytest.raises(pz.PzError) as e:
        zs.step([float("nan")])
    assert e.value.code == H.PZ_EDEGENERATE


def test_large_n_properties(pz):
    """BASELINE size (N = 10^8 would need the oracle for minutes): size-independent
    properties at N = 2^24 -- ancestors sorted in CDF order, offspring counts within 1 of N w_i, total
    mass conservation, and equality of two independent runs."""
    n = 1 << 24
    d = pz.rw1d()
    obs, _ = H.gen_obs(H.model_rw1d(), 3)
    zs = pz.zparticles(n, d, seed=5)
    zs.run(obs)
    lw = zs.read_logw().astype(np.float64)
    anc = zs.read_ancestors()
    assert anc.min() >= 0 and anc.max() < n and np.all(np.diff(H.cdf_position(anc)) >= 0)   # sorted in the spec's CDF order
    w = np.exp(lw - lw.max()); w /= w.sum()
    cnt = np.bincount(anc, minlength=n)
    assert np.all(np.abs(cnt - n * w) < 1 + 1e-3 * n * w + 1e-4)
    zs2 = pz.zparticles(n, d, seed=5)
    zs2.run(obs)
    assert np.array_equal(zs2.read_ancestors(), anc)


@pytest.mark.parametrize("f64", [False, True], ids=["f32", "f64"])
@pytest.mark.parametrize("model", ["rw1d", "mtt_track", "stt"])
def test_independent_of_tile_size_and_specialisation(pz, model, f64):
    """Launch-shape independence: the small / large output-tile instantiations and the dense /
    position-velocity specialised model code give the same bits (separate processes: the choice is
    read once from the environment)."""
    import hashlib
    import subprocess
    import sys
    code = (
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import __graft_entry__ as g, pzhelpers as H\n"
        "pz = g.load_package()\n"
        "hd = {'rw1d': H.model_rw1d, 'mtt_track': H.model_mtt_track, 'stt': H.model_stt}[%r]()\n"
        "obs, _ = H.gen_obs(hd, 5)\n"
        "zs = pz.zparticles(50000 + 77, getattr(pz, %r)(f64=%r), seed=31)\n"
        "posts = zs.run(obs)\n"
        "print(hashlib.sha256(zs.read_state().tobytes()).hexdigest(), posts[-1].total_weight)\n"
    ) % (H.ROOT, H.ROOT + "/tests", model, model, f64)
    outs = set()
    for env in ({}, {"PZ_TILE_OUT": "small"}, {"PZ_NO_KRON": "1"}, {"PZ_TILE_OUT": "small", "PZ_NO_KRON": "1"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env={**__import__("os").environ, **env})
        assert r.returncode == 0, r.stderr[-2000:]
        outs.add(r.stdout.strip().split("\n")[-1])
    assert len(outs) == 1, outs
