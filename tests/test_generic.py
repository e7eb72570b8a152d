"""The stored-log-weight path (pz_generic.cu) and the primitives around it: the MStream `particles` filter with
carried weights (Inference.hs:73-86), `resample` / `importanceResampling` (a7, Inference.hs:51-54,
68-71), the distribution primitives as device functions (a11, Distributions.hs:127-212), adaptive (ESS-triggered)
resampling and the Walker model (8f rank 4, Examples/Walker.hs), Beta-Bernoulli conjugacy under delayed sampling
(8f rank 2, DelayedSampling.hs:140,150-151).

CPU tests pin the oracle's restatements (closed forms, scipy, the exact Kalman filter); GPU tests compare the
CUDA path with the oracle through the C ABI: bit-exact for states, log-weights and ancestors of the
linear-Gaussian variants; integer-exact / 1e-12 relative for the primitives (libm's log and exp differ from
CUDA's in the last ulp); 1e-9 relative for the walker's double state."""
import ctypes as C
import math

import numpy as np
import pytest

import pzhelpers as H


# ---------------------------------------------------------------------------------------------------------
# CPU: the oracle's restatements
# ---------------------------------------------------------------------------------------------------------
def test_oracle_primitive_scores_match_closed_forms():
    from scipy import stats
    x = np.array([0.0, 1.0, 2.0, 5.0, 11.0])
    assert np.allclose(H.oracle_dist_score("poisson", [3.0], x), stats.poisson.logpmf(x, 3.0), rtol=1e-12)
    assert np.allclose(H.oracle_dist_score("bernoulli", [0.3], [0.0, 1.0]), [math.log(0.7), math.log(0.3)], rtol=1e-14)
    ps = [0.2, 0.5, 0.3]
    assert np.allclose(H.oracle_dist_score("categorical", ps, [0, 1, 2]), np.log(ps), rtol=1e-14)
    assert H.oracle_dist_score("categorical", ps, [3])[0] == -np.inf
    assert np.allclose(H.oracle_dist_score("uniform", [2.0, 7.0], [2.0, 4.5, 7.0]), -math.log(5.0))
    assert np.all(H.oracle_dist_score("uniform", [2.0, 7.0], [1.99, 7.01]) == -np.inf)
    xs = np.linspace(-3, 4, 9)
    # gaussian_ll' is TWICE the log-density (Distributions.hs:160-161)
    assert np.allclose(H.oracle_dist_score("normal", [0.5, 3.0], xs), 2 * stats.norm.logpdf(xs, 0.5, math.sqrt(3.0)), rtol=1e-12)
    assert np.allclose(H.oracle_dist_score("exponential", [1.7], [0.0, 0.4, 3.0]), stats.expon.logpdf([0.0, 0.4, 3.0], scale=1 / 1.7), rtol=1e-12)


def test_oracle_primitive_samples_have_the_right_law():
    n = 400000
    k = H.oracle_dist_sample("poisson", [3.0], 7, n)
    assert abs(k.mean() - 3.0) < 5 * math.sqrt(3.0 / n) and abs(k.var() - 3.0) < 0.03
    b = H.oracle_dist_sample("bernoulli", [0.3], 7, n)
    assert set(np.unique(b)) == {0.0, 1.0} and abs(b.mean() - 0.3) < 5 * math.sqrt(0.21 / n)
    c = H.oracle_dist_sample("categorical", [0.2, 0.5, 0.3], 9, n)
    assert np.allclose(np.bincount(c.astype(int), minlength=3) / n, [0.2, 0.5, 0.3], atol=5e-3)
    c2 = H.oracle_dist_sample("categorical", [2.0, 5.0, 3.0], 9, n)     # unnormalised weights: scaled by the total
    assert np.array_equal(c, c2)
    u = H.oracle_dist_sample("uniform", [2.0, 7.0], 3, n)
    assert u.min() > 2.0 and u.max() < 7.0 and abs(u.mean() - 4.5) < 0.01
    e = H.oracle_dist_sample("exponential", [1.7], 3, n)
    assert e.min() > 0 and abs(e.mean() - 1 / 1.7) < 0.005
    z = H.oracle_dist_sample("normal", [0.5, 3.0], 3, n)
    assert abs(z.mean() - 0.5) < 0.02 and abs(z.var() - 3.0) < 0.03


def test_oracle_importance_resampling_draws_by_weight():
    rng = np.random.default_rng(0)
    lw = rng.standard_normal(64).astype(np.float32)
    rc, idx = H.oracle_importance_resampling(np.tile(lw, 1), 5, 0, 64)
    assert rc == 0 and idx.min() >= 0 and idx.max() < 64
    # many independent single draws (importanceResampling returns ONE sample): frequencies follow the weights
    w = np.exp(lw.astype(np.float64)); w /= w.sum()
    big = np.tile(lw, 2048)                       # 131072 particles, weights w / 2048 each
    rc, draws = H.oracle_importance_resampling(big, 11, 3, 100000)
    freq = np.bincount(draws % 64, minlength=64) / draws.size
    assert rc == 0 and np.max(np.abs(freq - w)) < 6 * np.sqrt(w.max() / draws.size)


@pytest.mark.parametrize("model", ["rw1d", "lg42"])
def test_oracle_adaptive_resampling_stays_on_the_kalman_posterior(model):
    """ess_threshold = 0.5: resample only when ESS < N / 2.  The filter is still a consistent estimator of the
    Kalman posterior (weights accumulate between resamples) and really skips resamples."""
    mk = {"rw1d": H.model_rw1d, "lg42": H.model_mtt_track}[model]
    hd = mk()
    T, n, seeds = 25, 4000, 12
    obs, _ = H.gen_obs(hd, T)
    km, kc, _ = H.kalman(hd, obs)
    means = np.zeros((seeds, T, hd.nx)); skipped = 0
    for s in range(seeds):
        d = mk(); d.ess_threshold = 0.5
        f = H.OracleFilter(d, n, 100 + s)
        for t in range(T):
            rc, summ = f.step(obs[t])
            assert rc == 0
            means[s, t] = summ.mean[:hd.nx]
            skipped += 0 if summ.reserved else 1
        f.close()
    assert skipped > seeds            # some steps kept their particles
    H.assert_within_mc_tolerance(means, km, "adaptive mean")


def test_oracle_carried_weights_double_count_the_likelihood():
    """carry_weights = 1 (MStream `particles`): the children keep the parent's cumulative weight, so a heavy
    particle is favoured twice -- the effective sample size collapses relative to the canonical filter.  (This is the
    reference's non-canonical variant; the test pins that the flag changes the filter the way the
    reference's streamWeights does, not that it is a good filter.)"""
    hd = H.model_rw1d()
    obs, _ = H.gen_obs(hd, 30)
    ess = {}
    for carry in (0, 1):
        d = H.model_rw1d(); d.carry_weights = carry
        f = H.OracleFilter(d, 5000, 3)
        e = []
        for y in obs:
            rc, s = f.step(y)
            assert rc == 0
            e.append(s.ess)
        ess[carry] = np.mean(e[5:])
        f.close()
    assert ess[1] < 0.8 * ess[0]


def test_oracle_walker_tracks_the_walker():
    hd = H.model_walker()
    T = 30
    obs, truth = H.walker_obs(hd, T, seed=21)
    assert np.all(np.isin(truth[:, 4], [0, 1, 2])) and np.std(obs - truth[:, :2]) < 8.0   # measured with variance 10 (Walker.hs:96-101)
    # adaptive resampling (canonical weights) tracks the position to within the posterior spread
    d = H.model_walker(carry_weights=0, ess_threshold=0.5)
    f = H.OracleWalker(d, 20000, 5)
    err = []
    for t in range(T):
        rc, ess, _ = f.step(obs[t])
        assert rc == 0 and ess > 1
        x, lw, _ = f.state()
        w = np.exp(lw - lw.max()); w /= w.sum()
        err.append(np.hypot(w @ x[0] - truth[t, 0], w @ x[1] - truth[t, 1]))
        assert set(np.unique(x[4])) <= {0.0, 1.0, 2.0}
    assert np.median(err) < 15.0      # observation std 10 per axis
    f.close()


def test_descriptor_layout_matches_the_header():
    """ABI 2 fields sit where include/pzgpu.h puts them (compiled with the C compiler)."""
    import subprocess, tempfile, os
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "pzgpu.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(pz_model_desc), offsetof(pz_model_desc, state_f64), offsetof(pz_model_desc, ess_threshold),
         offsetof(pz_model_desc, carry_weights), offsetof(pz_model_desc, walker), offsetof(pz_model_desc, beta_a), sizeof(pz_walker_params));
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "l.c")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(H.ROOT, "include"), c, "-o", os.path.join(td, "l")])
        got = [int(v) for v in subprocess.check_output([os.path.join(td, "l")]).split()]
    D = H.ModelDesc
    assert got == [C.sizeof(D), D.state_f64.offset, D.ess_threshold.offset, D.carry_weights.offset, D.walker.offset, D.beta_a.offset,
                   C.sizeof(H.WalkerParams)]


# ---------------------------------------------------------------------------------------------------------
# GPU: the CUDA path against the oracle
# ---------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu
MK = {"rw1d": (H.model_rw1d, "rw1d"), "lg42": (H.model_mtt_track, "mtt_track"), "lg63": (H.model_stt, "stt")}


@gpu
@pytest.mark.parametrize("model", ["rw1d", "lg42", "lg63"])
def test_stored_weight_path_reproduces_the_fused_kernel(pz, model):
    """ess_threshold = 1 resamples at every step and never carries a weight: the separate-kernel path must give the
    fused kernel's population, log-weights and totals bit for bit."""
    hd = MK[model][0]()
    obs, _ = H.gen_obs(hd, 7)
    n = 30000 + 19
    a = pz.zparticles(n, getattr(pz, MK[model][1])(), seed=12)
    b = pz.zparticles(n, getattr(pz, MK[model][1])(ess_threshold=1.0), seed=12)
    for y in obs:
        pa, pb = a.step(y), b.step(y)
        assert np.array_equal(a.read_state().view(np.uint32), b.read_state().view(np.uint32))
        assert np.array_equal(a.read_logw().view(np.uint32), b.read_logw().view(np.uint32))
        assert pa.total_weight == pb.total_weight and pa.e_max == pb.e_max and pb.resampled
        assert np.allclose(pa.mean, pb.mean, rtol=1e-5, atol=1e-5) and abs(pa.log_evidence_inc - pb.log_evidence_inc) < 1e-4
    a.close(); b.close()


@gpu
@pytest.mark.parametrize("variant", ["carry", "ess", "carry+ess", "carry+multinomial"])
@pytest.mark.parametrize("model", ["rw1d", "lg42", "lg63"])
def test_carried_and_adaptive_filters_bit_exact(pz, model, variant):
    carry = "carry" in variant
    ess = 0.5 if "ess" in variant else 0.0
    rs = H.PZ_RESAMPLE_MULTINOMIAL_REF if "multinomial" in variant else H.PZ_RESAMPLE_SYSTEMATIC
    hd = MK[model][0](resampler=rs); hd.carry_weights = int(carry); hd.ess_threshold = ess
    d = getattr(pz, MK[model][1])(resampler=rs, carry_weights=carry, ess_threshold=ess)
    n, T = 20000 + 77, 12
    obs, _ = H.gen_obs(hd, T)
    zs = pz.zparticles(n, d, seed=99)
    orc = H.OracleFilter(hd, n, 99)
    kept = 0
    for t in range(T):
        anc_gpu = zs.read_ancestors()
        post = zs.step(obs[t])
        rc, s = orc.step(obs[t])
        assert rc == 0
        x_ref, lw_ref, anc_ref = orc.state()
        assert np.array_equal(anc_gpu, anc_ref), (model, variant, t)
        assert np.array_equal(zs.read_state().view(np.uint32), x_ref.view(np.uint32)), (model, variant, t)
        assert np.array_equal(zs.read_logw().view(np.uint32), lw_ref.view(np.uint32)), (model, variant, t)
        assert post.total_weight == s.total_weight and post.e_max == s.e_max
        assert post.resampled == bool(s.reserved)
        kept += 0 if post.resampled else 1
        nx = hd.nx
        assert np.allclose(post.mean, s.mean[:nx], rtol=1e-9, atol=1e-9) and np.allclose(post.var, s.var[:nx], rtol=1e-7, atol=1e-9)
        assert abs(post.ess - s.ess) < 1e-9 * s.ess and abs(post.log_evidence_inc - s.log_evidence_inc) < 1e-8
    if ess:
        assert kept > 0       # the adaptive rule really skipped resamples
    # the equally weighted output list of a no-resample step is the population itself
    zs.close(); orc.close()


@gpu
def test_adaptive_filter_matches_kalman_on_the_gpu(pz):
    hd = H.model_mtt_track()
    T, n, seeds = 25, 50000, 12
    obs, _ = H.gen_obs(hd, T)
    km, _, kl = H.kalman(hd, obs)
    means = np.zeros((seeds, T, 4)); ll = np.zeros(seeds)
    for s in range(seeds):
        zs = pz.zparticles(n, pz.mtt_track(ess_threshold=0.5), seed=500 + s)
        posts = zs.run(obs)
        means[s] = [p.mean for p in posts]
        ll[s] = sum(p.log_evidence_inc for p in posts)
        zs.close()
    H.assert_within_mc_tolerance(means, km, "adaptive mean (GPU)")
    assert abs(ll.mean() - kl.sum()) < 4 * ll.std(ddof=1) / np.sqrt(seeds) + 0.05    # the evidence of an adaptive filter


@gpu
@pytest.mark.parametrize("variant", ["particles", "adaptive", "multinomial"])
def test_walker_matches_the_oracle(pz, variant):
    carry, ess, rs = {"particles": (1, 0.0, 0), "adaptive": (0, 0.5, 0), "multinomial": (1, 0.0, 1)}[variant]
    hd = H.model_walker(carry_weights=carry, ess_threshold=ess, resampler=rs)
    d = pz.walker(resampler=rs, carry_weights=bool(carry), ess_threshold=ess)
    assert bytes(d.walker) == bytes(hd.walker)
    n, T = 20000, 8
    obs, _ = H.walker_obs(hd, T, seed=21)
    zs = pz.zparticles(n, d, seed=8)
    orc = H.OracleWalker(hd, n, 8)
    x0, _, _ = orc.state()
    assert np.array_equal(zs.read_state()[4], x0[4]) and np.allclose(zs.read_state(), x0, rtol=1e-12, atol=1e-12)   # walkerInit
    for t in range(T):
        anc_gpu = zs.read_ancestors()
        post = zs.step(obs[t])
        rc, ess_ref, rs_ref = orc.step(obs[t])
        assert rc == 0
        x_ref, lw_ref, anc_ref = orc.state()
        # CUDA's and glibc's log / sin / cos differ in the last ulp: a float log-weight may round the other way once
        # in ~10^8 draws and move a handful of ancestors; the data-dependent loop counts are equal
        assert np.mean(anc_gpu != anc_ref) < 1e-3, (variant, t)
        same = anc_gpu == anc_ref
        x = zs.read_state()
        assert np.array_equal(x[4][same], x_ref[4][same])
        assert np.allclose(x[:4, same], x_ref[:4, same], rtol=1e-9, atol=1e-9)
        assert np.allclose(zs.read_logw()[same], lw_ref[same], rtol=1e-5, atol=1e-4)
        assert abs(post.ess - ess_ref) < 1e-3 * ess_ref + 1e-6 and post.resampled == rs_ref
        w = np.exp(lw_ref.astype(np.float64) - lw_ref.max()); w /= w.sum()
        assert np.allclose(post.mean[:4], x_ref[:4] @ w, rtol=1e-3, atol=1e-2)
        assert abs(post.mean[4] - w @ (x_ref[4] == 1)) < 1e-3 and abs(post.mean[5] - w @ (x_ref[4] == 2)) < 1e-3
    parts = post.particles(stride=100)
    assert parts.shape == (n // 100, 5)
    zs.close(); orc.close()


@gpu
def test_device_primitives_match_the_oracle(pz):
    n = 100000
    cases = [("bernoulli", [0.3]), ("categorical", [0.2, 0.5, 0.3]), ("categorical", [2.0, 5.0, 3.0, 1.0, 0.5]), ("poisson", [3.0]),
             ("poisson", [0.1]), ("uniform", [2.0, 7.0]), ("exponential", [1.7]), ("normal", [0.5, 3.0])]
    for kind, p in cases:
        got, ref = pz.dist_sample(kind, p, 77, n), H.oracle_dist_sample(kind, p, 77, n)
        if kind in ("bernoulli", "categorical", "poisson"):
            assert np.array_equal(got, ref), kind
        elif kind == "uniform":
            assert np.array_equal(got, ref)          # a + (b - a) u: no transcendental
        else:
            assert np.allclose(got, ref, rtol=1e-13, atol=1e-15), kind
        x = ref[:1000] if kind != "uniform" else np.concatenate([ref[:998], [1.0, 8.0]])
        assert np.allclose(pz.dist_score(kind, p, x), H.oracle_dist_score(kind, p, x), rtol=1e-12, atol=1e-14, equal_nan=True), kind


@gpu
def test_importance_resampling_bit_exact(pz):
    rng = np.random.default_rng(4)
    lw = rng.standard_normal(5000) * 2
    for k in (1, 7, 5000):
        rc, ref = H.oracle_importance_resampling(lw.astype(np.float32), 31, 2, k)
        assert rc == 0 and np.array_equal(pz.importance_resampling(lw, 31, 2, k), ref)
    assert np.array_equal(pz.importance_resampling(lw, 31, 2, 5000), pz.resample_multinomial(lw, 31, 2))   # `resample` = the k = n case


@gpu
def test_beta_bernoulli_conjugacy(pz):
    """betaBernoulli 1 1 (cycle [True, True, False]) (DelayedSamplingAbstract.hs:57-69): MBeta (a + #true) (b + #false),
    marginal score log(a / (a + b)) or log(b / (a + b))."""
    zs = pz.zparticles(1000, pz.beta_bernoulli(1.0, 1.0), seed=0)
    a, b = 1.0, 1.0
    for t, y in enumerate([True, True, False] * 5):
        post = zs.step([1.0 if y else 0.0])
        ll = math.log(a / (a + b)) if y else math.log(b / (a + b))
        a, b = (a + 1, b) if y else (a, b + 1)
        mean, cov = zs.read_posterior()
        assert mean[0] == a and cov[0, 0] == pytest.approx(a * b / ((a + b) ** 2 * (a + b + 1)))
        assert post.step == t and post.ess == 1000 and post.log_evidence_inc == pytest.approx(ll, abs=1e-14)
        assert post.mean[0] == pytest.approx(a / (a + b))
    assert np.allclose(post.particles(stride=100), a / (a + b))
    with pytest.raises(pz.PzError):
        zs.step([0.5])
    zs.reset()
    assert zs.step([1.0]).log_evidence_inc == pytest.approx(math.log(0.5))
    zs.close()


@gpu
def test_reset_after_a_degenerate_step(pz):
    """PZ_EDEGENERATE leaves the populations overwritten; pz_filter_reset returns the handle to the prior."""
    hd = H.model_rw1d()
    obs, _ = H.gen_obs(hd, 5)
    for kw in ({}, {"f64": True}, {"ess_threshold": 0.5}):
        zs = pz.zparticles(20000, pz.rw1d(**kw), seed=3)
        zs.step(obs[0])
        with pytest.raises(pz.PzError) as e:
            zs.step([float("nan")])
        assert e.value.code == H.PZ_EDEGENERATE
        zs.reset(seed=3)
        fresh = pz.zparticles(20000, pz.rw1d(**kw), seed=3)
        for y in obs:
            pa, pb = zs.step(y), fresh.step(y)
            assert pa.total_weight == pb.total_weight and pa.step == pb.step
        assert np.array_equal(zs.read_state(), fresh.read_state())
        zs.reset(seed=4)             # a new seed gives a new trajectory
        zs.step(obs[0])
        fresh2 = pz.zparticles(20000, pz.rw1d(**kw), seed=4)
        fresh2.step(obs[0])
        assert np.array_equal(zs.read_state(), fresh2.read_state())
        zs.close(); fresh.close(); fresh2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mk,f64", [("rw1d", True), ("rw1d", False), ("stt", True), ("mtt_track", False)])
def test_steps_queued_behind_a_degenerate_step_are_harmless(pz, mk, f64):
    """pz_filter_run queues its steps without a host check in between: a step whose weights are all zero leaves a plan that
    covers no output slot, and the step kernels behind it (which gather parents by index) must not touch memory through it.
    The run reports PZ_EDEGENERATE, the context stays healthy and pz_filter_reset returns the handle to the prior."""
    d = getattr(pz, mk)(f64=f64)
    hd = {"rw1d": H.model_rw1d, "stt": H.model_stt, "mtt_track": H.model_mtt_track}[mk]()
    obs, _ = H.gen_obs(hd, 12)
    bad = obs.copy(); bad[3] = float("nan")
    zs = pz.zparticles(100000, d, seed=3)
    with pytest.raises(pz.PzError) as e:
        zs.run(bad)
    assert e.value.code == H.PZ_EDEGENERATE
    zs.reset(seed=3)
    fresh = pz.zparticles(100000, d, seed=3)
    a, b = zs.run(obs), fresh.run(obs)
    assert [p.total_weight for p in a] == [p.total_weight for p in b]
    assert np.array_equal(zs.read_state(), fresh.read_state())
    zs.close(); fresh.close()
