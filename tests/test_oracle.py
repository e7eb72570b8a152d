"""CPU tests of the oracle (the checker): known-answer vectors, math accuracy, resampler
invariants, and the exact-Kalman pin (the reference has no golden vectors, so
the oracle is pinned against Random123 KATs, libm and the reference's own Kalman algebra)."""
import ctypes as C

import numpy as np
import pytest

import pzhelpers as H


def philox(ctr, key):
    lib = H.oracle()
    c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
    lib.pzo_philox4x32_10(H.ptr(c, C.c_uint32), H.ptr(k, C.c_uint32), H.ptr(o, C.c_uint32))
    return [int(v) for v in o]


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_golden_fixture():
    g = np.load(H.ROOT + "/tests/golden/philox_normal_kat.npz")
    lib = H.oracle()
    r = g["r"]
    z = np.zeros(r.size, np.float32)
    lib.pzo_icdf_normals(H.ptr(r, C.c_uint32), r.size, H.ptr(z, C.c_float))
    assert np.array_equal(z.view(np.uint32), g["z"].view(np.uint32))
    q, w, eb = H.block_weights(g["lw"])
    assert np.array_equal(q, g["q"]) and np.array_equal(w.view(np.uint32), g["w"].view(np.uint32)) and np.array_equal(eb, g["eb"])


def test_icdf_normal_against_ndtri():
    """One Philox word -> N(0,1): the piecewise-cubic inverse CDF against scipy's ndtri."""
    from scipy.special import ndtri
    lib = H.oracle()
    rng = np.random.default_rng(0)
    r = rng.integers(0, 2 ** 32, 400000, dtype=np.uint32)
    r[:8] = [0, 1, 2, 2 ** 31 - 1, 2 ** 31, 2 ** 31 + 1, 2 ** 32 - 1, 0x40000000]
    # every dyadic segment, including the far tails random words never reach
    tails = np.array([(1 << s) | (rng.integers(0, 1 << s) if s else 0) for s in range(31) for _ in range(64)], dtype=np.uint32)
    r = np.concatenate([r, tails, tails | np.uint32(0x80000000)])
    z = np.zeros(r.size, np.float32)
    lib.pzo_icdf_normals(H.ptr(r, C.c_uint32), r.size, H.ptr(z, C.c_float))
    # the tail probability the spec assigns to the word: the top 3 + 23 bits below the leading one of v = (r << 1) | 1
    v = ((r.astype(np.uint64) << 1) | 1) & 0xFFFFFFFF
    k = 31 - np.floor(np.log2(v.astype(np.float64))).astype(np.int64)
    vn = (v << k.astype(np.uint64)) & 0xFFFFFFFF
    frac = ((vn >> 5) & 0x3FFFFFF).astype(np.float64) / 2.0 ** 26      # 26 bits below the leading one
    u = (1.0 + frac) * 2.0 ** -(k + 2.0)
    ref = -ndtri(u) * np.where(r >> 31, -1.0, 1.0)
    assert np.max(np.abs(z - ref)) < 1.5e-6
    assert np.all(np.isfinite(z)) and np.max(np.abs(z)) < 6.35
    zz = z[:400000].astype(np.float64)
    assert abs(zz.mean()) < 0.006 and abs(zz.var() - 1) < 0.008 and abs((zz ** 4).mean() - 3) < 0.06
    assert abs(np.mean(zz < -1.0) - 0.158655) < 0.003


def test_block_weights_against_libm():
    rng = np.random.default_rng(1)
    lw = np.concatenate([-rng.random(100000) * 9, [0.0, -1e-30, -0.5, -5.0]]).astype(np.float32)
    q, w, eb = H.block_weights(lw)
    scale = 2.0 ** eb[np.arange(lw.size) // 256].astype(np.float64)
    rel = np.abs(w.astype(np.float64) * scale / np.exp(lw.astype(np.float64)) - 1)
    assert np.max(rel) < 2e-6                    # fp32 weight relative to exp(lw)
    assert np.all(np.abs(q.astype(np.float64) - w.astype(np.float64) * 32768.0) <= 0.5)   # q = rint(w * 2^15)
    bmax = np.array([q[b * 256:(b + 1) * 256].max() for b in range(eb.size)])
    assert bmax.min() >= 22800 and bmax.max() <= 46990   # block maximum in [2^14.48, 2^15.52]
    # special values -> zero weight; a block of them is empty
    sp = np.array([-np.inf, np.inf, np.nan, -np.inf] * 64, np.float32)
    q, w, eb = H.block_weights(sp)
    assert np.all(q == 0) and eb[0] == -(1 << 30)
    # weights far below the block maximum flush to zero instead of wrapping
    lw = np.zeros(256, np.float32); lw[1:] = -np.arange(1, 256, dtype=np.float32) * 3
    q, w, eb = H.block_weights(lw)
    assert q[0] == 32768 and np.all(np.diff(q.astype(np.int64)) <= 0) and q[-1] == 0 and np.all(w > 0)


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 1000, 4097, 50000])
def test_systematic_invariants(n):
    rng = np.random.default_rng(n)
    lw = (rng.standard_normal(n) * 4).astype(np.float32)
    for U in [0, 1, 12345, 2 ** 31, 2 ** 32 - 1]:
        rc, anc = H.oracle_systematic(lw, U)
        assert rc == 0
        assert anc.min() >= 0 and anc.max() < n
        assert np.all(np.diff(H.cdf_position(anc)) >= 0)  # systematic ancestors are sorted in CDF order
        # offspring counts within 1 of N * w_i (the defining property of systematic resampling);
        # the weights carry 15 bits relative to their block's maximum
        w = np.exp(lw.astype(np.float64) - lw.max())
        w /= w.sum()
        cnt = np.bincount(anc, minlength=n)
        q, _, eb = H.block_weights(lw)
        wq = q.astype(np.float64) * 2.0 ** (eb[np.arange(n) // 256] - eb.max()).astype(np.float64)
        wq /= wq.sum()
        assert np.all(np.abs(cnt - n * wq) < 1 + 1e-6 * n)
        assert np.all(np.abs(cnt - n * w) < 1.05 + 1e-3 * n * w)


def test_systematic_matches_f64_definition():
    """The fixed-point resampler agrees with the floating-point thresholds (u+j)/N definition
    (monad-bayes Population.systematic lineage) except at near-ties."""
    lib = H.oracle()
    rng = np.random.default_rng(3)
    n = 20480   # whole blocks: CDF position == rank in the spec's order
    lw = (rng.standard_normal(n) * 2).astype(np.float32)
    U = 987654321
    rc, anc = H.oracle_systematic(lw, U)
    a2 = np.zeros(n, np.int32)
    order = np.argsort(H.cdf_position(np.arange(n)))       # the f64 walk in the spec's CDF order
    lwd = lw.astype(np.float64)[order]
    lib.pzo_resample_systematic_f64(H.ptr(lwd, C.c_double), n, U / 2.0 ** 32, H.ptr(a2, C.c_int32))
    pa = H.cdf_position(anc)
    assert np.mean(pa != a2) < 3e-2   # near-ties only: the weights carry 15 bits relative to the block maximum
    assert np.max(np.abs(pa - a2)) <= 4


def test_systematic_edge_cases():
    lw = np.full(1000, -np.inf, np.float32)
    lw[777] = -3.0
    rc, anc = H.oracle_systematic(lw, 4242)
    assert rc == 0 and np.all(anc == 777)  # one-hot
    rc, _ = H.oracle_systematic(np.full(10, -np.inf, np.float32), 5)
    assert rc == H.PZ_EDEGENERATE
    rc, _ = H.oracle_systematic(np.full(10, np.nan, np.float32), 5)
    assert rc == H.PZ_EDEGENERATE
    # equal weights: every particle survives exactly once for a mid-range uniform
    rc, anc = H.oracle_systematic(np.zeros(5000, np.float32), 2 ** 31)
    assert rc == 0 and np.array_equal(np.sort(anc), np.arange(5000))
    # huge dynamic range across blocks
    lw = np.full(2048, -600.0, np.float32)
    lw[5] = 0.0
    lw[1500] = -0.6931472
    rc, anc = H.oracle_systematic(lw, 1)
    cnt = np.bincount(anc, minlength=2048)
    assert cnt[5] + cnt[1500] == 2048 and abs(cnt[5] - 2048 * 2 / 3) <= 1


def test_multinomial_distribution():
    rng = np.random.default_rng(5)
    n = 4000
    lw = (rng.standard_normal(n)).astype(np.float32)
    w = np.exp(lw.astype(np.float64)); w /= w.sum()
    cnt = np.zeros(n)
    reps = 50
    for r in range(reps):
        rc, anc = H.oracle_multinomial(lw, 99, r)
        assert rc == 0 and anc.min() >= 0 and anc.max() < n
        cnt += np.bincount(anc, minlength=n)
    # chi-square-ish: standardised residuals of multinomial counts
    z = (cnt - reps * n * w) / np.sqrt(reps * n * w)
    assert abs(z.mean()) < 0.1 and 0.85 < z.std() < 1.15


def test_kalman_matches_closed_form_rw1d():
    # steady-state variance of x' ~ N(x,1), y ~ N(x', 1.5): P solves P^2 + P - 1.5 = 0
    d = H.model_rw1d()
    obs, _ = H.gen_obs(d, 200)
    means, covs, _ = H.kalman(d, obs)
    P = (-1 + np.sqrt(1 + 6)) / 2
    assert abs(covs[-1, 0, 0] - P) < 1e-12
    # gaussianConditioning (DelayedSampling.hs:88-95) for the first step: prior N(0,1), obs var 1.5
    mu1 = (1 / 1.0 * 0 + 1 / 1.5 * obs[0, 0]) / (1 / 1.0 + 1 / 1.5)
    assert abs(means[0, 0] - mu1) < 1e-12 and abs(covs[0, 0, 0] - 1 / (1 + 1 / 1.5)) < 1e-12


@pytest.mark.parametrize("model", ["rw1d", "rw1d_1x", "mtt_track", "stt"])
def test_canonical_filter_matches_kalman(model):
    """Posterior mean/variance of the canonical filter within 3 standard errors of the exact
    Kalman posterior, SE estimated over 20 seeds (gate 2)."""
    d = {"rw1d": H.model_rw1d, "rw1d_1x": lambda: H.model_rw1d(scalar_ll_2x=0), "mtt_track": H.model_mtt_track,
         "stt": H.model_stt}[model]()
    T, n, seeds = 25, 4000, 20
    obs, _ = H.gen_obs(d, T)
    km, kc, kl = H.kalman(d, obs)
    nx = d.nx
    means = np.zeros((seeds, T, nx)); vars_ = np.zeros((seeds, T, nx)); lls = np.zeros((seeds, T))
    for s in range(seeds):
        f = H.OracleFilter(d, n, 100 + s)
        for t in range(T):
            rc, summ = f.step(obs[t])
            assert rc == 0
            means[s, t] = summ.mean[:nx]; vars_[s, t] = summ.var[:nx]; lls[s, t] = summ.log_evidence_inc
        f.close()
    se_m = means.std(axis=0, ddof=1) / np.sqrt(seeds)
    zm = (means.mean(axis=0) - km) / np.maximum(se_m, 1e-12)
    kvar = np.array([np.diag(c) for c in kc])
    se_v = vars_.std(axis=0, ddof=1) / np.sqrt(seeds)
    zv = (vars_.mean(axis=0) - kvar) / np.maximum(se_v, 1e-12)
    # 3 SE per statistic; allow the expected handful of >3 sigma events among T*nx*2 tests
    H.assert_within_mc_tolerance(means, km, "mean")
    H.assert_within_mc_tolerance(vars_, kvar, "var", rel_bias_allow=100.0 / n)
    if model != "rw1d":  # the doubled scalar score is not a normalised density
        se_l = lls.sum(axis=1).std(ddof=1) / np.sqrt(seeds)
        assert abs(lls.sum(axis=1).mean() - kl.sum()) < 4 * se_l + 0.05


def test_refpf_faithful_matches_kalman_and_fast_mode():
    """The reference-faithful multinomial filter (Inference.hs:57-66,101-104) in double: the
    O(N^2) CDF walk and the O(N log N) search give identical ancestors, and the filter tracks
    the exact posterior."""
    lib = H.oracle()
    d = H.model_rw1d()
    T, n = 30, 1000
    obs, _ = H.gen_obs(d, T)
    km, kc, _ = H.kalman(d, obs)
    m0, v0, m1, v1 = (np.zeros((T, 1)) for _ in range(4))
    assert lib.pzo_refpf_run(C.byref(d), n, 11, H.ptr(obs, C.c_double), T, 0, H.ptr(m0, C.c_double), H.ptr(v0, C.c_double), 1) == 0
    assert lib.pzo_refpf_run(C.byref(d), n, 11, H.ptr(obs, C.c_double), T, 1, H.ptr(m1, C.c_double), H.ptr(v1, C.c_double), 1) == 0
    assert np.allclose(m0, m1, atol=1e-9)
    assert np.max(np.abs(m0 - km)) < 6 * np.sqrt(kc[-1, 0, 0] / n) * 2.5


def test_refpf_optimised_cpu_mode_tracks_the_exact_posterior():
    """Mode 2 of the CPU port (the "optimised CPU" baseline row of SURVEY.md 8(d)(ii): max-shifted weights,
    one O(N) systematic resample over a chunked parallel prefix sum) is a correct filter, for every model and
    for any thread count (the chunking must not change which ancestors are drawn)."""
    lib = H.oracle()
    for d, n, T in ((H.model_rw1d(), 4000, 25), (H.model_mtt_track(), 6000, 12), (H.model_stt(), 6000, 12)):
        obs, _ = H.gen_obs(d, T)
        km, kc, _ = H.kalman(d, obs)
        nx = d.nx
        runs = []
        for threads in (1, 3):
            m, v = np.zeros((T, nx)), np.zeros((T, nx))
            assert lib.pzo_refpf_run(C.byref(d), n, 21, H.ptr(obs, C.c_double), T, 2, H.ptr(m, C.c_double), H.ptr(v, C.c_double), threads) == 0
            runs.append(m)
            sd = np.sqrt(np.array([np.diag(kc[t])[:nx] for t in range(T)]))
            assert np.max(np.abs(m - km[:, :nx]) / sd) < 6 * 2.5 / np.sqrt(n) * np.sqrt(nx) * 2
        assert np.allclose(runs[0], runs[1], rtol=0, atol=1e-9)


def test_canonical_vs_refpf_distribution():
    """PF-versus-PF (gate 3): canonical fp32 systematic filter vs the
    reference-faithful double multinomial filter agree within Monte Carlo error."""
    lib = H.oracle()
    d = H.model_rw1d()
    T, n, seeds = 15, 2000, 16
    obs, _ = H.gen_obs(d, T)
    a = np.zeros((seeds, T)); b = np.zeros((seeds, T))
    for s in range(seeds):
        f = H.OracleFilter(d, n, 500 + s)
        for t in range(T):
            _, summ = f.step(obs[t]); a[s, t] = summ.mean[0]
        f.close()
        m = np.zeros((T, 1))
        lib.pzo_refpf_run(C.byref(d), n, 900 + s, H.ptr(obs, C.c_double), T, 1, H.ptr(m, C.c_double), None, 1)
        b[s] = m[:, 0]
    se = np.sqrt(a.var(axis=0, ddof=1) / seeds + b.var(axis=0, ddof=1) / seeds)
    z = (a.mean(axis=0) - b.mean(axis=0)) / se
    assert np.max(np.abs(z)) < 4.5 and np.mean(np.abs(z) > 3) < 0.15
