"""MultiCamera tracker (Examples/MultiCamera.hs:40-227): the CPU restatement against closed forms, the CUDA
kernel against the restatement (exactly where the association is forced, statistically elsewhere), the C ABI around it."""
import ctypes as C

import numpy as np
import pytest

import pzhelpers as H


def _gpu_model(pz, d):
    m = pz.multicam()
    m.mtt.clutter_lambda = d.mtt.clutter_lambda
    return m


def _forced_stream(T, seed=3):
    """One track per camera from step 0 on, no clutter, no deaths: every association is forced (one observation per camera),
    so the posterior of each box is the scalar Kalman recursion and each appearance follows recursive least squares."""
    rng = np.random.default_rng(seed)
    app = rng.normal(size=(2, 10))
    box = np.array([[0.3, -0.2, 1.1, 0.9], [-0.5, 0.4, 0.8, 1.2]])
    rows = []
    for t in range(T):
        box = box + rng.normal(size=(2, 4)) * np.sqrt([1, 1, 0.1, 0.1])
        obs = np.zeros((2, 16))
        for c in range(2):
            pose = rng.normal(size=10); pose /= np.linalg.norm(pose)
            obs[c, 0] = c; obs[c, 1:11] = pose; obs[c, 11:15] = box[c] + rng.normal(size=4); obs[c, 15] = pose @ app[c] + rng.normal()
        rows.append(obs)
    return rows


def _forced_desc():
    d = H.model_multicam()
    d.mtt.survival_prob = 1.0      # no deaths
    d.mtt.clutter_lambda = 1.0
    return d


def _closed_form(rows, births):
    """Box: 4 independent scalar Kalman filters (F = H = 1); appearance: RLS.  births[c] = step at which camera c's track
    appeared (its first update follows the prior directly: a newborn track is not predicted in its birth step)."""
    q = np.array([1, 1, 0.1, 0.1]); out = {}
    for c in range(2):
        m = np.array([0, 0, 1, 1.0]); v = np.array([1, 1, 0.2, 0.2]); am = np.zeros(10); aP = np.eye(10)
        for t in range(births[c], len(rows)):
            if t > births[c]:
                v = v + q
            z = rows[t][c, 11:15]
            K = v / (v + 1); m = m + K * (z - m); v = v - K * v
            h = rows[t][c, 1:11]; Ph = aP @ h; s = h @ Ph + 1
            am = am + Ph / s * (rows[t][c, 15] - h @ am); aP = aP - np.outer(Ph, Ph) / s
        out[c] = (m, v, am, aP)
    return out


def test_oracle_forced_association_is_kalman_and_rls():
    """Particles that carry exactly one track per camera (born at step 0) must hold the closed-form posteriors."""
    rows = _forced_stream(12)
    d = _forced_desc()
    d.mtt.new_track_lambda = 3.0   # births are likely, so some particles start with one track in each camera
    orc = H.OracleMcam(d, 4000, 7)
    rc, m, v, lz = orc.step(rows[0])
    assert rc == 0
    d2 = orc  # keep stepping with the same filter: further births make a camera hold 2 tracks for 1 observation -> weight 0
    for t in range(1, 12):
        rc, m, v, lz = orc.step(rows[t])
        assert rc == 0
    ref = _closed_form(rows, {0: 0, 1: 0})
    lw = orc.logw()
    found = 0
    for i in np.flatnonzero(np.isfinite(lw))[:200]:
        tracks, na = orc.particle(int(i))
        if len(tracks) != 2 or {c for _, _, c, _, _ in tracks} != {0, 1}:
            continue
        if tracks[0][1] == tracks[1][1]:
            continue   # both tracks share one appearance: not the independent closed form
        for tid, idn, cam, mean, cov in tracks:
            rm, rv, ram, raP = ref[cam]
            assert np.allclose(mean, rm, atol=1e-9) and np.allclose(np.diag(cov), rv, atol=1e-9)
            assert np.allclose(cov - np.diag(np.diag(cov)), 0, atol=1e-12)   # the box covariance stays diagonal
            am, aP = orc.appearance(int(i), idn)
            assert np.allclose(am, ram, atol=1e-9) and np.allclose(aP, raP, atol=1e-9)
        found += 1
    assert found > 0
    orc.close()


def test_oracle_generator_and_filter_track_the_scene():
    d, obs, truth = H.mcam_scenario(40)
    assert all(o.shape[1] == 16 for o in obs) and all(set(np.unique(o[:, 0])) <= {0.0, 1.0} for o in obs if len(o))
    for o, tr in zip(obs, truth):   # every live track is observed (pd = 1): at least as many rows as tracks, per camera
        for c in (0, 1):
            assert (o[:, 0] == c).sum() >= sum(1 for _, _, cam, _ in tr if cam == c)
    orc = H.OracleMcam(d, 4000, 2)
    cnt = []
    for o in obs:
        rc, m, v, lz = orc.step(o)
        assert rc == 0 and np.isfinite(lz)
        cnt.append(m)
    true_cnt = np.array([len(t) for t in truth])
    assert abs(np.mean(cnt[-15:]) - true_cnt[-15:].mean()) < 1.0
    orc.close()


def test_reference_constants_starve_the_population():
    """At the reference's amtClutter = 1 the model explains clutter rows by newborn tracks (the pose of a track hypothesis is
    scored by a constant, e^11 above the clutter's Gaussian pose density) which the following step cannot feed: the
    population dies out.  Stated as a property of the model, not of this build."""
    d, obs, truth = H.mcam_scenario(40, clutter=1.0)
    orc = H.OracleMcam(d, 3000, 2)
    rcs = [orc.step(o)[0] for o in obs]
    assert -4 in rcs   # PZ_EDEGENERATE
    orc.close()


def test_descriptor_matches_reference_constants(pz):
    mine, ref = pz.multicam(), H.model_multicam()
    assert (mine.kind, mine.nx, mine.ny) == (ref.kind, ref.nx, ref.ny) == (6, 4, 0)
    for a, b in zip(H.desc_arrays(mine), H.desc_arrays(ref)):
        assert np.array_equal(a, b)
    assert list(mine.x0)[:4] == [0, 0, 1, 1]
    for k in ("clutter_lambda", "new_track_lambda", "pd", "survival_prob", "clutter_var", "birth_pos_var", "birth_vel_var", "meas_var", "k_max", "m_max"):
        assert getattr(mine.mtt, k) == getattr(ref.mtt, k), k


@pytest.mark.gpu
def test_gpu_forced_association_matches_closed_form_and_oracle(pz):
    """Where nothing is left to chance the kernel must reproduce the closed forms (fp32 tolerance) and the oracle's evidence."""
    rows = _forced_stream(12)
    d = pz.multicam()
    d.mtt.survival_prob = 1.0
    d.mtt.new_track_lambda = 3.0
    n = 20000
    zs = pz.zparticles(n, d, seed=7)
    for t in range(12):
        post = zs.step(rows[t])
        assert np.isfinite(post.log_evidence_inc)
    ref = _closed_form(rows, {0: 0, 1: 0})
    parts = zs.read_mcam_tracks(stride=7, appearances=True)
    found = 0
    for p in parts:
        tr = p["tracks"]
        if len(tr) != 2 or {c for _, _, c, _, _, _ in tr} != {0, 1} or tr[0][1] == tr[1][1] or any(s != 0.0 for _, _, _, s, _, _ in tr):
            continue
        for tid, idn, cam, start, mean, var in tr:
            rm, rv, ram, raP = ref[cam]
            assert np.allclose(mean, rm, atol=2e-4) and np.allclose(var, rv, atol=2e-5)
            am, aP = p["app"][idn]
            assert np.allclose(am, ram, atol=5e-4) and np.allclose(aP, raP, atol=5e-5) and np.allclose(aP, aP.T)
        found += 1
    assert found > 0   # after 12 steps the survivors are the particles born with one track per camera
    zs.close()


@pytest.mark.gpu
def test_gpu_multicam_matches_cpu_restatement(pz):
    """Statistical parity on a generated scene (different uniforms): expected track count curve and stream log-evidence."""
    d, obs, truth = H.mcam_scenario(40)
    n, seeds = 20000, 5
    gpu = np.zeros((seeds, len(obs))); cpu = np.zeros((seeds, len(obs)))
    glz = np.zeros((seeds, len(obs))); clz = np.zeros((seeds, len(obs)))
    for s in range(seeds):
        zs = pz.zparticles(n, _gpu_model(pz, d), seed=20 + s)
        orc = H.OracleMcam(d, n // 5, 200 + s)
        for t, o in enumerate(obs):
            post = zs.step(o)
            rc, m, v, lz = orc.step(o)
            assert rc == 0
            gpu[s, t], cpu[s, t] = post.mean[0], m
            glz[s, t], clz[s, t] = post.log_evidence_inc, lz
        zs.close(); orc.close()
    se = np.sqrt(gpu.var(axis=0, ddof=1) / seeds + cpu.var(axis=0, ddof=1) / seeds) + 0.03
    z = (gpu.mean(axis=0) - cpu.mean(axis=0)) / se
    assert np.mean(np.abs(z) > 3) < 0.1 and np.max(np.abs(z)) < 6, z
    a, b = glz.sum(axis=1), clz.sum(axis=1)
    assert abs(a.mean() - b.mean()) < 4 * np.sqrt(a.var(ddof=1) / seeds + b.var(ddof=1) / seeds) + 0.5, (a, b)
    true_cnt = np.array([len(t) for t in truth])
    assert abs(gpu.mean(axis=0)[-15:].mean() - true_cnt[-15:].mean()) < 1.0


@pytest.mark.gpu
def test_gpu_multicam_resampling_readout_and_errors(pz, monkeypatch):
    d, obs, truth = H.mcam_scenario(25)
    n = 30000
    zs = pz.zparticles(n, _gpu_model(pz, d), seed=4)
    lib = H.oracle()
    for t, o in enumerate(obs):
        post = zs.step(o)
        if t in (5, 24):   # the tracker's stored log-weights go through the canonical resampler: ancestors bit-exact
            lw = zs.read_logw(); anc = zs.read_ancestors()
            rc, ref = H.oracle_systematic(lw, lib.pzo_resample_uniform(4, t + 1))
            assert rc == 0 and np.array_equal(anc, ref), t
    parts = zs.read_mcam_tracks(stride=100, appearances=True)
    assert len(parts) == n // 100
    cnts = np.array([len(p["tracks"]) for p in parts])
    assert abs(cnts.mean() - post.mean[0]) < 0.6
    for p in parts:
        ids = [tid for tid, *_ in p["tracks"]]
        assert ids == sorted(ids) and len(set(ids)) == len(ids)
        for tid, idn, cam, start, mean, var in p["tracks"]:
            assert 0 <= idn < p["n_app"] and cam in (0, 1) and np.all(np.isfinite(mean)) and np.all(var > 0)
        for am, aP in p["app"]:
            assert np.allclose(aP, aP.T) and np.all(np.linalg.eigvalsh(aP) > -1e-4) and np.all(np.diag(aP) <= 1 + 1e-5)
    with pytest.raises(pz.PzError):
        zs.read_tracks(stride=100)          # the MTT reader does not apply
    bad = obs[-1].copy(); bad[0, 0] = 2.0
    with pytest.raises(pz.PzError):
        zs.step(bad)                        # camera id out of range
    with pytest.raises(pz.PzError):
        zs.step(np.zeros((33, 16)))         # more rows than m_max
    zs.close()
    # the sharded path (3 shards sharing device 0) reproduces the unsharded run bit for bit
    monkeypatch.setenv("PZ_MULTI_SAME_DEVICE", "1")
    one = pz.zparticles(20000, _gpu_model(pz, d), seed=9)
    many = pz.zparticles_multi(20000, _gpu_model(pz, d), [0, 0, 0], seed=9)
    for o in obs[:12]:
        a, b = one.step(o), many.step(o)
        assert (a.total_weight, a.e_max, a.ess, a.log_evidence_inc, a.mean[0]) == (b.total_weight, b.e_max, b.ess, b.log_evidence_inc, b.mean[0])
    pa, pb = one.read_mcam_tracks(stride=53), many.read_mcam_tracks(stride=53)
    assert [[x[:4] for x in p["tracks"]] for p in pa] == [[x[:4] for x in p["tracks"]] for p in pb]
    assert all(np.array_equal(x[4], y[4]) for p, q in zip(pa, pb) for x, y in zip(p["tracks"], q["tracks"]))
    one.close(); many.close()


@pytest.mark.gpu
def test_multi_device_multicam_equals_single_device(pz):
    """The MultiCamera tracker over 2 real devices of one process (parent records read over NVLink) equals the one-device run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d, obs, truth = H.mcam_scenario(20)
    one = pz.zparticles(30000, _gpu_model(pz, d), seed=6)
    many = pz.zparticles_multi(30000, _gpu_model(pz, d), [0, 1], seed=6)
    for o in obs:
        a, b = one.step(o), many.step(o)
        assert (a.total_weight, a.e_max, a.ess, a.log_evidence_inc, a.mean[0]) == (b.total_weight, b.e_max, b.ess, b.log_evidence_inc, b.mean[0])
    pa, pb = one.read_mcam_tracks(stride=61, appearances=True), many.read_mcam_tracks(stride=61, appearances=True)
    for p, q in zip(pa, pb):
        assert [x[:4] for x in p["tracks"]] == [x[:4] for x in q["tracks"]] and p["n_app"] == q["n_app"]
        assert all(np.array_equal(x[4], y[4]) and np.array_equal(x[5], y[5]) for x, y in zip(p["tracks"], q["tracks"]))
        assert all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(p["app"], q["app"]))
    one.close(); many.close()


def _first_step_logweight(d, rows):
    """Closed form of the step-0 log-weight of a particle born with exactly one track in each camera (no clutter row):
    per camera log N(box; prior mean, prior var + 1) - logNSphereVolume(10) + log N(reading; 0, |pose|^2 + 1), the count prior
    -(lgamma(nT + nC + 1) - lgamma(nC + 1)) with nT = 1, nC = 0, and the Poisson(clutter) probability of no clutter row."""
    from scipy.special import gammaln
    from scipy.stats import norm
    pm = np.array([0, 0, 1, 1.0]); pv = np.array([d.mtt.birth_pos_var, d.mtt.birth_pos_var, d.mtt.birth_vel_var, d.mtt.birth_vel_var])
    log_vol = np.log(10.0) + 5.0 * np.log(np.pi) - gammaln(6.0)
    lw = 0.0
    for c in range(2):
        r = rows[c]
        lw += norm.logpdf(r[11:15], pm, np.sqrt(pv + d.mtt.meas_var)).sum() - log_vol
        lw += norm.logpdf(r[15], 0.0, np.sqrt(r[1:11] @ r[1:11] + 1.0))
        lw += -(gammaln(2.0) - gammaln(1.0)) - d.mtt.clutter_lambda
    return lw


def test_oracle_first_step_logweight_against_scipy():
    """Pins every term of the score (box predictive, sphere constant, appearance predictive, count prior, clutter count)."""
    rows = _forced_stream(1)[0]
    d = _forced_desc(); d.mtt.new_track_lambda = 2.0
    orc = H.OracleMcam(d, 3000, 11)
    assert orc.step(rows)[0] == 0
    lw, ref, found = orc.logw(), _first_step_logweight(d, rows), 0
    for i in range(3000):
        tracks, na = orc.particle(i)
        if len(tracks) == 2 and sorted(c for _, _, c, _, _ in tracks) == [0, 1] and tracks[0][1] != tracks[1][1]:
            # (two tracks of ONE identity share the appearance: the second camera then scores against the updated marginal)
            assert abs(lw[i] - ref) < 1e-5 * abs(ref), (lw[i], ref)   # (the oracle returns its log-weights as fp32)
            found += 1
    assert found > 10
    orc.close()


@pytest.mark.gpu
def test_gpu_first_step_logweight_against_scipy(pz):
    rows = _forced_stream(1)[0]
    d = pz.multicam(); d.mtt.survival_prob = 1.0; d.mtt.new_track_lambda = 2.0
    zs = pz.zparticles(3000, d, seed=11)
    zs.step(rows)
    lw = zs.read_logw()
    # the pending population (pre-resample) is what the log-weights belong to: find its one-track-per-camera particles by weight
    ref = _first_step_logweight(d, rows)
    close = np.abs(lw - ref) < 2e-3
    assert close.sum() > 10   # P(two births, one per camera, distinct identities) = Poisson(2; 2) / 2 / 2 = 0.068: ~200 of 3000
    # every other finite log-weight belongs to a different configuration and differs by far more than fp32 rounding
    others = lw[np.isfinite(lw) & ~close]
    assert others.size == 0 or np.min(np.abs(others - ref)) > 0.05
    zs.close()


def test_json_line_is_the_aeson_encoding_of_the_reference_triple(pz):
    """(groundTruth, particles, obs) as `encode` writes it (MultiCamera.hs:221-227): generic record objects in declaration
    order, bare arrays for R n, {"mean","cov"} for a marginal, 4-tuples as arrays."""
    import json
    d, obs, truth = H.mcam_scenario(30)
    t = max(range(30), key=lambda i: len(truth[i]))
    parts = [{"tracks": [(3, 0, 1, 2.0, np.array([0.1, 0.2, 1.0, 1.1]), np.array([0.5, 0.5, 0.2, 0.2]))], "n_app": 1},
             {"tracks": [], "n_app": 0}]
    line = pz.json_line_mcam(truth[t], parts, obs[t])
    arr = json.loads(line)
    assert len(arr) == 3 and len(arr[0]) == len(truth[t]) and len(arr[1]) == 2 and len(arr[2]) == len(obs[t])
    for g, (tid, idn, cam, box) in zip(arr[0], truth[t]):
        assert list(g.keys()) == ["posWH", "startTime", "trackID", "identity", "camera"]
        assert (g["trackID"], g["identity"], g["camera"]) == (tid, idn, cam) and np.allclose(g["posWH"], box, rtol=1e-8)
    m = arr[1][0][0]
    assert list(m.keys()) == ["posWH", "startTime", "trackID", "identity", "camera"] and m["startTime"] == 2
    assert m["posWH"]["mean"] == [0.1, 0.2, 1, 1.1] and np.array_equal(np.diag(m["posWH"]["cov"]), [0.5, 0.5, 0.2, 0.2])
    assert arr[1][1] == []
    for o, row in zip(arr[2], obs[t]):
        assert o[0] == int(row[0]) and len(o[1]) == 10 and len(o[2]) == 4 and len(o[3]) == 1 and np.allclose(o[2], row[11:15], rtol=1e-8)
