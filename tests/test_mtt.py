"""Multi-target tracker (Examples/MultiTargetTracking.hs:26-145).

CPU: the oracle's term-by-term weight equals the reduced form the kernel computes; a
single-target / no-clutter scenario collapses to the exact Kalman filter of the (4,2) track
model; the JSON line has the shape App.java:347-399 reads.
GPU (-m gpu): the CUDA tracker against the CPU restatement on the same detection stream
(statistical: different RNG streams and fp32 vs fp64), its resampling against the oracle's
canonical resampler on the device's own log-weights (bit-exact), and the track read-out."""
import json

import numpy as np
import pytest

import pzhelpers as H


def test_oracle_weight_reduction_and_tracking():
    d, obs, truth = H.mtt_scenario(60)
    f = H.OracleMtt(d, 2000, 3)
    cnt = []
    for o in obs:
        rc, m, v, lz = f.step(o)
        assert rc == 0 and np.isfinite(lz)
        cnt.append(m)
    # -log q + log prior + log likelihood == sum_j log L_j + n_undet log(1 - ps pd) - (lc + lb) - log M!
    assert f.simplification_error() < 1e-9
    true_cnt = np.array([len(t) for t in truth])
    # over the last 20 steps the filter's expected track count follows the truth
    assert abs(np.mean(cnt[-20:]) - np.mean(true_cnt[-20:])) < 1.0
    f.close()


def test_oracle_single_target_reduces_to_kalman():
    """pd = 1, (almost) no clutter or births after the first detection: the surviving track's
    Rao-Blackwellised posterior is the exact Kalman posterior of the track model started from the
    birth prior conditioned on the first detection."""
    d = H.model_mtt()
    d.mtt.pd, d.mtt.survival_prob = 0.999999, 0.999999
    d.mtt.clutter_lambda, d.mtt.new_track_lambda = 1e-9, 1e-3
    lg = H.model_mtt_track()
    T = 12
    yobs, _ = H.gen_obs(lg, T, seed=77)
    f = H.OracleMtt(d, 64, 1)
    F, Q, Hm, R = H.desc_arrays(lg)
    mean = np.zeros(4); P = np.diag([5.0, 5.0, 0.1, 0.1])
    for t in range(T):
        rc, m, v, lz = f.step(yobs[t][None, :])
        assert rc == 0
        if t > 0:
            mean = F @ mean; P = F @ P @ F.T + Q
        S = Hm @ P @ Hm.T + R; K = P @ Hm.T @ np.linalg.inv(S)
        mean = mean + K @ (yobs[t] - Hm @ mean); P = P - K @ Hm @ P
        post, frac = f.track_posterior(0)
        assert frac > 0.99 and abs(m - 1.0) < 0.02
        assert np.allclose(post, mean, atol=1e-9)
    f.close()


def test_json_line_matches_java_parser(pz):
    """App.java:347-399: arr[0] = [[id,[x,y,..]]..]; arr[1][i][j] = [id, {"mean":[..],"cov":[[..],[..],..]}] or
    [id,[..]] (RConst); arr[2] = [[ox,oy]..]; only elements 0 and 1 of mean / cov rows are read."""
    gt = [(0, [1.0, 2.0, 0.1, 0.2]), (3, [-1.5, 0.25, 0.0, 0.0])]
    cov = np.arange(16.0).reshape(4, 4)
    parts = [[(0, [1.1, 2.1, 0.0, 0.0], cov), (3, [0.0, 0.0, 0.0, 0.0], None)], []]
    line = pz.json_line_tracks(gt, parts, [[0.5, -0.5], [3.0, 4.0]])
    assert "\n" not in line
    arr = json.loads(line)
    assert len(arr) == 3
    for node in arr[0]:
        assert isinstance(node[0], int) and float(node[1][0]) == float(node[1][0]) and len(node[1]) >= 2
    p0 = arr[1][0]
    assert p0[0][0] == 0 and p0[0][1]["mean"][:2] == [1.1, 2.1]
    assert p0[0][1]["cov"][0][0] == 0.0 and p0[0][1]["cov"][0][1] == 1.0 and p0[0][1]["cov"][1][1] == 5.0
    assert isinstance(p0[1][1], list) and len(p0[1][1]) == 4      # RConst: bare array
    assert arr[1][1] == [] and arr[2] == [[0.5, -0.5], [3.0, 4.0]]
    flat = json.loads(pz.json_line_generic1d([1.0, 2.5, -3.0]))     # Generic1D.java:333-348
    assert flat == [1.0, 2.5, -3.0]
    # a degenerate step (NaN mean) must still be a JSON line: aeson writes non-finite doubles as null
    assert json.loads(pz.json_line_generic1d([float("nan"), float("inf"), 1.0])) == [None, None, 1.0]


@pytest.mark.gpu
def test_gpu_tracker_matches_cpu_restatement(pz):
    d, obs, truth = H.mtt_scenario(50)
    n, seeds = 20000, 6
    gpu = np.zeros((seeds, len(obs))); cpu = np.zeros((seeds, len(obs)))
    glz = np.zeros((seeds, len(obs))); clz = np.zeros((seeds, len(obs)))
    for s in range(seeds):
        zs = pz.zparticles(n, pz.mtt(), seed=10 + s)
        orc = H.OracleMtt(d, n // 4, 100 + s)
        for t, o in enumerate(obs):
            post = zs.step(o)
            rc, m, v, lz = orc.step(o)
            assert rc == 0
            gpu[s, t], cpu[s, t] = post.mean[0], m
            glz[s, t], clz[s, t] = post.log_evidence_inc, lz
        zs.close(); orc.close()
    # expected track count: seed-averaged curves agree within Monte Carlo error
    se = np.sqrt(gpu.var(axis=0, ddof=1) / seeds + cpu.var(axis=0, ddof=1) / seeds) + 0.02
    z = (gpu.mean(axis=0) - cpu.mean(axis=0)) / se
    assert np.mean(np.abs(z) > 3) < 0.1 and np.max(np.abs(z)) < 6, z
    # log-evidence of the whole stream
    a, b = glz.sum(axis=1), clz.sum(axis=1)
    assert abs(a.mean() - b.mean()) < 4 * np.sqrt(a.var(ddof=1) / seeds + b.var(ddof=1) / seeds) + 0.5
    true_cnt = np.array([len(t) for t in truth])
    assert abs(gpu.mean(axis=0)[-15:].mean() - true_cnt[-15:].mean()) < 1.0


@pytest.mark.gpu
def test_gpu_tracker_resampling_bit_exact_and_readout(pz):
    d, obs, truth = H.mtt_scenario(30)
    n = 30000
    zs = pz.zparticles(n, pz.mtt(), seed=2)
    lib = H.oracle()
    for t, o in enumerate(obs):
        post = zs.step(o)
        if t in (3, 17, 29):
            lw = zs.read_logw()
            anc = zs.read_ancestors()
            U = lib.pzo_resample_uniform(2, t + 1)   # the uniform of the NEXT step's resample
            rc, ref = H.oracle_systematic(lw, U)
            assert rc == 0 and np.array_equal(anc, ref), t
    tracks = zs.read_tracks(stride=100)
    assert len(tracks) == n // 100
    cnts = np.array([len(p) for p in tracks])
    assert abs(cnts.mean() - post.mean[0]) < 0.5
    for p in tracks:
        ids = [tid for tid, _, _ in p]
        assert ids == sorted(ids) and len(set(ids)) == len(ids)
        for tid, mean, cov in p:
            assert np.all(np.isfinite(mean)) and np.allclose(cov, cov.T) and np.all(np.linalg.eigvalsh(cov) > -1e-4)
    # the line the visualiser would read
    line = pz.json_line_tracks(truth[-1], tracks[:100], obs[-1])
    arr = json.loads(line)
    assert len(arr[1]) == 100 and len(arr[2]) == len(obs[-1])
    # a confidently tracked target: weighted posterior position close to the truth
    if truth[-1]:
        tid, x = truth[-1][0]
        pos = np.array([m[:2] for p in tracks for i, m, c in p if i >= 0])
        assert pos.size == 0 or np.min(np.linalg.norm(pos - x[:2], axis=1)) < 3.0
    zs.close()


@pytest.mark.gpu
def test_gpu_tracker_capacity_and_errors(pz):
    zs = pz.zparticles(4096, pz.mtt(), seed=1)
    with pytest.raises(pz.PzError) as e:
        zs.step(np.zeros((40, 2)))           # more than m_max detections
    assert e.value.code == -5
    post = zs.step(np.zeros((0, 2)))          # an empty scan is a valid step
    assert post.mean[0] == 0.0
    with pytest.raises(pz.PzError):
        zs.read_state()
    zs.close()


def _three_target_scenario(T, seed=5):
    """Three well separated targets seen from step 0 on, every frame (pd ~ 1), essentially no clutter: births take
    the ids 0, 1, 2 in detection order on the GPU and in the oracle alike, and every later detection has one
    overwhelmingly likely track, so the Rao-Blackwellised posterior of each id is the exact Kalman posterior of its
    detection sequence."""
    lg = H.model_mtt_track()
    F, Q, Hm, R = H.desc_arrays(lg)
    rng = np.random.default_rng(seed)
    x = np.array([[-8.0, -6.0, 0.2, 0.1], [0.5, 7.0, -0.1, 0.0], [9.0, -2.0, 0.0, -0.2]])
    obs = []
    for _ in range(T):
        x = x @ F.T + rng.standard_normal((3, 4)) * np.sqrt(np.diag(Q))
        obs.append(x[:, :2] + rng.standard_normal((3, 2)))
    d = H.model_mtt()
    d.mtt.pd, d.mtt.survival_prob = 0.999999, 0.999999
    d.mtt.clutter_lambda, d.mtt.new_track_lambda = 1e-9, 1e-3
    return d, obs, (F, Q, Hm, R)


@pytest.mark.gpu
def test_gpu_tracker_per_track_posteriors_match_oracle_and_kalman(pz):
    """BASELINE.md gate 4: per-track posterior means AND covariances, GPU vs the CPU restatement vs the exact Kalman
    recursion, track by track (ids are deterministic in this scenario).  Tolerance: fp32 state on the GPU."""
    T, n = 13, 4096          # the targets stay more than 10 units apart over these steps (they wander afterwards)
    d, obs, (F, Q, Hm, R) = _three_target_scenario(T)
    gd = pz.mtt()
    gd.mtt.pd, gd.mtt.survival_prob, gd.mtt.clutter_lambda, gd.mtt.new_track_lambda = d.mtt.pd, d.mtt.survival_prob, d.mtt.clutter_lambda, d.mtt.new_track_lambda
    zs = pz.zparticles(n, gd, seed=3)
    orc = H.OracleMtt(d, 512, 4)
    mean = np.zeros((3, 4)); P = np.stack([np.diag([5.0, 5.0, 0.1, 0.1])] * 3)
    for t in range(T):
        post = zs.step(obs[t])
        rc, m, v, lz = orc.step(obs[t])
        assert rc == 0 and abs(post.mean[0] - 3.0) < 1e-3 and abs(m - 3.0) < 1e-3
        for k in range(3):   # the exact recursion of track k on its own detections
            if t > 0:
                mean[k] = F @ mean[k]; P[k] = F @ P[k] @ F.T + Q
            S = Hm @ P[k] @ Hm.T + R; K = P[k] @ Hm.T @ np.linalg.inv(S)
            mean[k] = mean[k] + K @ (obs[t][k] - Hm @ mean[k]); P[k] = P[k] - K @ Hm @ P[k]
        assert abs(post.log_evidence_inc - lz) < 5e-3 * max(1.0, abs(lz))
        if t in (0, 1, 5, 9, T - 1):
            tracks = zs.read_tracks(stride=64)
            for p in tracks:
                assert [tid for tid, _, _ in p] == [0, 1, 2]
                for tid, gm, gc in p:
                    assert np.allclose(gm, mean[tid], rtol=2e-4, atol=2e-4), (t, tid)
                    assert np.allclose(gc, P[tid], rtol=2e-3, atol=2e-5) and np.allclose(gc, gc.T)
            for k in range(3):
                om, frac = orc.track_posterior(k)
                assert frac > 0.999 and np.allclose(om, mean[k], atol=1e-9)
    zs.close(); orc.close()


@pytest.mark.gpu
def test_gpu_tracker_long_stream_stays_symmetric_psd_and_on_the_oracle_evidence(pz):
    """10^3 steps of predict / update in fp32 (P - K P updates): every covariance read back stays symmetric positive
    definite, the track count stays with the truth, and the log-evidence of the first 300 steps agrees with the CPU
    restatement (double) within 4 standard errors + 0.05 nat per 100 steps."""
    T, n = 1000, 20000
    d, obs, truth = H.mtt_scenario(T, seed=9)
    seeds, Tz = 5, 300
    glz = np.zeros(seeds); clz = np.zeros(seeds)
    for s in range(seeds):
        zs = pz.zparticles(n, pz.mtt(), seed=40 + s)
        orc = H.OracleMtt(d, 3000, 140 + s)
        for t in range(Tz):
            glz[s] += zs.step(obs[t]).log_evidence_inc
            rc, m, v, lz = orc.step(obs[t])
            assert rc == 0
            clz[s] += lz
        orc.close()
        if s == 0:   # one filter runs the whole stream
            cnt = []
            for t in range(Tz, T):
                post = zs.step(obs[t])
                assert np.isfinite(post.log_evidence_inc) and post.ess > 1
                cnt.append(post.mean[0])
                if t % 100 == 99:
                    for p in zs.read_tracks(stride=200):
                        for tid, mean, cov in p:
                            assert np.all(np.isfinite(mean)) and np.allclose(cov, cov.T, rtol=0, atol=0)
                            ev = np.linalg.eigvalsh(cov)
                            assert ev.min() > 1e-4 and ev.max() < 50.0, (t, tid, ev)
            true_cnt = np.array([len(x) for x in truth[Tz:]])
            assert abs(np.mean(cnt[-200:]) - true_cnt[-200:].mean()) < 1.5
        zs.close()
    se = np.sqrt(glz.var(ddof=1) / seeds + clz.var(ddof=1) / seeds)
    assert abs(glz.mean() - clz.mean()) < 4 * se + 0.05 * Tz / 100, (glz, clz)


def _tracker_runs_equal(pz, devices, n, steps, stride):
    d, obs, truth = H.mtt_scenario(steps)
    one = pz.zparticles(n, pz.mtt(), seed=5)
    many = pz.zparticles_multi(n, pz.mtt(), devices, seed=5)
    assert many.n == n
    for t, o in enumerate(obs):
        a, b = one.step(o), many.step(o)
        for k in ("total_weight", "e_max", "ess", "log_evidence_inc", "step"):
            assert getattr(a, k) == getattr(b, k), (t, k, getattr(a, k), getattr(b, k))
        assert a.mean[0] == b.mean[0] and a.var[0] == b.var[0], t
    ta, tb = one.read_tracks(stride=stride), many.read_tracks(stride=stride)
    assert len(ta) == len(tb)
    for pa, pb in zip(ta, tb):
        assert [i for i, _, _ in pa] == [i for i, _, _ in pb]
        for (_, ma, ca), (_, mb, cb) in zip(pa, pb):
            assert np.array_equal(ma, mb) and np.array_equal(ca, cb)
    one.close(); many.close()


@pytest.mark.gpu
def test_multi_device_tracker_equals_single_device(pz):
    """configs[4] 'sharded tracker': the tracker over 2 devices of one process (pz_filter_create_multi) -- own output tiles
    per device, parent records read from their owners, block statistics all-gathered by peer stores -- equals the
    one-device run bit for bit: summaries every step, track records at the end."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _tracker_runs_equal(pz, [0, 1], 50000, 25, 41)


@pytest.mark.gpu
def test_sharded_tracker_path_on_one_device(pz, monkeypatch):
    """The same sharded code path with 3 shards sharing device 0 (test hook PZ_MULTI_SAME_DEVICE): every kernel, push and
    event dependency of the multi-device tracker runs on a one-GPU box and must reproduce the unsharded bits."""
    monkeypatch.setenv("PZ_MULTI_SAME_DEVICE", "1")
    _tracker_runs_equal(pz, [0, 0, 0], 30000, 20, 29)
    with pytest.raises(pz.PzError):   # too few particles for the device count (whole 2048-slot tiles per device)
        pz.zparticles_multi(3000, pz.mtt(), [0, 0, 0], seed=1)


@pytest.mark.gpu
def test_taps_on_a_multi_device_tracker_are_refused(pz, monkeypatch):
    """The per-device taps (state, log-weights, ancestors, particles, reset) return PZ_ESTATE on a multi-device tracker handle
    instead of touching one shard's memory; the track readers work."""
    monkeypatch.setenv("PZ_MULTI_SAME_DEVICE", "1")
    zs = pz.zparticles_multi(3 * 2048 + 100, pz.mtt(), [0, 0], seed=1)
    zs.step(np.array([[0.5, -0.5]]))
    for call in (zs.read_state, zs.read_logw, zs.read_ancestors, lambda: zs.read_particles(10), lambda: zs.reset(1)):
        with pytest.raises(pz.PzError) as e:
            call()
        assert e.value.code == -7
    assert len(zs.read_tracks(stride=100)) == (3 * 2048 + 100 + 99) // 100
    zs.close()
