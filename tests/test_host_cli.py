"""examples-gpu: the reference's CLI driver (haskell/src/test/Examples.hs:26-40) over the C ABI.
CPU: it builds, links libpzgpu.so, documents the reference's sub-commands and fails loudly without
a GPU.  GPU: every line it prints parses the way App.java:347-399 / Generic1D.java:333-348 do."""
import json
import os
import subprocess

import pytest

import pzhelpers as H

EXE = os.path.join(H.ROOT, "host", "examples-gpu")


def build():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(g.PKG_DIR, "libpzgpu.so")):
        g.build()
    subprocess.check_call(["make", "-C", os.path.join(H.ROOT, "host")], stdout=subprocess.DEVNULL)


def test_cli_builds_and_reports_usage():
    build()
    r = subprocess.run([EXE, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "mtt [N]" in r.stderr and "stt [delay] [N]" in r.stderr and "walker [N]" in r.stderr
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2
    r = subprocess.run([EXE, "stt", "Maybe", "100"], capture_output=True, text=True)
    assert r.returncode == 2   # the delay flag is a Haskell Bool


def test_cli_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    build()
    r = subprocess.run([EXE, "rw1d", "1000", "--steps", "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and r.stdout == ""


def java_app_reads(line):
    """The access pattern of App.TrackManager.run (App.java:347-399)."""
    arr = json.loads(line)
    for node in arr[0]:
        int(node[0]); float(node[1][0]); float(node[1][1])
    n_tracks = 0
    for particle in arr[1]:
        for node in particle:
            if isinstance(node[1], dict):
                pv = node[1]
                float(pv["mean"][0]); float(pv["mean"][1])
                float(pv["cov"][0][0]); float(pv["cov"][0][1]); float(pv["cov"][1][1])
            else:
                float(node[1][0]); float(node[1][1])
            n_tracks += 1
    for o in arr[2]:
        float(o[0]); float(o[1])
    return arr, n_tracks


@pytest.mark.gpu
def test_cli_streams_lines_the_visualiser_parses():
    build()
    r = subprocess.run([EXE, "mtt", "20000", "--steps", "25", "--emit", "50"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    assert len(lines) == 25
    total = 0
    for ln in lines:
        arr, n = java_app_reads(ln)
        assert len(arr[1]) == 50
        total += n
    assert total > 0  # tracks were reported
    r = subprocess.run([EXE, "stt", "False", "10000", "--steps", "5", "--emit", "20"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for ln in r.stdout.strip().split("\n"):
        arr, n = java_app_reads(ln)
        assert len(arr[0]) == 1 and len(arr[0][0][1]) == 6 and len(arr[1]) == 20 and len(arr[2]) == 1 and len(arr[2][0]) == 3
    # the reference's default `stt` (delay=True): every particle is the marginal {"mean":..,"cov":..}
    r = subprocess.run([EXE, "stt", "True", "1000", "--steps", "6", "--emit", "10"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for ln in r.stdout.strip().split("\n"):
        arr, n = java_app_reads(ln)
        assert len(arr[1]) == 10 and all(isinstance(p[0][1], dict) and len(p[0][1]["cov"]) == 6 for p in arr[1])
    r = subprocess.run([EXE, "rw1d", "100000", "--steps", "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = [json.loads(ln) for ln in r.stdout.strip().split("\n")]
    assert len(rows) == 8 and all(len(v) == 3 for v in rows)        # Generic1D.java: one double per series
    assert all(abs(v[2] - v[1]) < 6.0 for v in rows)                 # the posterior mean follows the truth


@pytest.mark.gpu
def test_cli_walker_prints_the_reference_summary():
    """`examples walker` (Walker.run, Examples/Walker.hs:137-148): one Haskell-`show`n tuple per measurement:
    ([(frequency,MotionType)..],(mean x,mean y),(mean vx,mean vy)), frequencies of the resampled particles' motion types."""
    import re
    build()
    r = subprocess.run([EXE, "walker", "2000", "--steps", "12", "--seed", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    assert len(lines) == 12
    num = r"-?(?:\d+\.\d+(?:e-?\d+)?|NaN|Infinity)"
    pat = re.compile(rf"^\(\[((?:\({num},(?:Stationary|Walking|Running)\),?)+)\],\(({num}),({num})\),\(({num}),({num})\)\)$")
    for ln in lines:
        m = pat.match(ln)
        assert m, ln
        freqs = [float(x) for x in re.findall(rf"\(({num}),", m.group(1))]
        assert abs(sum(freqs) - 1.0) < 1e-9 and all(0 < f <= 1 for f in freqs)
        names = re.findall(r"(Stationary|Walking|Running)", m.group(1))
        assert names == sorted(names, key=["Stationary", "Walking", "Running"].index) and len(set(names)) == len(names)
        assert all(float(m.group(k)) == float(m.group(k)) and abs(float(m.group(k))) < 1e9 for k in (2, 3, 4, 5))


MIRROR = os.path.join(H.ROOT, "host", "shim-mirror")


def test_shim_mirror_covers_every_haskell_import():
    """haskell/PzGpu.hs cannot be compiled here (no GHC); host/shim_mirror.c makes the same calls in C.  Every symbol the
    Haskell module imports must be called by the mirror, with the same descriptor / summary sizes and byte offsets."""
    import re
    hs = open(os.path.join(H.ROOT, "haskell", "PzGpu.hs")).read()
    c = open(os.path.join(H.ROOT, "host", "shim_mirror.c")).read()
    for sym in set(re.findall(r'foreign import ccall \w+ "&?(pz_\w+)"', hs)):
        assert re.search(rf"\b{sym}\(", c), f"shim_mirror.c never calls {sym}"
    assert int(re.search(r"sizeofModelDesc = (\d+)", hs).group(1)) == int(re.search(r"SIZEOF_MODEL_DESC (\d+)", c).group(1))
    assert int(re.search(r"sizeofSummary = (\d+)", hs).group(1)) == int(re.search(r"SIZEOF_SUMMARY (\d+)", c).group(1))
    for off in (4, 20, 24):   # nx, delayed, state_f64: the descriptor bytes the Haskell module reads / pokes
        assert re.search(rf"(peekByteOff|pokeByteOff) p {off} ", hs) and re.search(rf"(\+ {off},|, {off},)", c)
    build()
    assert os.path.exists(MIRROR)


@pytest.mark.gpu
def test_shim_mirror_runs_the_haskell_call_sequence(pz):
    """The C mirror of the Haskell module against the ctypes host on the same inputs: same summaries (bitwise: same
    kernels), particle rows chunked by the descriptor's nx, the tracker's track maps, the Beta posterior."""
    import numpy as np
    import torch
    build()
    ndev = min(torch.cuda.device_count(), 2)
    r = subprocess.run([MIRROR, str(ndev)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    sc = {d["scenario"]: d for d in (json.loads(ln) for ln in r.stdout.strip().split("\n"))}
    obs = lambda t: [0.3 * (t + 1), -0.2 * t, 0.1]
    for name, mk, kw in (("rw1d", pz.rw1d, {}), ("rw1d_double", pz.rw1d, {"f64": True}), ("mtt_track", pz.mtt_track, {}), ("stt", pz.stt, {}),
                         ("stt_delayed", pz.stt, {"delay": True}), ("walker", pz.walker, {})):
        d = mk(**kw)
        zs = pz.zparticles(10000, d, seed=7)
        assert sc[name]["nx"] == d.nx
        for t in range(4):
            post = zs.step(obs(t)[:d.ny])
            got = sc[name]["steps"][t]
            assert got["step"] == t and got["le"] == post.log_evidence_inc and got["ess"] == post.ess
            assert got["mean"] == list(post.mean)[:d.nx] and got["var"] == list(post.var)[:d.nx]   # the Haskell side reads nx entries
        stride = {"mtt_track": 999, "stt_delayed": 5000}.get(name, 1000)
        parts = zs.read_particles(stride)
        assert sc[name]["rows"] == parts.shape[0] and len(sc[name]["first_particle"]) == d.nx
        assert np.array_equal(sc[name]["first_particle"], parts[0]) and np.array_equal(sc[name]["last_particle"], parts[-1])
        zs.close()
    zs = pz.zparticles(10000, pz.rw1d(), seed=7)
    posts = zs.run(np.array([[0.3], [0.6], [0.9], [1.2]]))
    assert [s["mean"][0] for s in sc["run_all"]["steps"]] == [p.mean[0] for p in posts]
    assert sc["beta_bernoulli"] == {"scenario": "beta_bernoulli", "alpha": 3.0, "beta": 2.0}
    m = sc["mtt"]
    assert m["rows"] == 8 and len(m["particles"]) == 8 and any(len(p) > 0 for p in m["particles"])
    for p in m["particles"]:
        for tid, mean, var in p:
            assert tid >= 0 and all(np.isfinite(mean)) and var > 0
    mc = sc["multicam"]   # multiCameraGPU: same call sequence through ctypes gives the same marginal tracks
    zs = pz.zparticles(4000, pz.multicam(), seed=3)
    for t in range(3):
        rows = np.zeros((2, 16))
        for c in range(2):
            rows[c, 0] = c; rows[c, 1 + (t + c) % 10] = 1.0
            rows[c, 11:16] = [0.3 * t - 0.2 * c, 0.1 * c, 1.0, 1.1, 0.4 - 0.3 * c]
        zs.step(rows)
    mine = zs.read_mcam_tracks(stride=500)
    assert mc["rows"] == 8 == len(mine)
    for got, p in zip(mc["particles"], mine):
        assert [(g[0], g[1], g[2]) for g in got] == [(t[0], t[1], t[2]) for t in p["tracks"]]
        for g, t in zip(got, p["tracks"]):
            assert np.allclose(g[3], t[4][:2], rtol=1e-8) and np.isclose(g[4], t[5][0], rtol=1e-8)
    zs.close()
    if ndev > 1:   # withDevices: the posterior of the multi-device handle equals the single-device one
        a, b = sc["multi_mtt_track"], sc["single_mtt_track_same_n"]
        assert a["n_local"] == 256 * 40 * ndev
        for sa, sb in zip(a["steps"], b["steps"]):
            # (fp32 per-lane partial sums grouped by tile: the shards' tiles differ from the single filter's)
                assert np.allclose(sa["mean"], sb["mean"], rtol=1e-5, atol=1e-7) and abs(sa["ess"] - sb["ess"]) < 1e-6 * sb["ess"]
