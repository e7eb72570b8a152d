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
    assert r.returncode == 0 and "mtt [N]" in r.stderr and "stt [delay] [N]" in r.stderr
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
