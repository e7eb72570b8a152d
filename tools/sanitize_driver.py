"""Small-N steps of every kernel family, for compute-sanitizer (tools/gpu_sanitize.sh).  argv[1]: all | multi"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g

pz = g.load_package()
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
rng = np.random.default_rng(0)


def lg(mk, kw, n=6000, T=3):
    d = mk(**kw)
    zs = pz.zparticles(n, d, seed=1)
    for t in range(T):
        zs.step(rng.standard_normal(d.ny))
    zs.read_particles(50)
    zs.close()


if mode == "all":
    for mk in (pz.rw1d, pz.mtt_track, pz.stt):
        for kw in ({}, {"f64": True}, {"ess_threshold": 0.5}, {"carry_weights": True}):
            lg(mk, kw)
    lg(pz.rw1d, {"resampler": pz.PZ_RESAMPLE_MULTINOMIAL_REF})
    lg(pz.stt, {"delay": True}, n=100)
    zs = pz.zparticles(4000, pz.walker(), seed=2)
    for t in range(2):
        zs.step(rng.standard_normal(2) * 10)
    zs.close()
    zs = pz.zparticles(4096, pz.mtt(), seed=3)
    for t in range(3):
        zs.step(rng.standard_normal((3, 2)) * 3)
    zs.read_tracks(100)
    zs.close()
    # MultiCamera tracker: one row per camera, plus a clutter-like row
    zs = pz.zparticles(4096, pz.multicam(), seed=4)
    for t in range(3):
        rows = np.zeros((3, 16)); rows[:, 0] = [0, 1, 1]
        rows[:, 1:11] = rng.standard_normal((3, 10)); rows[:, 1:11] /= np.linalg.norm(rows[:, 1:11], axis=1, keepdims=True)
        rows[:, 11:15] = rng.standard_normal((3, 4)) + [0, 0, 1, 1]; rows[:, 15] = rng.standard_normal(3)
        try:
            zs.step(rows)
        except pz.PzError:
            break
    zs.read_mcam_tracks(100, appearances=True)
    zs.close()
    # the sharded tracker path with every shard on device 0 (test hook): ancestors per tile range, peer-pointer parent
    # reads, slice pushes, event-ordered global plan
    os.environ["PZ_MULTI_SAME_DEVICE"] = "1"
    zs = pz.zparticles_multi(3 * 2048 + 500, pz.mtt(), [0, 0, 0], seed=3)
    for t in range(3):
        zs.step(rng.standard_normal((3, 2)) * 3)
    zs.read_tracks(100)
    zs.close()
    zs = pz.zparticles(100, pz.beta_bernoulli(), seed=0)
    zs.step([1.0]); zs.close()
    pz.resample_systematic(rng.standard_normal(5000), 0.3)
    pz.importance_resampling(rng.standard_normal(5000), 1, 0, 7)
    pz.dist_sample("poisson", [3.0], 1, 1000)
    print("sanitize driver: all families stepped")
else:
    import torch
    nd = min(torch.cuda.device_count(), 2)
    for mk, kw in ((pz.rw1d, {}), (pz.mtt_track, {"f64": True})):
        zs = pz.zparticles_multi(nd * 256 * 40, mk(**kw), list(range(nd)), seed=5)
        for t in range(3):
            zs.step(rng.standard_normal(mk().ny))
        zs.read_state()
        zs.close()
    print("sanitize driver: multi-device filters stepped on", nd, "devices")
