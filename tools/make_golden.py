"""Generate tests/golden/philox_normal_kat.npz from the CPU oracle.

The reference has no golden vectors; these pin the canonical math spec so
that any later edit of the oracle or of the coefficients is caught by the CPU suite, and
give the GPU suite a fixture that does not need the oracle's arithmetic at run time."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pzhelpers as H  # noqa: E402

lib = H.oracle()
rng = np.random.default_rng(20261019)
ra = rng.integers(0, 2 ** 32, 4096, dtype=np.uint32)
rb = rng.integers(0, 2 ** 32, 4096, dtype=np.uint32)
ra[:6] = [0, 1, 2, 2 ** 32 - 1, 2 ** 31, 0x3F3504F3]
rb[:6] = [0, 2 ** 30 - 1, 2 ** 30, 2 ** 31, 2 ** 32 - 1, 3 << 30]
z0, z1 = np.zeros(ra.size, np.float32), np.zeros(ra.size, np.float32)
lib.pzo_normal_pairs(H.ptr(ra, C.c_uint32), H.ptr(rb, C.c_uint32), ra.size, H.ptr(z0, C.c_float), H.ptr(z1, C.c_float))
lw = np.concatenate([-rng.random(4000) * 100, rng.standard_normal(90) * 5, [0, -0.0, 1e-20, -1e-20, -745.0, 88.0]]).astype(np.float32)
n, m = np.zeros(lw.size, np.int32), np.zeros(lw.size, np.uint32)
lib.pzo_weight_split(H.ptr(lw, C.c_float), lw.size, H.ptr(n, C.c_int32), H.ptr(m, C.c_uint32))
# resampler fixture: ancestors for a fixed weight vector and uniform
lw_r = (rng.standard_normal(3000) * 3).astype(np.float32)
lw_r[100:140] = -np.inf
rc, anc = H.oracle_systematic(lw_r, 0x9E3779B9)
assert rc == 0
# filter fixture: 6 steps of the rw1d and mtt-track filters, state checksums
fx = {}
for name, d in (("rw1d", H.model_rw1d()), ("lg42", H.model_mtt_track()), ("lg63", H.model_stt())):
    obs, _ = H.gen_obs(d, 6)
    f = H.OracleFilter(d, 3000, 42)
    for t in range(6):
        rc, s = f.step(obs[t]); assert rc == 0
    x, lwf, a = f.state()
    fx[name + "_x"] = x; fx[name + "_anc"] = a; fx[name + "_obs"] = obs
    fx[name + "_mean"] = np.array(s.mean[: d.nx]); fx[name + "_total"] = np.uint64(s.total_weight)
np.savez_compressed(os.path.join(ROOT, "tests/golden/philox_normal_kat.npz"), ra=ra, rb=rb, z0=z0, z1=z1, lw=lw, n=n, m=m,
                    lw_r=lw_r, anc_r=anc, U_r=np.uint32(0x9E3779B9), **fx)
print("wrote golden fixture")
