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
r = rng.integers(0, 2 ** 32, 8192, dtype=np.uint32)
r[:8] = [0, 1, 2, 2 ** 31 - 1, 2 ** 31, 2 ** 31 + 1, 2 ** 32 - 1, 0x40000000]
r[8:39] = [1 << s for s in range(31)]
z = np.zeros(r.size, np.float32)
lib.pzo_icdf_normals(H.ptr(r, C.c_uint32), r.size, H.ptr(z, C.c_float))
lw = np.concatenate([-rng.random(4000) * 100, rng.standard_normal(90) * 5, [0, -0.0, 1e-20, -1e-20, -745.0, 88.0]]).astype(np.float32)
q, w, eb = H.block_weights(lw)
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
np.savez_compressed(os.path.join(ROOT, "tests/golden/philox_normal_kat.npz"), r=r, z=z, lw=lw, q=q, w=w, eb=eb,
                    lw_r=lw_r, anc_r=anc, U_r=np.uint32(0x9E3779B9), **fx)
print("wrote golden fixture")
