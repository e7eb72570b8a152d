import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import __graft_entry__ as g
import pzhelpers as H
pz = g.load_package()
hd = H.model_rw1d()
obs, _ = H.gen_obs(hd, 6)
n = 2 * 256 * 101
one = pz.zparticles(n, pz.rw1d(), seed=99)
many = pz.zparticles_multi(n, pz.rw1d(), [0, 1], seed=99)
print("created", flush=True)
for t in range(6):
    pa = one.step(obs[t])
    try:
        pb = many.step(obs[t])
    except Exception as e:
        print("step", t, "failed:", e, flush=True); break
    print(t, pa.total_weight == pb.total_weight, np.array_equal(one.read_state(), many.read_state()), pb.n_migrated, flush=True)
