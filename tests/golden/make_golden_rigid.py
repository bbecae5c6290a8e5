"""Golden fixtures of the RigidBody path (World::step on rigid bodies: src/world.hpp:293-363, src/rigid_body.hpp,
src/rb_constraint_solver.hpp) from the UNMODIFIED reference compiled in place (oracle/_ref, oracle/ref/ref_rigid.cpp).

    python tests/golden/make_golden_rigid.py

Writes tests/golden/rigid_<world>.npz (tds_b200.workloads.rigid_world): bodies, states, forces, parameters and the reference's
states after 1 and after 5 steps.  Build container only (needs /root/reference)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
import tds_b200.workloads as wl  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N = 64

if __name__ == "__main__":
    for kind in wl.RIGID_WORLDS:
        w = wl.rigid_world(kind, N)
        rw = ref.RefRigidWorld(w["bodies"])
        rw.set_params(**w["params"])
        out = dict(bodies=w["bodies"], state=w["state"], force=w["force"])
        for k, v in w["params"].items():
            out["param_" + k] = np.asarray(v)
        s1, s5, nc = [], [], []
        for i in range(N):
            o, c = rw.step(w["state"][i], w["force"][i], 1)
            s1.append(o); nc.append(c)
            s5.append(rw.step(w["state"][i], w["force"][i], 5)[0])
        out.update(state_1=np.array(s1), state_5=np.array(s5), n_contacts=np.array(nc, dtype=np.int32))
        np.savez_compressed(os.path.join(HERE, "rigid_" + kind + ".npz"), **out)
        print(kind, "bodies", w["bodies"].shape[0], "contacts per world", np.mean(nc))
