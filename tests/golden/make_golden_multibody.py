"""Golden fixtures for worlds of several multibodies (contacts between multibodies, src/world.hpp:206-282; one LCP per list of
World::mb_contacts_, :351-355) from the UNMODIFIED reference compiled in place (oracle/_ref/libtds_ref.so, oracle/ref/ref_world.cpp).

    python tests/golden/make_golden_multibody.py [world ...]

Writes tests/golden/mb_<world>.npz: the merged flat model (tds_b200.workloads.multibody_world_model), seeded inputs and the
reference's (fp64) outputs of one full step and of World::step alone, with the contact lists of the step.
Run in the build container only (needs /root/reference); the GPU box reads the committed files."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
import tds_b200.workloads as wl  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N = 64
CAP = 32


def main():
    for kind in (sys.argv[1:] or wl.MULTIBODY_WORLDS):      # optionally only the named worlds
        w = wl.multibody_world(kind, N)
        rw = ref.RefWorld(w["model"])
        rw.set_params(**w["params"])
        out = dict(model=w["model"], q_in=w["q"], qd_in=w["qd"], tau=w["tau"], mode=2)
        for k, v in w["params"].items():
            out["param_" + k] = np.asarray(v)
        q_out, qd_out, qd_world = [], [], []
        n_con = np.zeros(N, dtype=np.int32)
        idx = np.full((N, CAP, 3), -9, dtype=np.int32)
        dat = np.zeros((N, CAP, 10))
        lists = set()
        for i in range(N):
            r = rw.step(2, w["q"][i], w["qd"][i], w["tau"][i], contact_cap=CAP)
            q_out.append(r["q"]); qd_out.append(r["qd"])
            assert r["n_contacts"] <= CAP
            n_con[i] = r["n_contacts"]
            idx[i, :r["n_contacts"]] = r["contact_idx"]
            dat[i, :r["n_contacts"]] = r["contact_data"]
            lists.add(tuple(sorted(set(r["contact_idx"][r["contact_data"][:, 9] < 0, 0].tolist()))))
            qd_world.append(rw.step(3, w["q"][i], w["qd"][i])["qd"])
        out.update(q_out=np.array(q_out), qd_out=np.array(qd_out), qd_world_step=np.array(qd_world), n_contacts=n_con,
                   contact_idx=idx, contact_data=dat)
        np.savez_compressed(os.path.join(HERE, "mb_" + kind + ".npz"), **out)
        print(kind, "links", int(w["model"][1]), "dofs", int(w["model"][4]), "combinations of penetrating lists:", len(lists))


if __name__ == "__main__":
    main()
