""""additionally report agreement vs the fp32 oracle build".  The reference compiled in place on
TinyAlgebra<float, FloatUtils> (oracle/_ref, prec=32) against its own fp64 build, next to the kernel sources (host-compiled,
tests/cpp) in their mixed arithmetic - same golden inputs, one step.  CPU only.
    python tests/reports/fp32_reference_report.py > profiles/r02_fp32_reference_agreement.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu  # noqa: E402
from oracle import ref  # noqa: E402
from tds_b200.model import fixture_path, load_model  # noqa: E402
from test_kernel_source_on_host import params_from_golden, rel_err  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
print("# One step from the golden inputs; error against the reference's fp64 build (the golden outputs), max |x - ref| / max(1, |ref|).")
print("# 'reference fp32' = the unmodified reference instantiated on TinyAlgebra<float, FloatUtils> (it stores M^-1 and runs ABA, CRBA,")
print("# the Cholesky inverse and the PGS sweep in fp32); 'kernel mixed' = csrc/tds_stepw.cu, articulated inertias fp32, kinematics and")
print("# composite inertias fp64, solver fp32, executed from its source on the CPU; 'spec mixed' = csrc/tds_steps.cu likewise.\n")
print(f"{'fixture':20s} {'quantity':5s} {'reference fp32':>15s} {'kernel mixed':>13s} {'spec mixed':>11s}")
for name in ("cartpole", "pendulum5", "sphere2", "laikago", "humanoid", "ant", "box", "cartpole_plane"):
    g = np.load(os.path.join(G, name + ".npz"))
    model = load_model(fixture_path(name))
    mode = int(g["mode"])
    params = params_from_golden(g)
    tau = g["tau"] if "tau" in g.files else None
    n_tau = int(model[4]) - (6 if int(model[2]) else 0)
    t = tau[:, -n_tau:] if (tau is not None and tau.shape[1] != n_tau) else tau
    rs = ref.RefSim.from_model(model, prec=32)
    rs.set_params(**params)
    n = g["q_in"].shape[0]
    r32 = [rs.step(mode, g["q_in"][i], g["qd_in"][i], None if tau is None else tau[i]) for i in range(n)]
    k = emu.step(model, mode, g["q_in"], g["qd_in"], t, precision=0, **params)
    sp = emu.step_spec(name, mode, g["q_in"], g["qd_in"], t, precision=0, **params) if name in ("laikago", "ant") else None
    if mode == 0:
        rows = [("qdd", np.array([r["qdd"] for r in r32]), g["qdd"], k["qdd"], None if sp is None else sp["qdd"])]
    else:
        rows = [("qd'", np.array([r["qd"] for r in r32]), g["qd_out"], k["qd"], None if sp is None else sp["qd"])]
    for lab, a32, gold, kk, ss in rows:
        print(f"{name:20s} {lab:5s} {rel_err(a32, gold):15.2e} {rel_err(kk, gold):13.2e} {('%11.2e' % rel_err(ss, gold)) if ss is not None else '          -'}")
