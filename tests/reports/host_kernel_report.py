"""Evidence for profiles/: errors of the kernel SOURCES executed on the CPU (tests/cpp/*_host.cpp) against the reference's golden
vectors - the specialised step kernel, the generic world-frame kernel (incl. worlds of several multibodies) and the rigid-body
world kernel.  No GPU involved; says so in its header.    python tests/reports/host_kernel_report.py > profiles/r02_host_compiled_kernels.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu  # noqa: E402
import tds_b200.workloads as wl  # noqa: E402
from tds_b200.model import fixture_path, load_model  # noqa: E402
from test_kernel_source_on_host import params_from_golden, rel_err  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
print("# Kernel sources compiled for the host (g++), executed on the CPU of the build container, against golden vectors of the")
print("# reference.  NOT a GPU measurement: it checks the arithmetic of the CUDA sources, not their nvcc build.")
print("# rel err = max |x - ref| / max(1, |ref|) over 64 states; precision: mixed = RA fp32 / RC fp64 / RS fp32, f64 = all fp64\n")
print("## csrc/tds_steps.cu (model-specialised step kernel; -DTDS_B200_EXACT_RCP)")
for name in ("laikago", "ant"):
    g = np.load(os.path.join(G, name + ".npz"))
    model = load_model(fixture_path(name))
    tau = g["tau"][:, -int(model[4]):]
    for prec, lab in ((0, "mixed"), (1, "f64"), (2, "f32")):
        o = emu.step_spec(name, 2, g["q_in"], g["qd_in"], tau, precision=prec, **params_from_golden(g))
        print(f"{name:22s} {lab:6s} q' {rel_err(o['q'], g['q_out']):.2e}  qd' {rel_err(o['qd'], g['qd_out']):.2e}")
print("\n## csrc/tds_stepw.cu (generic world-frame kernel)")
for name in ("cartpole", "pendulum5", "sphere2", "laikago", "humanoid", "ant", "box", "cartpole_plane", "pendulum5spherical", "humanoid_spherical"):
    g = np.load(os.path.join(G, name + ".npz"))
    model = load_model(fixture_path(name))
    mode = int(g["mode"])
    tau = g["tau"] if "tau" in g.files else None
    n_tau = int(model[4]) - (6 if int(model[2]) else 0)
    if tau is not None and tau.shape[1] != n_tau:
        tau = tau[:, -n_tau:]
    for prec, lab in ((0, "mixed"), (1, "f64")):
        o = emu.step(model, mode, g["q_in"], g["qd_in"], tau, precision=prec, **params_from_golden(g))
        if mode == 0:
            print(f"{name:22s} {lab:6s} qdd {rel_err(o['qdd'], g['qdd']):.2e}")
        else:
            print(f"{name:22s} {lab:6s} q' {rel_err(o['q'], g['q_out']):.2e}  qd' {rel_err(o['qd'], g['qd_out']):.2e}")
print("\n## csrc/tds_stepw.cu, worlds of several multibodies (contacts between multibodies, one LCP per pair in sequence)")
for kind in wl.MULTIBODY_WORLDS:
    g = np.load(os.path.join(G, "mb_" + kind + ".npz"))
    for prec, lab in ((0, "mixed"), (1, "f64")):
        o = emu.step(g["model"], 2, g["q_in"], g["qd_in"], g["tau"], precision=prec, **params_from_golden(g))
        ws = emu.step(g["model"], 3, g["q_in"], g["qd_in"], None, precision=prec, **params_from_golden(g))
        k = o["contact_dist"].shape[1]
        print(f"{'mb_' + kind:22s} {lab:6s} q' {rel_err(o['q'], g['q_out']):.2e}  qd' {rel_err(o['qd'], g['qd_out']):.2e}  World::step alone qd' "
              f"{rel_err(ws['qd'], g['qd_world_step']):.2e}  candidate distances {np.max(np.abs(o['contact_dist'] - g['contact_data'][:, :k, 9])):.1e}")
print("\n## csrc/tds_rigid.cu (RigidBody path of World::step; fp64 on both sides, absolute error)")
for kind in wl.RIGID_WORLDS:
    g = np.load(os.path.join(G, "rigid_" + kind + ".npz"))
    p = params_from_golden(g)
    p["num_solver_iterations"] = int(p["num_solver_iterations"])
    e1 = np.max(np.abs(emu.rigid_step(g["bodies"], g["state"], g["force"], 1, **p) - g["state_1"]))
    e5 = np.max(np.abs(emu.rigid_step(g["bodies"], g["state"], g["force"], 5, **p) - g["state_5"]))
    print(f"{'rigid_' + kind:22s} f64    1 step {e1:.1e}   5 steps {e5:.1e}   ({int(g['n_contacts'].mean())} contacts per world)")
