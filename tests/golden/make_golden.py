"""Generates the committed golden fixtures from the UNMODIFIED reference compiled in place
(oracle/_ref/libtds_ref.so, built by oracle/build_ref.sh from /root/reference).

    python tests/golden/make_golden.py

Writes
  tests/golden/models/<config>.json   flat models exported from the reference's own URDF loader
  tests/golden/<config>.npz           seeded inputs + reference (fp64) single-step outputs
Run in the build container only (needs /root/reference); the GPU box reads the committed files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
import tds_b200.workloads as wl  # noqa: E402
from tds_b200.model import save_model  # noqa: E402

D = os.environ.get("TDS_REFERENCE_ROOT", "/root/reference") + "/data/"
HERE = os.path.dirname(os.path.abspath(__file__))
N = 64

CONFIGS = {
    # name: (plane urdf, urdf, floating, workload generator)
    "cartpole": (None, D + "cartpole.urdf", False, wl.cartpole),
    "pendulum5": (None, D + "pendulum5.urdf", False, wl.pendulum5),
    "sphere2": (D + "plane_implicit.urdf", D + "sphere2.urdf", True, wl.sphere2),
    "laikago": (D + "plane_implicit.urdf", D + "laikago/laikago_toes_zup_xyz_xyzrot.urdf", False, wl.laikago_perturbed),
    "humanoid": (D + "plane_implicit.urdf", D + "humanoid.urdf", True, wl.humanoid),
    "ant": (D + "plane_implicit.urdf", D + "gym/ant_org_xyz_xyzrot.urdf", False, wl.ant_perturbed),
    # box shapes against the plane (contact_plane_box): our own one-box fixture and the reference's cartpole on the plane
    "box": (D + "plane_implicit.urdf", os.path.join(HERE, "urdf", "box.urdf"), True, wl.box),
    "cartpole_plane": (D + "plane_implicit.urdf", D + "cartpole.urdf", False, wl.cartpole_plane),
    # spherical joints (forward_dynamics.hpp:56-109, integrator.hpp:97-122)
    "pendulum5spherical": (None, D + "pendulum5spherical.urdf", False, wl.pendulum5spherical),
    "humanoid_spherical": (D + "plane_implicit.urdf", D + "humanoid_xyz_spherical.urdf", False, wl.humanoid_spherical),
}

ANT_POSES, ANT_KP, ANT_KD, ANT_MAX = np.array([0.0, -0.5] * 4), 15.0, 0.3, 3.0   # ant_environment2.h:43-66


def pd_tau(w):
    """PD torques of locomotion_contact_simulation.h:168-258 for the laikago workload (host restatement used
    only to feed the raw (q, qd, tau) reference step; the env-level golden below uses the reference's own PD)."""
    q, qd, a = w["q"], w["qd"], np.clip(w["action"], -0.4, 0.4)
    init = np.array([0.2, 0.0, -0.7] * 4)
    tau = np.zeros((q.shape[0], 18))
    f = 100.0 * (init + a - q[:, 6:18]) + 2.0 * (0.0 - qd[:, 6:18])
    tau[:, 6:18] = np.clip(f, -50.0, 50.0)
    return tau


def main():
    for name, (plane, urdf, floating, gen) in CONFIGS.items():
        sim = ref.RefSim.from_urdf(urdf, plane, floating)
        model = sim.export_model()
        meta = dict(source=os.path.relpath(urdf, D) if urdf.startswith(D) else os.path.relpath(urdf, HERE), plane=bool(plane),
                    floating=floating, exported_by="reference UrdfCache::construct via oracle/_ref")
        for dest in (os.path.join(HERE, "models"), os.path.join(ROOT, "tiny-differentiable-simulator_b200", "models")):
            save_model(os.path.join(dest, name + ".json"), model, meta=meta)
        w = gen(N)
        sim.set_params(**w["params"])
        if name == "humanoid":
            # second half of the batch: lower the base until the lowest contact point penetrates by 0..3 cm
            # (the base z that does this is found with the reference's own contact distances)
            rng = np.random.default_rng(99)
            for i in range(N // 2, N):
                o = sim.step(2, w["q"][i], w["qd"][i], w["tau"][i], contact_cap=64)
                w["q"][i, 6] -= o["contact_data"][:, 9].min() + rng.uniform(0.0, 0.03)
            w["q"] = w["q"].astype(np.float32).astype(np.float64)
        mode = w["mode"]
        tau = w.get("tau")
        if name == "laikago":
            tau = pd_tau(w)
        if name == "ant":   # PD torques of the env (host restatement only to feed the raw reference step)
            f = ANT_KP * (ANT_POSES + np.clip(w["action"], -0.4, 0.4) - w["q"][:, 6:14]) + ANT_KD * (0.0 - w["qd"][:, 6:14])
            tau = np.zeros((N, 14)); tau[:, 6:14] = np.clip(f, -ANT_MAX, ANT_MAX)
        outs = dict(q=[], qd=[], qdd=[], dist=[], link_a=[], link_b=[], n_contacts=[])
        for i in range(N):
            o = sim.step(mode, w["q"][i], w["qd"][i], None if tau is None else tau[i], contact_cap=64)
            outs["q"].append(o["q"]); outs["qd"].append(o["qd"]); outs["qdd"].append(o["qdd"])
            outs["n_contacts"].append(o["n_contacts"])
            outs["dist"].append(o["contact_data"][:, 9] if o["n_contacts"] else np.zeros(0))
            outs["link_a"].append(o["contact_idx"][:, 0] if o["n_contacts"] else np.zeros(0, dtype=np.int32))
            outs["link_b"].append(o["contact_idx"][:, 1] if o["n_contacts"] else np.zeros(0, dtype=np.int32))
        save = dict(q_in=w["q"], qd_in=w["qd"], mode=mode,
                    q_out=np.array(outs["q"]), qd_out=np.array(outs["qd"]), qdd=np.array(outs["qdd"]),
                    n_contacts=np.array(outs["n_contacts"]), contact_dist=np.array(outs["dist"]),
                    contact_link_a=np.array(outs["link_a"]), contact_link_b=np.array(outs["link_b"]))
        if tau is not None:
            save["tau"] = tau
        for k, v in w["params"].items():
            save["param_" + k] = np.asarray(v, dtype=np.float64)
        if name == "laikago":
            save["action"] = w["action"]
            L = ref.LaikagoRef(1)
            x = np.zeros((N, 51))
            x[:, :18], x[:, 18:36], x[:, 36:48] = w["q"], w["qd"], w["action"]
            x[:, 48:51] = [100.0, 2.0, 50.0]
            save["env_input"] = x
            save["env_output_templated"] = L.step(x, ref.LaikagoRef.IMPL_TEMPLATED)
            save["env_output_codegen"] = L.step(x, ref.LaikagoRef.IMPL_CODEGEN)
            rd = [L.reward_done(o) for o in save["env_output_templated"]]
            save["env_reward"] = np.array([r for r, _ in rd])
            save["env_done"] = np.array([float(d) for _, d in rd])
        if name == "ant":
            save["action"] = w["action"]
            A = ref.AntRef(1)
            x = np.zeros((N, A.input_dim))
            x[:, :14], x[:, 14:28], x[:, 28:36] = w["q"], w["qd"], w["action"]
            x[:, 36:39] = [ANT_KP, ANT_KD, ANT_MAX]
            save["env_input"] = x
            save["env_output_templated"] = A.step(x, ref.AntRef.IMPL_TEMPLATED)
            save["env_output_codegen"] = A.step(x, ref.AntRef.IMPL_CODEGEN)
            rd = [A.reward_done(x[i, :28], save["env_output_templated"][i, :28]) for i in range(N)]
            save["env_reward"] = np.array([r for r, _ in rd])
            save["env_done"] = np.array([float(d) for _, d in rd])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        pen = [int((d < 0).sum()) for d in outs["dist"]]
        print(f"{name}: model {model.size} doubles, {N} vectors, contacts/env {outs['n_contacts'][0]}, "
              f"penetrating min/max {min(pen) if pen else 0}/{max(pen) if pen else 0}")


if __name__ == "__main__":
    main()
