"""The product's headline kernel - the model-specialised env step of csrc/tds_steps.cu (Laikago and Ant instances) - executed on
the CPU by compiling its SOURCE for the host (tests/cpp/steps_host.cpp: four host threads play the four role warps of one lane
and meet at a host barrier where the kernel has __syncthreads).  Golden vectors of the reference, its env step (PD, reward, done),
fresh inputs and a solver-parameter sweep against the C oracle - without a GPU.  tests/test_parity_gpu.py checks the same kernel
as nvcc builds it; this file keeps the kernel source honest in a container that has no GPU (rcp.approx is replaced by an exact
division here: -DTDS_B200_EXACT_RCP)."""
import os

import numpy as np
import pytest

import tds_b200.envs as envs
import tds_b200.workloads as wl
from tds_b200.model import fixture_path, load_model
from oracle import port
import emu
from test_kernel_source_on_host import GOLDEN, TOL, params_from_golden, rel_err

ENV = {
    "laikago": dict(poses=envs.LAIKAGO_INITIAL_POSES, kp=envs.LAIKAGO_KP, kd=envs.LAIKAGO_KD, max_force=envs.LAIKAGO_MAX_FORCE, reward_kind=1,
                    params=dict(dt=1e-3, friction=1.0, keep_all_points=True)),
    "ant": dict(poses=envs.ANT_INITIAL_POSES, kp=envs.ANT_KP, kd=envs.ANT_KD, max_force=envs.ANT_MAX_FORCE, reward_kind=3,
                params=dict(dt=envs.ANT_DT, friction=1.0, keep_all_points=True)),
}


def env_vector(name, reward_kind=None):
    e = ENV[name]
    return np.array([len(e["poses"]), 6, e["kp"], e["kd"], e["max_force"], 0.4, e["reward_kind"] if reward_kind is None else reward_kind, *e["poses"]])


@pytest.mark.parametrize("name", ["laikago", "ant"])
@pytest.mark.parametrize("precision", [0, 1, 2])
@pytest.mark.parametrize("other_lane_in_contact", [False, True])
def test_golden_vectors_through_the_specialised_kernel_source(name, precision, other_lane_in_contact):
    """precision 0: the shipped mixed instance (articulated inertias fp32, kinematics / composite inertias fp64, solver fp32);
    1: all fp64; 2: all fp32.  other_lane_in_contact: the tile-uniform "any contact" flag forced on, i.e. the solve path taken by an
    environment whose own contact set may be empty."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = load_model(fixture_path(name))
    n_tau = int(model[4])
    tau = g["tau"][:, -n_tau:]
    out = emu.step_spec(name, 2, g["q_in"], g["qd_in"], tau, precision=precision, other_lane_in_contact=other_lane_in_contact, **params_from_golden(g))
    tol = TOL if precision < 2 else 1e-4
    assert rel_err(out["q"], g["q_out"]) <= tol and rel_err(out["qd"], g["qd_out"]) <= tol


@pytest.mark.parametrize("name", ["laikago", "ant"])
@pytest.mark.parametrize("var", [0, 1])
def test_env_step_vs_reference_env(name, var):
    """PD controller + full step + reward / done against the reference's own env step (golden env_output_templated;
    locomotion_contact_simulation.h:168-299, laikago_environment2.h:130-171, ant_environment2.h:75-105), through the general (0)
    and the lean (1) instance of the kernel."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n_q = g["q_in"].shape[1]
    out = emu.step_spec(name, 2, g["q_in"], g["qd_in"], g["action"], precision=0, var=var, use_pd=True, env=env_vector(name), **ENV[name]["params"])
    ref = g["env_output_templated"]
    assert rel_err(out["q"], ref[:, :n_q]) <= TOL and rel_err(out["qd"], ref[:, n_q:2 * n_q]) <= TOL
    assert np.array_equal(out["done"], g["env_done"].astype(np.float64))
    assert np.max(np.abs(out["reward"] - g["env_reward"])) <= 1e-5 * max(1.0, np.max(np.abs(g["env_reward"])))


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("precision", [0, 1])
def test_fresh_laikago_states_vs_c_oracle(mode, precision):
    n = 96   # three tiles
    model = load_model(fixture_path("laikago"))
    w = wl.laikago_perturbed(n, seed=4242)
    rng = np.random.default_rng(7)
    tau = rng.uniform(-5, 5, (n, 18)); tau[:, :6] = 0
    P = port.make_params(**w["params"])
    out = emu.step_spec("laikago", mode, w["q"], w["qd"], tau, precision=precision, **w["params"])
    q32, qd32, t32 = (a.astype(np.float32).astype(np.float64) for a in (w["q"], w["qd"], tau))
    refs = [port.step(model, P, mode, q32[i], qd32[i], t32[i]) for i in range(n)]
    if mode == 0:
        ref = np.array([r["qdd"] for r in refs])
        # accelerations of O(100) rad/s^2 through fp32 articulated inertias in the mixed instance
        assert rel_err(out["qdd"], ref) <= (1e-4 if precision == 0 else TOL)
        return
    assert rel_err(out["q"], np.array([r["q"] for r in refs])) <= TOL
    assert rel_err(out["qd"], np.array([r["qd"] for r in refs])) <= TOL


@pytest.mark.parametrize("name", ["laikago", "ant"])
@pytest.mark.parametrize("sweep", [dict(pgs_iterations=5), dict(pgs_iterations=20, friction=0.3), dict(restitution=0.5, erp=0.1, cfm=1e-3),
                                   dict(keep_all_points=False), dict(dt=4e-3, gravity=(0.5, 0.0, -9.0))])
def test_solver_parameter_sweep_vs_c_oracle(name, sweep):
    """iterations, friction, restitution, erp, cfm, keep_all_points, dt and gravity away from the env defaults
    (the looped PGS sweep, the non-penetrating candidate filter and the Baumgarte term of the specialised kernel)."""
    n = 32
    model = load_model(fixture_path(name))
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    params = dict(ENV[name]["params"]); params.update(sweep)
    q, qd = g["q_in"][:n], g["qd_in"][:n]
    out = emu.step_spec(name, 2, q, qd, g["action"][:n], precision=0, use_pd=True, env=env_vector(name, 0), **params)
    P = port.make_params(**params)
    x = np.zeros((n, 2 * q.shape[1] + g["action"].shape[1] + 3))
    nq = q.shape[1]
    x[:, :nq], x[:, nq:2 * nq], x[:, 2 * nq:-3], x[:, -3:] = q, qd, g["action"][:n], [ENV[name]["kp"], ENV[name]["kd"], ENV[name]["max_force"]]
    ref = port.locomotion_step(model, P, ENV[name]["poses"], 6, x, g["env_output_templated"].shape[1])
    assert rel_err(out["q"], ref[:, :nq]) <= TOL and rel_err(out["qd"], ref[:, nq:2 * nq]) <= 2 * TOL


def test_rollout_of_the_kernel_source_tracks_the_oracle():
    """40 env steps of 8 Laikago environments dropped from the reset pose: state carried in fp32 between steps as on the device,
    against the fp64 oracle stepping from the kernel's own previous state (per-step error, not trajectory divergence)."""
    n, steps = 8, 40
    model = load_model(fixture_path("laikago"))
    params = ENV["laikago"]["params"]
    P = port.make_params(**params)
    rng = np.random.default_rng(11)
    q = np.tile(envs.laikago_reset_pose(), (n, 1)); q[:, 6:18] += 0.05 * rng.uniform(-1, 1, (n, 12)); q[:, 2] -= 0.02
    qd = np.zeros((n, 18))
    worst = 0.0
    for s in range(steps):
        act = 0.3 * rng.uniform(-1, 1, (n, 12))
        q32, qd32, a32 = (a.astype(np.float32).astype(np.float64) for a in (q, qd, act))
        out = emu.step_spec("laikago", 2, q32, qd32, a32, precision=0, var=1, use_pd=True, env=env_vector("laikago"), **params)
        x = np.zeros((n, 51)); x[:, :18], x[:, 18:36], x[:, 36:48], x[:, 48:] = q32, qd32, a32, [envs.LAIKAGO_KP, envs.LAIKAGO_KD, envs.LAIKAGO_MAX_FORCE]
        ref = port.locomotion_step(model, P, envs.LAIKAGO_INITIAL_POSES, 6, x, 411)
        worst = max(worst, rel_err(out["q"], ref[:, :18]), rel_err(out["qd"], ref[:, 18:36]))
        q, qd = out["q"], out["qd"]
    assert worst <= 2 * TOL and np.all(np.isfinite(q))


@pytest.mark.parametrize("name", ["laikago", "ant"])
def test_contact_distances_and_link_transforms_of_the_general_instance(name):
    """The general instance also reports what the reference exposes after a step: the signed distance of every candidate point
    (golden contact_dist, from the reference's own contact list) and the world transform of every link (forward kinematics of
    the step, against the oracle)."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = load_model(fixture_path(name))
    n, n_links = 16, int(model[1])
    params = params_from_golden(g)
    tau = g["tau"][:n, -int(model[4]):]
    out = emu.step_spec(name, 2, g["q_in"][:n], g["qd_in"][:n], tau, precision=0, **params)
    ref_d = np.stack(list(g["contact_dist"]))[:n]
    assert out["contact_dist"].shape == ref_d.shape and np.max(np.abs(out["contact_dist"] - ref_d)) < 2e-6
    xf = emu.link_xf_of(out, n, n_links)
    P = port.make_params(**params)
    q32, qd32, t32 = (a.astype(np.float32).astype(np.float64) for a in (g["q_in"][:n], g["qd_in"][:n], tau))
    for i in range(n):
        r = port.step(model, P, 2, q32[i], qd32[i], t32[i])
        assert np.max(np.abs(xf[i] - r["link_xf"])) < 5e-6


@pytest.mark.parametrize("name,n", [("laikago", 40), ("ant", 32)])
def test_host_layout_instance_on_whole_tiles(name, n):
    """The instance behind tds_b200_env_step_host: actions arrive environment-major, the tile stages them through shared memory,
    and observations | reward | done leave environment-major with coalesced (float4 on full tiles, scalar on the ragged last tile)
    stores - 128 host threads per tile.  Must equal the lean device-layout instance bit for bit."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = dict(precision=0, use_pd=True, env=env_vector(name), **ENV[name]["params"])
    a = emu.step_spec(name, 2, g["q_in"][:n], g["qd_in"][:n], g["action"][:n], var=1, whole_tile=True, **kw)
    b = emu.step_spec(name, 2, g["q_in"][:n], g["qd_in"][:n], g["action"][:n], var=2, **kw)
    n_q = g["q_in"].shape[1]
    assert np.array_equal(b["obs"][:, :n_q], a["q"]) and np.array_equal(b["obs"][:, n_q:], a["qd"])
    assert np.array_equal(b["q"], a["q"]) and np.array_equal(b["qd"], a["qd"])   # the SoA state is written as well
    assert np.array_equal(b["obs_reward"], a["reward"]) and np.array_equal(b["obs_done"], a["done"])
    ref = g["env_output_templated"][:n]
    assert rel_err(b["obs"], ref[:, :2 * n_q]) <= TOL and np.array_equal(b["obs_done"], g["env_done"][:n].astype(np.float64))


def test_tile_wide_contact_flag_whole_tile_equals_lane_by_lane():
    """A tile mixing environments in contact with environments in the air: the tile-uniform "any contact" branch (whole tile, exact
    __syncthreads_or) against each environment on its own."""
    g = np.load(os.path.join(GOLDEN, "laikago.npz"))
    n = 32
    q = g["q_in"][:n].copy(); q[::2, 2] += 1.0   # every other robot lifted a metre: no contact
    kw = dict(precision=0, var=1, use_pd=True, env=env_vector("laikago"), **ENV["laikago"]["params"])
    a = emu.step_spec("laikago", 2, q, g["qd_in"][:n], g["action"][:n], whole_tile=True, **kw)
    b = emu.step_spec("laikago", 2, q, g["qd_in"][:n], g["action"][:n], **kw)
    assert np.max(np.abs(a["qd"] - b["qd"])) <= 1e-6 and np.max(np.abs(a["q"] - b["q"])) <= 1e-7


def test_auto_reset_inside_the_kernel():
    """laikago_environment2.h:130-171 + VectorizedEnvironment auto-reset: a robot reported done leaves the step at the reset pose
    with zero velocity (the reward / done of the finished step are still reported)."""
    g = np.load(os.path.join(GOLDEN, "laikago.npz"))
    n = 8
    q = g["q_in"][:n].copy(); qd = g["qd_in"][:n].copy()
    q[:3, 3:6] = [1.2, 0.0, 0.0]   # rolled over: the up axis leaves the done cone
    reset_q = envs.laikago_reset_pose()
    kw = dict(precision=0, var=1, use_pd=True, env=env_vector("laikago"), **ENV["laikago"]["params"])
    keep = emu.step_spec("laikago", 2, q, qd, g["action"][:n], **kw)
    out = emu.step_spec("laikago", 2, q, qd, g["action"][:n], auto_reset=True, reset_q=reset_q, **kw)
    assert np.array_equal(out["done"], keep["done"]) and np.array_equal(out["reward"], keep["reward"])
    d = out["done"] > 0
    assert d[:3].all() and not d.all()
    assert np.allclose(out["q"][d], reset_q.astype(np.float32)) and np.all(out["qd"][d] == 0)
    assert np.array_equal(out["q"][~d], keep["q"][~d]) and np.array_equal(out["qd"][~d], keep["qd"][~d])


def _random_quadruped_urdf(rng, leg_len, trunk_geom):
    """A random robot of the class the specialised kernel serves: xyz + xyz-rotation root chain (the reference's fixed-base
    emulation of a free trunk), four legs of `leg_len` revolute links each - the same joint kinds / shapes at the same depth in every
    leg (one structural class), every length, axis offset, mass, inertia and shape size drawn per leg - and optionally fixed toes."""
    v3 = lambda lo, hi: " ".join("%.6g" % x for x in rng.uniform(lo, hi, 3))
    inert = lambda m, I: f'<mass value="{m:.6g}"/><inertia ixx="{I[0]:.6g}" iyy="{I[1]:.6g}" izz="{I[2]:.6g}" ixy="0" ixz="0" iyz="0"/>'
    parts = ['<?xml version="1.0"?>', '<robot name="q">', '<link name="world"/>']
    prev = "world"
    for k, (ax, ty) in enumerate([("1 0 0", "prismatic"), ("0 1 0", "prismatic"), ("0 0 1", "prismatic"), ("1 0 0", "continuous"),
                                  ("0 1 0", "continuous"), ("0 0 1", "continuous")]):
        last = k == 5
        name = "body" if last else f"c{k}"
        col = (f'<collision><origin xyz="0 0 0"/><geometry><sphere radius="{rng.uniform(0.1, 0.2):.4g}"/></geometry></collision>'
               if (last and trunk_geom) else "")
        body = inert(rng.uniform(2, 8), rng.uniform(0.02, 0.3, 3)) if last else inert(0.0, np.zeros(3))
        parts.append(f'<link name="{name}"><inertial><origin xyz="0 0 0"/>{body}</inertial>{col}</link>')
        parts.append(f'<joint name="j{k}" type="{ty}"><parent link="{prev}"/><child link="{name}"/><origin xyz="0 0 0"/><axis xyz="{ax}"/></joint>')
        prev = name
    axes = [rng.choice(["1 0 0", "0 1 0", "0 0 1", "0 -1 0", "0.6 0 0.8"]) for _ in range(leg_len)]
    toe_fixed = rng.random() < 0.7
    geom_kind = [int(rng.integers(0, 3)) for _ in range(leg_len)]   # per depth: none / sphere / capsule
    for leg in range(4):
        par = "body"
        for d in range(leg_len):
            name = f"l{leg}_{d}"
            col = ""
            if geom_kind[d] == 1:
                col = f'<collision><origin xyz="{v3(-0.05, 0.05)}"/><geometry><sphere radius="{rng.uniform(0.03, 0.08):.4g}"/></geometry></collision>'
            if geom_kind[d] == 2:
                col = (f'<collision><origin xyz="{v3(-0.05, 0.05)}" rpy="{v3(-1, 1)}"/><geometry><capsule radius="{rng.uniform(0.03, 0.06):.4g}" '
                       f'length="{rng.uniform(0.1, 0.3):.4g}"/></geometry></collision>')
            parts.append(f'<link name="{name}"><inertial><origin xyz="{v3(-0.05, 0.05)}" rpy="{v3(-0.5, 0.5)}"/>'
                         f'{inert(rng.uniform(0.2, 1.5), rng.uniform(1e-3, 0.05, 3))}</inertial>{col}</link>')
            parts.append(f'<joint name="j{leg}_{d}" type="continuous"><parent link="{par}"/><child link="{name}"/>'
                         f'<origin xyz="{v3(-0.3, 0.3)}" rpy="{v3(-0.5, 0.5)}"/><axis xyz="{axes[d]}"/></joint>')
            par = name
        if toe_fixed:
            parts.append(f'<link name="toe{leg}"><inertial><origin xyz="0 0 0"/>{inert(0.0, np.zeros(3))}</inertial><collision><origin xyz="0 0 0"/>'
                         f'<geometry><sphere radius="{rng.uniform(0.02, 0.05):.4g}"/></geometry></collision></link>')
            parts.append(f'<joint name="jt{leg}" type="fixed"><parent link="{par}"/><child link="toe{leg}"/><origin xyz="{v3(-0.1, 0.1)}"/></joint>')
    parts.append("</robot>")
    return "\n".join(parts)


@pytest.mark.parametrize("seed", [2, 3])
def test_model_compiler_and_specialised_kernel_on_random_robots(seed, tmp_path):
    """Not only Laikago and Ant: a random quadruped goes through our URDF compiler, the ahead-of-time model compiler of the
    specialised kernel (csrc/gen_spec.cpp -> constexpr tables) and the kernel source itself (host build with that model), and must
    step like the reference built from the same flat model - forward dynamics, contact-free step and full step with dozens of
    penetrating points, 1 and 2 Gauss-Seidel sweeps.  fp64 arithmetic meets the 1e-5 bar with three orders of margin; the mixed
    arithmetic is reported against a looser bound (its error is model- and state-dependent: violent random states here)."""
    import ctypes
    import subprocess
    from oracle import ref
    from tds_b200.model import compile_urdf
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(seed)
    leg_len = int(rng.integers(2, 4))
    (tmp_path / "q.urdf").write_text(_random_quadruped_urdf(rng, leg_len, bool(seed % 2)))
    model = compile_urdf(str(tmp_path / "q.urdf"), os.path.join(GOLDEN, "urdf", "plane.urdf"), False)
    with open(tmp_path / "model.inc", "w") as f:
        vals = [repr(float(v)) for v in model]
        f.write("// flat model\n" + "".join(", ".join(vals[i:i + 6]) + ",\n" for i in range(0, len(vals), 6)))
    inc = ["-I" + emu.CSRC, "-I" + os.path.join(emu.ROOT, "include")]
    subprocess.check_call(["g++", "-std=c++17", "-O1"] + inc + [os.path.join(emu.CSRC, "gen_spec.cpp"), "-o", str(tmp_path / "gen_spec")])
    subprocess.check_call([str(tmp_path / "gen_spec"), "SpecTest", str(tmp_path / "model.inc"), str(4 * leg_len), "6", str(tmp_path / "spec_test.h")])
    so = str(tmp_path / "_steps_test.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-shared", "-fPIC", "-w", "-DTDSEMU_SPEC_TEST", "-I" + str(tmp_path)] + inc +
                          ["-I/usr/local/cuda/include", emu.SRC_S, "-o", so])
    L = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    L.tdsemu_steps.restype = ctypes.c_int
    L.tdsemu_steps.argtypes = [ctypes.c_int, dp, dp] + [ctypes.c_int] * 6 + [dp] * 12
    L.tdsemu_steps_tile_bytes.restype = ctypes.c_long
    assert 0 < L.tdsemu_steps_tile_bytes(1, 0) <= 227 * 1024        # the mixed instance of this robot fits a CTA's shared memory
    rs = ref.RefSim.from_model(model)
    params = dict(dt=1e-3, friction=0.9, keep_all_points=bool(seed % 3 == 0), pgs_iterations=1 + seed % 2)
    rs.set_params(**params)
    n, nq = 24, rs.n_q
    q = np.zeros((n, nq))
    q[:, 0:2], q[:, 2], q[:, 3:6], q[:, 6:] = rng.uniform(-0.3, 0.3, (n, 2)), rng.uniform(0.05, 0.6, n), rng.uniform(-0.6, 0.6, (n, 3)), rng.uniform(-1, 1, (n, nq - 6))
    qd, tau = rng.uniform(-1, 1, (n, nq)), rng.uniform(-3, 3, (n, nq))
    q, qd, tau = (a.astype(np.float32).astype(np.float64) for a in (q, qd, tau))
    pv = np.array([params["dt"], 0.0, 0.0, -9.81, params["friction"], 0.0, 0.2, 1e-5, params["pgs_iterations"], int(params["keep_all_points"])])
    penetrating = 0
    for precision, tol in ((1, TOL), (0, 1e-4)):
        for mode in (0, 1, 2):
            oq, oqd, oqdd = np.zeros((n, nq)), np.zeros((n, nq)), np.zeros((n, nq))
            rc = L.tdsemu_steps(1, emu._dp(pv), None, precision, 0, mode, 0, 0, n, emu._dp(q), emu._dp(qd), emu._dp(tau), emu._dp(oq), emu._dp(oqd),
                                emu._dp(oqdd), None, None, None, None, None, None)
            assert rc >= 0
            for i in range(n):
                r = rs.step(mode, q[i], qd[i], tau[i])
                if mode == 0:
                    assert rel_err(oqdd[i], r["qdd"]) <= tol
                else:
                    assert rel_err(oq[i], r["q"]) <= tol and rel_err(oqd[i], r["qd"]) <= tol
                if mode == 2 and precision == 1:
                    penetrating += int(np.sum(r["contact_data"][:, 9] < 0))
    assert penetrating >= 20
