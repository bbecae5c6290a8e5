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
