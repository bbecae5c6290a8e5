"""The product's generic step kernel (csrc/tds_stepw.cu) executed on the CPU by compiling its SOURCE for the host
(tests/cpp/stepw_host.cpp): the golden vectors of the reference, the spring-damper branch against the oracle and the
dual-number Jacobian against central differences, without a GPU.  The GPU tests (tests/test_parity_gpu.py) check the same
kernel as nvcc builds it; this file keeps the kernel source honest in a container that has no GPU."""
import os

import numpy as np
import pytest

import tds_b200.workloads as wl
from tds_b200.model import fixture_path, load_model
from oracle import port
import emu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = ["cartpole", "pendulum5", "sphere2", "laikago", "humanoid", "ant", "box", "cartpole_plane", "pendulum5spherical", "humanoid_spherical"]
TOL = 1e-5


def rel_err(a, ref):
    return float(np.max(np.abs(a - ref) / np.maximum(1.0, np.abs(ref)))) if ref.size else 0.0


def params_from_golden(g):
    kw = {}
    for k in g.files:
        if k.startswith("param_"):
            v = g[k]
            kw[k[6:]] = tuple(v.tolist()) if v.ndim else (bool(v) if k == "param_keep_all_points" else float(v))
    return kw


@pytest.mark.parametrize("name", CONFIGS)
@pytest.mark.parametrize("precision", [0, 1])
def test_golden_vectors_through_the_kernel_source(name, precision):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = load_model(fixture_path(name))
    mode = int(g["mode"])
    tau = g["tau"] if "tau" in g.files else None
    n_tau = int(model[4]) - (6 if int(model[2]) else 0)
    if tau is not None and tau.shape[1] != n_tau:
        tau = tau[:, -n_tau:]
    out = emu.step(model, mode, g["q_in"], g["qd_in"], tau, precision=precision, **params_from_golden(g))
    if mode == 0:
        assert rel_err(out["qdd"], g["qdd"]) <= (TOL if precision == 1 else 5e-5)
        return
    tol = 5e-4 if (name in ("humanoid", "humanoid_spherical", "pendulum5spherical") and precision == 0) else TOL
    assert rel_err(out["q"], g["q_out"]) <= tol and rel_err(out["qd"], g["qd_out"]) <= tol
    if mode == 2:
        ref_d = np.stack(list(g["contact_dist"]))
        assert out["contact_dist"].shape == ref_d.shape and np.max(np.abs(out["contact_dist"] - ref_d)) < 2e-6


@pytest.mark.parametrize("name", ["sphere2", "laikago", "box"])
def test_spring_damper_branch_vs_oracle(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = load_model(fixture_path(name))
    params = params_from_golden(g)
    law = dict(spring_k=40000.0, damper_d=3000.0, exponent_n=1.5, v_transition=0.02, hard_contact_condition=True)
    tau = g["tau"] if "tau" in g.files else None
    out = emu.step(model, 2, g["q_in"], g["qd_in"], tau, precision=1, contact_model=1, **law, **params)
    P = port.make_params(contact_model=1, **law, **params)
    refs = [port.step(model, P, 2, g["q_in"][i], g["qd_in"][i], None if tau is None else tau[i]) for i in range(g["q_in"].shape[0])]
    assert rel_err(out["qd"], np.array([r["qd"] for r in refs])) <= TOL


@pytest.mark.parametrize("name,gen", [("pendulum5", wl.pendulum5), ("cartpole", wl.cartpole), ("sphere2", wl.sphere2)])
def test_dual_number_jacobian_vs_central_differences(name, gen):
    n = 6
    model = load_model(fixture_path(name))
    w = gen(n, seed=2718)
    mode, tau = w["mode"], w.get("tau")
    n_q, n_qd = int(model[3]), int(model[4])
    n_tau = n_qd - (6 if int(model[2]) else 0)
    t = None if tau is None or not n_tau else tau[:, -n_tau:]
    J = emu.step(model, mode, w["q"], w["qd"], t, jacobian=True, **w["params"])["jac"]
    P = port.make_params(**w["params"])
    ok = []
    for e in range(n):
        def f(x):
            r = port.step(model, P, mode, x[:n_q], x[n_q:n_q + n_qd], x[n_q + n_qd:] if n_tau else None)
            return r["qdd"] if mode == 0 else np.concatenate([r["q"], r["qd"]])
        x0 = np.concatenate([w["q"][e], w["qd"][e], t[e] if t is not None else np.zeros(0)])
        y0 = f(x0)
        Jr = np.zeros((y0.size, x0.size))
        for j in range(x0.size):
            xp, xm = x0.copy(), x0.copy(); xp[j] += 1e-6; xm[j] -= 1e-6
            Jr[:, j] = (f(xp) - f(xm)) / 2e-6
        ok.append(np.max(np.abs(J[e] - Jr) / np.maximum(1.0, np.abs(Jr))) <= 1e-4)
    assert np.mean(ok) >= (1.0 if mode != 2 else 0.8)


def _stiff_spherical_model():
    m = np.array(load_model(fixture_path("pendulum5spherical")))
    for i in range(5):
        m[16 + 13 + i * 34 + 32] = 3.0 + i        # Link::stiffness (acts through the quaternion's axis-angle, tiny_algebra.hpp:509-527)
        m[16 + 13 + i * 34 + 33] = 0.2 + 0.1 * i  # Link::damping
    return m


def test_spherical_joint_stiffness_and_damping_vs_reference():
    """forward_dynamics.hpp:56-109 with non-zero Link::stiffness / damping on spherical joints, incl. the small-angle branch of
    quaternion_axis_angle, against the reference itself (the C oracle does not restate spherical joints)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    m = _stiff_spherical_model()
    rs = ref.RefSim.from_model(m)
    w = wl.pendulum5spherical(10, seed=5)
    q = w["q"].copy()
    q[0, :4] = [0, 0, 0, 1]
    q[1, :4] = np.array([1e-5, 0, 0, 1]) / np.sqrt(1 + 1e-10)
    q = q.astype(np.float32).astype(np.float64)
    for mode in (0, 1):
        out = emu.step(m, mode, q, w["qd"], w["tau"], precision=1)
        for i in range(10):
            r = rs.step(mode, q[i], w["qd"][i], w["tau"][i])
            a, b = (out["qdd"][i], r["qdd"]) if mode == 0 else (np.concatenate([out["q"][i], out["qd"][i]]), np.concatenate([r["q"], r["qd"]]))
            assert rel_err(a, b) <= TOL


def test_spherical_dual_number_jacobian_vs_reference_differences():
    """The differentiable step through spherical joints (quaternion coordinates as inputs): dual numbers in the kernel
    source against central differences of the reference's own forward dynamics."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    m = _stiff_spherical_model()
    rs = ref.RefSim.from_model(m)
    w = wl.pendulum5spherical(3, seed=9)
    J = emu.step(m, 0, w["q"], w["qd"], w["tau"], jacobian=True)["jac"]
    assert J.shape == (3, 15, 20 + 15 + 15)
    for e in range(3):
        x0 = np.concatenate([w["q"][e], w["qd"][e], w["tau"][e]])
        f = lambda x: rs.step(0, x[:20], x[20:35], x[35:])["qdd"]
        Jr = np.zeros((15, x0.size))
        for j in range(x0.size):
            xp, xm = x0.copy(), x0.copy(); xp[j] += 1e-6; xm[j] -= 1e-6
            Jr[:, j] = (f(xp) - f(xm)) / 2e-6
        assert np.max(np.abs(J[e] - Jr) / np.maximum(1.0, np.abs(Jr))) <= 1e-4


@pytest.mark.parametrize("seed", range(16))
def test_random_models_through_the_kernel_source_vs_live_reference(seed, tmp_path):
    """Differential fuzzing of the any-model kernel: random trees (every joint kind incl. spherical, unit / negative / oblique axes,
    rotated joint and inertial origins, spheres / capsules / boxes with their own origins, fixed or floating base, with and
    without the plane, keep_all_points both ways, 1-3 Gauss-Seidel sweeps) compiled by OUR URDF compiler, stepped by the kernel
    source in fp64 arithmetic and by the reference itself (oracle/_ref on the same flat model): forward dynamics, the contact-free
    step and the full step."""
    import ctypes
    from oracle import ref
    from tds_b200 import _lib
    from tds_b200.model import compile_urdf
    from test_model_compiler import _random_urdf
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(9000 + seed)
    text = _random_urdf(rng, int(rng.integers(2, 11)), massless_links=False, boxes=True)
    if seed % 4 == 0:
        text = text.replace('type="continuous"', 'type="spherical"', 2)
    path = tmp_path / "rnd.urdf"
    path.write_text(text)
    floating = bool(seed % 2) and "spherical" not in text
    plane = os.path.join(GOLDEN, "urdf", "plane.urdf") if seed % 3 else None
    model = compile_urdf(str(path), plane, floating)
    if _lib.lib().tds_b200_validate_model(model.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), model.size):
        pytest.skip("beyond the capacity of the flat format: " + _lib.last_error())
    rs = ref.RefSim.from_model(model)
    params = dict(dt=1e-3, friction=0.8, keep_all_points=bool(seed % 3 == 0), pgs_iterations=1 + seed % 3)
    rs.set_params(**params)
    n, nq, nqd, nt = 6, rs.n_q, rs.n_qd, rs.n_tau
    q, qd, tau = rng.uniform(-0.8, 0.8, (n, nq)), rng.uniform(-1, 1, (n, nqd)), rng.uniform(-2, 2, (n, max(nt, 1)))[:, :nt]
    for l in model[16 + 13:16 + 13 + int(model[1]) * 34].reshape(-1, 34):
        if int(l[1]) == 8:                                   # JOINT_SPHERICAL: unit quaternion
            k = int(l[2]); v = rng.normal(size=(n, 4)); q[:, k:k + 4] = v / np.linalg.norm(v, axis=1, keepdims=True)
    if floating:
        v = rng.normal(size=(n, 4)); q[:, :4] = v / np.linalg.norm(v, axis=1, keepdims=True)
        q[:, 4:6] = rng.uniform(-1, 1, (n, 2)); q[:, 6] = rng.uniform(0.0, 0.6, n)
    q, qd, tau = (a.astype(np.float32).astype(np.float64) for a in (q, qd, tau))
    for mode in (0, 1, 2):
        out = emu.step(model, mode, q, qd, tau if nt else None, precision=1, **params)
        for i in range(n):
            r = rs.step(mode, q[i], qd[i], tau[i] if nt else None)
            if mode == 0:
                assert rel_err(out["qdd"][i], r["qdd"]) <= TOL
            else:
                assert rel_err(out["q"][i], r["q"]) <= TOL and rel_err(out["qd"][i], r["qd"]) <= TOL


@pytest.mark.parametrize("seed", range(6))
def test_dual_number_jacobian_on_random_models_vs_reference(seed, tmp_path):
    """The differentiable instance on random trees (spherical joints, oblique axes, floating bases: derivatives with respect to the
    raw quaternion components, as the reference's own differentiation would see them) against central differences of the
    reference: forward dynamics and the contact-free step (smooth maps: every environment must agree)."""
    from oracle import ref
    from tds_b200.model import compile_urdf
    from test_model_compiler import _random_urdf
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(7000 + seed)
    text = _random_urdf(rng, int(rng.integers(2, 7)), massless_links=False, boxes=False)
    if seed % 2 == 0:
        text = text.replace('type="continuous"', 'type="spherical"', 1)
    path = tmp_path / "rnd.urdf"
    path.write_text(text)
    floating = bool(seed % 2) and "spherical" not in text
    model = compile_urdf(str(path), None, floating)
    rs = ref.RefSim.from_model(model)
    rs.set_params()
    n, nq, nqd, nt = 2, rs.n_q, rs.n_qd, rs.n_tau
    if nqd == 0:
        pytest.skip("a tree of fixed joints")
    q, qd, tau = rng.uniform(-0.8, 0.8, (n, nq)), rng.uniform(-1, 1, (n, nqd)), rng.uniform(-2, 2, (n, max(nt, 1)))[:, :nt]
    for l in model[16 + 13:16 + 13 + int(model[1]) * 34].reshape(-1, 34):
        if int(l[1]) == 8:
            k = int(l[2]); v = rng.normal(size=(n, 4)); q[:, k:k + 4] = v / np.linalg.norm(v, axis=1, keepdims=True)
    if floating:
        v = rng.normal(size=(n, 4)); q[:, :4] = v / np.linalg.norm(v, axis=1, keepdims=True)
    q, qd, tau = (a.astype(np.float32).astype(np.float64) for a in (q, qd, tau))
    for mode in (0, 1):
        J = emu.step(model, mode, q, qd, tau if nt else None, jacobian=True)["jac"]
        for e in range(n):
            def f(x):
                r = rs.step(mode, x[:nq], x[nq:nq + nqd], x[nq + nqd:] if nt else None)
                return r["qdd"] if mode == 0 else np.concatenate([r["q"], r["qd"]])
            x0 = np.concatenate([q[e], qd[e], tau[e]])
            Jr = np.zeros((f(x0).size, x0.size))
            for j in range(x0.size):
                xp, xm = x0.copy(), x0.copy(); xp[j] += 1e-6; xm[j] -= 1e-6
                Jr[:, j] = (f(xp) - f(xm)) / 2e-6
            assert np.max(np.abs(J[e] - Jr) / np.maximum(1.0, np.abs(Jr))) <= 1e-6
