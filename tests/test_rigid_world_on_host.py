"""The RigidBody path of World::step (src/world.hpp:293-363, src/rigid_body.hpp, src/rb_constraint_solver.hpp) - the
kernel csrc/tds_rigid.cu executed on the CPU from its SOURCE (tests/cpp/rigid_host.cpp) against golden vectors of the reference
(tests/golden/rigid_*.npz, make_golden_rigid.py) and the live reference (oracle/ref/ref_rigid.cpp).  Both sides compute in fp64:
the bar is 1e-12.  GPU twins: tests/test_parity_gpu.py::test_rigid_world_*."""
import ctypes
import os

import numpy as np
import pytest

import tds_b200.workloads as wl
from tds_b200 import rigid as rg
import emu
from test_kernel_source_on_host import GOLDEN, params_from_golden

TOL = 1e-12


@pytest.mark.parametrize("kind", wl.RIGID_WORLDS)
def test_rigid_golden_vectors_through_the_kernel_source(kind):
    g = np.load(os.path.join(GOLDEN, "rigid_" + kind + ".npz"))
    w = wl.rigid_world(kind, g["state"].shape[0])
    assert np.array_equal(w["bodies"], g["bodies"]) and np.array_equal(w["state"], g["state"])   # the committed inputs are the package's
    params = params_from_golden(g)
    params["num_solver_iterations"] = int(params["num_solver_iterations"])
    one = emu.rigid_step(g["bodies"], g["state"], g["force"], 1, **params)
    five = emu.rigid_step(g["bodies"], g["state"], g["force"], 5, **params)
    assert np.max(np.abs(one - g["state_1"])) <= TOL and np.max(np.abs(five - g["state_5"])) <= TOL
    # the contacts matter in these fixtures: without solver sweeps the velocities differ in most worlds
    free = emu.rigid_step(g["bodies"], g["state"], g["force"], 1, **dict(params, num_solver_iterations=0))
    assert np.mean(np.max(np.abs(free - one), axis=(1, 2)) > 1e-6) > 0.5


@pytest.mark.parametrize("kind", wl.RIGID_WORLDS)
@pytest.mark.parametrize("sweep", [dict(), dict(friction=0.0), dict(restitution=0.8, erp=0.3, num_solver_iterations=4), dict(dt=1e-3, gravity=(0.5, 0.0, -9.0))])
def test_rigid_fresh_states_and_parameters_vs_live_reference(kind, sweep):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    n = 32
    w = wl.rigid_world(kind, n, seed=808)
    params = dict(w["params"]); params.update(sweep)
    rw = ref.RefRigidWorld(w["bodies"])
    rw.set_params(**params)
    for steps, force in ((1, w["force"]), (3, None), (20, w["force"])):
        out = emu.rigid_step(w["bodies"], w["state"], force, steps, **params)
        for i in range(n):
            r, _ = rw.step(w["state"][i], None if force is None else force[i], steps)
            assert np.max(np.abs(out[i] - r)) <= (TOL if steps < 20 else 1e-9)   # 20 chained steps: round-off through the contact branches


def test_rigid_jacobian_by_dual_numbers_vs_central_differences():
    """What python/examples/billiard_optimization.py differentiates: the state after a few steps with respect to the force on the
    white ball (and everything else), forward-mode in the kernel against central differences of the reference."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    n, steps = 4, 3
    w = wl.rigid_world("billiard", n, seed=5)
    rw = ref.RefRigidWorld(w["bodies"])
    rw.set_params(**w["params"])
    out, J = emu.rigid_step(w["bodies"], w["state"], w["force"], steps, jacobian=True, **w["params"])
    assert J.shape == (n, 91, 112)
    ok = []
    for e in range(n):
        assert np.max(np.abs(out[e] - rw.step(w["state"][e], w["force"][e], steps)[0])) <= TOL
        x0 = np.concatenate([w["state"][e].ravel(), w["force"][e].ravel()])
        f = lambda x: rw.step(x[:91].reshape(7, 13), x[91:].reshape(7, 3), steps)[0].ravel()
        Jr = np.zeros((91, 112))
        for j in range(112):
            xp, xm = x0.copy(), x0.copy(); xp[j] += 1e-6; xm[j] -= 1e-6
            Jr[:, j] = (f(xp) - f(xm)) / 2e-6
        ok.append(np.max(np.abs(J[e] - Jr) / np.maximum(1.0, np.abs(Jr))) <= 1e-5)
    assert np.mean(ok) >= 0.75    # a contact switching inside the difference stencil spoils the finite differences, not the duals


def test_rigid_create_refuses_bad_descriptions_on_the_host():
    """tds_b200_rigid_create validates before it touches the GPU: unknown shapes, too many bodies / candidate contacts -> NULL + reason."""
    from tds_b200 import _lib
    L = _lib.lib()
    def create(bodies):
        d = np.ascontiguousarray(bodies, dtype=np.float64)
        return L.tds_b200_rigid_create(ctypes.c_void_p(d.ctypes.data), d.shape[0], 1, 0)
    assert not create([[1.0, 3, 0.1, 0, 0, 0]]) and "shapes" in _lib.last_error()          # mesh
    assert not create([rg.sphere(1.0, 0.1)] * 17) and "bodies" in _lib.last_error()
    assert not create([rg.plane()] + [rg.box(1.0, (1, 1, 1))] * 7) and "candidate" in _lib.last_error()   # 7 x 8 corners


@pytest.mark.parametrize("seed", range(12))
def test_random_rigid_worlds_vs_live_reference(seed):
    """Differential fuzzing: 2-8 random bodies (spheres, capsules, boxes, tilted planes with a constant, static bodies), random
    World parameters (dt, gravity, friction, restitution, erp, 0-30 solver iterations), 3 steps with an external force."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    r = np.random.default_rng(31000 + seed)
    nb = int(r.integers(2, 9))
    bodies = []
    for _ in range(nb):
        k, m = int(r.integers(0, 5)), (0.0 if r.random() < 0.15 else float(r.uniform(0.3, 4)))
        if k in (0, 4):
            bodies.append(rg.sphere(m, float(r.uniform(0.1, 0.4))))
        elif k == 1:
            bodies.append(rg.capsule(m, float(r.uniform(0.05, 0.2)), float(r.uniform(0.2, 0.8))))
        elif k == 2:
            bodies.append(rg.box(m, tuple(r.uniform(0.1, 0.6, 3))))
        else:
            bodies.append(rg.plane(tuple(r.normal(size=3) * 0.2 + np.array([0, 0, 1])), float(r.uniform(-0.2, 0.2))))
    n = 6
    s = np.zeros((n, nb, 13))
    s[:, :, 0:3] = r.uniform(-0.5, 0.5, (n, nb, 3))
    q = r.normal(size=(n, nb, 4)); s[:, :, 3:7] = q / np.linalg.norm(q, axis=2, keepdims=True)
    s[:, :, 7:13] = r.uniform(-2, 2, (n, nb, 6))
    f = r.uniform(-30, 30, (n, nb, 3))
    params = dict(dt=float(r.choice([1 / 60, 1e-3, 5e-3])), gravity=tuple(r.uniform(-1, 1, 2)) + (-9.81,), friction=float(r.uniform(0, 1)),
                  restitution=float(r.uniform(0, 0.9)), erp=float(r.uniform(0, 0.4)), num_solver_iterations=int(r.integers(0, 30)))
    try:
        out = emu.rigid_step(bodies, s, f, 3, **params)
    except RuntimeError:
        pytest.skip("more candidate contact points than the kernel's list holds (tds_b200_rigid_create refuses the world)")
    rw = ref.RefRigidWorld(bodies)
    rw.set_params(**params)
    for i in range(n):
        o, _ = rw.step(s[i], f[i], 3)
        # (two static bodies in contact divide by inv_mass_a + inv_mass_b + ang = 0 in the reference: NaN there, NaN here)
        assert np.array_equal(np.isnan(o), np.isnan(out[i]))
        ok = ~np.isnan(o)
        assert np.max(np.abs(o[ok] - out[i][ok]) / np.maximum(1.0, np.abs(o[ok]))) <= 1e-10
