"""Worlds of several multibodies (contacts between multibodies, src/world.hpp:206-282 - sphere-sphere and
capsule-sphere through the dispatcher, one LCP per list of World::mb_contacts_ solved one after the other, :351-355): the
contact stage of the world-frame kernel csrc/tds_stepw.cu, executed on the CPU from its SOURCE (tests/cpp/stepw_host.cpp),
against golden vectors of the reference (tests/golden/mb_*.npz, tests/golden/make_golden_multibody.py) and the live reference
(oracle/_ref, oracle/ref/ref_world.cpp).  The GPU twins are in tests/test_parity_gpu.py."""
import os

import numpy as np
import pytest

import tds_b200.workloads as wl
from tds_b200.model import merge_models, model_dims
from oracle import port
import emu
from test_kernel_source_on_host import GOLDEN, TOL, params_from_golden, rel_err


@pytest.mark.parametrize("kind", wl.MULTIBODY_WORLDS)
@pytest.mark.parametrize("precision", [0, 1])
def test_multibody_golden_vectors_through_the_kernel_source(kind, precision):
    g = np.load(os.path.join(GOLDEN, "mb_" + kind + ".npz"))
    assert np.array_equal(g["model"], wl.multibody_world_model(kind))   # the committed model is the one the package builds
    params = params_from_golden(g)
    tol = TOL if precision == 1 else 3e-5
    out = emu.step(g["model"], 2, g["q_in"], g["qd_in"], g["tau"], precision=precision, **params)
    assert rel_err(out["q"], g["q_out"]) <= tol and rel_err(out["qd"], g["qd_out"]) <= tol
    out = emu.step(g["model"], 3, g["q_in"], g["qd_in"], None, precision=precision, **params)   # World::step alone
    assert rel_err(out["qd"], g["qd_world_step"]) <= tol
    # the fixtures do exercise what they are for: contacts between multibodies, with and without plane contacts around them
    pen = g["contact_data"][..., 9] < 0
    lists = g["contact_idx"][..., 0]
    n_plane_lists = int(g["model"][12])
    assert np.any(pen & (lists >= n_plane_lists)) and np.any(pen & (lists >= 0) & (lists < n_plane_lists))


@pytest.mark.parametrize("kind", wl.MULTIBODY_WORLDS)
@pytest.mark.parametrize("sweep", [dict(), dict(pgs_iterations=5), dict(keep_all_points=True, friction=0.3),
                                   dict(restitution=0.4, erp=0.1, cfm=1e-3, pgs_iterations=3), dict(dt=4e-3, gravity=(0.3, 0.0, -9.0))])
def test_multibody_fresh_states_and_solver_parameters_vs_live_reference(kind, sweep):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    n = 48
    w = wl.multibody_world(kind, n, seed=2024)
    params = dict(w["params"]); params.update(sweep)
    rw = ref.RefWorld(w["model"])
    rw.set_params(**params)
    for mode in (2, 3):
        out = emu.step(w["model"], mode, w["q"], w["qd"], w["tau"], precision=1, **params)
        for i in range(n):
            r = rw.step(mode, w["q"][i], w["qd"][i], w["tau"][i])
            assert rel_err(out["qd"][i], r["qd"]) <= TOL
            if mode == 2:
                assert rel_err(out["q"][i], r["q"]) <= TOL


def test_multibodies_far_apart_step_like_separate_worlds():
    """No contact between the multibodies: the merged world must reproduce each multibody stepped alone on the plane (the C
    oracle, single multibody) - forest handling of the dynamics passes and the plane LCP shared by uncoupled multibodies."""
    n = 32
    a = wl.free_body_model(1.0, (0.036,) * 3, [("sphere", 0.3, (0, 0, 0))])
    b = wl.free_body_model(2.0, (0.05, 0.05, 0.01), [("capsule", 0.15, 0.6, (0, 0, 0.05), wl._rot_y(0.3))], arm=(0.5, 0.3, 0.12))
    world = merge_models([a, b])
    assert model_dims(world)["n_links"] == 13 and int(world[12]) == 2
    r = np.random.default_rng(5)
    q = np.zeros((n, 13)); qd = r.uniform(-1, 1, (n, 13)); tau = r.uniform(-1, 1, (n, 13))
    q[:, 0:2] = r.uniform(-0.3, 0.3, (n, 2)); q[:, 2] = r.uniform(0.2, 0.4, n); q[:, 3:6] = r.uniform(-1, 1, (n, 3))
    q[:, 6:8] = 5.0 + r.uniform(-0.3, 0.3, (n, 2)); q[:, 8] = r.uniform(0.1, 0.5, n); q[:, 9:13] = r.uniform(-1, 1, (n, 4))
    q, qd, tau = wl._f32(q), wl._f32(qd), wl._f32(tau)
    params = dict(friction=0.7, keep_all_points=False)
    out = emu.step(world, 2, q, qd, tau, precision=1, **params)
    P = port.make_params(**params)
    for i in range(n):
        ra = port.step(a, P, 2, q[i, :6], qd[i, :6], tau[i, :6])
        rb = port.step(b, P, 2, q[i, 6:], qd[i, 6:], tau[i, 6:])
        assert rel_err(out["q"][i], np.concatenate([ra["q"], rb["q"]])) <= TOL
        assert rel_err(out["qd"][i], np.concatenate([ra["qd"], rb["qd"]])) <= TOL


def test_dual_number_jacobian_through_contacts_between_multibodies():
    """d(q', qd') / d(q, qd, tau) of the full step by forward-mode dual numbers in the kernel against central differences of the
    reference's step (environments whose contact set changes inside the difference stencil are excluded by the 80 % bar)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    n = 6
    w = wl.multibody_world("capsule_sphere", n, seed=77)
    rw = ref.RefWorld(w["model"])
    rw.set_params(**w["params"])
    J = emu.step(w["model"], 2, w["q"], w["qd"], w["tau"], jacobian=True, **w["params"])["jac"]
    ok = []
    for e in range(n):
        def f(x):
            r = rw.step(2, x[:12], x[12:24], x[24:])
            return np.concatenate([r["q"], r["qd"]])
        x0 = np.concatenate([w["q"][e], w["qd"][e], w["tau"][e]])
        Jr = np.zeros((24, 36))
        for j in range(36):
            xp, xm = x0.copy(), x0.copy(); xp[j] += 1e-6; xm[j] -= 1e-6
            Jr[:, j] = (f(xp) - f(xm)) / 2e-6
        ok.append(np.max(np.abs(J[e] - Jr) / np.maximum(1.0, np.abs(Jr))) <= 1e-4)
    assert np.mean(ok) >= 0.8


def test_merge_models_refuses_what_the_reference_world_cannot_mean():
    a = wl.free_body_model(1.0, (0.036,) * 3, [("sphere", 0.3, (0, 0, 0))])
    with pytest.raises(ValueError):
        merge_models([a])
    floating = a.copy(); floating[2] = 1
    with pytest.raises(ValueError):
        merge_models([a, floating])
    two_roots = merge_models([a, a])   # a world is not a multibody: it cannot be merged again as one
    with pytest.raises(ValueError):
        merge_models([two_roots, a])


def _reference_lists(g):
    """(body_a, link_a, body_b, link_b) per contact of the reference's step, from the list index: World::mb_contacts_ holds one list
    per pair of multibodies (i < j), the plane being multibody 0 (src/world.hpp:212-281)."""
    k = int(g["model"][12])
    pairs = [(i, j) for i in range(k + 1) for j in range(i + 1, k + 1)]
    out = []
    for e in range(g["q_in"].shape[0]):
        rows = g["contact_idx"][e, :g["n_contacts"][e]]
        out.append(np.array([[pairs[l][0], a, pairs[l][1], b] for l, a, b in rows], dtype=np.int32).reshape(-1, 4))
    return out


@pytest.mark.parametrize("kind", wl.MULTIBODY_WORLDS)
def test_multibody_candidate_lists_and_distances_match_the_reference(kind):
    """Contact-pair index lists (north_star: bit-exact): the static candidate list of the C-ABI (host-only call) equals what the
    reference's World::mb_contacts_ holds after a step - multibody and link indices, list after list - and the distances the
    kernel reports per candidate are the reference's contact distances."""
    import ctypes
    from tds_b200 import _lib
    g = np.load(os.path.join(GOLDEN, "mb_" + kind + ".npz"))
    m = np.ascontiguousarray(g["model"])
    L = _lib.lib()
    t = np.zeros((128, 4), dtype=np.int32)
    k = L.tds_b200_model_contact_pairs(m.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), m.size, ctypes.c_void_p(t.ctypes.data), 128)
    assert k > 0
    ref_lists = _reference_lists(g)
    for e, r in enumerate(ref_lists):
        assert np.array_equal(t[:k], r), e    # every candidate emits a point in these fixtures (no coincident centres)
    out = emu.step(g["model"], 2, g["q_in"], g["qd_in"], g["tau"], precision=1, **params_from_golden(g))
    assert out["contact_dist"].shape == (g["q_in"].shape[0], k)
    ref_d = np.array([g["contact_data"][e, :k, 9] for e in range(g["q_in"].shape[0])])
    assert np.max(np.abs(out["contact_dist"] - ref_d)) < 2e-6


def test_world_of_urdf_multibodies_vs_live_reference():
    """Two URDF multibodies (tests/golden/urdf/free_*.urdf) through the model compiler, merged into one world: the compiled models
    equal the reference loader's, and the step of the merged world equals the reference World holding the same multibodies."""
    from oracle import ref
    from tds_b200.model import compile_urdf
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    here = os.path.join(GOLDEN, "urdf")
    plane = os.path.join(here, "plane.urdf")
    models = []
    for name in ("free_capsule", "free_sphere"):
        m = compile_urdf(os.path.join(here, name + ".urdf"), plane, False)
        assert np.array_equal(m, ref.RefSim.from_urdf(os.path.join(here, name + ".urdf"), plane, False).export_model())
        models.append(m)
    world = merge_models(models)
    w = wl.multibody_world("capsule_sphere", 32, seed=31)    # same shapes of state: two free bodies of 6 coordinates
    rw = ref.RefWorld(world)
    rw.set_params(**w["params"])
    out = emu.step(world, 2, w["q"], w["qd"], w["tau"], precision=1, **w["params"])
    hit = 0
    for i in range(32):
        r = rw.step(2, w["q"][i], w["qd"][i], w["tau"][i])
        assert rel_err(out["q"][i], r["q"]) <= TOL and rel_err(out["qd"][i], r["qd"]) <= TOL
        hit += int(np.any((r["contact_idx"][:, 0] == 2) & (r["contact_data"][:, 9] < 0)))
    assert hit >= 8   # contacts between the two multibodies do occur


@pytest.mark.parametrize("seed", range(8))
def test_random_worlds_of_multibodies_vs_live_reference(seed):
    """Differential fuzzing of the contact stage between multibodies: two to four random free bodies (1-3 spheres / capsules / boxes /
    plane shapes each at random offsets, some on xyz + spherical joints, some with an extra revolute arm), random solver parameters."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(4400 + seed)
    spherical = bool(seed % 2)
    bodies, dofs = [], []
    for b in range(int(rng.integers(2, 5))):
        geoms = []
        for _ in range(int(rng.integers(1, 4))):
            off, u = tuple(rng.uniform(-0.25, 0.25, 3)), rng.random()
            if u < 0.35:
                geoms.append(("sphere", float(rng.uniform(0.1, 0.3)), off))
            elif u < 0.65:
                geoms.append(("capsule", float(rng.uniform(0.08, 0.15)), float(rng.uniform(0.2, 0.6)), off, wl._rot_y(float(rng.uniform(-1.5, 1.5)))))
            elif u < 0.85:      # boxes only meet the plane and plane shapes
                geoms.append(("box", tuple(rng.uniform(0.1, 0.5, 3)), off, wl._rot_y(float(rng.uniform(-1, 1)))))
            else:               # a plane shape on the body link (tilted)
                nrm = rng.normal(size=3) * 0.3 + np.array([0.0, 0.0, 1.0])
                geoms.append(("plane", tuple(nrm / np.linalg.norm(nrm))))
        arm = (0.4, 0.3, 0.1) if (not spherical and rng.random() < 0.4) else None
        m = float(rng.uniform(0.5, 3.0))
        bodies.append(wl.free_body_model(m, tuple(rng.uniform(0.02, 0.1, 3)), geoms, arm=arm, spherical=spherical))
        dofs.append((int(bodies[-1][3]), int(bodies[-1][4])))
    world = merge_models(bodies)
    n, nq, nqd = 12, int(world[3]), int(world[4])
    q, qd, tau = np.zeros((n, nq)), rng.uniform(-1, 1, (n, nqd)), rng.uniform(-1, 1, (n, nqd))
    centre = rng.uniform(-0.3, 0.3, (n, 2))
    o = 0
    for bq, _ in dofs:
        q[:, o:o + 2] = centre + rng.uniform(-0.25, 0.25, (n, 2)); q[:, o + 2] = rng.uniform(0.1, 0.5, n)
        if spherical:
            v = rng.normal(size=(n, 4)); q[:, o + 3:o + 7] = v / np.linalg.norm(v, axis=1, keepdims=True)
        else:
            q[:, o + 3:o + bq] = rng.uniform(-1, 1, (n, bq - 3))
        o += bq
    q, qd, tau = wl._f32(q), wl._f32(qd), wl._f32(tau)
    params = dict(friction=float(rng.uniform(0.2, 1.0)), restitution=float(rng.uniform(0, 0.5)), pgs_iterations=int(rng.integers(1, 5)),
                  keep_all_points=bool(seed % 3 == 0))
    rw = ref.RefWorld(world)
    rw.set_params(**params)
    try:
        out = emu.step(world, 2, q, qd, tau, precision=1, **params)
    except RuntimeError:
        pytest.skip("more candidate points than the flat format holds (boxes: 8 each); tds_b200_create refuses such a world")
    for i in range(n):
        r = rw.step(2, q[i], qd[i], tau[i], contact_cap=256)
        assert rel_err(out["q"][i], r["q"]) <= TOL and rel_err(out["qd"][i], r["qd"]) <= TOL
        assert r["n_contacts"] == out["contact_dist"].shape[1]          # same number of candidate points as the reference emits


def _enumerate_like_the_reference(model):
    """Independent restatement of the loops of World::compute_contacts_multi_body_internal (src/world.hpp:212-281) with the plane as
    multibody 0 and the dispatcher's point counts: (mb_a, link_a, geom_a, mb_b, link_b, geom_b) per emitted point."""
    m = np.asarray(model)
    n_links, n_geoms, has_plane, k = int(m[1]), int(m[5]), int(m[7]), max(int(m[12]), 1)
    L = m[16 + 13:16 + 13 + n_links * 34].reshape(n_links, 34)
    G = m[16 + 13 + n_links * 34:16 + 13 + n_links * 34 + n_geoms * 18].reshape(n_geoms, 18)
    body_of, first = [], {}
    for i in range(n_links):
        b = (len(first) if int(L[i, 0]) < 0 else body_of[int(L[i, 0])]) if k > 1 else 0
        first.setdefault(b, i)
        body_of.append(b)
    bodies = []      # per multibody: list of (local link, [(geom index in link, type)])
    for b in range(k):
        links = [(-1, [(gi, int(g[1])) for gi, g in enumerate(G[G[:, 0] < 0])] if (b == 0 and k == 1) else [])]
        for i in range(n_links):
            if body_of[i] == b:
                gs = G[G[:, 0] == i]
                links.append((i - first[b], [(gi, int(g[1])) for gi, g in enumerate(gs)]))
        bodies.append(links)
    world = ([[(-1, [(0, 1)])]] if has_plane else []) + bodies     # the plane: one PLANE geom on its base
    off = 0 if has_plane else 1
    pts = {(1, 0): 1, (1, 2): 2, (1, 4): 8, (0, 0): 1, (2, 0): 2}   # [type a][type b] -> points (contact_point.hpp:468-473)
    out = []
    for i in range(len(world)):
        for j in range(i + 1, len(world)):
            for la, ga in world[i]:
                for gia, ta in ga:
                    for lb, gb in world[j]:
                        for gib, tb in gb:
                            n = pts.get((ta, tb), pts.get((tb, ta), 0))
                            out += [[i + off, la, gia, j + off, lb, gib]] * n
    return np.array(out, dtype=np.int32).reshape(-1, 6)


@pytest.mark.parametrize("name", ["sphere2", "laikago", "humanoid", "ant", "box", "cartpole_plane", "humanoid_spherical"] + ["mb_" + k for k in wl.MULTIBODY_WORLDS])
def test_contact_tuples_with_geometry_indices(name):
    """(mb_a, link_a, geom_a, mb_b, link_b, geom_b) per candidate through the C-ABI (host-only call) against an independent
    enumeration of the reference's loops; the first, second, fourth and fifth columns are also what the goldens pin."""
    import ctypes
    from tds_b200 import _lib
    from tds_b200.model import fixture_path, load_model
    model = np.load(os.path.join(GOLDEN, name + ".npz"))["model"] if name.startswith("mb_") else load_model(fixture_path(name))
    m = np.ascontiguousarray(model, dtype=np.float64)
    L = _lib.lib()
    t6 = np.zeros((128, 6), dtype=np.int32)
    k = L.tds_b200_model_contact_tuples(ctypes.c_void_p(m.ctypes.data), m.size, ctypes.c_void_p(t6.ctypes.data), 128)
    want = _enumerate_like_the_reference(m)
    assert k == want.shape[0] and k > 0 and np.array_equal(t6[:k], want)
    t4 = np.zeros((128, 4), dtype=np.int32)
    assert L.tds_b200_model_contact_pairs(m.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), m.size, ctypes.c_void_p(t4.ctypes.data), 128) == k
    assert np.array_equal(t4[:k], t6[:k][:, [0, 1, 3, 4]])


def test_rollout_of_a_world_of_multibodies_tracks_the_reference():
    """80 chained steps of two free bodies on xyz + spherical joints (no Euler-angle singularity to run into: with the x-y-z revolute
    emulation a body tumbling through pitch = 90 degrees makes the reference's own mass matrix singular and its velocities explode -
    which the kernel reproduces digit for digit, but is no test), falling onto the plane and onto each other; state carried in fp32
    between steps as on the device: per-step error against the reference stepping from the kernel's own previous state."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    n, steps = 6, 80
    w = wl.multibody_world("spherical_pair", n, seed=12)
    params = dict(w["params"], dt=2e-3)
    rw = ref.RefWorld(w["model"])
    rw.set_params(**params)
    q, qd = w["q"], 0.2 * w["qd"]
    worst, pair_steps = 0.0, 0
    for s in range(steps):
        q32, qd32 = wl._f32(q), wl._f32(qd)
        out = emu.step(w["model"], 2, q32, qd32, None, precision=1, **params)
        for i in range(n):
            r = rw.step(2, q32[i], qd32[i], None, contact_cap=128)
            worst = max(worst, rel_err(out["q"][i], r["q"]), rel_err(out["qd"][i], r["qd"]))
            pair_steps += int(np.any((r["contact_idx"][:, 0] >= 2) & (r["contact_data"][:, 9] < 0)))
        q, qd = out["q"], out["qd"]
    assert worst <= TOL and np.all(np.isfinite(q)) and np.max(np.abs(qd)) < 100.0 and pair_steps >= 20
