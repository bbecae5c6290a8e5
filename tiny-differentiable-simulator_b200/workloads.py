"""Synthetic, deterministic inputs for the BASELINE.json configs.
seed 12345 = ARSConfig::env_seed (examples/ars/ars_config.h:10).  Values are rounded to fp32 so the
fp64 oracle and the fp32-state GPU path see identical inputs."""
import numpy as np

SEED = 12345


def _f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def cartpole(n, seed=SEED):
    """C1: cartpole.urdf, q,qd ~ U(-0.05,0.05), tau0 ~ U(-10,10), dt=1/60, g=-10 (cartpole_environment2.h)."""
    r = np.random.default_rng(seed)
    q = r.uniform(-0.05, 0.05, (n, 2))
    qd = r.uniform(-0.05, 0.05, (n, 2))
    tau = np.zeros((n, 2))
    tau[:, 0] = r.uniform(-10, 10, n)
    return dict(q=_f32(q), qd=_f32(qd), tau=_f32(tau), params=dict(dt=1.0 / 60.0, gravity=(0.0, 0.0, -10.0)), mode=1)


def pendulum5(n, seed=SEED):
    """C2: pendulum5.urdf, forward_dynamics only."""
    r = np.random.default_rng(seed)
    return dict(q=_f32(r.uniform(-np.pi, np.pi, (n, 5))), qd=_f32(r.uniform(-2, 2, (n, 5))),
                tau=_f32(r.uniform(-1, 1, (n, 5))), params=dict(), mode=0)


def sphere2(n, seed=SEED):
    """C3: sphere2.urdf (floating) on plane_implicit: ~50% penetrating, friction 0.5, keep_all_points false."""
    r = np.random.default_rng(seed)
    q = np.zeros((n, 7))
    q[:, 3] = 1.0
    q[:, 4:6] = r.uniform(-1, 1, (n, 2))
    q[:, 6] = r.uniform(0.45, 0.55, n)
    qd = r.uniform(-1, 1, (n, 6))
    return dict(q=_f32(q), qd=_f32(qd), tau=None, params=dict(friction=0.5, keep_all_points=False), mode=2)


def laikago(n, seed=SEED):
    """C4: Laikago on plane; LaikagoContactSimulation::reset state + noise, actions ~ U(-0.4, 0.4)."""
    r = np.random.default_rng(seed)
    q = np.zeros((n, 18))
    q[:, 2] = 0.48
    q[:, 6:18] = np.array([0.2, 0.0, -0.7] * 4) + 0.05 * (r.random((n, 12)) - 0.5) * 2.0
    qd = np.zeros((n, 18))
    act = r.uniform(-0.4, 0.4, (n, 12))
    return dict(q=_f32(q), qd=_f32(qd), action=_f32(act), params=dict(friction=1.0, keep_all_points=True), mode=2)


def laikago_perturbed(n, seed=SEED):
    """Laikago states spread over the reachable set (for parity tests): base height such that 0-4 toes
    penetrate, non-zero base orientation and velocities."""
    r = np.random.default_rng(seed)
    q = np.zeros((n, 18))
    q[:, 0:2] = r.uniform(-0.2, 0.2, (n, 2))
    q[:, 2] = r.uniform(0.36, 0.50, n)
    q[:, 3:6] = r.uniform(-0.25, 0.25, (n, 3))
    q[:, 6:18] = np.array([0.2, 0.0, -0.7] * 4) + r.uniform(-0.15, 0.15, (n, 12))
    qd = r.uniform(-1.0, 1.0, (n, 18))
    act = r.uniform(-0.5, 0.5, (n, 12))
    return dict(q=_f32(q), qd=_f32(qd), action=_f32(act), params=dict(friction=1.0, keep_all_points=True), mode=2)


def ant_perturbed(n, seed=SEED):
    """Ant (gym/ant_org_xyz_xyzrot.urdf, fixed-base emulation; AntContactSimulation2: dt 0.01, kp 15, kd 0.3, max 3):
    torso height such that 0-4 legs touch, tilted torso, non-zero velocities."""
    r = np.random.default_rng(seed)
    q = np.zeros((n, 14))
    q[:, 0:2] = r.uniform(-0.5, 0.5, (n, 2))
    q[:, 2] = r.uniform(0.30, 0.55, n)
    q[:, 3:6] = r.uniform(-0.2, 0.2, (n, 3))
    q[:, 6:14] = np.array([0.0, -0.5] * 4) + r.uniform(-0.2, 0.2, (n, 8))
    qd = r.uniform(-0.5, 0.5, (n, 14))
    act = r.uniform(-0.5, 0.5, (n, 8))
    return dict(q=_f32(q), qd=_f32(qd), action=_f32(act), params=dict(dt=0.01, friction=1.0, keep_all_points=True), mode=2)


def humanoid(n, seed=SEED, z_range=(0.9, 1.45)):
    """C5: humanoid.urdf floating base; identity-ish base orientation, joints U(+-0.05)."""
    r = np.random.default_rng(seed)
    q = np.zeros((n, 28))
    quat = np.zeros((n, 4)); quat[:, 3] = 1.0
    quat[:, :3] = r.uniform(-0.1, 0.1, (n, 3))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    q[:, :4] = quat
    q[:, 4:6] = r.uniform(-0.5, 0.5, (n, 2))
    q[:, 6] = r.uniform(z_range[0], z_range[1], n)
    q[:, 7:] = r.uniform(-0.05, 0.05, (n, 21))
    qd = r.uniform(-0.5, 0.5, (n, 27))
    tau = r.uniform(-1, 1, (n, 21))
    # re-normalise after fp32 rounding so the reference's unit-quaternion assert holds
    q = _f32(q)
    return dict(q=q, qd=_f32(qd), tau=_f32(tau), params=dict(friction=1.0, keep_all_points=False), mode=2)


def box(n, seed=SEED):
    """A free box (tests/golden/urdf/box.urdf) over the plane: random orientation, height such that 0-4 corner spheres
    penetrate (contact_plane_box)."""
    r = np.random.default_rng(seed)
    q = np.zeros((n, 7))
    quat = r.normal(size=(n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    q[:, :4] = quat
    q[:, 4:6] = r.uniform(-1, 1, (n, 2))
    q[:, 6] = r.uniform(0.08, 0.27, n)
    qd = r.uniform(-1, 1, (n, 6))
    return dict(q=_f32(q), qd=_f32(qd), tau=None, params=dict(friction=0.6, keep_all_points=False), mode=2)


def cartpole_plane(n, seed=SEED):
    """cartpole.urdf (box shapes) with the ground plane: the cart rests in the plane, the pole swings."""
    r = np.random.default_rng(seed)
    q = np.zeros((n, 2))
    q[:, 0] = r.uniform(-0.5, 0.5, n)
    q[:, 1] = r.uniform(-1.0, 1.0, n)
    qd = r.uniform(-1, 1, (n, 2))
    tau = np.zeros((n, 2)); tau[:, 0] = r.uniform(-10, 10, n)
    return dict(q=_f32(q), qd=_f32(qd), tau=_f32(tau), params=dict(friction=0.5, keep_all_points=True), mode=2)


def _unit_quats(r, n, spread=None):
    q = r.normal(size=(n, 4)) if spread is None else np.concatenate([r.uniform(-spread, spread, (n, 3)), np.ones((n, 1))], axis=1)
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def pendulum5spherical(n, seed=SEED):
    """pendulum5spherical.urdf: five spherical joints (quaternion xyzw each: n_q 20, n_qd 15), FD -> integrate_euler."""
    r = np.random.default_rng(seed)
    q = np.concatenate([_unit_quats(r, n) for _ in range(5)], axis=1)
    return dict(q=_f32(q), qd=_f32(r.uniform(-1, 1, (n, 15))), tau=_f32(r.uniform(-1, 1, (n, 15))), params=dict(), mode=1)


def humanoid_spherical(n, seed=SEED):
    """humanoid_xyz_spherical.urdf on the plane: xyz prismatic root + one spherical joint (fixed-base emulation of a free
    torso), the limbs revolute; heights such that 0-14 candidate points penetrate."""
    r = np.random.default_rng(seed)
    q = np.zeros((n, 28))
    q[:, 0:2] = r.uniform(-0.3, 0.3, (n, 2))
    q[:, 2] = r.uniform(0.6, 1.3, n)
    q[:, 3:7] = _unit_quats(r, n, spread=0.2)
    q[:, 7:] = r.uniform(-0.1, 0.1, (n, 21))
    return dict(q=_f32(q), qd=_f32(r.uniform(-0.5, 0.5, (n, 27))), tau=_f32(r.uniform(-1, 1, (n, 27))),
                params=dict(friction=1.0, keep_all_points=False), mode=2)


# ---- worlds of several multibodies (contacts between multibodies: src/world.hpp:206-282) ----------------------------------------
def free_body_model(mass, inertia_diag, geoms, plane=True, arm=None, spherical=False):
    """Flat model (include/tds_b200_model.h) of one fixed-base multibody emulating a free rigid body the way the reference's
    `*_xyz_xyzrot.urdf` files do: three massless prismatic links (x, y, z), then three revolute links (x, y, z) - or ONE
    spherical joint - the last of which carries the mass and the collision shapes.
    geoms: ("sphere", radius, (x, y, z)) or ("capsule", radius, length, (x, y, z), R 3x3 row-major).
    arm: optional (length, mass, radius): a further link on a revolute y joint at the body origin carrying a sphere at its tip."""
    from .model import GEOM, HEADER, BASE, LINK, MAGIC
    links, gl = [], []
    eye = np.eye(3).ravel()

    def link(parent, jtype, qi, qdi, axis, t=(0, 0, 0), m=0.0, com=(0, 0, 0), inertia=(0, 0, 0)):
        r = np.zeros(LINK)
        r[0], r[1], r[2], r[3] = parent, jtype, qi, qdi
        r[4:7], r[7:16], r[16:19] = axis, eye, t
        r[19], r[20:23] = m, com
        r[23:32] = np.diag(inertia).ravel()
        links.append(r)

    for k in range(3):
        link(k - 1, k, k, k, np.eye(3)[k])                       # JOINT_PRISMATIC_X / Y / Z
    if spherical:
        link(2, 8, 3, 3, (0, 0, 0), m=mass, inertia=inertia_diag)   # JOINT_SPHERICAL: q = quaternion xyzw
        nq, nqd = 7, 6
    else:
        for k in range(3):
            last = k == 2
            link(2 + k, 4 + k, 3 + k, 3 + k, np.eye(3)[k], m=mass if last else 0.0, inertia=inertia_diag if last else (0, 0, 0))
        nq, nqd = 6, 6
    body = len(links) - 1
    for g in geoms:
        r = np.zeros(GEOM)
        r[0] = body
        if g[0] == "sphere":
            r[1], r[2], r[5:14], r[14:17] = 0, g[1], eye, g[2]
        elif g[0] == "box":          # ("box", extents (x, y, z), offset[, R])
            r[1], r[2:5], r[14:17] = 4, g[1], g[2]
            r[5:14] = np.asarray(g[3], dtype=np.float64).ravel() if len(g) > 3 else eye
        elif g[0] == "plane":        # ("plane", unit normal): a plane shape on the body link (the reference ignores the link's pose for it)
            r[1], r[2:5], r[5:14] = 1, g[1], eye
        else:
            r[1], r[2], r[3], r[14:17] = 2, g[1], g[2], g[3]
            r[5:14] = np.asarray(g[4], dtype=np.float64).ravel() if len(g) > 4 else eye
        gl.append(r)
    if arm is not None:
        length, am, ar = arm
        link(body, 5, nq, nqd, (0, 1, 0), m=am, com=(length, 0, 0), inertia=(0.4 * am * ar * ar,) * 3)   # JOINT_REVOLUTE_Y
        r = np.zeros(GEOM)
        r[0], r[1], r[2], r[5:14], r[14:17] = len(links) - 1, 0, ar, eye, (length, 0, 0)
        gl.append(r)
        nq += 1; nqd += 1
    head = np.zeros(HEADER)
    head[0], head[1], head[3], head[4], head[5] = MAGIC, len(links), nq, nqd, len(gl)
    if plane:
        head[7], head[8:11] = 1, (0, 0, 1)
    return np.concatenate([head, np.zeros(BASE), np.concatenate(links), np.concatenate(gl) if gl else np.zeros(0)])


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


MULTIBODY_WORLDS = ("spheres2", "capsule_sphere", "sphere_capsule", "three_bodies", "spherical_pair", "racket", "racket_last")


def multibody_world_model(kind):
    """Merged flat model of one of the test worlds (every multibody a free body on the ground plane)."""
    from .model import merge_models
    sph = lambda r, m: free_body_model(m, (0.4 * m * r * r,) * 3, [("sphere", r, (0, 0, 0))])
    cap = lambda: free_body_model(2.0, (0.05, 0.05, 0.01), [("capsule", 0.15, 0.6, (0, 0, 0.05), _rot_y(0.3))])
    if kind == "spheres2":
        return merge_models([sph(0.3, 1.0), sph(0.2, 0.5)])
    if kind == "capsule_sphere":
        return merge_models([cap(), sph(0.25, 1.0)])
    if kind == "sphere_capsule":       # the dispatcher's swapped call (src/contact_point.hpp:478-492)
        return merge_models([sph(0.25, 1.0), cap()])
    if kind == "three_bodies":         # three lists of contacts between multibodies, solved one after the other; an articulated one
        a = free_body_model(1.5, (0.06, 0.06, 0.06), [("sphere", 0.3, (0, 0, 0)), ("sphere", 0.15, (0.35, 0, 0))], arm=(0.5, 0.3, 0.12))
        return merge_models([a, cap(), sph(0.2, 0.5)])
    if kind == "spherical_pair":       # free bodies on xyz + one spherical joint
        a = free_body_model(1.0, (0.04, 0.05, 0.06), [("sphere", 0.3, (0.05, 0, 0))], spherical=True)
        b = free_body_model(2.0, (0.05, 0.05, 0.01), [("capsule", 0.15, 0.6, (0, 0, 0), _rot_y(0.2))], spherical=True)
        return merge_models([a, b])
    if kind in ("racket", "racket_last"):
        # a <plane> collision shape on a link of a multibody (the reference's data has one: franka_panda/panda_racket.urdf) against
        # a ball, a capsule and a box of other multibodies; "racket_last": the plane's multibody comes last (dispatcher's swapped call)
        racket = free_body_model(1.2, (0.05, 0.04, 0.03), [("plane", (0.0, 0.0, 1.0)), ("sphere", 0.1, (0.2, 0, 0))])
        ball = sph(0.2, 0.4)
        boxy = free_body_model(1.5, (0.05, 0.06, 0.07), [("box", (0.4, 0.3, 0.2), (0.02, 0, 0))])
        return merge_models([racket, ball, cap(), boxy] if kind == "racket" else [ball, cap(), boxy, racket])
    raise ValueError(kind)


def multibody_world(kind, n, seed=SEED):
    """States of the test worlds: the multibodies are dropped close enough to each other and to the ground that every kind of
    contact list occurs (none, plane only, multibody pair only, both)."""
    r = np.random.default_rng(seed)
    model = multibody_world_model(kind)
    n_q, n_qd = int(model[3]), int(model[4])
    nb = int(model[12])
    spherical = kind == "spherical_pair"
    q = np.zeros((n, n_q)); qd = r.uniform(-1.0, 1.0, (n, n_qd))
    centre = r.uniform(-0.5, 0.5, (n, 2))
    off = 0
    arm_extra = {"three_bodies": (1, 0, 0)}.get(kind, (0,) * nb)
    for b in range(nb):
        q[:, off:off + 2] = centre + r.uniform(-0.3, 0.3, (n, 2))
        q[:, off + 2] = r.uniform(0.1, 0.6, n)
        if spherical:
            q[:, off + 3:off + 7] = _unit_quats(r, n)
            off += 7
        else:
            q[:, off + 3:off + 6] = r.uniform(-1.0, 1.0, (n, 3))
            off += 6
        if arm_extra[b]:
            q[:, off] = r.uniform(-1.0, 1.0, n)
            off += 1
    tau = r.uniform(-1.0, 1.0, (n, n_qd))
    return dict(model=model, q=_f32(q), qd=_f32(qd), tau=_f32(tau), params=dict(friction=0.7, keep_all_points=False), mode=2)


# ---- worlds of rigid bodies (the RigidBody path of World::step) -------------------------------------------------------------------
RIGID_WORLDS = ("billiard", "stack", "swapped")


def rigid_world(kind, n, seed=SEED):
    """bodies (tds_b200.rigid records), states [n][n_bodies][13], forces [n][n_bodies][3] and World parameters of the test worlds.
    billiard: the seven balls of python/examples/billiard_optimization.py (no gravity, 50 solver iterations); stack: plane, two
    spheres, a capsule and a box under gravity; swapped: bodies listed so that the dispatcher's swapped calls are taken, tilted
    plane with a non-zero constant."""
    from . import rigid as rg
    r = np.random.default_rng(seed)
    if kind == "billiard":
        bodies = [rg.sphere(1.0, 0.5)] * 7
        params = dict(gravity=(0.0, 0.0, 0.0), num_solver_iterations=50)
    elif kind == "stack":
        bodies = [rg.plane(), rg.sphere(1.0, 0.3), rg.sphere(2.0, 0.2), rg.capsule(1.5, 0.15, 0.6), rg.box(2.0, (0.4, 0.3, 0.2))]
        params = dict(num_solver_iterations=50, friction=0.6)
    elif kind == "swapped":
        bodies = [rg.sphere(1.0, 0.3), rg.capsule(1.5, 0.15, 0.6), rg.plane((0.1, 0.0, 1.0), 0.05), rg.box(2.0, (0.4, 0.3, 0.2))]
        params = dict(num_solver_iterations=10, restitution=0.3, erp=0.2)
    else:
        raise ValueError(kind)
    desc = np.asarray(bodies, dtype=np.float64)
    nb = desc.shape[0]
    s = np.zeros((n, nb, 13))
    s[:, :, 0:2] = r.uniform(-0.6, 0.6, (n, nb, 2)) * (2.5 if kind == "billiard" else 1.0)
    s[:, :, 2] = 0.0 if kind == "billiard" else r.uniform(0.05, 0.5, (n, nb))
    q = r.normal(size=(n, nb, 4))
    s[:, :, 3:7] = q / np.linalg.norm(q, axis=2, keepdims=True)
    s[:, :, 7:10] = r.uniform(-1, 1, (n, nb, 3))
    s[:, :, 10:13] = r.uniform(-1, 1, (n, nb, 3))
    for b in range(nb):
        if desc[b, 1] == rg.PLANE:
            s[:, b, :] = 0.0
            s[:, b, 6] = 1.0
    f = r.uniform(-50, 50, (n, nb, 3))
    return dict(bodies=desc, state=s, force=f, params=params)
