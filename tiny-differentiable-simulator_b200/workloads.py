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
