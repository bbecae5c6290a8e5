"""TEST INFRASTRUCTURE: ctypes binding of tests/cpp/_stepw_host.so - the product's generic step kernel (csrc/tds_stepw.cu)
compiled for the host (see tests/cpp/stepw_host.cpp) and of tests/cpp/_steps_host.so - the same for the model-specialised kernel
(csrc/tds_steps.cu, tests/cpp/steps_host.cpp).  Used only by the CPU test-suite to execute the kernel SOURCE without a
GPU; the package never loads it."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "cpp", "_stepw_host.so")
SRC = os.path.join(HERE, "cpp", "stepw_host.cpp")
CSRC = os.path.join(ROOT, "tiny-differentiable-simulator_b200", "csrc")
_lib = None


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("tds_stepw.cu", "tds_wcommon.cuh", "tds_math.cuh", "tds_dual.cuh", "tds_model.h", "tds_types.h")]
    if os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-w", "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
                           "-I/usr/local/cuda/include", SRC, "-o", SO])


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(SO)
        dp = ctypes.POINTER(ctypes.c_double)
        L.tdsemu_stepw.restype = ctypes.c_int
        L.tdsemu_stepw.argtypes = [dp, ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [dp] * 8
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def step(model, mode, q, qd, tau=None, precision=1, use_pd=False, env=None, jacobian=False, dt=1e-3, gravity=(0.0, 0.0, -9.81),
         friction=0.5, restitution=0.0, erp=0.2, cfm=1e-5, pgs_iterations=1, keep_all_points=False, contact_model=0,
         spring_k=50000.0, damper_d=5000.0, exponent_n=1.5, v_transition=0.01, hard_contact_condition=True):
    """One step of every row of q / qd through the host-compiled kernel.  env = (n_act, start_link, kp, kd, max_force,
    action_limit, poses...) for use_pd.  Returns dict(q, qd, qdd, contact_dist[, jac])."""
    m = np.ascontiguousarray(model, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64); qd = np.ascontiguousarray(qd, dtype=np.float64)
    n, n_q, n_qd = q.shape[0], int(m[3]), int(m[4])
    t = None if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
    params = np.array([dt, *gravity, friction, restitution, erp, cfm, pgs_iterations, int(keep_all_points), contact_model, spring_k,
                       damper_d, exponent_n, v_transition, int(hard_contact_condition)], dtype=np.float64)
    e = None if env is None else np.ascontiguousarray(env, dtype=np.float64)
    out = dict(q=np.zeros((n, n_q)), qd=np.zeros((n, n_qd)), qdd=np.zeros((n, n_qd)), contact_dist=np.zeros((n, 64)))
    n_tau = n_qd - (6 if int(m[2]) else 0)
    rows = n_qd if mode == 0 else n_q + n_qd
    cols = n_q + n_qd + ((int(e[0]) + 3) if use_pd else n_tau)
    jac = np.zeros((n, rows, cols)) if jacobian else None
    # contact_dist is written [n][n_points]: size it after the call from the return value
    cdbuf = np.zeros(n * 128)
    rc = lib().tdsemu_stepw(_dp(m), m.size, _dp(params), _dp(e), precision, mode, int(use_pd), n, _dp(q), _dp(qd), _dp(t),
                            _dp(out["q"]), _dp(out["qd"]), _dp(out["qdd"]), _dp(cdbuf), _dp(jac))
    if rc < 0:
        raise RuntimeError(f"tdsemu_stepw rc={rc}")
    if jacobian:
        out["jac"] = jac
        del out["contact_dist"]
    else:
        out["contact_dist"] = cdbuf[:n * rc].reshape(n, rc) if rc else np.zeros((n, 0))
    return out


# ---- the model-specialised kernel (csrc/tds_steps.cu) ------------------------------------------------------------------------------
SO_S = os.path.join(HERE, "cpp", "_steps_host.so")
SRC_S = os.path.join(HERE, "cpp", "steps_host.cpp")
_lib_s = None
SPEC = {"laikago": 0, "ant": 1}


def build_spec():
    gen = os.path.join(CSRC, "generated")
    deps = [SRC_S] + [os.path.join(CSRC, f) for f in ("tds_steps.cu", "tds_math.cuh", "tds_types.h")] + \
        [os.path.join(gen, f) for f in os.listdir(gen)]
    if os.path.exists(SO_S) and all(os.path.getmtime(d) <= os.path.getmtime(SO_S) for d in deps):
        return
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-shared", "-fPIC", "-w", "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
                           "-I/usr/local/cuda/include", SRC_S, "-o", SO_S + ".tmp"])
    os.replace(SO_S + ".tmp", SO_S)


def lib_spec():
    global _lib_s
    if _lib_s is None:
        build_spec()
        L = ctypes.CDLL(SO_S)
        dp = ctypes.POINTER(ctypes.c_double)
        L.tdsemu_steps.restype = ctypes.c_int
        L.tdsemu_steps.argtypes = [ctypes.c_int, dp, dp] + [ctypes.c_int] * 6 + [dp] * 12
        _lib_s = L
    return _lib_s


def step_spec(name, mode, q, qd, tau=None, precision=0, var=0, use_pd=False, env=None, other_lane_in_contact=False, whole_tile=False,
              auto_reset=False, reset_q=None, dt=1e-3, gravity=(0.0, 0.0, -9.81), friction=0.5, restitution=0.0, erp=0.2, cfm=1e-5,
              pgs_iterations=1, keep_all_points=False):
    """One step of every row of q / qd through the host-compiled specialised kernel of model `name` (state is carried in fp32
    as on the device).  env = (n_act, start_link, kp, kd, max_force, action_limit, reward_kind, poses...) for use_pd; tau then
    holds the actions.  var 0: general instance (qdd, contact distances, link transforms), 1: lean instance, 2: lean instance
    with the host layouts (whole_tile only).  whole_tile: 128 host threads per tile (exact, slow) instead of 4 per lane.
    Returns dict(q, qd, qdd, reward, done[, contact_dist, link_xf][, obs, obs_reward, obs_done])."""
    q = np.ascontiguousarray(q, dtype=np.float64); qd = np.ascontiguousarray(qd, dtype=np.float64)
    n, n_q, n_qd = q.shape[0], q.shape[1], qd.shape[1]
    t = None if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
    params = np.array([dt, *gravity, friction, restitution, erp, cfm, pgs_iterations, int(keep_all_points)], dtype=np.float64)
    e = None
    if env is not None:
        env = np.asarray(env, dtype=np.float64)
        rq = np.zeros(n_q) if reset_q is None else np.asarray(reset_q, dtype=np.float64)
        e = np.ascontiguousarray(np.concatenate([env[:7], [float(auto_reset)], env[7:], rq]))
    out = dict(q=np.zeros((n, n_q)), qd=np.zeros((n, n_qd)), qdd=np.zeros((n, n_qd)), reward=np.zeros(n), done=np.zeros(n))
    cd = np.zeros((n, 64)) if var == 0 else None
    xf = np.zeros((n, 64 * 12)) if var == 0 else None
    obs = np.zeros((n, n_q + n_qd)) if var == 2 else None
    tail = np.zeros(2 * n) if var == 2 else None
    rc = lib_spec().tdsemu_steps(SPEC[name], _dp(params), _dp(e), precision, var, mode, int(use_pd),
                                 int(other_lane_in_contact) | (int(whole_tile or var == 2) << 1), n,
                                 _dp(q), _dp(qd), _dp(t), _dp(out["q"]), _dp(out["qd"]), _dp(out["qdd"]), _dp(out["reward"]), _dp(out["done"]),
                                 _dp(cd), _dp(xf), _dp(obs), _dp(tail))
    if rc < 0:
        raise RuntimeError(f"tdsemu_steps rc={rc}")
    if var == 0:
        out["contact_dist"] = cd.ravel()[:n * rc].reshape(n, rc)
        out["_xf_flat"] = xf.ravel()
    if var == 2:
        out.update(obs=obs, obs_reward=tail[:n], obs_done=tail[n:])
    return out


def link_xf_of(out, n, n_links):
    """[n][n_links][12] world transforms (R row-major 9, p 3) from step_spec(var=0)."""
    return out["_xf_flat"][:n * n_links * 12].reshape(n, n_links, 12)


# ---- the rigid-body world kernel (csrc/tds_rigid.cu) ---------------------------------------------------------------------------------
SO_R = os.path.join(HERE, "cpp", "_rigid_host.so")
SRC_R = os.path.join(HERE, "cpp", "rigid_host.cpp")
_lib_r = None


def lib_rigid():
    global _lib_r
    if _lib_r is None:
        deps = [SRC_R] + [os.path.join(CSRC, f) for f in ("tds_rigid.cu", "tds_math.cuh", "tds_dual.cuh")]
        if not (os.path.exists(SO_R) and all(os.path.getmtime(d) <= os.path.getmtime(SO_R) for d in deps)):
            subprocess.check_call(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-w", "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
                                   "-I/usr/local/cuda/include", SRC_R, "-o", SO_R + ".tmp"])
            os.replace(SO_R + ".tmp", SO_R)
        L = ctypes.CDLL(SO_R)
        dp = ctypes.POINTER(ctypes.c_double)
        L.tdsemu_rigid.restype = ctypes.c_int
        L.tdsemu_rigid.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int, dp, dp, ctypes.c_int, dp, dp]
        _lib_r = L
    return _lib_r


def rigid_step(desc, state, force=None, steps=1, jacobian=False, dt=1.0 / 60.0, gravity=(0.0, 0.0, -9.81), friction=0.5, restitution=0.0,
               erp=0.1, num_solver_iterations=1):
    """`steps` World::step calls of every world of state [n][n_bodies][13] through the host-compiled rigid-body kernel."""
    d = np.ascontiguousarray(desc, dtype=np.float64)
    s = np.ascontiguousarray(state, dtype=np.float64)
    n, nb = s.shape[0], d.shape[0]
    f = None if force is None else np.ascontiguousarray(force, dtype=np.float64)
    params = np.array([dt, *gravity, friction, restitution, erp, num_solver_iterations], dtype=np.float64)
    out = np.zeros_like(s)
    jac = np.zeros((n, 13 * nb, 16 * nb)) if jacobian else None
    rc = lib_rigid().tdsemu_rigid(_dp(d), nb, _dp(params), n, _dp(s), _dp(f), steps, _dp(out), _dp(jac))
    if rc:
        raise RuntimeError(f"tdsemu_rigid rc={rc}")
    return (out, jac) if jacobian else out
