"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

ctypes binding of oracle/libtds_oracle.so (oracle/tds_oracle.c): the plain-C fp64
restatement of the reference hot path.  Used only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtds_oracle.so")
_lib = None

MODE_FD, MODE_NOCONTACT, MODE_FULL = 0, 1, 2
MAX_CONTACTS = 48


class Params(ctypes.Structure):
    _fields_ = [("dt", ctypes.c_double), ("gravity", ctypes.c_double * 3), ("friction", ctypes.c_double),
                ("restitution", ctypes.c_double), ("erp", ctypes.c_double), ("cfm", ctypes.c_double),
                ("pgs_iterations", ctypes.c_int), ("keep_all_points", ctypes.c_int),
                ("contact_model", ctypes.c_int), ("spring_k", ctypes.c_double), ("damper_d", ctypes.c_double),
                ("exponent_n", ctypes.c_double), ("v_transition", ctypes.c_double), ("hard_contact_condition", ctypes.c_int)]


def make_params(dt=1e-3, gravity=(0.0, 0.0, -9.81), friction=0.5, restitution=0.0, erp=0.2, cfm=1e-5,
                pgs_iterations=1, keep_all_points=False, contact_model=0, spring_k=50000.0, damper_d=5000.0, exponent_n=1.5,
                v_transition=0.01, hard_contact_condition=True):
    p = Params()
    p.dt = dt
    p.gravity[:] = gravity
    p.friction, p.restitution, p.erp, p.cfm = friction, restitution, erp, cfm
    p.pgs_iterations, p.keep_all_points = pgs_iterations, int(keep_all_points)
    p.contact_model, p.spring_k, p.damper_d, p.exponent_n = int(contact_model), spring_k, damper_d, exponent_n
    p.v_transition, p.hard_contact_condition = v_transition, int(hard_contact_condition)
    return p


def build():
    subprocess.check_call(["bash", os.path.join(_HERE, "build_oracle.sh")], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        pp = ctypes.POINTER(Params)
        L.tdso_step.restype = ctypes.c_int
        L.tdso_step.argtypes = [dp, pp, ctypes.c_int, dp, dp, dp, dp, dp, dp, ip, ip, dp, ctypes.c_int, dp]
        L.tdso_mass_matrix.restype = ctypes.c_int
        L.tdso_mass_matrix.argtypes = [dp, dp, dp]
        L.tdso_locomotion_step.restype = ctypes.c_int
        L.tdso_locomotion_step.argtypes = [dp, pp, dp, ctypes.c_int, ctypes.c_int, dp, dp, ctypes.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def step(model, params, mode, q, qd, tau=None):
    """One step from (q, qd, tau); returns dict(q, qd, qdd, n_contacts, contact_idx, contact_data, link_xf)."""
    m = np.ascontiguousarray(model, dtype=np.float64)
    n_links, n_q, n_qd = int(m[1]), int(m[3]), int(m[4])
    q = np.ascontiguousarray(q, dtype=np.float64)
    qd = np.ascontiguousarray(qd, dtype=np.float64)
    tau = np.zeros(max(n_qd, 1)) if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
    out = dict(q=np.zeros(n_q), qd=np.zeros(n_qd), qdd=np.zeros(n_qd))
    nc = ctypes.c_int(0)
    cidx = np.zeros((MAX_CONTACTS, 2), dtype=np.int32)
    cdat = np.zeros((MAX_CONTACTS, 10))
    xf = np.zeros((max(n_links, 1), 12))
    rc = lib().tdso_step(_dp(m), ctypes.byref(params), mode, _dp(q), _dp(qd), _dp(tau), _dp(out["q"]),
                         _dp(out["qd"]), _dp(out["qdd"]), ctypes.byref(nc),
                         cidx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _dp(cdat), MAX_CONTACTS, _dp(xf))
    if rc:
        raise RuntimeError(f"tdso_step failed rc={rc}")
    n = nc.value
    out.update(n_contacts=n, contact_idx=cidx[:n].copy(), contact_data=cdat[:n].copy(), link_xf=xf[:n_links])
    return out


def mass_matrix(model, q):
    m = np.ascontiguousarray(model, dtype=np.float64)
    n = int(m[4])
    q = np.ascontiguousarray(q, dtype=np.float64)
    M = np.zeros((n, n))
    rc = lib().tdso_mass_matrix(_dp(m), _dp(q), _dp(M))
    if rc:
        raise RuntimeError(f"tdso_mass_matrix failed rc={rc}")
    return M


def locomotion_step(model, params, initial_poses, base_dof, inputs, output_dim):
    """LocomotionContactSimulation::step_forward_original restated; inputs [n][in_dim]."""
    m = np.ascontiguousarray(model, dtype=np.float64)
    ip = np.ascontiguousarray(initial_poses, dtype=np.float64)
    x = np.ascontiguousarray(inputs, dtype=np.float64)
    out = np.zeros((x.shape[0], output_dim))
    for i in range(x.shape[0]):
        rc = lib().tdso_locomotion_step(_dp(m), ctypes.byref(params), _dp(ip), ip.size, base_dof,
                                        _dp(x[i]), _dp(out[i]), output_dim)
        if rc:
            raise RuntimeError(f"tdso_locomotion_step failed rc={rc}")
    return out
