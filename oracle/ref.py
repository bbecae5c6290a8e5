"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

ctypes binding of oracle/_ref/libtds_ref.so: the UNMODIFIED reference hot path compiled in
place from /root/reference by oracle/build_ref.sh (shims: oracle/ref/ref_core.cpp,
oracle/ref/ref_laikago.cpp).  Used only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs - as the checker and the CPU baseline, never as a
product path.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libtds_ref.so")

_lib = None


class _quiet_stdout:
    """The reference prints "Loading URDF ..." on stdout from C++; keep it out of JSON-producing callers."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *a):
        os.dup2(self._saved, 1)
        os.close(self._saved)
        os.close(self._null)


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{LIB_PATH} missing: run oracle/build_ref.sh (needs /root/reference)")
        L = ctypes.CDLL(LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        vp = ctypes.c_void_p
        L.tdsref_create_from_urdf.restype = vp
        L.tdsref_create_from_urdf.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        L.tdsref_create_from_model.restype = vp
        L.tdsref_create_from_model.argtypes = [dp, ctypes.c_int, ctypes.c_int]
        L.tdsref_destroy.argtypes = [vp]
        L.tdsref_export_model.restype = ctypes.c_int
        L.tdsref_export_model.argtypes = [vp, dp, ctypes.c_int]
        for f in ("tdsref_dof_q", "tdsref_dof_qd", "tdsref_dof_tau", "tdsref_num_links"):
            getattr(L, f).restype = ctypes.c_int
            getattr(L, f).argtypes = [vp]
        L.tdsref_set_params.argtypes = [vp, ctypes.c_double, dp, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        L.tdsref_step.argtypes = [vp, ctypes.c_int, dp, dp, dp, dp, dp, dp, dp, ip, ip, dp, ctypes.c_int]
        if hasattr(L, "tdsref_step_batch"):
            L.tdsref_step_batch.argtypes = [vp, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, dp]
        L.tdsref_link_transforms.argtypes = [vp, dp]
        L.tdsref_mass_matrix.argtypes = [vp, dp, dp]
        L.tdsref_point_jacobian.argtypes = [vp, dp, ctypes.c_int, dp, dp]
        if hasattr(L, "tdsrefw_create"):
            L.tdsrefw_create.restype = vp
            L.tdsrefw_create.argtypes = [dp, ctypes.c_int]
            L.tdsrefw_destroy.argtypes = [vp]
            L.tdsrefw_num_bodies.restype = ctypes.c_int
            L.tdsrefw_num_bodies.argtypes = [vp]
            L.tdsrefw_set_params.argtypes = [vp, ctypes.c_double, dp, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_double, ctypes.c_double]
            L.tdsrefw_step.argtypes = [vp, ctypes.c_int, dp, dp, dp, dp, dp, ip, ip, dp, ctypes.c_int]
        if hasattr(L, "tdsrefr_create"):
            L.tdsrefr_create.restype = vp
            L.tdsrefr_create.argtypes = [dp, ctypes.c_int]
            L.tdsrefr_destroy.argtypes = [vp]
            L.tdsrefr_set_params.argtypes = [vp, ctypes.c_double, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int]
            L.tdsrefr_step.argtypes = [vp, dp, dp, ctypes.c_int, dp, ip]
        L.tdsref_laikago_create.restype = vp
        L.tdsref_laikago_create.argtypes = [ctypes.c_int]
        L.tdsref_laikago_destroy.argtypes = [vp]
        for f in ("tdsref_laikago_input_dim", "tdsref_laikago_output_dim", "tdsref_laikago_num_threads"):
            getattr(L, f).restype = ctypes.c_int
            getattr(L, f).argtypes = [vp]
        L.tdsref_laikago_step.argtypes = [vp, ctypes.c_int, ctypes.c_int, dp, dp]
        L.tdsref_laikago_reward_done.argtypes = [vp, dp, dp, ip]
        if hasattr(L, "tdsref_ant_create"):
            L.tdsref_ant_create.restype = vp
            L.tdsref_ant_create.argtypes = [ctypes.c_int]
            L.tdsref_ant_destroy.argtypes = [vp]
            for f in ("tdsref_ant_input_dim", "tdsref_ant_output_dim", "tdsref_ant_state_dim", "tdsref_ant_action_dim",
                      "tdsref_ant_num_threads"):
                getattr(L, f).restype = ctypes.c_int
                getattr(L, f).argtypes = [vp]
            L.tdsref_ant_step.argtypes = [vp, ctypes.c_int, ctypes.c_int, dp, dp]
            L.tdsref_ant_reward_done.argtypes = [vp, dp, dp, dp, ip]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int)) if a is not None else None


class RefSim:
    """One reference World (optional ground plane + one MultiBody)."""

    MODE_FD, MODE_NOCONTACT, MODE_FULL = 0, 1, 2

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference sim creation failed")
        self._h = handle
        L = lib()
        self.n_q = L.tdsref_dof_q(handle)
        self.n_qd = L.tdsref_dof_qd(handle)
        self.n_tau = L.tdsref_dof_tau(handle)
        self.n_links = L.tdsref_num_links(handle)

    @classmethod
    def from_urdf(cls, urdf, plane_urdf=None, floating=False, prec=64):
        with _quiet_stdout():
            h = lib().tdsref_create_from_urdf((plane_urdf or "").encode(), urdf.encode(), int(floating), prec)
        return cls(h)

    @classmethod
    def from_model(cls, model, prec=64):
        m = np.ascontiguousarray(model, dtype=np.float64)
        return cls(lib().tdsref_create_from_model(_dp(m), m.size, prec))

    def close(self):
        if self._h:
            lib().tdsref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export_model(self):
        n = lib().tdsref_export_model(self._h, None, 0)
        out = np.zeros(n, dtype=np.float64)
        lib().tdsref_export_model(self._h, _dp(out), n)
        return out

    def set_params(self, dt=1e-3, gravity=(0.0, 0.0, -9.81), friction=0.5, restitution=0.0,
                   keep_all_points=False, pgs_iterations=1, erp=0.2, cfm=1e-5):
        g = np.asarray(gravity, dtype=np.float64)
        lib().tdsref_set_params(self._h, dt, _dp(g), friction, restitution, int(keep_all_points),
                                pgs_iterations, erp, cfm)

    def step(self, mode, q, qd, tau=None, contact_cap=64):
        q = np.ascontiguousarray(q, dtype=np.float64)
        qd = np.ascontiguousarray(qd, dtype=np.float64)
        tau = np.zeros(self.n_tau) if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
        assert q.size == self.n_q and qd.size == self.n_qd and tau.size == self.n_tau
        out = dict(q=np.zeros(self.n_q), qd=np.zeros(self.n_qd), qdd=np.zeros(self.n_qd),
                   qd_pre=np.zeros(self.n_qd))
        nc = ctypes.c_int(0)
        cidx = np.zeros((contact_cap, 2), dtype=np.int32)
        cdat = np.zeros((contact_cap, 10), dtype=np.float64)
        lib().tdsref_step(self._h, mode, _dp(q), _dp(qd), _dp(tau), _dp(out["q"]), _dp(out["qd"]),
                          _dp(out["qdd"]), _dp(out["qd_pre"]), ctypes.byref(nc), _ip(cidx), _dp(cdat),
                          contact_cap)
        n = nc.value
        assert n <= contact_cap
        out["n_contacts"] = n
        out["contact_idx"] = cidx[:n].copy()
        out["contact_data"] = cdat[:n].copy()
        return out

    def step_batch(self, mode, q, qd, tau=None):
        """n independent steps in one foreign call (bench.py's reference arm); returns (q', qd', qdd)."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        qd = np.ascontiguousarray(qd, dtype=np.float64)
        n = q.shape[0]
        t = None if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
        qo, qdo, qddo = np.zeros_like(q), np.zeros_like(qd), np.zeros_like(qd)
        lib().tdsref_step_batch(self._h, mode, n, _dp(q), _dp(qd), _dp(t), _dp(qo), _dp(qdo), _dp(qddo))
        return qo, qdo, qddo

    def link_transforms(self):
        out = np.zeros((self.n_links, 12))
        lib().tdsref_link_transforms(self._h, _dp(out))
        return out

    def mass_matrix(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        M = np.zeros((self.n_qd, self.n_qd))
        lib().tdsref_mass_matrix(self._h, _dp(q), _dp(M))
        return M

    def point_jacobian(self, q, link, point):
        q = np.ascontiguousarray(q, dtype=np.float64)
        p = np.ascontiguousarray(point, dtype=np.float64)
        J = np.zeros((3, self.n_qd))
        lib().tdsref_point_jacobian(self._h, _dp(q), link, _dp(p), _dp(J))
        return J


class RefWorld:
    """A reference World of several fixed-base multibodies (+ the plane) built from a merged flat model
    (tds_b200.model.merge_models; oracle/ref/ref_world.cpp)."""

    def __init__(self, model):
        m = np.ascontiguousarray(model, dtype=np.float64)
        self._L = lib()
        self._h = self._L.tdsrefw_create(_dp(m), m.size)
        if not self._h:
            raise RuntimeError("tdsrefw_create failed")
        self.n_q, self.n_qd = int(m[3]), int(m[4])
        self.set_params()

    def close(self):
        if self._h:
            self._L.tdsrefw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, dt=1e-3, gravity=(0.0, 0.0, -9.81), friction=0.5, restitution=0.0, keep_all_points=False,
                   pgs_iterations=1, erp=0.2, cfm=1e-5):
        g = np.asarray(gravity, dtype=np.float64)
        self._L.tdsrefw_set_params(self._h, dt, _dp(g), friction, restitution, int(keep_all_points), pgs_iterations, erp, cfm)

    def step(self, mode, q, qd, tau=None, contact_cap=128):
        """mode 2: full step of every multibody; 3: World::step alone.  Returns dict(q, qd, n_contacts, contact_idx
        [(list, link_a, link_b)], contact_data [normal on b, point on a, point on b, distance])."""
        q = np.ascontiguousarray(q, dtype=np.float64); qd = np.ascontiguousarray(qd, dtype=np.float64)
        t = None if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
        qo, qdo = np.zeros(self.n_q), np.zeros(self.n_qd)
        nc = ctypes.c_int(0)
        idx = np.zeros((contact_cap, 3), dtype=np.int32)
        dat = np.zeros((contact_cap, 10))
        with _quiet_stdout():
            self._L.tdsrefw_step(self._h, mode, _dp(q), _dp(qd), _dp(t), _dp(qo), _dp(qdo), ctypes.byref(nc), _ip(idx), _dp(dat), contact_cap)
        n = min(nc.value, contact_cap)
        return dict(q=qo, qd=qdo, n_contacts=nc.value, contact_idx=idx[:n].copy(), contact_data=dat[:n].copy())


class RefRigidWorld:
    """A reference World of RigidBodys (oracle/ref/ref_rigid.cpp).  desc [n_bodies][6]: mass, shape, p0..p3."""

    def __init__(self, desc):
        d = np.ascontiguousarray(desc, dtype=np.float64)
        self.n_bodies = d.shape[0]
        self._L = lib()
        self._h = self._L.tdsrefr_create(_dp(d), self.n_bodies)
        if not self._h:
            raise RuntimeError("tdsrefr_create failed")

    def close(self):
        if self._h:
            self._L.tdsrefr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, dt=1.0 / 60.0, gravity=(0.0, 0.0, -9.81), friction=0.5, restitution=0.0, erp=0.1, num_solver_iterations=1):
        g = np.asarray(gravity, dtype=np.float64)
        self._L.tdsrefr_set_params(self._h, dt, _dp(g), friction, restitution, erp, num_solver_iterations)

    def step(self, state, force=None, steps=1):
        s = np.ascontiguousarray(state, dtype=np.float64)
        f = None if force is None else np.ascontiguousarray(force, dtype=np.float64)
        out = np.zeros_like(s)
        nc = ctypes.c_int(0)
        self._L.tdsrefr_step(self._h, _dp(s), _dp(f), steps, _dp(out), ctypes.byref(nc))
        return out, nc.value


class LaikagoRef:
    """The reference's Laikago env step (51 doubles in, 411 out), both of its CPU paths."""

    IMPL_TEMPLATED, IMPL_CODEGEN = 0, 1

    def __init__(self, num_threads=1):
        L = lib()
        with _quiet_stdout():
            self._h = L.tdsref_laikago_create(num_threads)
        self.input_dim = L.tdsref_laikago_input_dim(self._h)
        self.output_dim = L.tdsref_laikago_output_dim(self._h)
        self.num_threads = L.tdsref_laikago_num_threads(self._h)

    def close(self):
        if self._h:
            lib().tdsref_laikago_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, inputs, impl=0, out=None):
        x = np.ascontiguousarray(inputs, dtype=np.float64)
        n = x.shape[0]
        assert x.shape[1] == self.input_dim
        if out is None:
            out = np.zeros((n, self.output_dim))
        lib().tdsref_laikago_step(self._h, impl, n, _dp(x), _dp(out))
        return out

    def reward_done(self, state):
        s = np.ascontiguousarray(state, dtype=np.float64)
        r = ctypes.c_double(0)
        d = ctypes.c_int(0)
        lib().tdsref_laikago_reward_done(self._h, _dp(s), ctypes.byref(r), ctypes.byref(d))
        return r.value, bool(d.value)


class AntRef:
    """The reference's Ant env step (AntContactSimulation2, examples/environments/ant_environment2.h), both CPU paths."""

    IMPL_TEMPLATED, IMPL_CODEGEN = 0, 1

    def __init__(self, num_threads=1):
        L = lib()
        with _quiet_stdout():
            self._h = L.tdsref_ant_create(num_threads)
        self.input_dim = L.tdsref_ant_input_dim(self._h)
        self.output_dim = L.tdsref_ant_output_dim(self._h)
        self.state_dim = L.tdsref_ant_state_dim(self._h)
        self.action_dim = L.tdsref_ant_action_dim(self._h)
        self.num_threads = L.tdsref_ant_num_threads(self._h)

    def close(self):
        if self._h:
            lib().tdsref_ant_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, inputs, impl=0, out=None):
        x = np.ascontiguousarray(inputs, dtype=np.float64)
        n = x.shape[0]
        assert x.shape[1] == self.input_dim
        if out is None:
            out = np.zeros((n, self.output_dim))
        lib().tdsref_ant_step(self._h, impl, n, _dp(x), _dp(out))
        return out

    def reward_done(self, prev_state, cur_state):
        a = np.ascontiguousarray(prev_state, dtype=np.float64)
        b = np.ascontiguousarray(cur_state, dtype=np.float64)
        r = ctypes.c_double(0)
        d = ctypes.c_int(0)
        lib().tdsref_ant_reward_done(self._h, _dp(a), _dp(b), ctypes.byref(r), ctypes.byref(d))
        return r.value, bool(d.value)
