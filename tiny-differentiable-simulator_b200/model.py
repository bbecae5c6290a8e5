"""Flat model handling (layout: include/tds_b200_model.h): URDF compile, load / save fixtures."""
import ctypes
import json
import os

import numpy as np

from . import _lib

MAGIC = 20250200
HEADER, BASE, LINK, GEOM, VIS = 16, 13, 34, 18, 13


def compile_urdf(urdf, plane_urdf=None, floating=False):
    """URDF file path or XML text -> flat model (np.float64 array).

    Mirrors UrdfCache::construct (src/urdf/urdf_cache.hpp:74-84) of the reference; `plane_urdf`
    adds the static ground plane body the locomotion environments create first.
    """
    L = _lib.lib()
    u = urdf.encode()
    p = (plane_urdf or "").encode()
    n = L.tds_b200_urdf_to_model(u, p, int(floating), None, 0)
    if n <= 0:
        raise ValueError("URDF compile failed: " + _lib.last_error())
    out = np.zeros(n, dtype=np.float64)
    n2 = L.tds_b200_urdf_to_model(u, p, int(floating), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n)
    if n2 != n:
        raise ValueError("URDF compile failed: " + _lib.last_error())
    return out


def model_dims(model):
    m = np.asarray(model)
    if int(m[0]) != MAGIC:
        raise ValueError("not a tds_b200 flat model")
    return dict(n_links=int(m[1]), floating=int(m[2]), n_q=int(m[3]), n_qd=int(m[4]), n_geoms=int(m[5]),
                n_vis=int(m[6]), has_plane=int(m[7]))


def merge_models(models):
    """A world of several multibodies: the flat models of K fixed-base multibodies (each ONE root link, e.g. the reference's
    `*_xyz_xyzrot.urdf` free-body emulations) -> one flat model with header field TDSM_H_NBODIES = K
    (include/tds_b200_model.h).  Mirrors a reference World holding the plane (if any model has one: multibody 0) and the K
    multibodies in the order given: geoms of different multibodies collide (src/world.hpp:206-282: sphere-sphere,
    capsule-sphere), geoms of one multibody never do; coordinates q / qd are the concatenation in the same order."""
    ms = [np.asarray(m, dtype=np.float64) for m in models]
    if len(ms) < 2:
        raise ValueError("merge_models needs at least two multibodies")
    dims = [model_dims(m) for m in ms]
    if any(d["floating"] for d in dims):
        raise ValueError("merge_models: fixed-base multibodies only (emulate a free body by prismatic + revolute / spherical joints)")
    head = np.zeros(HEADER)
    head[0] = MAGIC
    links, base_geoms, link_geoms, vis = [], [], [], []
    lo = qo = qdo = 0
    for m, d in zip(ms, dims):
        L0 = HEADER + BASE
        G0 = L0 + d["n_links"] * LINK
        V0 = G0 + d["n_geoms"] * GEOM
        l = m[L0:G0].reshape(d["n_links"], LINK).copy()
        if int(np.sum(l[:, 0] < 0)) != 1:
            raise ValueError("merge_models: every multibody must have exactly one root link")
        l[:, 0] = np.where(l[:, 0] >= 0, l[:, 0] + lo, -1)
        moving = l[:, 1] >= 0          # JOINT_FIXED = -1 carries no coordinate
        l[moving, 2] += qo
        l[moving, 3] += qdo
        links.append(l)
        g = m[G0:V0].reshape(d["n_geoms"], GEOM).copy()
        base_geoms.append(g[g[:, 0] < 0])
        gl = g[g[:, 0] >= 0]
        gl[:, 0] += lo
        link_geoms.append(gl)
        v = m[V0:V0 + d["n_vis"] * VIS].reshape(d["n_vis"], VIS).copy()
        v[:, 0] = np.where(v[:, 0] >= 0, v[:, 0] + lo, -1)
        vis.append(v)
        lo += d["n_links"]; qo += d["n_q"]; qdo += d["n_qd"]
        if d["has_plane"] and not head[7]:
            head[7] = 1
            head[8:12] = m[8:12]
    geoms = np.concatenate(base_geoms + link_geoms) if (base_geoms or link_geoms) else np.zeros((0, GEOM))
    vis = np.concatenate(vis) if vis else np.zeros((0, VIS))
    head[1], head[2], head[3], head[4], head[5], head[6], head[12] = lo, 0, qo, qdo, len(geoms), len(vis), len(ms)
    return np.concatenate([head, ms[0][HEADER:HEADER + BASE], np.concatenate(links).ravel(), geoms.ravel(), vis.ravel()])


def save_model(path, model, meta=None):
    with open(path, "w") as f:
        json.dump({"layout": "tds_b200_model.h", "meta": meta or {}, "model": [float(v) for v in model]}, f)


def load_model(path):
    with open(path) as f:
        d = json.load(f)
    return np.asarray(d["model"], dtype=np.float64)


def fixture_path(name):
    """Compiled models shipped with the package (`models/<name>.json`: flat models exported from the reference's own URDF
    loader by tests/golden/make_golden.py, which also keeps the copy under tests/golden/models the tests compare against)."""
    here = os.path.dirname(os.path.abspath(__file__))
    p = os.path.join(here, "models", name + ".json")
    if os.path.exists(p):
        return p
    return os.path.join(os.path.dirname(here), "tests", "golden", "models", name + ".json")
