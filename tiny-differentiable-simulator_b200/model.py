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
