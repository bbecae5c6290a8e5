"""C-ABI library: loads on a CPU box, exports every symbol include/tds_b200.h declares, and refuses
to create a simulator without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

import tds_b200
from tds_b200 import _lib
from tds_b200.model import fixture_path, load_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_in_header():
    text = open(os.path.join(ROOT, "include", "tds_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:tds_b200|cuda_model_laikago|cuda_model_ant|b200_laikago)_\w+|model_info)\s*\(", text)))


def test_header_symbols_exported():
    L = tds_b200.lib()
    names = declared_in_header()
    assert len(names) >= 18
    for nm in names:
        assert hasattr(L, nm), f"{nm} declared in include/tds_b200.h but not exported"
    assert sorted(_lib.DECLARED_SYMBOLS) == names


def test_v1_meta_matches_reference_dims():
    m = tds_b200.lib().cuda_model_laikago_forward_zero_meta()
    assert (m.input_dim, m.output_dim, m.global_dim) == (51, 411, 0)
    m = tds_b200.lib().cuda_model_ant_forward_zero_meta()
    assert (m.input_dim, m.output_dim, m.global_dim) == (39, 155, 0)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError) as e:
        tds_b200.BatchSim(load_model(fixture_path("laikago")), 8)
    assert "no CUDA device" in str(e.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the product package may import, include or load it."""
    pkg = os.path.dirname(tds_b200.lib_path())
    for dp, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                continue
            txt = open(os.path.join(dp, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
            assert not re.search(r"#include\s*[\"<][^\n]*oracle", txt), f
            assert "libtds_oracle" not in txt and "libtds_ref" not in txt and "tdso_" not in txt and "tdsref_" not in txt, f


def test_specialised_kernels_fit_two_tiles_per_sm():
    """Shared memory of a tile of the compiled models in the default (mixed) arithmetic: for the headline model two tiles
    must be co-resident on an SM (228 KB, 1 KB reserved per CTA) for batches with more tiles than SMs."""
    L = ctypes.CDLL(tds_b200.lib_path())
    L.tds_spec_smem_bytes.restype = ctypes.c_size_t
    L.tds_spec_smem_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
    L.tds_spec_name.restype = ctypes.c_char_p
    assert L.tds_spec_name(0) == b"laikago" and L.tds_spec_name(1) == b"ant"
    assert 2 * (L.tds_spec_smem_bytes(0, 0) + 1024) <= 228 * 1024          # headline model: two tiles per SM
    for spec in (0, 1):
        assert L.tds_spec_smem_bytes(spec, 0) <= 227 * 1024                 # default arithmetic fits one CTA
    # (an instance that does not fit - Ant in all-fp64: 235 KB - is not launched: the table-driven kernels take over)


def test_unsupported_models_are_refused_loudly():
    """Host-only model check: shapes / joints the step does not implement must fail the create, never be skipped silently."""
    import numpy as np
    L = tds_b200.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    L.tds_b200_last_error.restype = ctypes.c_char_p

    def check(m):
        m = np.ascontiguousarray(m, dtype=np.float64)
        return L.tds_b200_validate_model(m.ctypes.data_as(dp), int(m.size)), L.tds_b200_last_error().decode()

    for name in ("cartpole", "pendulum5", "sphere2", "laikago", "humanoid", "ant", "box", "cartpole_plane", "pendulum5spherical",
                 "humanoid_spherical"):
        assert check(load_model(fixture_path(name)))[0] == 0, name
    # a mesh shape against the ground plane: the reference collides it, the contact stage here does not -> refused
    m = np.array(load_model(fixture_path("box")), dtype=np.float64)
    m[16 + 13 + 1] = 3.0                           # the geom's type -> TDSG_MESH (no links: geoms follow the base record)
    rc, msg = check(m)
    assert rc == -6 and "mesh" in msg
    m = np.array(load_model(fixture_path("laikago")), dtype=np.float64)
    m[16 + 13 + 1] = 9.0                           # unknown joint type
    rc, msg = check(m)
    assert rc == -3 and "joint type" in msg
    assert check(m[:10])[0] == -1


def test_contact_pair_lists_match_reference_goldens():
    """Host-only: the candidate contact-pair list of every fixture model (enumeration order of
    World::compute_contacts_multi_body_internal, src/world.hpp:212-281) equals, bit for bit, the (link_a, link_b) list the
    reference itself produced for the golden vectors.  (The per-step lists come from the device: tests/test_parity_gpu.py.)"""
    import numpy as np
    L = tds_b200.lib()
    golden = os.path.join(ROOT, "tests", "golden")
    for name in ("sphere2", "laikago", "humanoid", "ant", "box", "cartpole_plane", "cartpole", "pendulum5", "humanoid_spherical",
                 "pendulum5spherical"):
        m = np.ascontiguousarray(load_model(fixture_path(name)), dtype=np.float64)
        t = np.zeros((64, 4), dtype=np.int32)
        k = L.tds_b200_model_contact_pairs(m.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(m.size), ctypes.c_void_p(t.ctypes.data), 64)
        g = np.load(os.path.join(golden, name + ".npz"))
        la, lb = np.stack(list(g["contact_link_a"])), np.stack(list(g["contact_link_b"]))
        assert k == la.shape[1] == int(g["n_contacts"][0]), name
        for e in range(la.shape[0]):
            assert np.array_equal(t[:k, 1], la[e]) and np.array_equal(t[:k, 3], lb[e]), name
        assert np.all(t[:k, 0] == 0) and np.all(t[:k, 2] == 1)


def test_packaged_models_equal_the_golden_fixtures():
    """The models the package ships (tds_b200/models) are the fixtures exported from the reference (tests/golden/models)."""
    import filecmp
    pkg = os.path.join(os.path.dirname(tds_b200.lib_path()), "models")
    gold = os.path.join(ROOT, "tests", "golden", "models")
    names = sorted(f for f in os.listdir(gold) if f.endswith(".json"))
    assert names and sorted(f for f in os.listdir(pkg) if f.endswith(".json")) == names
    for f in names:
        assert filecmp.cmp(os.path.join(pkg, f), os.path.join(gold, f), shallow=False), f
