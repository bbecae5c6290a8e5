"""Host logic of the kernels' tree decomposition (csrc/tds_team.h), checked on the CPU: invariants of the partition for the
fixture models and for random trees (tests/cpp/team_check.cpp compiled with g++)."""
import os
import subprocess

import numpy as np
import pytest

import tds_b200
from tds_b200.model import compile_urdf, fixture_path, load_model
from test_model_compiler import _random_urdf, PLANE


@pytest.fixture(scope="module")
def team_check(tmp_path_factory):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(os.path.dirname(tds_b200.lib_path()), "csrc")
    exe = str(tmp_path_factory.mktemp("team") / "team_check")
    subprocess.check_call([os.environ.get("TDS_CXX", "/usr/bin/g++"), "-std=c++17", "-O1", "-I", csrc, "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "team_check.cpp"), "-o", exe])
    return exe


def _run(exe, model, tmp_path, extra=()):
    path = str(tmp_path / "m.bin")
    np.asarray(model, dtype=np.float64).tofile(path)
    return subprocess.run([exe, path, *map(str, extra)], capture_output=True, text=True, timeout=60)


@pytest.mark.parametrize("name,extra,code", [("laikago", (12, 6), 0), ("ant", (8, 6), 0), ("humanoid", (), 0),
                                             ("pendulum5", (), 4), ("cartpole", (), 4)])
def test_fixture_models(team_check, tmp_path, name, extra, code):
    r = _run(team_check, load_model(fixture_path(name)), tmp_path, extra)
    assert r.returncode == code, r.stdout + r.stderr


@pytest.mark.parametrize("seed", range(40))
def test_random_trees(team_check, tmp_path, seed):
    rng = np.random.default_rng(9000 + seed)
    text = _random_urdf(rng, int(rng.integers(3, 24)), boxes=False)
    model = compile_urdf(text, PLANE if seed % 4 else None, floating=bool(seed % 2))
    r = _run(team_check, model, tmp_path)
    if r.returncode == 3 and "(-2)" in r.stderr:
        pytest.skip("random model exceeds the library's link / geom capacity")
    assert r.returncode in (0, 4), text + "\n" + r.stdout + r.stderr
