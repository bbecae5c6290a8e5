"""Host-side model compiler (csrc/urdf_model.cpp): URDF -> flat model.  CPU only."""
import os

import numpy as np
import pytest

import tds_b200
from tds_b200.model import compile_urdf, fixture_path, load_model, model_dims
from conftest import REFERENCE_ROOT, have_reference_tree

TINY_URDF = """<?xml version="1.0"?>
<!-- two-link arm with a fixed tool and a capsule -->
<robot name="arm">
  <link name="base"><inertial><mass value="0"/><inertia ixx="0" iyy="0" izz="0" ixy="0" ixz="0" iyz="0"/></inertial></link>
  <link name="upper"><inertial><origin xyz="0 0 0.25" rpy="0 0 0"/><mass value="2"/><inertia ixx="0.1" iyy="0.2" izz="0.3"/></inertial>
    <collision><origin xyz="0 0 0.25"/><geometry><capsule radius="0.05" length="0.5"/></geometry></collision></link>
  <link name="tool"><inertial><mass value="0.1"/><inertia ixx="1e-3" iyy="1e-3" izz="1e-3"/></inertial>
    <collision><geometry><sphere radius='0.02'/></geometry></collision>
    <visual><origin xyz="0 0 0.1" rpy="0 0 1.57"/><geometry><mesh filename="x.obj"/></geometry></visual></link>
  <link name="lower"><inertial><origin xyz="0 0 0.2"/><mass value="1"/><inertia ixx="0.05" iyy="0.05" izz="0.01"/></inertial></link>
  <joint name="j2" type="continuous"><parent link="upper"/><child link="lower"/><origin xyz="0 0 0.5"/><axis xyz="0 -1 0"/></joint>
  <joint name="j1" type="revolute"><parent link="base"/><child link="upper"/><origin xyz="0 0 0.1" rpy="0 0 0.3"/><axis xyz="1 0 0"/></joint>
  <joint name="jt" type="fixed"><parent link="lower"/><child link="tool"/><origin xyz="0 0 0.4"/></joint>
</robot>"""
PLANE = '<robot name="p"><link name="l"><collision><geometry><plane normal="0 0 2"/></geometry></collision></link></robot>'


def test_tiny_urdf_structure():
    m = compile_urdf(TINY_URDF, PLANE, floating=False)
    d = model_dims(m)
    assert d == dict(n_links=3, floating=0, n_q=2, n_qd=2, n_geoms=2, n_vis=1, has_plane=1)
    assert np.allclose(m[8:11], [0, 0, 1])                       # plane normal normalised, constant 0
    links = m[16 + 13:].reshape(-1)[:3 * 34].reshape(3, 34)
    # pre-order DFS, children in joint document order: upper(0) <- lower(1) <- tool(2)
    assert links[:, 0].tolist() == [-1, 0, 1]
    assert links[:, 1].tolist() == [4, 7, -1]                    # REVOLUTE_X, REVOLUTE_AXIS (axis -y != +1), FIXED
    assert links[:, 2].tolist() == [0, 1, -2] and links[:, 3].tolist() == [0, 1, -2]
    assert links[1, 4:7].tolist() == [0, -1, 0]
    assert np.allclose(links[0, 7:16].reshape(3, 3), [[np.cos(.3), -np.sin(.3), 0], [np.sin(.3), np.cos(.3), 0], [0, 0, 1]])
    geoms = m[16 + 13 + 3 * 34:][:2 * 18].reshape(2, 18)
    assert geoms[:, 0].tolist() == [0, 2] and geoms[:, 1].tolist() == [2, 0]     # capsule on link 0, sphere on link 2


def test_floating_indices_and_default_axis():
    u = TINY_URDF.replace('<axis xyz="1 0 0"/>', "")
    m = compile_urdf(u, None, floating=True)
    d = model_dims(m)
    assert (d["n_q"], d["n_qd"], d["has_plane"]) == (9, 8, 0)
    links = m[16 + 13:][:3 * 34].reshape(3, 34)
    assert links[0, 1] == 6 and links[0, 2] == 7 and links[0, 3] == 6           # default axis (0,0,1) -> REVOLUTE_Z; q starts at 7


def test_spherical_joint_indices():
    """A spherical joint takes 4 coordinates (quaternion) and 3 velocities (MultiBody::initialize, multi_body.hpp:324-349)."""
    m = compile_urdf(TINY_URDF.replace('type="continuous"', 'type="spherical"'), None, floating=False)
    d = model_dims(m)
    links = m[16 + 13:][:d["n_links"] * 34].reshape(d["n_links"], 34)
    sph = np.nonzero(links[:, 1] == 8)[0]
    assert sph.size >= 1 and d["n_q"] - d["n_qd"] == sph.size
    for i in sph:
        later = links[i + 1:][links[i + 1:, 1] != -1]
        if later.size:
            assert later[0, 2] == links[i, 2] + 4 and later[0, 3] == links[i, 3] + 3


@pytest.mark.parametrize("bad,msg", [
    ("<robot name='x'><link name='a'/><link name='b'/></robot>", "multiple parent links"),
    ("<robot name='x'><link name='a'></robot>", "XML error"),
    (TINY_URDF.replace('type="continuous"', 'type="planar"'), "unsupported type"),
    ("<robot><link name='a'/></robot>", "name"),
])
def test_errors(bad, msg):
    with pytest.raises(ValueError) as e:
        compile_urdf(bad)
    assert msg in str(e.value)


@pytest.mark.skipif(not have_reference_tree(), reason="reference URDF data only exists in the build container")
@pytest.mark.parametrize("name,urdf,plane,floating", [
    ("cartpole", "cartpole.urdf", None, False), ("pendulum5", "pendulum5.urdf", None, False),
    ("sphere2", "sphere2.urdf", "plane_implicit.urdf", True),
    ("laikago", "laikago/laikago_toes_zup_xyz_xyzrot.urdf", "plane_implicit.urdf", False),
    ("humanoid", "humanoid.urdf", "plane_implicit.urdf", True),
    ("ant", "gym/ant_org_xyz_xyzrot.urdf", "plane_implicit.urdf", False),
    ("cartpole_plane", "cartpole.urdf", "plane_implicit.urdf", False),
    ("pendulum5spherical", "pendulum5spherical.urdf", None, False),
    ("humanoid_spherical", "humanoid_xyz_spherical.urdf", "plane_implicit.urdf", False)])
def test_matches_reference_loader(name, urdf, plane, floating):
    """Our compiler on the reference's URDFs == the flat export of the reference's own loader
    (fixtures were exported from UrdfCache::construct by tests/golden/make_golden.py)."""
    D = os.path.join(REFERENCE_ROOT, "data")
    mine = compile_urdf(os.path.join(D, urdf), os.path.join(D, plane) if plane else None, floating)
    ref = load_model(fixture_path(name))
    assert mine.shape == ref.shape
    assert np.abs(mine - ref).max() < 1e-15


def test_generated_ant_table_is_the_fixture():
    """The Ant model the specialised kernel was generated from equals the reference-exported Ant model."""
    inc = os.path.join(os.path.dirname(tds_b200.lib_path()), "csrc", "generated", "ant_model.inc")
    vals = []
    for line in open(inc):
        if line.startswith("//"):
            continue
        vals += [float(x) for x in line.strip().rstrip(",").split(",") if x.strip()]
    assert np.array_equal(np.array(vals), load_model(fixture_path("ant")))


def test_generated_laikago_table_is_the_fixture():
    """The model table embedded in the C-ABI v1 drop-in equals the reference-exported Laikago model."""
    inc = os.path.join(os.path.dirname(tds_b200.lib_path()), "csrc", "generated", "laikago_model.inc")
    vals = []
    for line in open(inc):
        if line.startswith("//"):
            continue
        vals += [float(x) for x in line.strip().rstrip(",").split(",") if x.strip()]
    assert np.array_equal(np.array(vals), load_model(fixture_path("laikago")))


@pytest.mark.parametrize("name,inc,n_act", [("SpecLaikago", "laikago_model.inc", 12), ("SpecAnt", "ant_model.inc", 8)])
def test_committed_spec_headers_are_what_gen_spec_emits(name, inc, n_act, tmp_path):
    """The constexpr model tables the specialised kernel is compiled from are reproducible: gen_spec on the committed
    flat model gives the committed header byte for byte."""
    import subprocess
    csrc = os.path.join(os.path.dirname(tds_b200.lib_path()), "csrc")
    inc_dir = os.path.join(os.path.dirname(os.path.dirname(tds_b200.lib_path())), "include")
    exe = str(tmp_path / "gen_spec")
    subprocess.check_call([os.environ.get("TDS_CXX", "/usr/bin/g++"), "-std=c++17", "-O1", "-I", csrc, "-I", inc_dir,
                           os.path.join(csrc, "gen_spec.cpp"), "-o", exe])
    out = str(tmp_path / "spec.h")
    subprocess.check_call([exe, name, os.path.join(csrc, "generated", inc), str(n_act), "6", out])
    committed = os.path.join(csrc, "generated", "spec_" + name[4:].lower() + ".h")
    assert open(out).read() == open(committed).read()


def _random_urdf(rng, n_links, massless_links=True, boxes=True, extras=False):
    """A random tree of links: every joint type the step supports, unit / negative / oblique axes, joint and inertial
    origins with rotations, sphere / capsule / box collision shapes with their own origins, link visuals."""
    def v3(lo, hi):
        return " ".join("%.6g" % x for x in rng.uniform(lo, hi, 3))
    parts = ['<?xml version="1.0"?>', '<robot name="rnd">']
    for i in range(n_links):
        # (the reference's loader exits on a massless floating base: the root always has mass)
        mass = 0.0 if (massless_links and i > 0 and rng.random() < 0.15) else rng.uniform(0.1, 5.0)
        ixx, iyy, izz = (rng.uniform(1e-3, 0.5, 3) if mass > 0 else np.zeros(3))
        s = [f'<link name="l{i}">',
             f'<inertial><origin xyz="{v3(-0.2, 0.2)}" rpy="{v3(-1, 1)}"/><mass value="{mass:.6g}"/>'
             f'<inertia ixx="{ixx:.6g}" iyy="{iyy:.6g}" izz="{izz:.6g}" ixy="0" ixz="0" iyz="0"/></inertial>']
        for _ in range(rng.integers(0, 3)):
            kind = rng.integers(0, 3 if boxes else 2)
            geo = (f'<sphere radius="{rng.uniform(0.02, 0.2):.6g}"/>' if kind == 0 else
                   f'<capsule radius="{rng.uniform(0.02, 0.1):.6g}" length="{rng.uniform(0.1, 0.6):.6g}"/>' if kind == 1 else
                   f'<box size="{v3(0.05, 0.4)}"/>')
            if extras and rng.random() < 0.15:   # a plane shape on a link: kept, with its normal normalised (geometry.hpp:183)
                geo = f'<plane normal="{v3(-1, 1)}"/>'
            elif extras and rng.random() < 0.3:  # shapes the reference's loader drops (urdf_to_multi_body.hpp:234-277)
                geo = ('<mesh filename="part.obj" scale="1 1 1"/>' if rng.random() < 0.5 else
                       f'<cylinder radius="{rng.uniform(0.02, 0.1):.6g}" length="{rng.uniform(0.1, 0.6):.6g}"/>')
            s.append(f'<collision><origin xyz="{v3(-0.3, 0.3)}" rpy="{v3(-1.5, 1.5)}"/><geometry>{geo}</geometry></collision>')
        if rng.random() < 0.6:
            s.append(f'<visual><origin xyz="{v3(-0.1, 0.1)}" rpy="{v3(-1, 1)}"/><geometry><sphere radius="0.1"/></geometry></visual>')
        s.append('</link>')
        parts.append("".join(s))
    axes = ["1 0 0", "0 1 0", "0 0 1", "-1 0 0", "0 0 -1", "0.6 0 0.8", "0.3 -0.4 0.5", None]
    order = list(range(1, n_links))
    rng.shuffle(order)                       # joints in a random document order
    for i in order:
        parent = int(rng.integers(0, i))
        jt = ["revolute", "continuous", "prismatic", "fixed"][rng.integers(0, 4)]
        if extras and rng.random() < 0.2:
            jt = "spherical"
        ax = axes[rng.integers(0, len(axes))]
        axis = f'<axis xyz="{ax}"/>' if (ax is not None and jt != "fixed") else ""
        lim = '<limit lower="-1" upper="1" effort="10" velocity="10"/>' if jt in ("revolute", "prismatic") else ""
        parts.append(f'<joint name="j{i}" type="{jt}"><parent link="l{parent}"/><child link="l{i}"/>'
                     f'<origin xyz="{v3(-0.5, 0.5)}" rpy="{v3(-2, 2)}"/>{axis}{lim}</joint>')
    parts.append("</robot>")
    return "\n".join(parts)


@pytest.mark.skipif(not have_reference_tree(), reason="differential test against the reference's own URDF loader")
@pytest.mark.parametrize("seed", range(24))
def test_random_urdfs_match_reference_loader(seed, tmp_path):
    """Our URDF -> flat-model compiler against UrdfCache::construct + UrdfToMultiBody (through oracle/_ref) on random trees:
    link order, joint typing (axis exactly +1 -> REVOLUTE_X/Y/Z, otherwise *_AXIS), transforms, inertias, shapes, visuals."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1000 + seed)
    text = _random_urdf(rng, int(rng.integers(2, 12)))
    path = tmp_path / "rnd.urdf"
    path.write_text(text)
    floating = bool(seed % 2)
    plane = os.path.join(REFERENCE_ROOT, "data", "plane_implicit.urdf") if seed % 3 else None
    theirs = ref.RefSim.from_urdf(str(path), plane, floating).export_model()
    mine = compile_urdf(str(path), plane, floating)
    assert mine.shape == theirs.shape, text
    assert np.abs(mine - theirs).max() < 1e-14, text


@pytest.mark.parametrize("seed", range(12))
def test_random_urdfs_with_spherical_joints_and_dropped_shapes_match_reference_loader(seed, tmp_path):
    """As above with spherical joints (4 coordinates / 3 velocities in the index bookkeeping) and the collision shapes the
    reference's loader silently drops (mesh, cylinder); own plane URDF, so the test needs oracle/_ref only."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(2500 + seed)
    text = _random_urdf(rng, int(rng.integers(2, 12)), extras=True)
    path = tmp_path / "rnd.urdf"
    path.write_text(text)
    plane = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "urdf", "plane.urdf") if seed % 3 else None
    theirs = ref.RefSim.from_urdf(str(path), plane, False).export_model()
    mine = compile_urdf(str(path), plane, False)
    assert mine.shape == theirs.shape, text
    assert np.abs(mine - theirs).max() < 1e-14, text
