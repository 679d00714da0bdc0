"""The C++ StompRobotModel (cpp/stomp_motion_planner/stomp_robot_model_urdf.hpp: URDF -> segment / sphere / limit / inertia
tables and the robot bodies at the start state, host code above the C ABI) against urdf.py, which test_urdf_cpu.py holds to the
hand-written arm and test_reference_pinning.py to the compiled reference's generateLinkCollisionPoints.  Host-only C++."""
import math
import os
import subprocess

import numpy as np
import pytest

from stomp_motion_planner_icra2011_b200 import _abi
from stomp_motion_planner_icra2011_b200.urdf import robot_from_urdf
from tests.test_urdf_cpu import COLLISION_LINKS, GROUP, STATE, URDF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stomp_motion_planner_icra2011_b200")

# the same arm with <inertial> blocks and <collision> primitives on links outside (torso, base) and inside (upper arm) the group
EXTRA = """
  <link name="base_link"><collision><origin xyz="0 0 0.15"/><geometry><box size="0.65 0.65 0.3"/></geometry></collision></link>
  <link name="torso_lift_link"><inertial><mass value="36.2"/><origin xyz="-0.1 0 -0.08" rpy="0.1 0.2 0.3"/>
      <inertia ixx="2.7" ixy="0.004" ixz="0.17" iyy="2.5" iyz="0.02" izz="0.97"/></inertial>
    <collision><origin xyz="-0.1 0 -0.3" rpy="0 0 0.2"/><geometry><cylinder radius="0.18" length="0.7"/></geometry></collision>
    <collision><origin xyz="0 0 0.35"/><geometry><sphere radius="0.15"/></geometry></collision></link>
  <link name="r_upper_arm_link"><inertial><mass value="6.0"/><origin xyz="0.21 0 0"/><inertia ixx="0.015" iyy="0.077" izz="0.076"/></inertial>
    <collision><origin xyz="0.2 0 0" rpy="0 1.5707963267948966 0"/><geometry><cylinder radius="0.09" length="0.4"/></geometry></collision>
    <collision><geometry><mesh filename="package://upper_arm.stl"/></geometry></collision></link>
"""


def _urdf_with_geometry():
    out = URDF
    for name in ("base_link", "torso_lift_link", "r_upper_arm_link"):
        out = out.replace('<link name="%s"/>' % name, "", 1)
    return out.replace('<robot name="pr2_like_right_arm">', '<robot name="pr2_like_right_arm">' + EXTRA, 1)


def _run(tmp_path, urdf, spec_lines):
    exe = str(tmp_path / "urdf_model_test")
    if not os.path.exists(exe):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "cpp"), "-o", exe,
                               os.path.join(PKG, "cpp", "urdf_model_test.cpp"), "-L", PKG, "-lstomp_b200", "-Wl,-rpath," + PKG])
    (tmp_path / "robot.urdf").write_text(urdf)
    (tmp_path / "spec.txt").write_text("\n".join(spec_lines) + "\n")
    out = subprocess.run([exe, str(tmp_path / "robot.urdf"), str(tmp_path / "spec.txt")], capture_output=True, text=True, timeout=60)
    rows = [l.split() for l in out.stdout.splitlines()]
    return out.returncode, rows


def _spec(start=None, exclude=()):
    lines = ["group " + " ".join(GROUP), "reference base_link", "clearance 0.07", "chain torso_lift_link r_gripper_palm_link"]
    lines += ["collision %s %r %r" % (l, c["link_radius"], c.get("link_extension", 0.0)) for l, c in COLLISION_LINKS.items()]
    lines += ["state %s %r" % kv for kv in STATE.items()]
    if start is not None:
        lines.append("start " + " ".join(repr(float(v)) for v in start))
    if exclude:
        lines.append("exclude " + " ".join(exclude))
    lines.append("padding 1.0 0.01")
    return lines


def test_cpp_robot_model_matches_urdf_py(tmp_path):
    urdf = _urdf_with_geometry()
    want = robot_from_urdf(urdf, GROUP, "base_link", COLLISION_LINKS, 0.07, STATE, dynamics_chain=("torso_lift_link", "r_gripper_palm_link"))
    rc, rows = _run(tmp_path, urdf, _spec())
    assert rc == 0, rows
    segs = [r for r in rows if r[0] == "segment"]
    assert [r[1] for r in segs] == [g["name"] for g in want.segments]                    # DFS pre-order numbering
    for r, g in zip(segs, want.segments):
        assert (int(r[2]), int(r[3]), int(r[4])) == (g["parent"], g["type"], g["group"]), g["name"]
        v = np.array([float(x) for x in r[5:]])
        np.testing.assert_array_equal(v[:9], g["rot"])
        np.testing.assert_array_equal(v[9:12], g["pos"])
        np.testing.assert_array_equal(v[12:15], g["axis"])
        assert v[15] == g["fixed"]
    head = rows[0]
    assert head[0] == "reference" and int(head[1]) == want.reference_segment and (int(head[3]), int(head[4])) == tuple(want.chain)
    lims = [(int(r[1]), float(r[2]), float(r[3])) for r in rows if r[0] == "limit"]
    assert lims == [(int(h), lo, hi) for h, lo, hi in want.limits]
    sph = [r for r in rows if r[0] == "sphere"]
    assert len(sph) == len(want.spheres) == 47
    for r, s in zip(sph, want.spheres):
        assert int(r[1]) == s["segment"] and float(r[2]) == s["radius"] and float(r[3]) == s["clearance"]
        np.testing.assert_array_equal([float(x) for x in r[4:7]], s["pos"])
    inert = {int(r[1]): [float(x) for x in r[2:]] for r in rows if r[0] == "inertia"}
    assert set(inert) == set(want.inertias)
    for seg, (m, com, ic) in want.inertias.items():
        np.testing.assert_allclose(inert[seg], [m, *com, *ic], rtol=1e-14, atol=1e-18)


def test_cpp_robot_bodies_at_the_start_state(tmp_path):
    """the links' <collision> primitives placed at a joint state, in the reference frame: what setStartState voxelises"""
    urdf = _urdf_with_geometry()
    start = [0.3, 0.2, -0.4, -1.0, 0.5, -0.6, 0.1]
    rc, rows = _run(tmp_path, urdf, _spec(start=start, exclude=["r_upper_arm_link"]))
    assert rc == 0, rows
    bodies = [r for r in rows if r[0] == "body"]
    assert len(bodies) == 3                        # base box, torso cylinder + sphere; the arm link is excluded, the mesh skipped
    box, cyl, sph = ([float(x) for x in b[1:]] for b in bodies)
    assert box[0] == _abi.BODY_BOX and box[1:4] == [0.65, 0.65, 0.3] and box[4:7] == [0.0, 0.0, 0.15] and box[7:11] == [0.0, 0.0, 0.0, 1.0]
    torso = np.array([-0.05, 0.0, 0.739675 + 0.1])           # prismatic torso joint at its state value 0.1
    assert cyl[0] == _abi.BODY_CYLINDER and cyl[1:3] == [0.18, 0.7]
    np.testing.assert_allclose(cyl[4:7], torso + [-0.1, 0.0, -0.3], atol=1e-15)
    np.testing.assert_allclose(cyl[7:11], [0.0, 0.0, math.sin(0.1), math.cos(0.1)], atol=1e-15)
    assert sph[0] == _abi.BODY_SPHERE and sph[1] == 0.15
    np.testing.assert_allclose(sph[4:7], torso + [0.0, 0.0, 0.35], atol=1e-15)
    assert cyl[11:13] == [1.0, 0.01]
    # without the exclusion the upper-arm cylinder appears, moved by the first three group joints
    rc, rows = _run(tmp_path, urdf, _spec(start=start))
    arm = [[float(x) for x in b[1:]] for b in rows if b[0] == "body"][3]
    want = robot_from_urdf(urdf, GROUP, "base_link", COLLISION_LINKS, 0.07, STATE)
    # FK by hand: torso -> pan (z) -> lift (y, offset 0.1) -> roll (x): the body centre sits 0.2 m along the rolled x axis
    def rot(axis, q):
        c, s = math.cos(q), math.sin(q)
        x, y, z = axis
        return np.array([[c + (1 - c) * x * x, (1 - c) * x * y - s * z, (1 - c) * x * z + s * y],
                         [(1 - c) * x * y + s * z, c + (1 - c) * y * y, (1 - c) * y * z - s * x],
                         [(1 - c) * x * z - s * y, (1 - c) * y * z + s * x, c + (1 - c) * z * z]])
    Rpan, Rlift, Rroll = rot((0, 0, 1), start[0]), rot((0, 1, 0), start[1]), rot((1, 0, 0), start[2])
    p = torso + [0.0, -0.188, 0.0] + Rpan @ np.array([0.1, 0.0, 0.0]) + Rpan @ Rlift @ Rroll @ np.array([0.2, 0.0, 0.0])
    np.testing.assert_allclose(arm[4:7], p, atol=1e-14)
    assert len(want.segments) == 18


def test_cpp_attached_object_collision_points(tmp_path):
    """generateAttachedObjectCollisionPoints: one bounding-sphere collision point per attached shape, like scenes.Robot.add_attached_object
    (pinned to the compiled reference by tests/test_reference_pinning.py); an object on a link no group joint moves adds nothing."""
    want = robot_from_urdf(URDF, GROUP, "base_link", COLLISION_LINKS, 0.07, STATE)
    seg = [g["name"] for g in want.segments].index("r_gripper_palm_link")
    want.add_attached_object(seg, "box", (0.1, 0.06, 0.2), (0.12, 0.0, 0.01), padding=0.01, clearance=0.07)
    want.add_attached_object(seg, "cylinder", (0.03, 0.25), (0.15, 0.02, 0.0), padding=0.0, clearance=0.07)
    want.add_attached_object(seg, "sphere", (0.04,), (0.2, 0.0, 0.0), padding=0.005, clearance=0.07)
    lines = _spec() + ["attach r_gripper_palm_link 1 0.1 0.06 0.2 0.12 0.0 0.01 0.01", "attach r_gripper_palm_link 2 0.03 0.25 0 0.15 0.02 0.0 0.0",
                       "attach r_gripper_palm_link 0 0.04 0 0 0.2 0.0 0.0 0.005", "attach torso_lift_link 0 0.1 0 0 0 0 0 0"]
    rc, rows = _run(tmp_path, URDF, lines)
    assert rc == 0, rows
    sph = [r for r in rows if r[0] == "sphere"]
    assert len(sph) == len(want.spheres) == 50
    for r, s_ in zip(sph[47:], want.spheres[47:]):
        assert int(r[1]) == s_["segment"] and float(r[3]) == s_["clearance"]
        np.testing.assert_allclose(float(r[2]), s_["radius"], rtol=1e-15)
        np.testing.assert_array_equal([float(x) for x in r[4:7]], s_["pos"])


def test_cpp_urdf_errors(tmp_path):
    rc, rows = _run(tmp_path, URDF.replace("</robot>", ""), _spec())
    assert rc == 1 and any(r[0] == "error" for r in rows)
    rc, rows = _run(tmp_path, URDF, ["group nope_joint", "reference base_link"])
    assert rc == 1 and "not in the URDF" in " ".join(rows[0])
    rc, rows = _run(tmp_path, URDF, ["group " + " ".join(GROUP), "reference nowhere"])
    assert rc == 1 and "reference frame" in " ".join(rows[0])


def _box_triangles(half):
    hx, hy, hz = half
    c = np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    return np.array([[c[q[0]], c[q[1]], c[q[2]]] for q in quads] + [[c[q[0]], c[q[2]], c[q[3]]] for q in quads])


def test_cpp_mesh_link_geometry(tmp_path):
    """<mesh> collision geometry: file name and scale from the URDF, vertices from a binary or an ASCII STL file
    (loadLinkMeshes), placed at the joint state like the primitives (meshBodiesAtState) - the input of
    stomp_engine_build_sdf_meshes, which voxelises the vertices' convex hull like bodies::ConvexMesh."""
    import struct
    tris = _box_triangles((0.2, 0.05, 0.06))
    meshes = tmp_path / "meshes"
    meshes.mkdir()
    with open(meshes / "upper_arm.stl", "wb") as f:             # binary STL
        f.write(b"binary stl".ljust(80, b" ") + struct.pack("<I", len(tris)))
        for t in tris:
            f.write(struct.pack("<12fH", 0.0, 0.0, 0.0, *t.ravel(), 0))
    with open(meshes / "upper_arm_ascii.stl", "w") as f:        # ASCII STL
        f.write("solid arm\n")
        for t in tris:
            f.write(" facet normal 0 0 0\n  outer loop\n" + "".join("   vertex %r %r %r\n" % tuple(float(x) for x in v) for v in t) + "  endloop\n endfacet\n")
        f.write("endsolid arm\n")
    urdf = _urdf_with_geometry()
    start = [0.3, 0.2, -0.4, -1.0, 0.5, -0.6, 0.1]
    rc, rows = _run(tmp_path, urdf, _spec(start=start) + ["meshdir " + str(meshes)])
    assert rc == 0, rows
    mesh = [[float(x) for x in r[1:]] for r in rows if r[0] == "mesh"]
    assert len(mesh) == 1 and mesh[0][0] == 36
    np.testing.assert_allclose(mesh[0][10:16], [-0.2, -0.05, -0.06, 0.2, 0.05, 0.06], rtol=1e-7)     # float32 file
    assert mesh[0][8:10] == [1.0, 0.01]
    # same link as the upper-arm cylinder (4th primitive): same orientation up to the cylinder's own rpy, origin 0.2 m behind it
    arm = [[float(x) for x in b[1:]] for b in rows if b[0] == "body"][3]
    def quat_to_rot(q):
        x, y, z, w = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    Rlink = quat_to_rot(mesh[0][4:8])
    np.testing.assert_allclose(np.array(mesh[0][1:4]) + Rlink @ np.array([0.2, 0.0, 0.0]), arm[4:7], atol=1e-14)
    # the links of the excluded list carry no mesh either
    rc, rows = _run(tmp_path, urdf, _spec(start=start, exclude=["r_upper_arm_link"]) + ["meshdir " + str(meshes)])
    assert rc == 0 and not [r for r in rows if r[0] == "mesh"]
    # ASCII file + URDF scale
    scaled = urdf.replace('<mesh filename="package://upper_arm.stl"/>', '<mesh filename="package://upper_arm_ascii.stl" scale="2 1 0.5"/>')
    rc, rows = _run(tmp_path, scaled, _spec(start=start) + ["meshdir " + str(meshes)])
    assert rc == 0, rows
    mesh = [[float(x) for x in r[1:]] for r in rows if r[0] == "mesh"]
    np.testing.assert_allclose(mesh[0][10:16], [-0.4, -0.05, -0.03, 0.4, 0.05, 0.03], rtol=1e-15)
    # a missing file is an error, not a silently empty body
    rc, rows = _run(tmp_path, urdf, _spec(start=start) + ["meshdir " + str(tmp_path / "nowhere")])
    assert rc == 1 and any(r[0] == "error" for r in rows)
