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
    assert rc == 1 and rows[0][0] == "error"
    rc, rows = _run(tmp_path, URDF, ["group nope_joint", "reference base_link"])
    assert rc == 1 and "not in the URDF" in " ".join(rows[0])
    rc, rows = _run(tmp_path, URDF, ["group " + " ".join(GROUP), "reference nowhere"])
    assert rc == 1 and "reference frame" in " ".join(rows[0])
