"""URDF ingestion (StompRobotModel's job) reproduces the hand-written PR2-like joint table and sphere list."""
import numpy as np
import pytest

from stomp_motion_planner_icra2011_b200 import _abi, scenes
from stomp_motion_planner_icra2011_b200.urdf import robot_from_urdf

URDF = """<?xml version="1.0"?>
<robot name="pr2_like_right_arm">
  <link name="base_link"/><link name="torso_lift_link"/><link name="r_shoulder_pan_link"/><link name="r_shoulder_lift_link"/>
  <link name="r_upper_arm_roll_link"/><link name="r_upper_arm_link"/><link name="r_elbow_flex_link"/><link name="r_forearm_roll_link"/>
  <link name="r_forearm_link"/><link name="r_wrist_flex_link"/><link name="r_wrist_roll_link"/><link name="r_gripper_palm_link"/>
  <link name="r_gripper_l_finger_link"/><link name="r_gripper_l_finger_tip_link"/><link name="r_gripper_l_finger_tip_frame"/>
  <link name="r_gripper_r_finger_link"/><link name="r_gripper_r_finger_tip_link"/><link name="r_gripper_r_finger_tip_frame"/>
  <joint name="torso_lift_joint" type="prismatic"><parent link="base_link"/><child link="torso_lift_link"/>
    <origin xyz="-0.05 0 0.739675"/><axis xyz="0 0 1"/><limit lower="0" upper="0.31"/></joint>
  <joint name="r_shoulder_pan_joint" type="revolute"><parent link="torso_lift_link"/><child link="r_shoulder_pan_link"/>
    <origin xyz="0 -0.188 0"/><axis xyz="0 0 1"/><limit lower="-2.1353981634" upper="0.564601836603"/></joint>
  <joint name="r_shoulder_lift_joint" type="revolute"><parent link="r_shoulder_pan_link"/><child link="r_shoulder_lift_link"/>
    <origin xyz="0.1 0 0"/><axis xyz="0 1 0"/><limit lower="-0.3536" upper="1.2963"/></joint>
  <joint name="r_upper_arm_roll_joint" type="revolute"><parent link="r_shoulder_lift_link"/><child link="r_upper_arm_roll_link"/>
    <origin xyz="0 0 0"/><axis xyz="1 0 0"/><limit lower="-3.75" upper="0.65"/></joint>
  <joint name="r_upper_arm_joint" type="fixed"><parent link="r_upper_arm_roll_link"/><child link="r_upper_arm_link"/></joint>
  <joint name="r_elbow_flex_joint" type="revolute"><parent link="r_upper_arm_link"/><child link="r_elbow_flex_link"/>
    <origin xyz="0.4 0 0"/><axis xyz="0 1 0"/><limit lower="-2.1213" upper="-0.15"/></joint>
  <joint name="r_forearm_roll_joint" type="continuous"><parent link="r_elbow_flex_link"/><child link="r_forearm_roll_link"/>
    <axis xyz="1 0 0"/></joint>
  <joint name="r_forearm_joint" type="fixed"><parent link="r_forearm_roll_link"/><child link="r_forearm_link"/></joint>
  <joint name="r_wrist_flex_joint" type="revolute"><parent link="r_forearm_link"/><child link="r_wrist_flex_link"/>
    <origin xyz="0.321 0 0"/><axis xyz="0 1 0"/><limit lower="-2.0" upper="-0.1"/></joint>
  <joint name="r_wrist_roll_joint" type="continuous"><parent link="r_wrist_flex_link"/><child link="r_wrist_roll_link"/>
    <axis xyz="1 0 0"/></joint>
  <joint name="r_gripper_palm_joint" type="fixed"><parent link="r_wrist_roll_link"/><child link="r_gripper_palm_link"/></joint>
  <joint name="r_gripper_l_finger_joint" type="revolute"><parent link="r_gripper_palm_link"/><child link="r_gripper_l_finger_link"/>
    <origin xyz="0.07691 0.01 0"/><axis xyz="0 0 1"/><limit lower="0" upper="0.548"/></joint>
  <joint name="r_gripper_l_finger_tip_joint" type="revolute"><parent link="r_gripper_l_finger_link"/><child link="r_gripper_l_finger_tip_link"/>
    <origin xyz="0.09137 0.00495 0"/><axis xyz="0 0 -1"/><limit lower="0" upper="0.548"/></joint>
  <joint name="r_gripper_l_finger_tip_frame_joint" type="fixed"><parent link="r_gripper_l_finger_tip_link"/><child link="r_gripper_l_finger_tip_frame"/>
    <origin xyz="0.03 0 0" rpy="0 0 0.1"/></joint>
  <joint name="r_gripper_r_finger_joint" type="revolute"><parent link="r_gripper_palm_link"/><child link="r_gripper_r_finger_link"/>
    <origin xyz="0.07691 -0.01 0"/><axis xyz="0 0 -1"/><limit lower="0" upper="0.548"/></joint>
  <joint name="r_gripper_r_finger_tip_joint" type="revolute"><parent link="r_gripper_r_finger_link"/><child link="r_gripper_r_finger_tip_link"/>
    <origin xyz="0.09137 -0.00495 0"/><axis xyz="0 0 1"/><limit lower="0" upper="0.548"/></joint>
  <joint name="r_gripper_r_finger_tip_frame_joint" type="fixed"><parent link="r_gripper_r_finger_tip_link"/><child link="r_gripper_r_finger_tip_frame"/>
    <origin xyz="0.03 0 0" rpy="0 0 -0.1"/></joint>
</robot>
"""
GROUP = ["r_shoulder_pan_joint", "r_shoulder_lift_joint", "r_upper_arm_roll_joint", "r_elbow_flex_joint", "r_forearm_roll_joint",
         "r_wrist_flex_joint", "r_wrist_roll_joint"]
COLLISION_LINKS = {   # config/pr2_both_arms_stomp_config.yaml:3-19 (right arm), in kinematic order
    "r_upper_arm_link": {"link_radius": 0.10}, "r_forearm_link": {"link_radius": 0.065}, "r_gripper_palm_link": {"link_radius": 0.06},
    "r_gripper_l_finger_link": {"link_radius": 0.03, "link_extension": 0.01},
    "r_gripper_l_finger_tip_link": {"link_radius": 0.03, "link_extension": 0.01},
    "r_gripper_r_finger_link": {"link_radius": 0.03, "link_extension": 0.01},
    "r_gripper_r_finger_tip_link": {"link_radius": 0.03, "link_extension": 0.01},
}
STATE = {"torso_lift_joint": 0.1, "r_gripper_l_finger_joint": 0.25, "r_gripper_l_finger_tip_joint": 0.25,
         "r_gripper_r_finger_joint": 0.25, "r_gripper_r_finger_tip_joint": 0.25}


def test_urdf_reproduces_the_synthetic_arm():
    got = robot_from_urdf(URDF, GROUP, "base_link", COLLISION_LINKS, 0.07, STATE)
    want = scenes.pr2_right_arm()
    assert [g["name"] for g in got.segments] == [g["name"] for g in want.segments]      # DFS pre-order numbering
    for a, b in zip(got.segments, want.segments):
        assert (a["parent"], a["type"], a["group"]) == (b["parent"], b["type"], b["group"]), a["name"]
        np.testing.assert_allclose(a["rot"], b["rot"], atol=1e-15)
        np.testing.assert_allclose(a["pos"], b["pos"], atol=0)
        if a["type"] != _abi.JOINT_FIXED:
            np.testing.assert_allclose(a["axis"], b["axis"], atol=1e-15)
        assert a["fixed"] == b["fixed"]
    assert got.limits == want.limits and got.reference_segment == want.reference_segment
    assert len(got.spheres) == len(want.spheres) == 47
    for a, b in zip(got.spheres, want.spheres):
        assert a["segment"] == b["segment"] and a["radius"] == b["radius"] and a["clearance"] == b["clearance"]
        np.testing.assert_allclose(a["pos"], b["pos"], atol=0)


def test_urdf_errors():
    with pytest.raises(ValueError, match="not in the URDF"):
        robot_from_urdf(URDF, ["nope"], "base_link")
    with pytest.raises(ValueError, match="reference frame"):
        robot_from_urdf(URDF, GROUP, "nope")
    with pytest.raises(ValueError, match="one root"):
        robot_from_urdf("<robot><link name='a'/><link name='b'/></robot>", [], "a")


def test_urdf_rotated_axis_is_expressed_in_the_parent_frame():
    urdf = """<robot name="r"><link name="a"/><link name="b"/>
      <joint name="j" type="revolute"><parent link="a"/><child link="b"/><origin xyz="0 0 1" rpy="0 0 1.5707963267948966"/>
      <axis xyz="1 0 0"/><limit lower="-1" upper="1"/></joint></robot>"""
    rb = robot_from_urdf(urdf, ["j"], "a")
    np.testing.assert_allclose(rb.segments[1]["axis"], (0.0, 1.0, 0.0), atol=1e-15)   # R_pj * (1,0,0)


def test_attached_object_bounding_spheres():
    rb = scenes.pr2_right_arm()
    palm = [g["name"] for g in rb.segments].index("r_gripper_palm_link")
    n0 = len(rb.spheres)
    rb.add_attached_object(palm, "cylinder", (0.02, 0.3), (0.15, 0.0, 0.0), padding=0.01)
    rb.add_attached_object(palm, "box", (0.1, 0.2, 0.3), (0.2, 0.0, 0.0))
    rb.add_attached_object(palm, "sphere", (0.05,), (0.0, 0.1, 0.0))
    assert len(rb.spheres) == n0 + 3
    np.testing.assert_allclose([s["radius"] for s in rb.spheres[n0:]],
                               [np.hypot(0.03, 0.16), np.sqrt(0.05 ** 2 + 0.1 ** 2 + 0.15 ** 2), 0.05])
    assert all(s["segment"] == palm for s in rb.spheres[n0:])


def test_min_jerk_initial_trajectory():
    start, goal = np.array([0.3, -1.0]), np.array([1.1, 0.4])
    N, dt = 99, 0.05
    traj = scenes.fill_in_min_jerk(start, goal, N, dt)
    full = np.concatenate([start[:, None], traj, goal[:, None]], axis=1)
    s = np.arange(N + 2) / (N + 1)
    want = start[:, None] + (goal - start)[:, None] * (10 * s ** 3 - 15 * s ** 4 + 6 * s ** 5)     # the min-jerk polynomial
    np.testing.assert_allclose(full, want, rtol=1e-12, atol=1e-14)
    vel = np.diff(full, axis=1) / dt
    assert np.abs(vel[:, 0]).max() < 1e-3 and np.abs(vel[:, -1]).max() < 1e-3      # zero boundary velocity


def test_urdf_inertials_and_dynamics_chain():
    """<inertial> blocks -> the per-segment KDL::RigidBodyInertia table of stomp_engine_set_dynamics (mass, centre of mass and
    inertia about it in the LINK frame: the <origin> rotation turns the tensor), and the chain by link names."""
    urdf = URDF.replace('<link name="r_upper_arm_roll_link"/>', '<link name="r_upper_arm_roll_link"><inertial><mass value="6.0"/>'
                        '<origin xyz="0.2 0 0" rpy="0 0 1.5707963267948966"/>'
                        '<inertia ixx="0.01" iyy="0.08" izz="0.07" ixy="0" ixz="0" iyz="0"/></inertial></link>')
    urdf = urdf.replace('<link name="r_forearm_roll_link"/>', '<link name="r_forearm_roll_link"><inertial><mass value="2.5"/>'
                        '<inertia ixx="0.014" iyy="0.016" izz="0.015" ixy="0.001" ixz="0" iyz="0"/></inertial></link>')
    rb = robot_from_urdf(urdf, GROUP, "base_link", COLLISION_LINKS, 0.07, STATE, dynamics_chain=("torso_lift_link", "r_gripper_palm_link"))
    names = [g["name"] for g in rb.segments]
    up, fr = names.index("r_upper_arm_roll_link"), names.index("r_forearm_roll_link")
    assert set(rb.inertias) == {up, fr}
    m, com, ic = rb.inertias[up]
    assert m == 6.0 and com == (0.2, 0.0, 0.0)
    np.testing.assert_allclose(ic, (0.08, 0.01, 0.07, 0.0, 0.0, 0.0), atol=1e-15)     # a quarter turn about z swaps ixx and iyy
    np.testing.assert_allclose(rb.inertias[fr][2], (0.014, 0.016, 0.015, 0.001, 0.0, 0.0), atol=0)
    assert rb.chain == (names.index("torso_lift_link"), names.index("r_gripper_palm_link"))
    arr = rb.c_inertias()
    assert arr[up].mass == 6.0 and arr[names.index("r_elbow_flex_link")].mass == 0.0
    with pytest.raises(ValueError):
        robot_from_urdf(urdf, GROUP, "base_link", COLLISION_LINKS, 0.07, STATE, dynamics_chain=("torso_lift_link", "nope"))
    # the oracle accepts the table (chain joints = the group joints in order) and gravity loads the shoulder
    from oracle.oracle import Oracle
    sc = scenes.make_scenario("tiny", num_problems=1)
    sc.robot = rb
    o = Oracle(sc, 0)
    o.set_dynamics(1.0)
    o.execute(o.get_parameters(), 2)
    assert np.abs(o.last_torques()).max() > 1.0
