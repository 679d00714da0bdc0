"""URDF -> joint table + collision spheres: the part of StompRobotModel the engine's inputs come from
(src/stomp_robot_model.cpp:58-226,265-306), without ROS / urdf / kdl_parser (none exist here).

kdl_parser semantics restated: every URDF joint becomes one KDL segment named after its child
link, `pose(q) = Frame(Rot(R_pj a, q) R_pj, p_pj)` with (R_pj, p_pj) the joint <origin> and `a` the joint <axis>; segment
numbers are the DFS pre-order over children in document order (TreeFkSolverJointPosAxisPartial::assignSegmentNumber,
src/treefksolverjointposaxis_partial.cpp:180-191).  Continuous joints have no limits (src/stomp_robot_model.cpp:160-161).
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET

import numpy as np

from . import _abi
from .scenes import Robot


def _floats(text, n, default):
    if text is None:
        return list(default)
    vals = [float(v) for v in text.split()]
    if len(vals) != n:
        raise ValueError("expected %d numbers, got %r" % (n, text))
    return vals


def _rpy_matrix(r, p, y):
    """urdf::Rotation::setFromRPY -> KDL::Rotation (fixed-axis roll, pitch, yaw)."""
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return (cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
            sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
            -sp, cp * sr, cp * cr)


_TYPES = {"revolute": _abi.JOINT_REVOLUTE, "continuous": _abi.JOINT_REVOLUTE, "prismatic": _abi.JOINT_PRISMATIC,
          "fixed": _abi.JOINT_FIXED, "floating": _abi.JOINT_FIXED, "planar": _abi.JOINT_FIXED}


def _link_inertia(link_el):
    """<inertial> of a URDF link -> (mass, centre of mass, (ixx, iyy, izz, ixy, ixz, iyz)) in the LINK frame, the
    KDL::RigidBodyInertia kdl_parser attaches to the link's segment: the <origin> places the inertial frame, its rotation turns
    the inertia tensor (given about the centre of mass in the inertial frame's axes) into the link's axes."""
    el = link_el.find("inertial")
    if el is None or el.find("mass") is None:
        return None
    mass = float(el.find("mass").get("value", 0.0))
    origin = el.find("origin")
    xyz = _floats(origin.get("xyz") if origin is not None else None, 3, (0, 0, 0))
    rpy = _floats(origin.get("rpy") if origin is not None else None, 3, (0, 0, 0))
    it = el.find("inertia")
    g = (lambda k: float(it.get(k, 0.0))) if it is not None else (lambda k: 0.0)
    Ic = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
    R = np.asarray(_rpy_matrix(*rpy)).reshape(3, 3)
    Il = R @ Ic @ R.T
    return mass, tuple(xyz), (Il[0, 0], Il[1, 1], Il[2, 2], Il[0, 1], Il[0, 2], Il[1, 2])


def robot_from_urdf(urdf_xml: str, group_joints, reference_frame: str, collision_links=None, collision_clearance=0.07,
                    joint_state=None, dynamics_chain=None) -> Robot:
    """group_joints: ordered joint names of the planning group (planning_groups.yaml);
    collision_links: {link_name: {"link_radius": r, "link_extension": e}} (config/pr2_both_arms_stomp_config.yaml:3-33);
    joint_state: {joint_name: value} for joints outside the group (robot start state);
    dynamics_chain: (root link, tip link) of the inverse-dynamics chain (the reference hard-codes the PR2's
    ("torso_lift_link", "r_gripper_tool_frame"), src/stomp_robot_model.cpp:183); link <inertial> blocks fill Robot.inertias."""
    root = ET.fromstring(urdf_xml)
    links = [l.get("name") for l in root.findall("link")]
    link_inertia = {l.get("name"): _link_inertia(l) for l in root.findall("link")}
    joints = []
    for j in root.findall("joint"):
        origin = j.find("origin")
        xyz = _floats(origin.get("xyz") if origin is not None else None, 3, (0, 0, 0))
        rpy = _floats(origin.get("rpy") if origin is not None else None, 3, (0, 0, 0))
        axis_el = j.find("axis")
        axis = _floats(axis_el.get("xyz") if axis_el is not None else None, 3, (1, 0, 0))
        limit = j.find("limit")
        jtype = j.get("type")
        if jtype not in _TYPES:
            raise ValueError("unknown joint type %r" % jtype)
        joints.append(dict(name=j.get("name"), type=jtype, parent=j.find("parent").get("link"), child=j.find("child").get("link"),
                           xyz=xyz, rot=_rpy_matrix(*rpy), axis=axis,
                           lower=float(limit.get("lower", 0.0)) if limit is not None else 0.0,
                           upper=float(limit.get("upper", 0.0)) if limit is not None else 0.0))
    children_of = {}
    child_links = set()
    for j in joints:
        children_of.setdefault(j["parent"], []).append(j)
        child_links.add(j["child"])
    roots = [l for l in links if l not in child_links]
    if len(roots) != 1:
        raise ValueError("URDF must have exactly one root link, found %r" % roots)
    group_index = {name: i for i, name in enumerate(group_joints)}
    joint_state = joint_state or {}
    rb = Robot()
    seg_of_link = {}
    joint_by_name = {j["name"]: j for j in joints}
    missing = [n for n in group_joints if n not in joint_by_name]
    if missing:
        raise ValueError("planning group joints not in the URDF: %r" % missing)

    def add(link, parent_seg, joint):
        if joint is None:      # KDL tree root segment
            seg = rb.add_segment(link, -1, _abi.JOINT_FIXED, (0, 0, 0))
        else:
            a = np.asarray(joint["axis"], float)
            R = np.asarray(joint["rot"]).reshape(3, 3)
            a = R @ (a / (np.linalg.norm(a) or 1.0))        # kdl_parser: axis expressed in the parent frame
            jt = _TYPES[joint["type"]]
            seg = rb.add_segment(link, parent_seg, jt, joint["xyz"], a if jt != _abi.JOINT_FIXED else (0, 0, 1), rot=joint["rot"],
                                 group=group_index.get(joint["name"], -1), fixed=float(joint_state.get(joint["name"], 0.0)))
            rb.segments[seg]["joint"] = joint["name"]
        seg_of_link[link] = seg
        for cj in children_of.get(link, []):
            add(cj["child"], seg, cj)

    add(roots[0], -1, None)
    if reference_frame not in seg_of_link:
        raise ValueError("reference frame %r is not a link of the URDF" % reference_frame)
    rb.reference_segment = seg_of_link[reference_frame]
    rb.limits = []
    for name in group_joints:
        j = joint_by_name[name]
        if j["type"] == "continuous":
            rb.limits.append((0, 0.0, 0.0))
        else:
            rb.limits.append((1, j["lower"], j["upper"]))
    for link, cfg in (collision_links or {}).items():
        if link not in seg_of_link:
            continue                                         # links outside the model are ignored like in the reference
        # collision_links/<link>/{link_radius, link_clearance, link_extension} (src/stomp_robot_model.cpp:361-373)
        rb.add_link_spheres(seg_of_link[link], float(cfg["link_radius"]), float(cfg.get("link_clearance", collision_clearance)),
                            float(cfg.get("link_extension", 0.0)))
    rb.inertias = {seg_of_link[name]: val for name, val in link_inertia.items() if val is not None and name in seg_of_link}
    if dynamics_chain is not None:
        for name in dynamics_chain:
            if name not in seg_of_link:
                raise ValueError("dynamics chain link %r is not a link of the URDF" % name)
        rb.chain = (seg_of_link[dynamics_chain[0]], seg_of_link[dynamics_chain[1]])
    # a planning group only keeps the points some group joint moves (StompPlanningGroup::addCollisionPoint, :308-334)
    rb.spheres = [s for s in rb.spheres if rb.moved_by_group(s["segment"])]
    return rb
