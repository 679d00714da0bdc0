"""The inverse-dynamics (torque) term of StompOptimizer::execute (src/stomp_optimizer.cpp:1006-1061,1117-1142).

The reference calls KDL::ChainIdSolver_RNE, which it does not vendor; the oracle, the engine and the stand-in header that the
compiled reference links against (oracle/ref_shim/kdl/chainidsolver_recursive_newton_euler.hpp) all restate that published
algorithm.  So the PHYSICS is checked here against a formulation that shares nothing with recursive Newton-Euler: the
joint-space Lagrangian  tau = M(q) qdd + Mdot qd - 1/2 d(qd^T M qd)/dq + dV/dq  with M assembled from link Jacobians in NumPy
and its derivatives taken by central differences.  Parity of the reference's own part (finite-difference joint rates, the
sum of |tau_j|, the weights) is pinned by tests/test_reference_pinning.py against the compiled StompOptimizer.
"""
import numpy as np

from oracle import oracle
from stomp_motion_planner_icra2011_b200 import _abi, scenes


def _chain(rb):
    path, s = [], rb.chain[1]
    while s != rb.chain[0]:
        path.append(s)
        s = rb.segments[s]["parent"]
    return path[::-1]


def _rot(axis, angle):
    a = np.asarray(axis, float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def _mass_matrix_and_potential(rb, q, gravity):
    """M(q) [D][D] and V(q) of the chain, in the frame of the chain's root segment."""
    D = rb.num_dimensions
    R, p = np.eye(3), np.zeros(3)
    joints = []           # (group index, type, world axis, world origin)
    M, V = np.zeros((D, D)), 0.0
    for s in _chain(rb):
        g = rb.segments[s]
        Rj, pos, axis = np.array(g["rot"]).reshape(3, 3), np.array(g["pos"]), np.array(g["axis"])
        val = q[g["group"]] if g["group"] >= 0 else 0.0
        origin_w, axis_w = p + R @ pos, R @ axis
        if g["type"] == _abi.JOINT_REVOLUTE:
            R, p = R @ _rot(axis, val) @ Rj, origin_w
        elif g["type"] == _abi.JOINT_PRISMATIC:
            R, p = R @ Rj, origin_w + val * axis_w
        else:
            R, p = R @ Rj, origin_w
        if g["group"] >= 0:
            joints.append((g["group"], g["type"], axis_w, origin_w))
        if s in rb.inertias:
            m, com, ic = rb.inertias[s]
            pc = p + R @ np.asarray(com)
            Ic = np.array([[ic[0], ic[3], ic[4]], [ic[3], ic[1], ic[5]], [ic[4], ic[5], ic[2]]])
            Iw = R @ Ic @ R.T
            Jv, Jw = np.zeros((3, D)), np.zeros((3, D))
            for j, jt, aw, ow in joints:
                if jt == _abi.JOINT_REVOLUTE:
                    Jv[:, j], Jw[:, j] = np.cross(aw, pc - ow), aw
                else:
                    Jv[:, j] = aw
            M += m * Jv.T @ Jv + Jw.T @ Iw @ Jw
            V -= m * np.dot(gravity, pc)
    return M, V


def _lagrangian_torques(rb, q, qd, qdd, gravity, h=1e-5):
    D = len(q)
    M, _ = _mass_matrix_and_potential(rb, q, gravity)
    dM, dV = np.zeros((D, D, D)), np.zeros(D)
    for k in range(D):
        e = np.zeros(D)
        e[k] = h
        Mp, Vp = _mass_matrix_and_potential(rb, q + e, gravity)
        Mm, Vm = _mass_matrix_and_potential(rb, q - e, gravity)
        dM[k], dV[k] = (Mp - Mm) / (2 * h), (Vp - Vm) / (2 * h)
    Mdot = np.einsum("kij,k->ij", dM, qd)
    return M @ qdd + Mdot @ qd - 0.5 * np.einsum("kij,i,j->k", dM, qd, qd) + dV


def test_recursive_newton_euler_matches_the_lagrangian():
    sc = scenes.make_scenario("tiny", num_problems=1)
    rb = sc.robot
    rb.limits = [(0, 0.0, 0.0)] * rb.num_dimensions             # no joint-limit projection: the trajectory below is what is costed
    D, N, dt = rb.num_dimensions, sc.num_time_steps, sc.discretization
    rng = np.random.default_rng(3)
    q0, v0, a0 = rng.uniform(-1, 1, D), rng.uniform(-2, 2, D), rng.uniform(-6, 6, D)
    tt = np.arange(N) * dt
    params = q0[:, None] + v0[:, None] * tt + 0.5 * a0[:, None] * tt ** 2      # quadratic in time: the 7-tap rules are exact
    gravity = np.array([0.3, -0.2, -9.8])
    o = oracle.Oracle(sc, 0)
    w = 0.01
    o.set_dynamics(w, gravity)
    costs, _ = o.execute(params, 2)
    plain, _ = oracle.Oracle(sc, 0).execute(params, 2)
    tau = o.last_torques()
    for t in range(3, N - 3):                                    # points whose stencil stays off the constant padding
        want = _lagrangian_torques(rb, params[:, t], v0 + a0 * tt[t], a0, gravity)
        np.testing.assert_allclose(tau[t], want, rtol=2e-6, atol=2e-6 * np.abs(want).max())
        np.testing.assert_allclose(costs[0, t] - plain[0, t], w * np.abs(want).sum(), rtol=1e-5)
    assert np.abs(tau).max() > 5.0                               # gravity alone loads the shoulder with tens of N m


def test_static_arm_carries_only_gravity():
    """zero rates: tau = dV/dq, and a gravity-free arm at rest needs no torque."""
    sc = scenes.make_scenario("tiny", num_problems=1)
    rb = sc.robot
    rb.limits = [(0, 0.0, 0.0)] * rb.num_dimensions
    q = np.array([0.3, 0.4, -0.5, -1.0, 0.2, -0.7, 0.1])
    sc.start, sc.goal = q[None], q[None]
    params = np.repeat(q[:, None], sc.num_time_steps, axis=1)
    o = oracle.Oracle(sc, 0)
    o.set_dynamics(1.0, (0.0, 0.0, -9.8))
    o.execute(params, 2)
    want = _lagrangian_torques(rb, q, np.zeros(7), np.zeros(7), np.array([0.0, 0.0, -9.8]))
    np.testing.assert_allclose(o.last_torques()[5], want, rtol=1e-6, atol=1e-7)
    o.set_dynamics(1.0, (0.0, 0.0, 0.0))
    o.execute(params, 2)
    assert np.abs(o.last_torques()).max() < 1e-9


def test_chain_must_be_the_group_joints_in_order():
    import pytest
    sc = scenes.make_scenario("tiny", num_problems=1)
    o = oracle.Oracle(sc, 0)
    sc.robot.chain = (sc.robot.chain[0], sc.robot.chain[1] - 3)      # stops short of the wrist: 5 of the 7 group joints
    with pytest.raises(RuntimeError):
        o.set_dynamics(1.0)
