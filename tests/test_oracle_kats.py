"""CPU tests of the oracle: known answers and invariants derivable from the reference's formulas
 plus the committed golden vectors.  The reference has no unit tests of its own."""
import math
import os

import numpy as np
import pytest

from oracle.oracle import Oracle
from stomp_motion_planner_icra2011_b200 import _abi, scenes
from tests.helpers import correlated_noise

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_iterations.npz")


@pytest.fixture(scope="module")
def yaml_exact():
    """N=99, dt=0.05, acceleration cost only, ridge 0: config/params.yaml."""
    sc = scenes.make_scenario("C1", num_problems=1, num_time_steps=99)
    return sc, Oracle(sc, 0)


def test_control_cost_matrix_kat(yaml_exact):
    sc, o = yaml_exact
    R = o.get(_abi.FIELD_CONTROL_COST)
    assert np.allclose(R, R.T, rtol=1e-14)
    i, j = np.nonzero(R)
    assert np.abs(i - j).max() == 4                       # 9-banded
    dt = 0.05
    interior = np.array([1.0, -32.0, 316.0, -992.0, 1414.0, -992.0, 316.0, -32.0, 1.0]) / (144.0 * dt ** 4)
    np.testing.assert_allclose(R[50, 46:55], interior, rtol=1e-12)
    assert abs(R[50, 50] - 1.57111e6) / 1.57111e6 < 1e-5
    assert abs(np.linalg.cond(R) - 5.84e6) / 5.84e6 < 0.01


def test_inverse_and_projection_kat(yaml_exact):
    sc, o = yaml_exact
    R, Rinv, M = o.get(_abi.FIELD_CONTROL_COST), o.get(_abi.FIELD_INV_CONTROL_COST), o.get(_abi.FIELD_PROJECTION)
    np.testing.assert_allclose(R @ Rinv, np.eye(99), atol=1e-6)   # cond(R) ~ 5.8e6
    assert np.unravel_index(Rinv.argmax(), Rinv.shape) == (49, 49)
    assert abs(Rinv.max() - 0.0332248) < 1e-6
    # every column of M has maximum 1/N (src/policy_improvement.cpp:430-436)
    np.testing.assert_allclose(M.max(axis=0), 1.0 / 99, rtol=1e-12)
    L = o.get(_abi.FIELD_NOISE_CHOLESKY)
    assert np.allclose(np.triu(L, 1), 0.0)
    np.testing.assert_allclose(L @ L.T, Rinv, rtol=1e-9, atol=1e-14)
    # noise sigma at mid trajectory with noise_stddev = 2
    assert abs(2.0 * np.sqrt(Rinv[49, 49]) - 0.3646) < 1e-3


def test_quad_cost_inverse_scaled(yaml_exact):
    sc, o = yaml_exact
    Q = o.get(_abi.FIELD_QUAD_COST_INV)
    assert abs(Q.max() - 1.0) < 1e-12                     # scaled by the global max (stomp_optimizer.cpp:119-125)
    assert np.allclose(Q, Q.T, rtol=1e-6, atol=1e-12)


def test_min_control_cost_is_stationary(yaml_exact):
    sc, o = yaml_exact
    theta = o.get_parameters()
    D, N = theta.shape
    # rebuild R_all from the stencil definition and check the gradient on the free block is ~0
    dt = sc.movement_duration / (N + 1)
    rule = np.array([0, -1 / 12.0, 16 / 12.0, -30 / 12.0, 16 / 12.0, -1 / 12.0, 0]) / dt ** 2
    Nall = N + 12
    A = np.zeros((Nall, Nall))
    for i in range(Nall):
        for j in range(-3, 4):
            if 0 <= i + j < Nall:
                A[i, i + j] = rule[j + 3]
    Rall = A.T @ A
    for d in range(D):
        x = np.concatenate([np.full(6, sc.start[0, d]), theta[d], np.full(6, sc.goal[0, d])])
        g = (Rall @ x)[6:6 + N]
        assert np.abs(g).max() < 1e-6 * np.abs(Rall).max() * max(1.0, np.abs(x).max())


def test_constant_trajectory_has_zero_control_cost():
    sc = scenes.make_scenario("tiny", num_problems=1)
    sc.goal = sc.start.copy()
    o = Oracle(sc, 0)
    D, N = sc.robot.num_dimensions, sc.num_time_steps
    p = np.repeat(sc.start[0][:, None], N, axis=1)[None]
    c = o.compute_control_costs(p, np.zeros_like(p), 1.0)[0]
    assert np.abs(c[:, 1:-1]).max() < 1e-12
    # the first / last entries carry the folded padding rows, whose stencils lose their out-of-range taps
    # (rows 0 and 1 of the differentiation matrix do not sum to zero): a quirk of the reference that is kept
    # (src/covariant_trajectory_policy.cpp:213-221,245-250)
    dt = sc.movement_duration / (N + 1)
    edge = ((15.0 / 12.0) ** 2 + (1.0 / 12.0) ** 2) / dt ** 4
    np.testing.assert_allclose(c[:, 0], edge * sc.start[0] ** 2, rtol=1e-9)
    np.testing.assert_allclose(c[:, -1], edge * sc.start[0] ** 2, rtol=1e-9)
    p = p[None] if p.ndim == 2 else p
    np.testing.assert_allclose(o.get_parameters(), p[0], atol=1e-9)   # min-control-cost of start == goal is constant


def test_control_cost_matches_dense_formula():
    sc = scenes.make_scenario("tiny", num_problems=1)
    o = Oracle(sc, 0)
    D, N = sc.robot.num_dimensions, sc.num_time_steps
    rng = np.random.default_rng(0)
    p, e = rng.normal(size=(1, D, N)), rng.normal(size=(1, D, N))
    got = o.compute_control_costs(p, e, 0.25)[0]
    dt = sc.movement_duration / (N + 1)
    rule = np.array([0, -1 / 12.0, 16 / 12.0, -30 / 12.0, 16 / 12.0, -1 / 12.0, 0]) / dt ** 2
    for d in range(D):
        x = np.concatenate([np.full(6, sc.start[0, d]), p[0, d] + e[0, d], np.full(6, sc.goal[0, d])])
        acc = np.array([sum(rule[j + 3] * x[i + j] for j in range(-3, 4) if 0 <= i + j < N + 12) for i in range(N + 12)])
        c = 0.25 * acc ** 2
        want = c[6:6 + N].copy()
        want[0] += c[:6].sum()
        want[-1] += c[-6:].sum()
        np.testing.assert_allclose(got[d], want, rtol=1e-10)


def _iterate_with_taps(cumulative):
    sc = scenes.make_scenario("tiny", num_problems=1, use_cumulative_costs=cumulative)
    o = Oracle(sc, 0)
    rng = np.random.default_rng(1)
    L = o.get(_abi.FIELD_NOISE_CHOLESKY)
    eps = correlated_noise(L, rng, (sc.num_rollouts,), sc.noise_stddev)
    o.iterate(1, eps)
    return sc, o, eps


@pytest.mark.parametrize("cumulative", [0, 1])
def test_probabilities_and_cumulative_costs(cumulative):
    sc, o, eps = _iterate_with_taps(cumulative)
    P, C = o.get(_abi.FIELD_PROBABILITIES), o.get(_abi.FIELD_CUMULATIVE_COSTS)
    S, Cc = o.get(_abi.FIELD_STATE_COSTS), o.get(_abi.FIELD_CONTROL_COSTS)
    np.testing.assert_allclose(P.sum(axis=0), 1.0, rtol=1e-12)
    total = S[:, None, :] + Cc
    want = np.flip(np.cumsum(np.flip(total, -1), -1), -1) if cumulative else total
    np.testing.assert_allclose(C, want, rtol=1e-12)
    spread = C.max(axis=0) - C.min(axis=0)
    ratio = P.max(axis=0) / P.min(axis=0)
    ok = spread >= 1e-8
    np.testing.assert_allclose(ratio[ok], np.exp(10.0), rtol=1e-9)       # best/worst = e^10
    # update = M * sum_r P .* eps
    M = o.get(_abi.FIELD_PROJECTION)
    u = np.einsum("ij,dj->di", M, (P * o.get(_abi.FIELD_NOISE)).sum(axis=0))
    np.testing.assert_allclose(o.get(_abi.FIELD_UPDATES), u, rtol=1e-10, atol=1e-16)


def test_equal_costs_give_uniform_probabilities():
    sc = scenes.make_scenario("tiny", num_problems=1)
    sc.smoothness_cost_weight = 0.0
    o = Oracle(sc, 0)
    sigma = sc.noise_stddev
    o.get_rollouts(sigma, np.zeros((sc.num_rollouts, sc.robot.num_dimensions, sc.num_time_steps)))
    o.set_rollout_costs(np.ones((sc.num_rollouts, sc.num_time_steps)), 0.0)
    o.improve_policy()
    np.testing.assert_allclose(o.get(_abi.FIELD_PROBABILITIES), 1.0 / sc.num_rollouts, rtol=1e-12)


def test_rollout_reuse_keeps_the_cheapest():
    sc = scenes.make_scenario("tiny", num_problems=1)
    o = Oracle(sc, 0)
    R, Rre, D, N = sc.num_rollouts, sc.num_reused_rollouts, sc.robot.num_dimensions, sc.num_time_steps
    rng = np.random.default_rng(2)
    sigma = np.ones(D)
    ro1 = o.get_rollouts(sigma, rng.normal(size=(R, D, N)))
    assert ro1.shape[0] == R                                    # first iteration generates all R
    costs = rng.uniform(1.0, 2.0, size=(R, N))
    totals = o.set_rollout_costs(costs, 0.0)
    o.improve_policy()
    params1 = o.get(_abi.FIELD_PARAMETERS)
    o.add_extra_rollouts(np.full(N, 100.0))                     # expensive extra rollout: never reused
    ro2 = o.get_rollouts(sigma, rng.normal(size=(R - Rre, D, N)))
    assert ro2.shape[0] == R - Rre
    best = np.argsort(totals, kind="stable")[:Rre]
    params2 = o.get(_abi.FIELD_PARAMETERS)
    np.testing.assert_array_equal(params2[R - Rre:], params1[best])
    np.testing.assert_array_equal(o.get(_abi.FIELD_STATE_COSTS)[R - Rre:], costs[best])   # stale state costs are kept
    np.testing.assert_allclose(o.get(_abi.FIELD_NOISE)[R - Rre:], params1[best] - o.get_parameters(), rtol=0, atol=0)


def test_potential_is_continuous_and_three_piece():
    sc = scenes.make_scenario("tiny", num_problems=1)
    o = Oracle(sc, 0)
    dbg, _ = o.execute_debug(o.get_parameters())
    rb = sc.robot
    radius = np.array([s["radius"] for s in rb.spheres])
    clr = np.array([s["clearance"] for s in rb.spheres])
    tab = np.sqrt(np.arange(256.0)) * sc.sdf.resolution
    vox = dbg["voxel"]
    nx, ny, nz = sc.sdf.dims
    inside = np.all((vox >= 1) & (vox < np.array([nx, ny, nz]) - 1), axis=-1)
    dist = np.where(inside, tab[sc.sdf.voxels[np.clip(vox[..., 0], 0, nx - 1), np.clip(vox[..., 1], 0, ny - 1),
                                              np.clip(vox[..., 2], 0, nz - 1)]], 0.0)
    d = dist - radius
    want = np.where(d >= clr, 0.0, np.where(d >= 0.0, 0.5 * (d - clr) ** 2 / clr, -d + 0.5 * clr))
    np.testing.assert_allclose(dbg["potential"], want, rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(dbg["in_collision"], (dist <= radius).astype(np.int32))
    # continuity of the three pieces at d = 0 and d = clearance
    eps_, c = 1e-9, 0.07
    f = lambda d: 0.0 if d >= c else (0.5 * (d - c) ** 2 / c if d >= 0 else -d + 0.5 * c)
    assert abs(f(eps_) - f(-eps_)) < 1e-8 and abs(f(c - eps_) - f(c + eps_)) < 1e-8 and abs(f(0.0) - 0.5 * c) < 1e-15


def test_sphere_outside_grid_counts_as_collision():
    sc = scenes.make_scenario("tiny", num_problems=1)
    sc.sdf = scenes.bake_distance_field(size=(0.2, 0.2, 0.2), origin=(5.0, 5.0, 5.0), resolution=0.04)   # far away
    o = Oracle(sc, 0)
    costs, cf = o.execute(o.get_parameters()[None], iteration_number=2)
    assert cf[0] == 0 and costs.max() > 0.0


def test_voxel_index_rule():
    sc = scenes.make_scenario("tiny", num_problems=1)
    o = Oracle(sc, 0)
    dbg, _ = o.execute_debug(o.get_parameters())
    want = np.round((dbg["position"] - np.asarray(sc.sdf.origin)) / sc.sdf.resolution)
    want = np.where(np.abs(want) == 0, 0, want)       # np.round is half-even; positions never sit on .5 here
    np.testing.assert_array_equal(dbg["voxel"], want.astype(np.int32))


def test_fk_zero_configuration_is_product_of_fixed_transforms():
    sc = scenes.make_scenario("tiny", num_problems=1)
    rb = sc.robot
    sc.start = np.zeros((1, 7)); sc.goal = np.zeros((1, 7))
    rb.limits = [(0, 0.0, 0.0)] * 7
    o = Oracle(sc, 0)
    dbg, _ = o.execute_debug(np.zeros((7, sc.num_time_steps)))
    # at q = 0 every revolute pose is the pure parent->joint transform
    frames = []
    for g in rb.segments:
        Rm, p = np.array(g["rot"]).reshape(3, 3), np.array(g["pos"])
        if g["group"] < 0 and g["type"] == _abi.JOINT_REVOLUTE:
            a, q = np.array(g["axis"]), g["fixed"]
            K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            Rm = (np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * K @ K) @ Rm
        if g["group"] < 0 and g["type"] == _abi.JOINT_PRISMATIC:
            p = p + g["fixed"] * np.array(g["axis"])
        F = np.eye(4); F[:3, :3] = Rm; F[:3, 3] = p
        frames.append(F if g["parent"] < 0 else frames[g["parent"]] @ F)
    for j, s in enumerate(rb.spheres):
        want = (frames[s["segment"]] @ np.append(s["pos"], 1.0))[:3]
        np.testing.assert_allclose(dbg["position"][5, j], want, atol=1e-12)
    assert np.abs(dbg["vel_mag"]).max() < 1e-9           # nothing moves


def test_fd_velocity_exact_for_constant_joint_velocity():
    """a single prismatic joint moving at constant speed: the 4-tap rule is exact for linear motion."""
    rb = scenes.Robot()
    base = rb.add_segment("base", -1, _abi.JOINT_FIXED, (0, 0, 0))
    s = rb.add_segment("slide", base, _abi.JOINT_PRISMATIC, (0.1, 0.2, 0.3), (1, 0, 0), group=0)
    rb.spheres.append(dict(segment=s, radius=0.05, clearance=0.07, pos=(0.01, 0.02, 0.03)))
    rb.limits = [(0, 0.0, 0.0)]
    sc = scenes.make_scenario("tiny", num_problems=1)
    sc.robot = rb
    N = sc.num_time_steps
    speed = 0.37
    dtq = sc.discretization
    sc.start = np.array([[0.0]]); sc.goal = np.array([[speed * dtq * (N + 1)]])
    sc.noise_stddev = np.ones(1); sc.noise_decay = np.ones(1)
    o = Oracle(sc, 0)
    q = speed * dtq * np.arange(1, N + 1)[None]
    dbg, _ = o.execute_debug(q)
    # interior points only: the padding repeats start / goal, so the first and the last two steps see a kink
    np.testing.assert_allclose(dbg["vel_mag"][2:N - 1, 0], speed, rtol=1e-9)


def test_oracle_matches_committed_golden_vectors():
    g = np.load(GOLDEN)
    for cumulative in (0, 1):
        sc = scenes.make_scenario("tiny", num_problems=2, use_cumulative_costs=cumulative)
        ors = [Oracle(sc, b) for b in range(2)]
        for it in range(1, 5):
            key = "c%d_it%d_" % (cumulative, it)
            for b, o in enumerate(ors):
                cost, cf, _ = o.iterate(it, g[key + "eps"][b])
                np.testing.assert_allclose(cost, g[key + "noiseless_cost"][b], rtol=1e-9)
                assert cf == g[key + "collision_free"][b]
                np.testing.assert_allclose(o.get(_abi.FIELD_THETA), g[key + "theta"][b], rtol=1e-9, atol=1e-12)
                np.testing.assert_allclose(o.get(_abi.FIELD_PROBABILITIES), g[key + "probabilities"][b], rtol=1e-7, atol=1e-12)


def test_stomp_reduces_cost_on_the_shelf_scene():
    sc = scenes.make_scenario("C1", num_problems=1)
    o = Oracle(sc, 0)
    o.seed(1)
    first = o.iterate(1)[0]
    for it in range(2, 60):
        last = o.iterate(it)[0]
    assert last < first


def test_trilinear_extension_is_the_continuous_version_of_the_nearest_cell_field():
    """STOMP_SDF_TRILINEAR (engine extension): equals the reference's nearest-cell distance at cell centres, is the mean of two
    neighbouring cells half way between them, and is continuous where the nearest-cell field jumps."""
    from oracle.oracle import Oracle
    sc_n = scenes.make_scenario("tiny", num_problems=1)
    sc_t = scenes.make_scenario("tiny", num_problems=1)
    sc_t.sdf_mode = _abi.SDF_TRILINEAR
    rb = scenes.Robot()
    rb.add_segment("root", -1, _abi.JOINT_FIXED, (0, 0, 0))
    j = rb.add_segment("slide", 0, _abi.JOINT_PRISMATIC, (0.0, 0.0, 0.0), (1, 0, 0), group=0)
    rb.spheres = [dict(segment=j, radius=0.0, clearance=10.0, pos=(0.0, 0.0, 0.0))]   # potential = 0.5 (d - c)^2 / c: monotone in d
    rb.limits = [(0, 0.0, 0.0)]
    res, org = sc_n.sdf.resolution, np.asarray(sc_n.sdf.origin)
    cell = np.array([20, 12, 22])                          # inside the grid, next to the box of the tiny scene
    base = org + cell * res
    for sc in (sc_n, sc_t):
        sc.robot, sc.start, sc.goal = rb, np.zeros((1, 1)), np.zeros((1, 1))
    N = sc_n.num_time_steps

    def distances(sc, xs):
        """distance seen by a point sphere sliding along x through `base` (segment frame placed there by the prismatic joint)"""
        rb.segments[1]["pos"] = tuple(base)
        o = Oracle(sc, 0)
        params = np.zeros((1, N))
        params[0, :len(xs)] = xs
        dbg, _ = o.execute_debug(params)
        pot = dbg["potential"][1:1 + len(xs), 0]
        return 10.0 - np.sqrt(2.0 * 10.0 * pot)            # invert the middle branch of the potential

    xs = np.array([0.0, 0.25, 0.5, 0.75, 1.0]) * res
    dn, dt = distances(sc_n, xs), distances(sc_t, xs)
    np.testing.assert_allclose(dt[[0, 4]], dn[[0, 4]], rtol=0, atol=1e-12)           # cell centres: identical
    np.testing.assert_allclose(dt[2], 0.5 * (dn[0] + dn[4]), rtol=0, atol=1e-12)     # half way: the mean
    np.testing.assert_allclose(dt[1], 0.75 * dn[0] + 0.25 * dn[4], rtol=0, atol=1e-12)
    assert dn[1] == dn[0] and dn[3] == dn[4]                                         # the reference field is piecewise constant
    assert abs(dn[0] - dn[4]) > 1e-3, (dn, "pick a cell pair across which the field changes")


def test_body_voxelisation_restatement_is_strict_containment():
    """oracle/sdf_builder.py, bodies: the marked cells of a sphere / box / cylinder body are exactly the cells of the
    bounding-sphere-centred lattice points inside the scaled-then-padded primitive (StompCollisionSpace::getVoxelsInBody,
    src/stomp_collision_space.cpp:590-650): volume and extent checks against the analytic shapes."""
    from oracle import sdf_builder
    res = 0.02
    spec = dict(size=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), resolution=res, max_distance=0.1)
    ident = (0.0, 0.0, 0.0, 1.0)
    cases = [((0, (0.2,), (0.5, 0.5, 0.5), ident, 1.0, 0.01), 4.0 / 3.0 * math.pi * 0.21 ** 3),
             ((1, (0.3, 0.2, 0.4), (0.5, 0.5, 0.5), ident, 1.0, 0.0), 0.3 * 0.2 * 0.4),
             ((2, (0.1, 0.5), (0.5, 0.5, 0.5), ident, 1.2, 0.0), math.pi * 0.12 ** 2 * 0.6)]
    for body, volume in cases:
        _, occ = sdf_builder.build(bodies=[body], **spec)
        assert abs(occ.sum() * res ** 3 - volume) < 0.12 * volume
    # a box rotated by 90 degrees about z swaps its x / y extents
    q90 = (0.0, 0.0, math.sin(math.pi / 4), math.cos(math.pi / 4))
    _, occ = sdf_builder.build(bodies=[(1, (0.4, 0.1, 0.1), (0.5, 0.5, 0.5), q90, 1.0, 0.0)], **spec)
    ext = [np.ptp(np.nonzero(occ.any(axis=tuple(a for a in range(3) if a != k)))[0]) for k in range(3)]
    assert ext[1] > 3 * ext[0] and abs(ext[1] * res - 0.4) < 3 * res


def test_mesh_body_voxelisation_restatement():
    """oracle/sdf_builder.py, meshes (bodies::ConvexMesh + getVoxelsInBody's +z ray parity): a box given as a mesh of its
    eight corners (plus interior points the hull must drop) marks exactly the cells of the box primitive, posed and padded
    alike; an icosahedron's volume; a ray through a shared edge or a vertex is counted once (lattice-aligned octahedron)."""
    from oracle import sdf_builder
    res = 0.02
    spec = dict(size=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), resolution=res, max_distance=0.1)
    tilt = (math.sin(0.35) * 0.6, math.sin(0.35) * 0.8, 0.0, math.cos(0.35))
    half = np.array([0.15, 0.1, 0.2])
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float) * half
    inner = np.random.default_rng(2).uniform(-0.9, 0.9, (20, 3)) * half
    pos = (0.503, 0.497, 0.501)           # off the lattice: no point lies exactly on a face
    _, occ_mesh = sdf_builder.build(meshes=[(np.vstack([corners, inner]), pos, tilt, 1.0, 0.0)], **spec)
    _, occ_box = sdf_builder.build(bodies=[(1, tuple(2 * half), pos, tilt, 1.0, 0.0)], **spec)
    assert occ_box.sum() > 1000
    np.testing.assert_array_equal(occ_mesh, occ_box)
    # scale + padding move the corners along their rays from the centre: a cube stays a cube, half edge h * scale + padding / sqrt(3)
    h, scale, pad = 0.12, 1.2, 0.02
    cube = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float) * h
    _, occ_mesh = sdf_builder.build(meshes=[(cube, pos, tilt, scale, pad)], **spec)
    he = h * scale + pad / math.sqrt(3.0)
    _, occ_box = sdf_builder.build(bodies=[(1, (2 * he,) * 3, pos, tilt, 1.0, 0.0)], **spec)
    assert np.count_nonzero(occ_mesh != occ_box) <= 0.002 * occ_box.sum()      # the two lattices coincide; faces differ in rounding
    # icosahedron: volume 5 (3 + sqrt 5) / 12 a^3 with edge a = 2 (vertices (0, +-1, +-phi) ...), scaled to circumradius 0.2
    phi = (1 + math.sqrt(5)) / 2
    ico = np.array([(0, s1, s2 * phi) for s1 in (-1, 1) for s2 in (-1, 1)] + [(s1, s2 * phi, 0) for s1 in (-1, 1) for s2 in (-1, 1)] +
                   [(s2 * phi, 0, s1) for s1 in (-1, 1) for s2 in (-1, 1)], float)
    k = 0.2 / math.sqrt(1 + phi * phi)
    _, occ = sdf_builder.build(meshes=[(ico * k, pos, tilt, 1.0, 0.0)], **spec)
    volume = 5 * (3 + math.sqrt(5)) / 12 * (2 * k) ** 3
    assert abs(occ.sum() * res ** 3 - volume) < 0.05 * volume
    # octahedron centred ON a lattice point with lattice-aligned axes: rays pass through vertices and edges; every interior
    # lattice point |x| + |y| + |z| < r is marked exactly once, none outside
    r = 5 * res
    octa = np.array([(r, 0, 0), (-r, 0, 0), (0, r, 0), (0, -r, 0), (0, 0, r), (0, 0, -r)], float)
    c = (0.5, 0.5, 0.5)
    _, occ = sdf_builder.build(meshes=[(octa, c, (0.0, 0.0, 0.0, 1.0), 1.0, 0.0)], **spec)
    g = np.arange(50)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    l1 = np.abs(X - 25) + np.abs(Y - 25) + np.abs(Z - 25)
    assert occ[l1 < 5].all() and not occ[l1 > 5].any()


def test_upstream_propagation_differs_from_the_exact_transform_only_marginally():
    """The engine's distance-field rebuild computes the exact capped squared Euclidean transform; the reference propagates with
    distance_field::PropagationDistanceField (un-vendored), whose direction-restricted neighbourhoods make it a propagation, not
    an exact transform.  Its published algorithm restated (oracle/stomp_oracle.cpp, stomp_oracle_propagate_distance_field)
    against the exact transform on the benchmark scene (shelf + pole, 133 x 200 x 146 cells @ 15 mm, cap 12 cells) and on
    random clutter: never smaller, different in ~2e-5 of the cells, by at most 2 in the squared cell distance (2.6 mm)."""
    from oracle import sdf_builder
    ident = (0.0, 0.0, 0.0, 1.0)
    boxes = [(c, ident, d) for (c, d) in scenes.SHELF["boxes"]]
    cyls = [(c, ident, r, h) for (c, r, h) in scenes.POLE["cylinders"]]
    spec = dict(size=(2.0, 3.0, 2.2), origin=(-0.5, -1.5, -0.3), resolution=0.015, max_distance=0.17)
    exact, occ = sdf_builder.build(boxes=boxes, cylinders=cyls, **spec)
    cap = int(math.ceil(0.17 / 0.015))
    prop = sdf_builder.propagation_field(occ, cap)
    diff = prop.astype(np.int64) - exact.astype(np.int64)
    assert diff.min() == 0 and diff.max() <= 2
    assert 0 < np.count_nonzero(diff) <= 1e-4 * diff.size
    assert np.array_equal(prop == 0, occ)
    # dense random clutter: many obstacle fronts meet, which is where propagations lose exactness
    rng = np.random.default_rng(5)
    occ2 = rng.random((48, 48, 48)) < 0.004
    exact2 = np.minimum(np.rint(__import__("scipy.ndimage").ndimage.distance_transform_edt(~occ2) ** 2), 100).astype(np.int64)
    prop2 = sdf_builder.propagation_field(occ2, 10)
    d2 = prop2 - exact2
    assert d2.min() == 0 and np.count_nonzero(d2) < 0.02 * d2.size and d2.max() <= 6
