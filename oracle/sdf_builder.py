"""CPU restatement of the reference's distance-field construction for boxes and cylinders (TEST INFRASTRUCTURE ONLY).

Follows StompCollisionSpace::addCollisionObjectsToPoints (src/stomp_collision_space.cpp:238-293): lattice loops with
accumulated `x += resolution_`, point = pose * (position - lattice point), then distance_field's worldToGrid
(int(round((p - origin)/res)), points outside the grid dropped) and an exact squared Euclidean distance transform capped at
ceil(max_distance/res)^2 (scipy.ndimage; PropagationDistanceField is not vendored)."""
import math

import numpy as np


def _lattice(low, extent, res):
    out = []
    x = low
    while x <= low + extent + res:
        out.append(x)
        x += res
    return np.array(out)


def _rotation(q):
    x, y, z, w = q
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    return np.array([[w2 + x2 - y2 - z2, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y],
                     [2 * x * y + 2 * w * z, w2 - x2 + y2 - z2, 2 * y * z - 2 * w * x],
                     [2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, w2 - x2 - y2 + z2]])


def _round_half_away(t):
    return np.where(t >= 0, np.floor(t + 0.5), np.ceil(t - 0.5))


def mesh_body_geometry(vertices, pos, quat, scale, padding):
    """bodies::ConvexMesh restated (geometric_shapes is not vendored; include/stomp_b200.h states the contract): convex hull of
    the vertices (qhull through SciPy, where the engine has its own incremental hull), mesh centre = mean of the hull's
    vertices, bounding radius = largest distance from it, every hull vertex moved along its ray from the centre by
    scale + padding / distance, then posed.  Returns (world-frame triangles [T][3][3], bounding-sphere centre, radius)."""
    from scipy.spatial import ConvexHull
    V = np.asarray(vertices, float).reshape(-1, 3)
    hull = ConvexHull(V)
    on = np.zeros(len(V), bool)
    on[hull.simplices.ravel()] = True
    mc = np.zeros(3)
    for v in V[on]:                 # summed in vertex order, like the engine
        mc = mc + v
    mc = mc / float(on.sum())
    d = V - mc
    l = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    radius = float(l[on].max())
    fact = scale + np.where(l > 0.0, padding / np.where(l > 0.0, l, 1.0), 0.0)
    sv = mc + d * fact[:, None]
    Rm = _rotation(quat)
    pos = np.asarray(pos, float)

    def world(b):
        return np.stack([(Rm[r, 0] * b[..., 0] + Rm[r, 1] * b[..., 1] + Rm[r, 2] * b[..., 2]) + pos[r] for r in range(3)], axis=-1)

    return world(sv)[hull.simplices], world(mc), radius * scale + padding


def _ray_parity_inside(W, tris):
    """+z ray crossings of the triangles, per point of W [M][3] (src/stomp_collision_space.cpp:625-646): xy projection with the
    half-open edge rule, then the sign of the height above the point.  Same individually rounded operations as
    k_sdf_mark_meshes."""
    def owns(dx, dy):
        return (dy > 0.0) | ((dy == 0.0) & (dx < 0.0))

    count = np.zeros(len(W), np.int64)
    for T in tris:
        a, b, c = T[0] - W, T[1] - W, T[2] - W
        area = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
        flip = area < 0.0
        b2 = np.where(flip[:, None], c, b)
        c2 = np.where(flip[:, None], b, c)
        b, c = b2, c2
        wa = b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]
        wb = c[:, 0] * a[:, 1] - c[:, 1] * a[:, 0]
        wc = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
        inside = ((wa > 0.0) | ((wa == 0.0) & owns(c[:, 0] - b[:, 0], c[:, 1] - b[:, 1]))) & \
                 ((wb > 0.0) | ((wb == 0.0) & owns(a[:, 0] - c[:, 0], a[:, 1] - c[:, 1]))) & \
                 ((wc > 0.0) | ((wc == 0.0) & owns(b[:, 0] - a[:, 0], b[:, 1] - a[:, 1])))
        h = (wa * a[:, 2] + wb * b[:, 2]) + wc * c[:, 2]
        count += (area != 0.0) & inside & (h > 0.0)
    return (count & 1) == 1


def build(size, origin, resolution, max_distance, boxes=(), cylinders=(), points=None, bodies=(), meshes=()):
    from scipy import ndimage
    n = [int(size[i] / resolution) for i in range(3)]
    occ = np.zeros(n, dtype=bool)
    origin = np.asarray(origin, float)

    def mark(pos, quat, xs, ys, zs, radius):
        pos = np.asarray(pos, float)
        X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
        keep = np.ones(X.shape, bool)
        if radius > 0:
            keep = np.sqrt(np.abs(pos[0] - X) ** 2 + np.abs(pos[1] - Y) ** 2) <= radius
        P = np.stack([pos[0] - X[keep], pos[1] - Y[keep], pos[2] - Z[keep]], axis=-1)
        W = P @ _rotation(quat).T + pos
        c = _round_half_away((W - origin) / resolution).astype(np.int64)
        ok = np.all((c >= 0) & (c < np.array(n)), axis=1)
        c = c[ok]
        occ[c[:, 0], c[:, 1], c[:, 2]] = True

    for (p, q, d) in boxes:
        mark(p, q, _lattice(p[0] - d[0] / 2.0, d[0], resolution), _lattice(p[1] - d[1] / 2.0, d[1], resolution),
             _lattice(p[2] - d[2] / 2.0, d[2], resolution), 0.0)
    for (p, q, r, h) in cylinders:
        mark(p, q, _lattice(p[0] - r, r * 2.0, resolution), _lattice(p[1] - r, r * 2.0, resolution),
             _lattice(p[2] - h / 2.0, h, resolution), r)
    if points is not None and len(points):       # the "points" namespace: each collision-map point occupies its cell
        c = _round_half_away((np.asarray(points, float).reshape(-1, 3) - origin) / resolution).astype(np.int64)
        c = c[np.all((c >= 0) & (c < np.array(n)), axis=1)]
        occ[c[:, 0], c[:, 1], c[:, 2]] = True
    # robot / primitive bodies: StompCollisionSpace::getVoxelsInBody (src/stomp_collision_space.cpp:590-650) — the lattice
    # centre + k * res, |k| <= int(bounding radius / res) per axis; a point is kept when a +z ray from it crosses the surface
    # an odd number of times = strictly inside the convex scaled-then-padded primitive
    for (btype, dims, pos, quat, scale, padding) in bodies:
        c = np.asarray(pos, float)
        Rm = _rotation(quat)
        if btype == 0:
            par = [dims[0] * scale + padding]
            bound = par[0]
        elif btype == 1:
            par = [dims[k] / 2.0 * scale + padding for k in range(3)]
            bound = math.sqrt(sum(v * v for v in par))
        else:
            par = [dims[0] * scale + padding, dims[1] / 2.0 * scale + padding]
            bound = math.sqrt(par[0] ** 2 + par[1] ** 2)
        axes = []
        for k in range(3):
            gmin = int(((c[k] - bound) - c[k]) * (1.0 / resolution))
            gmax = int(((c[k] + bound) - c[k]) * (1.0 / resolution))
            axes.append(np.arange(gmin, gmax + 1) * resolution + c[k])
        X, Y, Z = np.meshgrid(*axes, indexing="ij")
        W = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=-1)
        d = W - c
        lx = Rm[0, 0] * d[:, 0] + Rm[1, 0] * d[:, 1] + Rm[2, 0] * d[:, 2]
        ly = Rm[0, 1] * d[:, 0] + Rm[1, 1] * d[:, 1] + Rm[2, 1] * d[:, 2]
        lz = Rm[0, 2] * d[:, 0] + Rm[1, 2] * d[:, 1] + Rm[2, 2] * d[:, 2]
        if btype == 0:
            inside = lx * lx + ly * ly + lz * lz < par[0] * par[0]
        elif btype == 1:
            inside = (np.abs(lx) < par[0]) & (np.abs(ly) < par[1]) & (np.abs(lz) < par[2])
        else:
            inside = (np.abs(lz) < par[1]) & (lx * lx + ly * ly < par[0] * par[0])
        cc = _round_half_away((W[inside] - origin) / resolution).astype(np.int64)
        cc = cc[np.all((cc >= 0) & (cc < np.array(n)), axis=1)]
        occ[cc[:, 0], cc[:, 1], cc[:, 2]] = True
    # mesh bodies: the same lattice walk, ray parity against the convex hull's triangles
    for (vertices, pos, quat, scale, padding) in meshes:
        tris, c, bound = mesh_body_geometry(vertices, pos, quat, scale, padding)
        axes = []
        for k in range(3):
            gmin = int(((c[k] - bound) - c[k]) * (1.0 / resolution))
            gmax = int(((c[k] + bound) - c[k]) * (1.0 / resolution))
            axes.append(np.arange(gmin, gmax + 1) * resolution + c[k])
        X, Y, Z = np.meshgrid(*axes, indexing="ij")
        W = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=-1)
        inside = _ray_parity_inside(W, tris)
        cc = _round_half_away((W[inside] - origin) / resolution).astype(np.int64)
        cc = cc[np.all((cc >= 0) & (cc < np.array(n)), axis=1)]
        occ[cc[:, 0], cc[:, 1], cc[:, 2]] = True
    cap = int(math.ceil(max_distance / resolution))
    if occ.any():
        d = ndimage.distance_transform_edt(~occ)
        d2 = np.minimum(np.rint(d * d), cap * cap).astype(np.int64)
    else:
        d2 = np.full(n, cap * cap, dtype=np.int64)
    return d2.astype(np.uint8 if cap * cap < 256 else np.uint16), occ


def propagation_field(occ, cap):
    """distance_field::PropagationDistanceField::addPointsToField restated (oracle/stomp_oracle.cpp,
    stomp_oracle_propagate_distance_field): squared cell distances [nx][ny][nz], capped at cap^2.  The upstream package is not in
    the reference repository; this follows its published algorithm (bucket queue over squared distances, closest-point
    propagation through direction-restricted neighbourhoods)."""
    import ctypes as C
    from oracle import oracle
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    out = np.empty(occ.shape, dtype=np.int32)
    L = oracle.lib()
    rc = L.stomp_oracle_propagate_distance_field(occ.shape[0], occ.shape[1], occ.shape[2], occ.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                 int(cap), out.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc:
        raise RuntimeError(L.stomp_oracle_last_error().decode())
    return out
