"""A second, independent restatement of the path in NumPy (written from the reference text, sharing no code with
oracle/stomp_oracle.cpp) checked against the C++ oracle on a tiny case.  The reference ships no golden vectors
(SURVEY.md §8c), so agreement of two independently written restatements - and of the CUDA engine with both - is the
strongest pin available for the oracle."""
import numpy as np

from oracle.oracle import Oracle
from stomp_motion_planner_icra2011_b200 import _abi, scenes
from tests.helpers import correlated_noise

RULES = np.array([[0, 0, -2 / 6.0, -3 / 6.0, 6 / 6.0, -1 / 6.0, 0],
                  [0, -1 / 12.0, 16 / 12.0, -30 / 12.0, 16 / 12.0, -1 / 12.0, 0],
                  [0, 1 / 12.0, -17 / 12.0, 46 / 12.0, -46 / 12.0, 17 / 12.0, -1 / 12.0]])   # stomp_utils.h:52-56


def diff_matrix(n, rule, scale):
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(-3, 4):
            if 0 <= i + j < n:
                A[i, i + j] = scale * rule[j + 3]
    return A


class NumpyStomp:
    """Follows policy_improvement.cpp / covariant_trajectory_policy.cpp / stomp_optimizer.cpp line by line in NumPy."""

    def __init__(self, sc, problem):
        self.sc = sc
        self.D, self.N, self.R = sc.robot.num_dimensions, sc.num_time_steps, sc.num_rollouts
        N = self.N
        self.Nall = N + 12
        dt = sc.movement_duration / (N + 1)                                  # covariant_trajectory_policy.cpp:152
        self.A = [diff_matrix(self.Nall, RULES[k], 1.0 / dt ** (k + 1)) for k in range(3)]
        R_all = sc.ridge_factor * np.eye(self.Nall) + sum(w * A.T @ A for w, A in zip(sc.derivative_costs, self.A))
        self.R_all = R_all
        Rf = R_all[6:6 + N, 6:6 + N]
        self.Rinv = np.linalg.inv(Rf)
        self.M = self.Rinv / (N * self.Rinv.max(axis=0))[None, :]             # policy_improvement.cpp:421-441
        Q = sum(w * sc.discretization ** (k + 1) * diff_matrix(self.Nall, RULES[k], 1.0).T @ diff_matrix(self.Nall, RULES[k], 1.0)
                for k, w in enumerate(sc.derivative_costs)) + sc.ridge_factor * np.eye(self.Nall)
        Qinv = np.linalg.inv(Q[6:6 + N, 6:6 + N])
        self.Qinv = Qinv / Qinv.max()                                         # stomp_cost.cpp:94-105, stomp_optimizer.cpp:119-125
        self.start, self.goal = sc.start[problem], sc.goal[problem]
        self.theta = np.zeros((self.D, N))
        for d in range(self.D):                                               # setToMinControlCost
            x = np.concatenate([np.full(6, self.start[d]), np.zeros(N), np.full(6, self.goal[d])])
            lin = 2.0 * (x[:6] @ R_all[:6, 6:6 + N] + x[-6:] @ R_all[-6:, 6:6 + N])
            self.theta[d] = -0.5 * self.Rinv @ lin
        vox = sc.sdf.voxels
        self.dist = (np.sqrt(vox.astype(np.float64)) * sc.sdf.resolution) if sc.sdf.voxel_dtype != _abi.VOXEL_F32 else vox.astype(np.float64)

    # -- cost plugin (stomp_optimizer.cpp:562-709,1063-1165) ------------------------------------------------------------
    def frames(self, q):
        out = []
        for g in self.sc.robot.segments:
            val = q[g["group"]] if g["group"] >= 0 else g["fixed"]
            Rj, p, a = np.array(g["rot"]).reshape(3, 3), np.array(g["pos"]), np.array(g["axis"])
            F = np.eye(4)
            if g["type"] == _abi.JOINT_REVOLUTE:
                K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
                F[:3, :3] = (np.eye(3) + np.sin(val) * K + (1 - np.cos(val)) * (np.outer(a, a) - np.eye(3))) @ Rj
                F[:3, 3] = p
            else:
                F[:3, :3] = Rj
                F[:3, 3] = p + (val * a if g["type"] == _abi.JOINT_PRISMATIC else 0.0)
            out.append(F if g["parent"] < 0 else out[g["parent"]] @ F)
        ref = np.linalg.inv(out[self.sc.robot.reference_segment])
        return [ref @ F for F in out]

    def execute(self, params):
        sc, N = self.sc, self.N
        traj = np.concatenate([np.repeat(self.start[:, None], 6, 1), params, np.repeat(self.goal[:, None], 6, 1)], axis=1)
        for j, (has, lo, hi) in enumerate(sc.robot.limits):                   # handleJointLimits
            if not has:
                continue
            for _ in range(11):
                free = traj[j, 6:6 + N]
                amount = np.where(free > hi, hi - free, np.where(free < lo, lo - free, 0.0))
                if not (np.abs(amount) > 1e-6).any():
                    break
                i = int(np.argmax(np.abs(amount)))                            # first maximum
                traj[j, 6:6 + N] += amount[i] / self.Qinv[i, i] * self.Qinv[:, i]
        K = len(sc.robot.spheres)
        pos = np.zeros((self.Nall, K, 3))
        pot = np.zeros((self.Nall, K))
        org, res, dims = np.array(sc.sdf.origin), sc.sdf.resolution, np.array(sc.sdf.dims)
        for i in range(self.Nall):
            Fs = self.frames(traj[:, i])
            for j, s in enumerate(sc.robot.spheres):
                p = (Fs[s["segment"]] @ np.append(s["pos"], 1.0))[:3]
                pos[i, j] = p
                t = (p - org) / res
                c = np.where(t >= 0, np.floor(t + 0.5), np.ceil(t - 0.5)).astype(int)
                dist = self.dist[tuple(c)] if np.all((c >= 1) & (c < dims - 1)) else 0.0
                d = dist - s["radius"]
                clr = s["clearance"]
                pot[i, j] = 0.0 if d >= clr else (0.5 * (d - clr) ** 2 / clr if d >= 0 else -d + 0.5 * clr)
        costs = np.zeros(N)
        for i in range(6, 6 + N):
            vel = sum(RULES[0][k + 3] / sc.discretization * pos[i + k] for k in range(-3, 4))
            contrib = pot[i] * np.linalg.norm(vel, axis=1)
            costs[i - 6] = sc.obstacle_cost_weight * np.cumsum(contrib).sum()  # cumulative over sphere index
        return costs

    def control_costs(self, params, noise, weight):
        out = np.zeros((self.D, self.N))
        for d in range(self.D):
            x = np.concatenate([np.full(6, self.start[d]), params[d] + noise[d], np.full(6, self.goal[d])])
            c = sum(weight * w * (A @ x) ** 2 for w, A in zip(self.sc.derivative_costs, self.A))
            out[d] = c[6:6 + self.N]
            out[d, 0] += c[:6].sum()
            out[d, -1] += c[-6:].sum()
        return out

    def iterate_first(self, eps):
        """first iteration (no reuse yet): returns (updated theta, state costs, probabilities)."""
        sc = self.sc
        params = self.theta[None] + eps
        S = np.stack([self.execute(p) for p in params])
        Mn = np.einsum("ij,rdj->rdi", self.M, eps)
        C = np.stack([self.control_costs(params[r], Mn[r], 0.5 * sc.smoothness_cost_weight) for r in range(self.R)])
        tot = S[:, None, :] + C
        if sc.use_cumulative_costs:
            tot = np.flip(np.cumsum(np.flip(tot, -1), -1), -1)
        mn, mx = tot.min(0), tot.max(0)
        P = np.exp(-10.0 * (tot - mn) / np.maximum(mx - mn, 1e-8))
        P /= P.sum(0)
        upd = np.einsum("ij,dj->di", self.M, (P * eps).sum(0))
        return self.theta + upd, S, P


def test_two_independent_restatements_agree():
    for cumulative in (0, 1):
        sc = scenes.make_scenario("tiny", num_problems=1, num_time_steps=12, num_rollouts=4, use_cumulative_costs=cumulative)
        o = Oracle(sc, 0)
        ref = NumpyStomp(sc, 0)
        np.testing.assert_allclose(ref.theta, o.get_parameters(), rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(ref.M, o.get(_abi.FIELD_PROJECTION), rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(ref.Qinv, o.get(_abi.FIELD_QUAD_COST_INV), rtol=1e-6, atol=1e-12)
        rng = np.random.default_rng(3)
        eps = correlated_noise(o.get(_abi.FIELD_NOISE_CHOLESKY), rng, (sc.num_rollouts,), sc.noise_stddev)
        theta1, S, P = ref.iterate_first(eps)
        o.iterate(1, eps)
        assert S.max() > 0.0
        np.testing.assert_allclose(S, o.get(_abi.FIELD_STATE_COSTS), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(P, o.get(_abi.FIELD_PROBABILITIES), rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(theta1, o.get(_abi.FIELD_THETA), rtol=1e-6, atol=1e-9)
