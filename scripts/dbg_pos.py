import numpy as np, sys
sys.path.insert(0, ".")
from stomp_motion_planner_icra2011_b200 import _abi, scenes
from stomp_motion_planner_icra2011_b200.engine import Engine
from oracle.oracle import Oracle
from tests.helpers import correlated_noise
sc = scenes.make_scenario("C1", num_problems=1)
eng = Engine(sc, keep_intermediates=1); o = Oracle(sc, 0)
rng = np.random.default_rng(5)
L = o.get(_abi.FIELD_NOISE_CHOLESKY)
eps = correlated_noise(L, rng, (1, 4), np.full(7, 2.0))
params = o.get_parameters()[None, None] + eps
for r in range(2):
    dbg = eng.execute_debug(params[0, r]); odbg, clipped = o.execute_debug(params[0, r])
    d = np.abs(dbg["position"] - odbg["position"]).max(axis=-1)   # [N+3][K]
    print("r", r, "max diff per t (first 30):", np.array2string(d.max(axis=1)[:30], precision=1))
    print("max diff per sphere:", np.array2string(d.max(axis=0), precision=1))
    viol = np.abs(clipped - params[0, r]).max(axis=0)
    print("clip delta per t (first 30)", np.array2string(viol[:30], precision=2))
# no-noise: plain theta
dbg = eng.execute_debug(o.get_parameters()); odbg, clipped = o.execute_debug(o.get_parameters())
print("theta only: max pos diff", np.abs(dbg["position"] - odbg["position"]).max(), "clip", np.abs(clipped - o.get_parameters()).max())
