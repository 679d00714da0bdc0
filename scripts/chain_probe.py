import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["STOMP_CHAIN_PROBE"] = "1"
from stomp_motion_planner_icra2011_b200 import scenes
from stomp_motion_planner_icra2011_b200.engine import Engine
eng = Engine(scenes.make_scenario(sys.argv[1] if len(sys.argv) > 1 else "C2"))
for it in range(1, 401):
    eng.iterate(it, stats=False)
eng.get(0)
eng.close()
