"""Two-stream timeline of a few overlapped iterations: python scripts/timeline.py [workload] [out.csv] (needs a B200).
Event-timed (stomp_engine_set_profiling(2) / stomp_engine_dump_timeline): begin = the stream reached the launch, end = kernel done."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stomp_motion_planner_icra2011_b200 import scenes          # noqa: E402
from stomp_motion_planner_icra2011_b200.engine import Engine   # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/timeline_%s.csv" % name
eng = Engine(scenes.make_scenario(name))
for it in range(1, 9):
    eng.iterate(it, stats=False)
eng.set_profiling(2)
for it in range(9, 15):
    eng.iterate(it, stats=False)
eng.dump_timeline(out)
eng.set_profiling(0)
rows = [l.strip().split(",") for l in open(out)][1:]
# one steady iteration: from the third k_update's end to the fourth's
ends = [float(r[4]) for r in rows if r[1] == "k_update"]
lo, hi = ends[2], ends[3]
print("iteration window %.1f us" % (hi - lo))
for r in rows:
    b, e = float(r[3]), float(r[4])
    if lo - 1 <= b and e <= hi + 1:
        print("%-16s stream %s  %8.1f -> %8.1f  (%6.1f us)" % (r[1], r[2], b - lo, e - lo, e - b))
