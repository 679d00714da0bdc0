"""The facade's request pre / post-processing (StompPlannerNode::planKinematicPath, src/stomp_planner_node.cpp:207-217,257-279):
shortest-angular-distance goals for wrap-around joints and the velocity-limited time_from_start of the response.  Host-only C++."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stomp_motion_planner_icra2011_b200")


def test_request_helpers(tmp_path):
    exe = str(tmp_path / "request_helpers_test")
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "cpp"), "-o", exe,
                           os.path.join(PKG, "cpp", "request_helpers_test.cpp"), "-L", PKG, "-lstomp_b200", "-Wl,-rpath," + PKG])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "request helpers ok" in out.stdout, out.stdout + out.stderr
