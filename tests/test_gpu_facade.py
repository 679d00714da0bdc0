"""Runs the C++ facade test binary (reference-shaped classes above the C ABI) on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "stomp_motion_planner_icra2011_b200", "cpp", "facade_test")


@pytest.mark.gpu
def test_cpp_facade_on_gpu():
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "facade ok" in out.stdout
