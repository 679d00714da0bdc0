import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the CUDA library and the oracle are built in-tree (git-ignored): build them when a fresh checkout is tested
    lib = os.path.join(ROOT, "stomp_motion_planner_icra2011_b200", "libstomp_b200.so")
    facade = os.path.join(ROOT, "stomp_motion_planner_icra2011_b200", "cpp", "facade_test")
    if not (os.path.exists(lib) and os.path.exists(facade) and os.path.exists(os.path.join(ROOT, "oracle", "libstomp_oracle.so"))):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def tiny():
    from stomp_motion_planner_icra2011_b200 import scenes
    return scenes.make_scenario("tiny")


@pytest.fixture(scope="session")
def c1():
    from stomp_motion_planner_icra2011_b200 import scenes
    return scenes.make_scenario("C1", num_problems=3)
