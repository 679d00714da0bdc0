"""CPU-side checks of the drop-in boundary: the CUDA library loads without a GPU, exports every symbol that
include/stomp_b200.h declares, its structs have the layout the ctypes mirror assumes, and the product fails
loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "stomp_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stomp_engine_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from stomp_motion_planner_icra2011_b200 import engine
    lib = engine.lib()
    declared = _declared_functions()
    assert len(declared) >= 30
    missing = [f for f in declared if not hasattr(lib, f)]
    assert not missing, missing
    assert sorted(engine.EXPORTS) == declared
    assert lib.stomp_engine_abi_version() == 1
    assert b"sm_100a" in lib.stomp_engine_build_info()


def test_struct_layouts_match_the_header(tmp_path):
    from stomp_motion_planner_icra2011_b200 import _abi
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "stomp_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(stomp_engine_desc),sizeof(stomp_segment),sizeof(stomp_sphere),sizeof(stomp_joint_limit),'
                    'sizeof(stomp_iter_stats),sizeof(stomp_sphere_debug),offsetof(stomp_engine_desc,movement_duration),'
                    'offsetof(stomp_segment,fixed_value));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_abi.EngineDesc), C.sizeof(_abi.Segment), C.sizeof(_abi.Sphere), C.sizeof(_abi.JointLimit),
            C.sizeof(_abi.IterStats), C.sizeof(_abi.SphereDebug), _abi.EngineDesc.movement_duration.offset,
            _abi.Segment.fixed_value.offset]
    assert got == want


def test_header_is_plain_c(tmp_path):
    prog = tmp_path / "c.c"
    prog.write_text('#include "stomp_b200.h"\nint main(void){return STOMP_F64;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c",
                           str(prog), "-o", str(tmp_path / "c.o")])


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from stomp_motion_planner_icra2011_b200 import scenes
    from stomp_motion_planner_icra2011_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CUDA device"):
        Engine(scenes.make_scenario("tiny"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "stomp_motion_planner_icra2011_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "scenes.py" or f == "distributed.py", (dirpath, f)
    out = subprocess.check_output([sys.executable, "-c", "import sys; import stomp_motion_planner_icra2011_b200.engine, "
                                   "stomp_motion_planner_icra2011_b200.scenes, stomp_motion_planner_icra2011_b200.distributed; "
                                   "print([m for m in sys.modules if 'oracle' in m])"], cwd=ROOT)
    assert out.strip() == b"[]"
