#!/bin/bash
# build; exit non-zero on any compiler error (used before spending GPU time)
set -e
cd "$(dirname "$0")"
out=$(nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC \
  -o stomp_motion_planner_icra2011_b200/libstomp_b200.so stomp_motion_planner_icra2011_b200/csrc/engine.cu 2>&1) || { echo "$out" | grep -E "error" | head; exit 1; }
make -s -C oracle libstomp_oracle.so
g++ -O2 -std=c++17 -I include -I stomp_motion_planner_icra2011_b200/cpp -o stomp_motion_planner_icra2011_b200/cpp/facade_test \
  stomp_motion_planner_icra2011_b200/cpp/facade_test.cpp -L stomp_motion_planner_icra2011_b200 -lstomp_b200 \
  -Wl,-rpath,$PWD/stomp_motion_planner_icra2011_b200
echo build ok
