#!/bin/bash
# Builds the CUDA engine (sm_100a) in-tree and the CPU oracle (test infrastructure).
set -e
cd "$(dirname "$0")"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC \
  -o stomp_motion_planner_icra2011_b200/libstomp_b200.so stomp_motion_planner_icra2011_b200/csrc/engine.cu
make -s -C oracle libstomp_oracle.so
