#!/bin/bash
# Build libstage_b200.so in-tree for sm_100a (also what __graft_entry__.build() runs).
set -e
cd "$(dirname "$0")/tvqaplus_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
     "$@" -o ../libstage_b200.so stage_b200.cu train_b200.cu
