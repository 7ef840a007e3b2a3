#!/bin/bash
# Development build of the CUDA library (same flags as __graft_entry__.build, plus ptxas -v into /tmp/ptxas.log).
set -e
CSRC=/root/repo/synference_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -Xptxas -v "$@" \
  $CSRC/capi.cu -o /tmp/libsb2_dev.so 2> /tmp/ptxas.log || { grep -i -B2 -A6 "error" /tmp/ptxas.log | head -60; exit 1; }
cp /tmp/libsb2_dev.so $CSRC/libsynference_b200.so
grep -c "Compiling entry" /tmp/ptxas.log
grep -B3 "spill stores" /tmp/ptxas.log | grep -A3 "synth3" | grep -v "0 bytes spill stores" | grep "spill" | head
