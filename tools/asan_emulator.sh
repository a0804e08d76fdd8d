#!/bin/bash
# AddressSanitizer run of the kernel sources on the SIMT emulator (tools/asan_emulator.py) -> profiles/
set -e
cd "$(dirname "$0")/.."
make -s -C oracle port
g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -shared -Wno-unknown-pragmas -x c++ \
    -o tests/hostsim/libhostsim_asan.so tests/hostsim/hostsim.cpp tests/hostsim/kernels_simt.cpp re2-modification_b200/csrc/rxm_plan.cpp
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=0 \
    python tools/asan_emulator.py 2>&1 | tee ${1:-/tmp/asan_emulator.log}
