#!/bin/bash
# The working tree's library with role-level cycle counters compiled in (-DVOC_TC_PROF) -> build/libvoc_prof.so
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/build"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DVOC_TC_PROF -Xcompiler -fPIC,-ffp-contract=off -shared \
    -o "$root/build/libvoc_prof.so" "$root"/qwen3-tts-axera-russian_b200/csrc/*.cu
ls -la "$root/build/libvoc_prof.so"
