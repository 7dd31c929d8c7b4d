#!/bin/bash
# Build the library as of a git ref into build/libvoc_<name>.so, for A/B runs inside ONE gpurun call
# (box-to-box variance on the pool is ~5 %, more than most kernel changes):
#   tools/ab_build.sh HEAD prev   ->  build/libvoc_prev.so
#   on the box:  tools/ab_run.sh prev -- python tools/gemm_bench.py ...   (swaps the library in, runs, swaps back)
set -e
ref=$1; name=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p "$tmp/csrc" "$tmp/include" "$root/build"
for f in $(git -C "$root" ls-tree --name-only "$ref" qwen3-tts-axera-russian_b200/csrc/); do
    git -C "$root" show "$ref:$f" > "$tmp/csrc/$(basename "$f")"
done
for f in $(git -C "$root" ls-tree --name-only "$ref" include/); do
    git -C "$root" show "$ref:$f" > "$tmp/include/$(basename "$f")"
done
mkdir -p "$tmp/pkg/csrc"; mv "$tmp/csrc"/* "$tmp/pkg/csrc/"; mkdir -p "$tmp/include2"
# the sources include ../../include/voc_b200.h relative to csrc/
mkdir -p "$tmp/top/pkg/csrc" "$tmp/top/include"
mv "$tmp/pkg/csrc"/* "$tmp/top/pkg/csrc/"; mv "$tmp/include"/* "$tmp/top/include/"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -shared \
    -o "$root/build/libvoc_$name.so" "$tmp/top/pkg/csrc"/*.cu 2>&1 | grep -i error || true
rm -rf "$tmp"
ls -la "$root/build/libvoc_$name.so"
