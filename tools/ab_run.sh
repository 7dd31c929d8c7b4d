#!/bin/bash
# Run a command with build/libvoc_<name>.so swapped in for the in-tree library (see tools/ab_build.sh).
set -e
name=$1; shift; [ "$1" == "--" ] && shift
root=$(cd "$(dirname "$0")/.." && pwd)
lib="$root/qwen3-tts-axera-russian_b200/libvoc_b200.so"
cp "$lib" "$root/build/libvoc_cur.so.bak"
cp "$root/build/libvoc_$name.so" "$lib"
"$@" || true
cp "$root/build/libvoc_cur.so.bak" "$lib"
