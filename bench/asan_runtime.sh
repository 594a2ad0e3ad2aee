#!/usr/bin/env bash
# Host ASAN + UBSAN build of the sampler's C++ runtime; CPU only.  Also run by tests/test_runtime_cpp.py.
set -eu
cd "$(dirname "$0")/../nanorlhf_b200/csrc"
out="${TMPDIR:-/tmp}/nrl_runtime_asan_$$"
g++ -std=c++17 -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=all -DNRL_RUNTIME_NO_PYBIND runtime_asan_test.cpp -o "$out"
ASAN_OPTIONS=detect_leaks=1 "$out"
rm -f "$out"
