#!/bin/sh
# Builds tests/emu/libfpt_emu.so: the product's kernel headers compiled for the CPU against cuda_emu.h.
# Test infrastructure only (see cuda_emu.h).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
CSRC="$HERE/../../fast-parallel-tools-for-genome-wide-analysis-of-genomic-divergence_b200/csrc"
# usage: build.sh [output.so [extra g++ flags...]]
OUT="${1:-$HERE/libfpt_emu.so}"
[ $# -gt 0 ] && shift
g++ -O2 -g -std=c++17 -ffp-contract=off -fPIC -shared -pthread "$@" -include "$HERE/cuda_emu.h" -I"$HERE" -I"$CSRC" \
    -Wall -Wno-unused-function -Wno-unknown-pragmas -Wno-sign-compare \
    "$HERE/emu_driver.cpp" -lm -o "$OUT"
