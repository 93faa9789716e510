#!/bin/sh
# TEST-ONLY: host build of the CUDA engine's per-instance source (G = 1).
set -e
cd "$(dirname "$0")"
mkdir -p _build
g++ -O2 -std=c++17 -fPIC -shared -x c++ -o _build/libahk_hostemu.so hostemu.cpp -lm
echo "built tests/hostemu/_build/libahk_hostemu.so"
