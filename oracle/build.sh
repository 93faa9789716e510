#!/bin/sh
# TEST INFRASTRUCTURE: builds the C restatement of the reference's algorithm.
# The reference itself is pure Python (no C sources), so there is nothing to
# compile into oracle/_ref/.
set -e
cd "$(dirname "$0")"
mkdir -p _build
gcc -O2 -std=gnu11 -fPIC -shared -fopenmp -ffp-contract=off -o _build/libahk_oracle.so ahk_oracle.c -lm
echo "built oracle/_build/libahk_oracle.so"
