#!/bin/sh
# Builds the plain-C oracle (TEST INFRASTRUCTURE ONLY) into oracle/_build/libpwc_oracle.so.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$here/_build"
gcc -O2 -fPIC -shared -fopenmp -std=c99 -o "$here/_build/libpwc_oracle.so" "$here/pwc_oracle.c" -lm
echo "built $here/_build/libpwc_oracle.so"
