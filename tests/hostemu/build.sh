#!/bin/sh
# Builds tests/hostemu/_build/libbu_hostemu.so (test infrastructure; see hostemu.cpp).
set -e
cd "$(dirname "$0")"
mkdir -p _build
g++ -std=c++17 -O2 -fPIC -shared -ffp-contract=off -fno-fast-math -Wall -Wno-unused-function -o _build/libbu_hostemu.so hostemu.cpp -lpthread
