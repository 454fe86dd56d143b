#!/bin/bash
# tools/build_variant.sh NAME [extra nvcc flags...] -> scratch/variants/NAME.so (development A/B builds; see tools/ab_bench.py)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p scratch/variants
cd basis_universal_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC -shared "$@" -o ../../scratch/variants/$name.so *.cu
