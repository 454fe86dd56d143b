#!/bin/bash
# tools/build_variant.sh NAME [extra nvcc flags...] -> scratch/variants/NAME.so (development A/B builds; see tools/ab_bench.py)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p scratch/variants
cd basis_universal_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC -shared "$@" -o ../../scratch/variants/$name.so b200_context.cu b200_uastc.cu b200_etc1s.cu b200_rdo.cu b200_tsvq.cu b200_dist.cu b200_image.cu
