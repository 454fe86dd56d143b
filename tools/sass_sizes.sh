#!/bin/bash
# tools/sass_sizes.sh [extra nvcc flags] : instruction count of every kernel and noinline device function of b200_uastc.cu
cd "$(dirname "$0")/../basis_universal_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 "$@" -cubin -o /tmp/_sz.cubin ${SRC:-b200_uastc.cu} || exit 1
nvdisasm -c /tmp/_sz.cubin | awk '
/^\$?_Z[A-Za-z0-9_$]*:$/ { if (name != "") printf "%7d %s\n", n, name; name=$1; n=0; next }
/^\.text\./ { next }
/^ +\/\*[0-9a-f]+\*\// { n++ }
END { printf "%7d %s\n", n, name }' | sed -e 's/\$_Z[0-9]*k_[a-z_]*[A-Za-z0-9_]*\$/  /' | c++filt | cut -c1-110
