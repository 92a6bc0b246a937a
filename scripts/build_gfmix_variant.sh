#!/bin/bash
# Development: build a variant of the library that differs only in gf_mix.cu (role-timeline counters on), for
# scripts/gfmix_ab.py.   usage: scripts/build_gfmix_variant.sh <name> [-D...]   ->  paif_b200/build/libv_<name>.so
# (the other objects come from the normal build: run `python -m paif_b200.build --force` first)
set -e
cd "$(dirname "$0")/../paif_b200"
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr --extended-lambda \
     -Xcompiler -fPIC -DPAIF_TC_PROFILE "$@" -c csrc/gf_mix.cu -o build/gf_mix_v_$name.o
objs=""
for s in abi stem gf conv_direct conv_tcgen05 pointwise glue fusion_net; do objs="$objs build/${s}.o"; done
nvcc -shared -o build/libv_$name.so build/gf_mix_v_$name.o $objs -lcudart
echo "built paif_b200/build/libv_$name.so"
