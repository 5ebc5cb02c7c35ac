#!/bin/bash
# usage: tools/build_slice_variant.sh <name> <k_bins.o to reuse> [-DFOO ...]   (rebuilds k_slice.cu + slicq_api.cu only)
name=$1; kb=$2; shift 2
d=build/variants/$name; mkdir -p $d
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v"
nvcc $F "$@" -c xumx_slicq_b200/csrc/k_slice.cu -o $d/k_slice.o > $d/k_slice.o.log 2>&1 &
nvcc $F "$@" -c xumx_slicq_b200/csrc/slicq_api.cu -o $d/slicq_api.o > $d/slicq_api.o.log 2>&1 &
wait
grep -h "error\|Used\|spill stores" $d/k_slice.o.log | grep -v " 0 bytes spill" 
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $d/libslicq.so $d/k_slice.o $d/slicq_api.o $kb && echo built $d/libslicq.so
