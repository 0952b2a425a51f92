#!/bin/bash
# usage: scripts/build_variant.sh TAG ELEMENT "-DFLAG ..."   -> ctdirect.jl_b200/lib/variants/libctd_TAG.so
set -e
cd "$(dirname "$0")/../ctdirect.jl_b200/csrc"
TAG=$1; ELEM=$2; shift 2
mkdir -p ../lib/variants /tmp/ctdv_$TAG
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -diag-suppress 177 $@"
nvcc $FLAGS -c ctd_model.cu -o /tmp/ctdv_$TAG/m.o &
nvcc $FLAGS -c ctd_capi.cu -o /tmp/ctdv_$TAG/c.o &
nvcc $FLAGS -c ctd_comm.cu -o /tmp/ctdv_$TAG/n.o &
nvcc $FLAGS -c ctd_kkt.cu -o /tmp/ctdv_$TAG/k.o &
nvcc $FLAGS -c generated/$ELEM.cu -o /tmp/ctdv_$TAG/e.o &
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o ../lib/variants/libctd_$TAG.so /tmp/ctdv_$TAG/m.o /tmp/ctdv_$TAG/c.o /tmp/ctdv_$TAG/n.o /tmp/ctdv_$TAG/k.o /tmp/ctdv_$TAG/e.o -ldl
echo built $TAG
