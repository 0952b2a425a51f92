#!/bin/bash
set -u
mkdir -p gpurun_out
V=$PWD/ctdirect.jl_b200/lib/variants
export KTIME_K=300 KTIME_FLAGS=7
for W in 2 8; do echo "base WPC=$W"; CTD_LIB=$V/libctd_base.so CTD_WPC=$W timeout 100 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep "F=" ; done | tee gpurun_out/e5_k.log
for W in 11 5 2; do echo "w11 WPC=$W"; CTD_LIB=$V/libctd_w11.so CTD_WPC=$W timeout 100 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep "F=" ; done | tee -a gpurun_out/e5_k.log
echo "--- steady trace WPC=11"
CTD_WPC=11 CTD_LIB=$V/libctd_w11t.so timeout 200 python scripts/trace_steady.py 7 11 2>&1 | tee gpurun_out/e5_trace_w11.log
echo "--- steady trace WPC=2"
CTD_WPC=2 CTD_LIB=$V/libctd_trace.so timeout 200 python scripts/trace_steady.py 7 2 2>&1 | tee gpurun_out/e5_trace_w2.log
