#!/bin/bash
set -u
mkdir -p gpurun_out
V=$PWD/ctdirect.jl_b200/lib/variants
export KTIME_K=300 KTIME_FLAGS=2,4,7
echo "GL3 staged loads + evict-first"; CTD_LIB=$V/libctd_gl3.so timeout 100 python scripts/ktime.py goddard gauss_legendre_3_constant_control 20000 2>&1 | grep "F="
echo "GL3 staged loads, no hint";    CTD_LIB=$V/libctd_gl3ne.so timeout 100 python scripts/ktime.py goddard gauss_legendre_3_constant_control 20000 2>&1 | grep "F="
export KTIME_FLAGS=7
echo "midpoint evict-first"; CTD_LIB=$V/libctd_gm.so timeout 100 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep "F="
echo "midpoint no hint";     CTD_LIB=$V/libctd_gmne.so timeout 100 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep "F="
echo "midpoint evict-first 3 streams"; KTIME_STREAMS=3 CTD_LIB=$V/libctd_gm.so timeout 100 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep "F="
echo "batch evict-first"; KTIME_K=20 CTD_LIB=$V/libctd_gm.so timeout 100 python scripts/ktime.py goddard midpoint 1000 4096 2>&1 | grep "F="
echo "batch no hint"; KTIME_K=20 CTD_LIB=$V/libctd_gmne.so timeout 100 python scripts/ktime.py goddard midpoint 1000 4096 2>&1 | grep "F="
echo "--- GL3 steady trace"
CTD_LIB=$V/libctd_gl3t.so timeout 200 python scripts/trace_steady.py 7 2 goddard gauss_legendre_3_constant_control 20000 2>&1 | grep -E "K: |period|gap"
