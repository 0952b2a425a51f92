#!/bin/bash
set -u
V=$PWD/ctdirect.jl_b200/lib/variants
CTD_NO_STAGES=1 CTD_LIB=$V/libctd_st.so timeout 300 python scripts/check_one.py goddard gauss_legendre_3_constant_control 2>&1 | tail -3
CTD_LIB=$V/libctd_st.so timeout 300 python scripts/check_one.py goddard gauss_legendre_3_constant_control 2>&1 | tail -2
export KTIME_K=300 KTIME_FLAGS=2,4,7,31
echo "stages"; CTD_LIB=$V/libctd_st.so timeout 100 python scripts/ktime.py goddard gauss_legendre_3_constant_control 20000 2>&1 | grep "F="
echo "parts";  CTD_NO_STAGES=1 CTD_LIB=$V/libctd_st.so timeout 100 python scripts/ktime.py goddard gauss_legendre_3_constant_control 20000 2>&1 | grep "F="
