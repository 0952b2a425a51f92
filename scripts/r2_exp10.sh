#!/bin/bash
set -u
V=$PWD/ctdirect.jl_b200/lib/variants
CTD_LIB=$V/libctd_g3s.so timeout 300 python scripts/check_one.py goddard gauss_legendre_3 2>&1 | tail -2
CTD_LIB=$V/libctd_g3c.so timeout 300 python scripts/check_one.py goddard gauss_legendre_3_constant_control 2>&1 | tail -1
CTD_LIB=$V/libctd_rob.so timeout 300 python scripts/check_one.py robot trapeze 2>&1 | tail -1
export KTIME_K=300 KTIME_FLAGS=2,4,7
CTD_LIB=$V/libctd_g3c.so timeout 100 python scripts/ktime.py goddard gauss_legendre_3_constant_control 20000 2>&1 | grep "F="
CTD_LIB=$V/libctd_g3s.so timeout 100 python scripts/ktime.py goddard gauss_legendre_3 20000 2>&1 | grep "F="
CTD_LIB=$V/libctd_rob.so timeout 100 python scripts/ktime.py robot trapeze 100000 2>&1 | grep "F="
