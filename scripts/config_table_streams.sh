#!/bin/bash
# Throughput with 3 independent evaluations in flight (3 handles / CUDA streams), same configurations.
export KTIME_FLAGS=7 KTIME_K=300 KTIME_STREAMS=3
python scripts/ktime.py goddard midpoint 100000 | grep -E "F="
python scripts/ktime.py goddard midpoint 100000 1 optimized | grep -E "F="
python scripts/ktime.py goddard gauss_legendre_3_constant_control 20000 | grep -E "F="
python scripts/ktime.py goddard gauss_legendre_3 20000 | grep -E "F="
python scripts/ktime.py goddard gauss_legendre_3 20000 1 optimized | grep -E "F="
python scripts/ktime.py robot trapeze 100000 | grep -E "F="
python scripts/ktime.py robot trapeze 100000 1 optimized | grep -E "F="
python scripts/ktime.py goddard trapeze 100000 | grep -E "F="
python scripts/ktime.py goddard euler 100000 | grep -E "F="
python scripts/ktime.py goddard gauss_legendre_2 100000 | grep -E "F="
