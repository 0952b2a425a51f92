#!/bin/bash
# Timing table over the BASELINE configurations (CUDA events, rotating buffers, one stream).
export KTIME_FLAGS=1,2,4,7,31 KTIME_K=200
python scripts/ktime.py double_integrator_minenergy trapeze 100 | grep -E "F=|copy_"
python scripts/ktime.py goddard midpoint 100000 | grep -E "F=|copy_"
python scripts/ktime.py goddard midpoint 100000 1 optimized | grep -E "F="
python scripts/ktime.py goddard midpoint 1000000 | grep -E "F=|copy_"
python scripts/ktime.py goddard gauss_legendre_3_constant_control 20000 | grep -E "F=|copy_"
python scripts/ktime.py goddard gauss_legendre_3 20000 | grep -E "F="
python scripts/ktime.py goddard gauss_legendre_3 20000 1 optimized | grep -E "F="
python scripts/ktime.py robot trapeze 100000 | grep -E "F=|copy_"
python scripts/ktime.py robot trapeze 100000 1 optimized | grep -E "F="
python scripts/ktime.py goddard euler 100000 | grep -E "F="
KTIME_K=20 python scripts/ktime.py goddard midpoint 1000 4096 | grep -E "F=|copy_"
KTIME_K=20 python scripts/ktime.py goddard midpoint 1000 4096 optimized | grep -E "F="
