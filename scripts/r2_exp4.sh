#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/e4_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/e4_pytest.log
export KTIME_K=300
KTIME_FLAGS=7,31 timeout 200 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep -E "F=" | tee gpurun_out/e4_k.log
KTIME_FLAGS=7 KTIME_STREAMS=3 timeout 200 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep -E "F=" | tee -a gpurun_out/e4_k.log
KTIME_FLAGS=7,31 timeout 200 python scripts/ktime.py robot trapeze 100000 2>&1 | grep -E "F=" | tee -a gpurun_out/e4_k.log
KTIME_FLAGS=7,31 timeout 200 python scripts/ktime.py goddard trapeze 100000 2>&1 | grep -E "F=" | tee -a gpurun_out/e4_k.log
KTIME_FLAGS=7,31 timeout 200 python scripts/ktime.py goddard gauss_legendre_2 100000 2>&1 | grep -E "F=" | tee -a gpurun_out/e4_k.log
KTIME_K=20 KTIME_FLAGS=7,31 timeout 200 python scripts/ktime.py goddard midpoint 1000 4096 2>&1 | grep -E "F=" | tee -a gpurun_out/e4_k.log
KTIME_K=20 KTIME_FLAGS=7,31 timeout 200 python scripts/ktime.py goddard midpoint 1000 512 2>&1 | grep -E "F=" | tee -a gpurun_out/e4_k.log
