#!/bin/bash
# round 2, experiment 2 (GPU box): full GPU test suite, warps per CTA, steady-state timeline
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/e2_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/e2_pytest.log
export KTIME_K=300 KTIME_FLAGS=7
for W in 8 4 2 1; do
  echo "CTD_WPC=$W"; CTD_WPC=$W timeout 200 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep -E "F=" | tee -a gpurun_out/e2_wpc.log
done
for W in 8 4 2; do
  echo "3 streams CTD_WPC=$W"; KTIME_STREAMS=3 CTD_WPC=$W timeout 200 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep -E "F=" | tee -a gpurun_out/e2_wpc.log
done
echo "--- steady-state timeline (trace build)"
for W in 8 2; do
CTD_WPC=$W CTD_LIB=$PWD/ctdirect.jl_b200/lib/variants/libctd_trace.so timeout 200 python scripts/trace_steady.py 7 $W 2>&1 | tee gpurun_out/e2_trace_w$W.log
done
