#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_julia_shim_sequence.py tests/test_mpmath_golden.py -m gpu -q > gpurun_out/e11_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/e11_pytest.log | tail -2
echo "--- GL3 steady trace"
CTD_LIB=$PWD/ctdirect.jl_b200/lib/variants/libctd_gl3t.so timeout 200 python scripts/trace_steady.py 7 2 goddard gauss_legendre_3_constant_control 20000 2>&1 | tee gpurun_out/e11_trace_gl3.log
export KTIME_FLAGS=7 KTIME_K=20
timeout 100 python scripts/ktime.py goddard gauss_legendre_3_constant_control 20000 > gpurun_out/k_gl3.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:ctd_eval_kernel -s 15 -c 2 -f \
    -o gpurun_out/r2_gl3 python scripts/ktime.py goddard gauss_legendre_3_constant_control 20000 > gpurun_out/n_gl3.log 2>&1
echo "ncu rc=$?"
