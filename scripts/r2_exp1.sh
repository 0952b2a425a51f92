#!/bin/bash
# round 2, experiment 1 (GPU box): parity of the rewritten kernels, timings, per-warp timeline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/e1_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/e1_pytest.log
export KTIME_K=300
KTIME_FLAGS=1,2,4,7,31 timeout 200 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep -E "F=|copy_" | tee gpurun_out/e1_k1.log
KTIME_FLAGS=7,31 KTIME_STREAMS=3 timeout 200 python scripts/ktime.py goddard midpoint 100000 2>&1 | grep -E "F=" | tee gpurun_out/e1_k3.log
echo "--- ticket mode (CTD_POLL_MAX_BATCH=0) vs collector mode, problems with a grid-wide sum"
for P in "bolza_freetf midpoint 100000" "double_integrator_minenergy trapeze 100000" "goddard gauss_legendre_3_constant_control 20000"; do
  KTIME_FLAGS=7,31 timeout 200 python scripts/ktime.py $P 2>&1 | grep -E "F=" | tee -a gpurun_out/e1_poll.log
  CTD_POLL_MAX_BATCH=0 KTIME_FLAGS=7,31 timeout 200 python scripts/ktime.py $P 2>&1 | grep -E "F=" | sed 's/^/ticket: /' | tee -a gpurun_out/e1_poll.log
done
KTIME_K=20 KTIME_FLAGS=7,31 timeout 200 python scripts/ktime.py goddard midpoint 1000 4096 2>&1 | grep -E "F=" | tee gpurun_out/e1_kb.log
echo "--- timeline"
CTD_LIB=$PWD/ctdirect.jl_b200/lib/variants/libctd_trace.so timeout 200 python scripts/trace_run.py 7 8 2>&1 | tee gpurun_out/e1_trace.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/e1_bench.json 2> gpurun_out/e1_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/e1_bench.json'))
print('value', d['value'], 'single', d['single_stream'], 'roofline', d['roofline']['frac'], d['roofline']['launch_us'], 'ov', d['roofline']['overlapped']['frac'])
print('e2e', d['e2e']['value'], 'batched', d['batched']['batched_evals_per_s'], d['batched']['hbm_frac_per_gpu'])
"
