#!/bin/bash
# Run ON THE GPU BOX (gpurun -- 'bash scripts/refresh_profiles.sh r2'): regenerates the raw material of profiles/
# under gpurun_out/ (GPU test log, bench line, configuration tables, per-warp timeline, ncu launch list, ncu --set full
# reports, DRAM traffic over consecutive launches).  Every ncu pass runs only after the same command has exited 0
# without ncu.  Afterwards, in the container:  python scripts/summarise_profile.py r2
set -u
R=${1:-r2}
mkdir -p gpurun_out
V=$PWD/ctdirect.jl_b200/lib/variants
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/${R}_pytest.log
timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/${R}_bench_reference.json 2>/dev/null; echo "reference arm rc=$?"
bash scripts/config_table.sh > gpurun_out/${R}_sizes.txt 2>&1
bash scripts/config_table_streams.sh > gpurun_out/${R}_sizes_streams.txt 2>&1
CTD_WPC=2 CTD_LIB=$V/libctd_trace.so timeout 200 python scripts/trace_steady.py 7 2 > gpurun_out/${R}_timeline.txt 2>&1; echo "timeline rc=$?"
CTD_LIB=$V/libctd_gl3t.so timeout 200 python scripts/trace_steady.py 7 2 goddard gauss_legendre_3_constant_control 20000 > gpurun_out/${R}_timeline_gl3.txt 2>&1
timeout 200 python bench.py --steps 50 --warmup 3 --no-cpu --no-configs > gpurun_out/b50.log 2>&1 && \
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${R}_launches.csv python bench.py --steps 50 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
export KTIME_FLAGS=7 KTIME_K=20
cap() {  # name, skip, count, ktime args...
  local name=$1 skip=$2 count=$3; shift 3
  timeout 100 python scripts/ktime.py "$@" > gpurun_out/k_$name.log 2>&1 && \
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:ctd_eval_kernel -s $skip -c $count -f \
      -o gpurun_out/${R}_$name python scripts/ktime.py "$@" > gpurun_out/n_$name.log 2>&1
  echo "$name rc=$?"
}
cap fused 15 3
cap batched 12 1 goddard midpoint 1000 4096
cap gl3 15 2 goddard gauss_legendre_3_constant_control 20000
# DRAM traffic in the regime the bench times: 12 CONSECUTIVE launches on the 6 rotating buffer sets, no cache flush
timeout 100 python scripts/ktime.py > gpurun_out/k_traffic.log 2>&1 && \
  timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,gpu__time_duration.sum \
    -k regex:ctd_eval_kernel -s 12 -c 12 --csv --log-file gpurun_out/${R}_traffic.csv python scripts/ktime.py > gpurun_out/n_traffic.log 2>&1
echo "traffic rc=$?"
