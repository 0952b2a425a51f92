#!/bin/bash
# Run ON THE GPU BOX (gpurun -- 'bash scripts/refresh_profiles.sh'): regenerates the raw material of profiles/
# under gpurun_out/ (bench line, configuration tables, ncu launch list, ncu --set full reports).  Every ncu pass
# runs only after the same command has exited 0 without ncu.  Afterwards, in the container:
#   cp gpurun_out/bench1.json profiles/rN_bench_line.json; cp gpurun_out/sizes*.txt profiles/ ...
#   python scripts/summarise_profile.py rN gpurun_out/rN_launches.csv gpurun_out/rN_fused.ncu-rep
set -u
R=${1:-r1}
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
bash scripts/config_table.sh > gpurun_out/sizes.txt 2>&1
bash scripts/config_table_streams.sh > gpurun_out/sizes_streams.txt 2>&1
timeout 200 python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/b50.log 2>&1 && \
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/${R}_launches.csv python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/ncu.log 2>&1
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
