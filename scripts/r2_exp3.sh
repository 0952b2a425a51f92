#!/bin/bash
# round 2, experiment 3 (GPU box): full GPU tests, bench line, configuration table
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/e3_pytest.log 2>&1; echo "pytest rc=$?"
tail -22 gpurun_out/e3_pytest.log
timeout 600 python bench.py > gpurun_out/e3_bench.json 2> gpurun_out/e3_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/e3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/e3_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'ov',d['overlapped']['value'],d['roofline']['overlapped']['frac'])
print('e2e',d['e2e']['value'],d['e2e']['pipelined']['value'])
for k,v in d['configs'].items(): print(k, round(v['us_per_eval'],2), round(v['frac'],3))
print('sep',{k:(round(v,2) if isinstance(v,float) else v) for k,v in d['separate_callbacks_us'].items() if k!='note'})
b=d['batched']; print('batched',b['batched_evals_per_s'],b['kernel_only'], 'e2e', b.get('e2e'))
print('cpu',d['cpu_baseline'])
PY
bash scripts/config_table.sh > gpurun_out/e3_sizes.txt 2>&1; cat gpurun_out/e3_sizes.txt
