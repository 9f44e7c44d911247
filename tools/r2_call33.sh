#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2u_pytest_gpu.log; tail -n 6 gpurun_out/r2u_pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2u_bench_c2_n1.json 2> gpurun_out/r2u_bench_c2_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u_bench_c2_n1.json'))
print('value',d['value']/1e6,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e6,'full',d['full_driver']['value'], d['roofline']['stage_ms_per_step'])
PY
tail -n 2 gpurun_out/r2u_bench_c2_n1.err
timeout 300 python tools/step_timeline.py c2_slice > gpurun_out/r2u_step_timeline_c2.txt 2>&1; tail -n 14 gpurun_out/r2u_step_timeline_c2.txt | cut -c1-140
