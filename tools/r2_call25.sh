#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "host_buffer or torch_ops" 2>&1 | tail -n 2
python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench_c2_n1.json 2> gpurun_out/r2q_bench_c2_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2q_bench_c2_n1.json'))
print('value',d['value']/1e6,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e6,'full',d['full_driver']['value'])
PY
