#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest_gpu.log; tail -n 4 gpurun_out/r2h_pytest_gpu.log
timeout 300 python tools/time_driver.py > gpurun_out/r2h_time_driver.txt 2>&1; tail -n 8 gpurun_out/r2h_time_driver.txt | cut -c1-400
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_c2_n1.json 2> gpurun_out/r2h_bench_c2_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2h_bench_c2_n1.json').read().strip().splitlines()[-1])
print("value %.1fM ms %.3f e2e %.2fM" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6), d["roofline"]["stage_ms_per_step"])
fd = d.get("full_driver"); print(fd["value"], fd["seconds"], fd["stage_ms"], fd["cpu_port"])
PY
