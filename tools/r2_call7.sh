#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest_gpu.log; tail -n 4 gpurun_out/r2g_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2g_smoke.log; tail -n 2 gpurun_out/r2g_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_c2_n1.json 2> gpurun_out/r2g_bench_c2_n1.err; echo "bench exit $?"
timeout 300 python tools/step_timeline.py c2_slice > gpurun_out/r2g_step_timeline_c2.txt 2>&1; tail -n 3 gpurun_out/r2g_step_timeline_c2.txt
timeout 300 python tools/time_driver.py > gpurun_out/r2g_time_driver.txt 2>&1; tail -n 12 gpurun_out/r2g_time_driver.txt
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2g_bench_c2_n1.json').read().strip().splitlines()[-1])
print("value %.1fM ms %.3f e2e %.2fM" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6), d["roofline"]["stage_ms_per_step"])
print(json.dumps(d.get("full_driver"))[:1500])
PY
