#!/bin/bash
# quick A/B: tensor-core parity tests + device-resident bench lines without the CPU / driver legs
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_scale_parity.py -m gpu -x -q 2>&1 | tail -n 3 | cut -c1-200
for c in ${CONFIGS:-c2 c3 c5}; do
  python bench.py --steps 10 --warmup 3 --config $c --no-cpu-baseline --no-full-driver 2>/dev/null | grep '^{' | tail -1 > gpurun_out/q_${c}.json
  python - <<PY
import json
d = json.load(open("gpurun_out/q_${c}.json"))
print("${c}", "value", round(d["value"] / 1e6, 1), "M/s  ms", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()}, d["clocks"]["sm_mhz"])
PY
done
