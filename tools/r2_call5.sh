#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest_gpu.log; tail -n 4 gpurun_out/r2e_pytest_gpu.log
grep "re-run by the three-pass" gpurun_out/r2e_pytest_gpu.log | head -40
CALIBRATE_COMPS=0,0.5,0.72,1.0 timeout 900 python tools/calibrate_gate.py c2_slice > gpurun_out/r2e_gate_calibration_comp.jsonl 2> gpurun_out/r2e_gate_calibration.err; echo "calibrate exit $?"
timeout 600 python tools/calibrate_gate.py c3_slice c5_slice >> gpurun_out/r2e_gate_calibration_comp.jsonl 2>> gpurun_out/r2e_gate_calibration.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-full-driver > gpurun_out/r2e_bench_c2_n1.json 2> gpurun_out/r2e_bench_c2_n1.err; echo "bench exit $?"
tail -c 600 gpurun_out/r2e_bench_c2_n1.json
