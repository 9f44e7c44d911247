#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_pytest_gpu.log; tail -n 3 gpurun_out/r2l_pytest_gpu.log
timeout 300 python tools/time_driver.py > gpurun_out/r2l_time_driver.txt 2>&1; tail -n 6 gpurun_out/r2l_time_driver.txt | cut -c1-330
timeout 300 python tools/run_round.py c2_slice 1000000 4 2>&1 | tail -4
timeout 300 python tools/step_timeline.py c2_slice 2>&1 | tail -22 | cut -c1-130
