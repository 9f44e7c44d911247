#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "sinkhorn or division or infer or oracle_driver or late or regroup" > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest_gpu.log; tail -n 3 gpurun_out/r2i_pytest_gpu.log
timeout 300 python tools/time_driver.py > gpurun_out/r2i_time_driver.txt 2>&1; tail -n 5 gpurun_out/r2i_time_driver.txt | cut -c1-300
