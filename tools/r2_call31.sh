#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2s_pytest_gpu.log; tail -n 3 gpurun_out/r2s_pytest_gpu.log
timeout 300 python tools/bench_kmeans.py > gpurun_out/r2s_kmeans.jsonl 2>&1; tail -n 1 gpurun_out/r2s_kmeans.jsonl | cut -c1-400
KM_N=200000 timeout 300 python tools/bench_kmeans.py >> gpurun_out/r2s_kmeans.jsonl 2>&1; tail -n 1 gpurun_out/r2s_kmeans.jsonl | cut -c1-300
