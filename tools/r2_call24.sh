#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2p_pytest_gpu.log; tail -n 4 gpurun_out/r2p_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; tail -n 2 gpurun_out/r2p_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench_c2_n1.json 2> gpurun_out/r2p_bench_c2_n1.err; tail -c 600 gpurun_out/r2p_bench_c2_n1.json; tail -n 3 gpurun_out/r2p_bench_c2_n1.err
