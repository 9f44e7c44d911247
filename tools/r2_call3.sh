#!/bin/bash
# GPU call 3 of round 2: regression gate for the new pass-2 skip logic, then the TF32 screening kernel (each risky step under its own timeout)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest_gpu.log; tail -n 3 gpurun_out/r2c_pytest_gpu.log
for cfg in c2_slice c5_slice; do
    timeout 300 python tools/check_tf32.py $cfg > gpurun_out/r2c_check_tf32_$cfg.txt 2>&1; echo "check_tf32 exit $?" >> gpurun_out/r2c_check_tf32_$cfg.txt
    tail -n 25 gpurun_out/r2c_check_tf32_$cfg.txt
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_c2_n1.json 2> gpurun_out/r2c_bench_c2_n1.err; echo "bench exit $?"; tail -c 3000 gpurun_out/r2c_bench_c2_n1.json
