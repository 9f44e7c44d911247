#!/bin/bash
# GPU call 4 of round 2: full gpu suite incl. the scale-parity tests, gate calibration of the TF32 screen, bench with the screen on by
# default + stage breakdown of the full driver, ncu launch list + full capture of the new first-layer kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest_gpu.log; tail -n 4 gpurun_out/r2d_pytest_gpu.log
grep "re-run by the three-pass" gpurun_out/r2d_pytest_gpu.log | head -30
timeout 600 python tools/calibrate_gate.py c2_slice c3_slice c5_slice > gpurun_out/r2d_gate_calibration.jsonl 2> gpurun_out/r2d_gate_calibration.err; echo "calibrate exit $?"
timeout 300 python tools/check_tf32.py c3_slice > gpurun_out/r2d_check_tf32_c3_slice.txt 2>&1; tail -n 9 gpurun_out/r2d_check_tf32_c3_slice.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench_c2_n1.json 2> gpurun_out/r2d_bench_c2_n1.err; echo "bench exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 --screen off --no-cpu-baseline --no-full-driver > gpurun_out/r2d_bench_c2_n1_noscreen.json 2>/dev/null
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline --no-full-driver > gpurun_out/r2d_bench_c3_n1.json 2>/dev/null
timeout 600 python bench.py --config c3 --screen tf32 --steps 10 --warmup 3 --no-cpu-baseline --no-full-driver > gpurun_out/r2d_bench_c3_n1_tf32.json 2>/dev/null
timeout 600 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu-baseline --no-full-driver > gpurun_out/r2d_bench_c5_n1.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-driver > gpurun_out/r2d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-driver > gpurun_out/r2d_ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-driver > gpurun_out/r2d_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:linear_tf32_kernel -s 3 -c 1 -o gpurun_out/r2d_ncu_full_linear_tf32 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-driver > gpurun_out/r2d_ncu_full.log 2>&1
tail -c 1500 gpurun_out/r2d_bench_c2_n1.json
