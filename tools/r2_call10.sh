#!/bin/bash
mkdir -p gpurun_out
python tools/run_round.py c2_slice 1000000 2 > gpurun_out/r2j_run_round_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:sinkhorn_regroup|linear_small_kernel|quantize_small_kernel" -c 12 \
    -o gpurun_out/r2j_ncu_full_round python tools/run_round.py c2_slice 1000000 2 > gpurun_out/r2j_ncu_round.log 2>&1
tail -n 4 gpurun_out/r2j_run_round_plain.log; tail -n 3 gpurun_out/r2j_ncu_round.log
