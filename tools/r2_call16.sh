#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:sinkhorn_regroup -c 8 -o gpurun_out/r2n_ncu_sk -f python tools/run_round.py c2_slice 1000000 2 > gpurun_out/r2n_ncu_sk.log 2>&1
tail -3 gpurun_out/r2n_ncu_sk.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
