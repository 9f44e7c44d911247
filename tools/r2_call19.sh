#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sinkhorn|small|memo|linear_exact|quantize|group|sk_class|radix|seg_|pack" --launch-skip 0 -c 400 --csv --log-file gpurun_out/r2o_round_launches.csv python tools/run_round.py c2_slice 1000000 5 > gpurun_out/r2o_round.log 2>&1
tail -2 gpurun_out/r2o_round.log | cut -c1-200
