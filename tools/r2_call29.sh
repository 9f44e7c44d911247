#!/bin/bash
# 1-GPU evidence refresh after the Sinkhorn / memo / pipeline changes
mkdir -p gpurun_out
for c in c3 c5; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2r_bench_${c}_n1.json 2> gpurun_out/r2r_bench_${c}_n1.err; echo "bench $c exit $?"; tail -n 2 gpurun_out/r2r_bench_${c}_n1.err
done
timeout 300 python tools/step_timeline.py c2_slice > gpurun_out/r2r_step_timeline_c2.txt 2>&1; tail -n 3 gpurun_out/r2r_step_timeline_c2.txt | cut -c1-150
timeout 300 python tools/time_driver.py > gpurun_out/r2r_time_driver.txt 2>&1; tail -n 4 gpurun_out/r2r_time_driver.txt | cut -c1-200
# ncu: launch list of one bench run, then --set full of the one-warp Sinkhorn kernel (4 rows) and the two-warp team kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 2 --warmup 1 --no-full-driver --no-cpu-baseline > gpurun_out/r2r_ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sinkhorn_regroup_own_kernel" --launch-skip 12 -c 9 -o gpurun_out/r2r_ncu_sk_own -f python tools/run_round.py c2_slice 1000000 3 > gpurun_out/r2r_ncu_sk_own.log 2>&1; echo "ncu sk exit $?"
ls -la gpurun_out/*.ncu-rep | tail -3
