#!/bin/bash
# 2-GPU validation: sharded driver (new per-round protocol) == single-GPU driver, peer-memory dedup, bench at N=2 with config.extra
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_shard_driver.py 200000 c2_slice,c1_slice > gpurun_out/r2f_shard_driver_n2.txt 2>&1; echo "shard driver exit $?" >> gpurun_out/r2f_shard_driver_n2.txt; tail -n 4 gpurun_out/r2f_shard_driver_n2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_shard_p2p.py > gpurun_out/r2f_shard_p2p_n2.txt 2>&1; echo "shard p2p exit $?" >> gpurun_out/r2f_shard_p2p_n2.txt; tail -n 4 gpurun_out/r2f_shard_p2p_n2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_c2_n2.json 2> gpurun_out/r2f_bench_c2_n2.err; echo "bench n2 exit $?"; tail -c 2500 gpurun_out/r2f_bench_c2_n2.json; tail -n 5 gpurun_out/r2f_bench_c2_n2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2>/dev/null; tail -c 700 gpurun_out/r2f_bench_reference.json
