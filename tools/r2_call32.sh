#!/bin/bash
# 2-GPU validation after the round-2 kernel changes: bench line (weak-scaled C2 + configs[2] strong-scaled), sharded driver, peer dedup
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2t_bench_c2_n2.json 2> gpurun_out/r2t_bench_c2_n2.err; echo "bench n2 exit $?"; tail -c 700 gpurun_out/r2t_bench_c2_n2.json; tail -n 2 gpurun_out/r2t_bench_c2_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/check_shard_driver.py 200000 c2_slice > gpurun_out/r2t_shard_driver_n2.txt 2>&1; echo "shard driver exit $?" >> gpurun_out/r2t_shard_driver_n2.txt; tail -n 4 gpurun_out/r2t_shard_driver_n2.txt | cut -c1-250
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_shard_p2p.py > gpurun_out/r2t_shard_p2p_n2.txt 2>&1; echo "shard p2p exit $?" >> gpurun_out/r2t_shard_p2p_n2.txt; tail -n 2 gpurun_out/r2t_shard_p2p_n2.txt | cut -c1-200
