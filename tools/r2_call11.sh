#!/bin/bash
# 8-GPU validation: bench at N=8 (weak-scaled C2 + config.extra: configs[2] 10 M strong-scaled, configs[4] 100 M), sharded driver, peer dedup
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2k_bench_c2_n8.json 2> gpurun_out/r2k_bench_c2_n8.err; echo "bench n8 exit $?"; tail -c 1200 gpurun_out/r2k_bench_c2_n8.json; tail -n 3 gpurun_out/r2k_bench_c2_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/check_shard_driver.py 400000 c2_slice > gpurun_out/r2k_shard_driver_n8.txt 2>&1; echo "shard driver exit $?" >> gpurun_out/r2k_shard_driver_n8.txt; tail -n 3 gpurun_out/r2k_shard_driver_n8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/check_shard_p2p.py > gpurun_out/r2k_shard_p2p_n8.txt 2>&1; echo "shard p2p exit $?" >> gpurun_out/r2k_shard_p2p_n8.txt; tail -n 3 gpurun_out/r2k_shard_p2p_n8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 10 --warmup 3 --no-extra > gpurun_out/r2k_bench_c2_n4.json 2> gpurun_out/r2k_bench_c2_n4.err; echo "bench n4 exit $?"
