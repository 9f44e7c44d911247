#!/bin/bash
# N-GPU evidence run (under gpurun --gpus N): the bench line with config.extra (configs[2] strong-scaled, configs[4] at N = 8,
# sharded driver / k-means checks) and the peer-memory dedup equality check.   usage: tools/evidence_ngpu.sh N
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_c2_n$N.out 2> gpurun_out/bench_c2_n$N.err
echo "bench exit $?"; grep '^{' gpurun_out/bench_c2_n$N.out | tail -1 > gpurun_out/bench_c2_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_c2_n$N.json"))
print("N=$N value", round(d["value"] / 1e6, 1), "M/s  ms", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"] / 1e6, 2), " ids==1gpu", d["config"]["multi_gpu_ids_equal_single_gpu"],
      {k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()})
for k, v in (d["config"].get("extra") or {}).items():
    print("  extra", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if not isinstance(vv, (dict, list, str))})
PY
[ -n "$SKIP_P2P" ] || timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/check_shard_p2p.py 2>&1 | grep -E "^case|peer_memory|nccl_all" > gpurun_out/shard_peer_dedup_n$N.txt; cat gpurun_out/shard_peer_dedup_n$N.txt
