#!/usr/bin/env python
"""bench.py — RQ-VAE semantic-ID encode items/s on B200 (BASELINE.json metric), one JSON line.

A "step" is one pass of the hot path over the whole synthetic catalogue shard of this rank:
get_indices(use_sk=False) over every item (encoder MLP + residual quantizer) followed by the suffix-code
dedup, inputs resident in HBM → final [N, L+1] int64 semantic ids resident in HBM.
  N=1 workload: BASELINE configs[1] — 1M items x 768-d, 3 levels x 256 codes, e_dim 32.
  N>1: weak scaling, every rank encodes its own 1M-item shard of an N x 1M catalogue; the dedup is global
       (packed codes stored into the key owner's buffer over NVLink peer memory by the partition kernel itself,
       ranks written back the same way) so ids are unique catalogue-wide.
`e2e` is the same work through the host-buffer C-ABI call (rqb200_generate_codes_host): pinned host
embeddings in, host semantic ids out, H2D/D2H inside the timed region.
`--impl reference` times the reference's own CPU implementation of the step on the box's host cores, on a bounded sample
of the same workload: the REAL PyTorch `RQVAE.get_indices` (reference RQ-VAE/models/rqvae.py:67-71, copied unmodified
into the git-ignored oracle/_ref/ by oracle/make_ref.py; all host threads, batch 65 536) followed by the suffix column
(oracle C restatement of infer.py:152-163 — the reference's own O(N·G) loop is timed separately on a bounded number of
duplicated codes and reported, not run in full).  Without oracle/_ref it falls back to the oracle C port (kind "port").
`config.extra` (N > 1): BASELINE configs[2] (10 M items, strong-scaled over the N GPUs) and, at N = 8, configs[4] (100 M).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rqvae_semantic_id_encode_items_per_s"
UNIT = "items/s"
WORKLOAD = "BASELINE configs[1]: 1M items x 768-d, 3 levels x 256 codes, e_dim 32, layers [256,128]"
N_PER_GPU = 1_000_000
SEED = 2024
# --config selects the shapes (default c2 = the configuration the metric is quoted on; c3 / c5 are documentation runs)
CONFIGS = {
    "c2": ("c2_slice", WORKLOAD),
    "c3": ("c3_slice", "BASELINE configs[2] shapes: 768-d, 4 levels x 256 codes, e_dim 64, layers [256,128]"),
    "c5": ("c5_slice", "BASELINE configs[4] shapes: 1024-d, 4 levels x 1024 codes, e_dim 64, layers [256,128]"),
}
GOLDEN = "c2_slice"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def golden_model(device, with_sk=False):
    from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden
    g, cfg, cbs = load_golden(GOLDEN)
    if not with_sk:
        cfg = dict(cfg, sk_epsilons=[0.0] * len(cfg["num_emb_list"]))  # bench step = pass 1 + suffix dedup (no Sinkhorn rounds)
    return build_model(cfg, cbs, device=device), cfg, cbs


def sample_rows_host(sample, in_dim):
    import numpy as np
    from ai_education_generative_recommendation_b200 import synth
    return np.concatenate([synth.synth_items(SEED, r0, min(65536, sample - r0), in_dim, N_PER_GPU)
                           for r0 in range(0, sample, 65536)])


_REF_MODEL = {}


def reference_torch_model():
    """The UNMODIFIED reference RQVAE (oracle/_ref/models, see oracle/make_ref.py) with the bench weights, or None."""
    if GOLDEN in _REF_MODEL:
        return _REF_MODEL[GOLDEN]
    model = None
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref_dir, "models", "rqvae.py")):
        import torch
        from ai_education_generative_recommendation_b200.fixtures import load_golden, synth_weights
        sys.path.insert(0, ref_dir)
        try:
            from models.rqvae import RQVAE as RefRQVAE
            g, cfg, cbs = load_golden(GOLDEN)
            sd, _, _ = synth_weights(cfg)
            for l, c in enumerate(cbs):
                sd[f"rq.vq_layers.{l}.embedding.weight"] = c
            model = RefRQVAE(in_dim=cfg["in_dim"], num_emb_list=cfg["num_emb_list"], e_dim=cfg["e_dim"], layers=cfg["layers"],
                             dropout_prob=0.0, bn=False, loss_type="mse", kmeans_init=False,
                             sk_epsilons=[0.0] * len(cfg["num_emb_list"]), sk_iters=cfg["sk_iters"])
            model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()})
            model.eval()
        finally:
            sys.path.remove(ref_dir)
    _REF_MODEL[GOLDEN] = model
    return model


def cpu_reference_leg(sample_rows, threads, x_host=None, kind="auto"):
    """The reference path on the host cores → (items/s, seconds, ids, kind).  kind "reference": the real PyTorch
    get_indices (all threads, batch 65 536) + suffix column; "port": the oracle C restatement of both."""
    import numpy as np
    from ai_education_generative_recommendation_b200.fixtures import load_golden, synth_weights
    from oracle import oracle as O
    O.build()
    g, cfg, cbs = load_golden(GOLDEN)
    if x_host is None:
        x_host = sample_rows_host(sample_rows, cfg["in_dim"])
    x = np.ascontiguousarray(x_host[:sample_rows])
    ref = reference_torch_model() if kind in ("auto", "reference") else None
    if ref is not None:
        import torch
        torch.set_num_threads(threads)
        xt = torch.from_numpy(x)
        with torch.no_grad():
            ref.get_indices(xt[:4096], use_sk=False)               # warm-up
            t0 = time.perf_counter()
            codes = torch.cat([ref.get_indices(xt[i:i + 65536], use_sk=False) for i in range(0, sample_rows, 65536)]).numpy()
        ids = O.suffix_dedup(codes)
        dt = time.perf_counter() - t0
        return sample_rows / dt, dt, ids, "reference"
    _, (ew, eb), _ = synth_weights(cfg)
    O.get_indices(x[:4096], ew, eb, cbs, threads=threads)     # warm-up
    t0 = time.perf_counter()
    codes = O.get_indices(x, ew, eb, cbs, threads=threads)
    ids = O.suffix_dedup(codes)
    dt = time.perf_counter() - t0
    return sample_rows / dt, dt, ids, "port"


def reference_dedup_lane(codes, max_dups=200):
    """B-dedup lane: the reference's own suffix loop (infer.py:152-163: np.unique + one np.where scan per
    duplicated code), verbatim, on at most `max_dups` duplicated codes → seconds per duplicated code at this N."""
    import numpy as np
    arr = np.hstack((codes, np.zeros((codes.shape[0], 1), dtype=int)))
    t0 = time.perf_counter()
    unique_codes, counts = np.unique(arr, axis=0, return_counts=True)
    t_unique = time.perf_counter() - t0
    duplicates = unique_codes[counts > 1]
    t0 = time.perf_counter()
    for duplicate in duplicates[:max_dups]:
        duplicate_indices = np.where((arr == duplicate).all(axis=1))[0]
        for i, idx in enumerate(duplicate_indices):
            arr[idx, -1] = i
    per_dup = (time.perf_counter() - t0) / max(1, min(max_dups, len(duplicates)))
    return {"n": int(codes.shape[0]), "np_unique_s": t_unique, "duplicated_codes": int(len(duplicates)),
            "s_per_duplicated_code": per_dup, "projected_total_s": t_unique + per_dup * len(duplicates)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 262144
    from ai_education_generative_recommendation_b200.fixtures import load_golden
    in_dim = json.loads(str(load_golden(GOLDEN)[0]["cfg"]))["in_dim"]
    x_host = sample_rows_host(sample, in_dim)          # generated once, outside the timed steps
    vals = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        v, dt, ids, kind = cpu_reference_leg(sample, threads, x_host=x_host)
        if i >= args.warmup:
            vals.append((v, dt))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(dt for _, dt in vals) / len(vals)
    what = ("the reference's own PyTorch RQVAE.get_indices (oracle/_ref, unmodified; batch 65536, all host threads) + suffix column "
            "(oracle C restatement of infer.py:152-163)" if kind == "reference"
            else "oracle C restatement of get_indices + suffix dedup, pthreads (oracle/_ref absent)")
    import torch
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{sample} rows of the same catalogue per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"{sample} rows/step: {what}", "torch": torch.__version__,
                             "reference_dedup_lane": reference_dedup_lane(ids[:, :-1])},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


KEEP_ALIVE = []


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off BEFORE the pinned host buffers are allocated and
    first touched, so that the e2e leg's H2D stream reads node-local memory (at N = 8 eight ranks stream 3 GB each)."""
    try:
        import torch
        props = torch.cuda.get_device_properties(local)
        bdf = None
        if hasattr(props, "pci_bus_id"):
            bdf = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, getattr(props, "pci_device_id", 0))
        else:
            q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                               capture_output=True, text=True, timeout=10).stdout.strip().lower()
            bdf = q[-12:] if q else None
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return {"numa_node": None, "note": "no NUMA information for the GPU"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as exc:          # placement is an optimisation
        return {"numa_node": None, "note": f"not bound ({type(exc).__name__})"}


def run_extra(key, n_total, rank, world, dev, group, args):
    """BASELINE configs[2] (10 M x 768, 4 x 256 codes, e 64) / configs[4] (100 M x 1024, 4 x 1024 codes, e 64): the catalogue
    is split into contiguous shards over the `world` GPUs (strong scaling), every rank encodes its shard (tensor-core
    route) and the suffix column is global (peer-memory dedup).  Returns the per-config record (rank 0) — items/s over all
    GPUs, per-stage ms of rank 0, and a sampled fast == exact check on this rank's first rows."""
    import torch
    import torch.distributed as dist
    import ai_education_generative_recommendation_b200 as rq
    from ai_education_generative_recommendation_b200 import _cabi, sharding
    global GOLDEN, WORKLOAD
    saved = (GOLDEN, WORKLOAD)
    GOLDEN, WORKLOAD = CONFIGS[key]
    try:
        lib = _cabi.lib()
        model, cfg, cbs = golden_model(dev)
        model.encode_mode = _cabi.ENCODE_FAST
        lo, hi = sharding.shard_range(n_total, rank, world)
        n = hi - lo
        in_dim, Ks = cfg["in_dim"], cfg["num_emb_list"]
        x = torch.empty((n, in_dim), dtype=torch.float32, device=dev)
        _cabi.check(lib.rqb200_synth_items(SEED, lo, n, in_dim, n_total, x.data_ptr(), _cabi.stream_ptr()))
        peer = sharding.PeerShardDedup(model, group, max_local_items=n + 1)

        def step():
            return peer(model.get_indices(x, use_sk=False), Ks)

        for _ in range(2):
            out = step()
        dist.barrier(); torch.cuda.synchronize()
        lib.rqb200_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            out = step()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        prof_ms = (ctypes.c_double * 12)()
        prof_cnt = (ctypes.c_longlong * 12)()
        lib.rqb200_profile_read(prof_ms, prof_cnt, 12)
        lib.rqb200_profile_enable(0)
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # sampled parity: tensor-core route == exact route on this rank's first rows; ids unique catalogue-wide
        ns = min(n, 200_000)
        fast_codes = model.get_indices(x[:ns], use_sk=False)
        model.encode_mode = _cabi.ENCODE_EXACT
        exact_codes = model.get_indices(x[:ns], use_sk=False)
        ok = torch.tensor([int(torch.equal(fast_codes, exact_codes))], dtype=torch.int64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        dup = torch.tensor([int((out[:, -1] > 0).sum())], dtype=torch.int64, device=dev)
        dist.all_reduce(dup, op=dist.ReduceOp.SUM)
        dist.barrier(); torch.cuda.synchronize()
        KEEP_ALIVE.append(peer)          # peer-mapped buffers stay mapped until the process exits (all ranks tear down together)
        del x, out, model
        torch.cuda.empty_cache()
        step_ms = float(ms.item())
        return {"workload": WORKLOAD, "global_items": n_total, "items_per_gpu": n, "scaling": "strong", "ms_per_step": step_ms,
                "value": n_total / (step_ms * 1e-3), "unit": UNIT,
                "stage_ms_rank0": {"tc_linear0": prof_ms[4] / reps, "tc_linear_rest": prof_ms[6] / reps,
                                   "quantize": prof_ms[2] / reps, "rerun_tier": prof_ms[7] / reps,
                                   "exact_rescue_tier": prof_ms[8] / reps, "dedup": prof_ms[3] / reps},
                "hbm_frac_layer1_rank0": (4.0 * in_dim * n / (prof_ms[4] / reps * 1e-3) / 1e9 / float(load_peaks()[0]["hbm_gbs"]))
                if prof_ms[4] > 0 else None,
                "fast_equals_exact_on_sample": bool(ok.item()), "sample_rows_per_gpu": ns,
                "items_with_nonzero_suffix": int(dup.item())}
    finally:
        GOLDEN, WORKLOAD = saved


def run_sharded_checks(rank, world, dev, group):
    """N > 1, driver-visible evidence for the two other cross-GPU steps of the path: the Sinkhorn re-encode
    rounds over a sharded catalogue (ids must equal the single-GPU driver's) and the sharded k-means statistics all-reduce
    (centres must equal a single-GPU fit of the gathered samples)."""
    import torch
    import torch.distributed as dist
    import ai_education_generative_recommendation_b200 as rq
    from ai_education_generative_recommendation_b200 import _cabi, sharding
    from ai_education_generative_recommendation_b200.fixtures import build_model, load_golden
    lib = _cabi.lib()
    out = {}
    g, cfg, cbs = load_golden("c2_slice")                    # sk_epsilons [0, 0, 0.003]: the re-encode rounds are live
    n_total = 200_000
    lo, hi = sharding.shard_range(n_total, rank, world)
    model = build_model(cfg, cbs, device=dev)
    x = torch.empty((hi - lo, cfg["in_dim"]), dtype=torch.float32, device=dev)
    _cabi.check(lib.rqb200_synth_items(SEED, lo, hi - lo, cfg["in_dim"], 1_000_000, x.data_ptr(), _cabi.stream_ptr()))
    # warm-up on a small slice: NCCL opens its point-to-point connections on the first all-to-all between each pair of ranks
    # (seconds at N = 8), which is set-up, not the driver
    sharding.generate_codes_sharded(build_model(cfg, cbs, device=dev), x[:min(2048, hi - lo)].contiguous(), group)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    mine, stats = sharding.generate_codes_sharded(model, x, group)
    torch.cuda.synchronize(); dist.barrier()
    t_sh = time.perf_counter() - t0
    pad = max(b - a for a, b in (sharding.shard_range(n_total, r, world) for r in range(world)))
    buf = torch.full((pad, mine.shape[1]), -7, dtype=torch.int64, device=dev)
    buf[:mine.shape[0]] = mine
    xb = torch.zeros((pad, x.shape[1]), dtype=torch.float32, device=dev)
    xb[:x.shape[0]] = x
    ids_all = [torch.empty_like(buf) for _ in range(world)]
    x_all = [torch.empty_like(xb) for _ in range(world)]
    dist.all_gather(ids_all, buf)
    dist.all_gather(x_all, xb)
    if rank == 0:
        sizes = [sharding.shard_range(n_total, r, world) for r in range(world)]
        full = torch.cat([ids_all[r][:b - a] for r, (a, b) in enumerate(sizes)])
        xs = torch.cat([x_all[r][:b - a] for r, (a, b) in enumerate(sizes)])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref, rstats = rq.generate_codes(build_model(cfg, cbs, device=dev), xs)
        torch.cuda.synchronize()
        out["sharded_driver"] = {"items": n_total, "rounds": stats["rounds"], "ids_equal_single_gpu": bool(torch.equal(ref, full)),
                                 "seconds_sharded": t_sh, "seconds_single_gpu": time.perf_counter() - t0,
                                 "what": "sharding.generate_codes_sharded: passes 1-3 incl. the <= 30 Sinkhorn re-encode rounds, codes / group "
                                         "sizes / residuals exchanged per round (NCCL all-to-all), vs rq.generate_codes on one GPU"}
    del ids_all, x_all, xb, buf
    # sharded k-means: per-cluster sums / counts all-reduced every Lloyd iteration
    e, K, per = 64, 256, 131072
    xs = torch.empty((per, e), dtype=torch.float32, device=dev)
    _cabi.check(lib.rqb200_synth_items(SEED, rank * per, per, e, per * world, xs.data_ptr(), _cabi.stream_ptr()))
    gathered = [torch.empty_like(xs) for _ in range(world)]
    dist.all_gather(gathered, xs)
    init = torch.cat(gathered)[:: (per * world) // K][:K].contiguous()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    c_sh = rq.kmeans(xs, K, 10, init=init, group=group, tol=0.0)
    torch.cuda.synchronize(); dist.barrier()
    t_km = time.perf_counter() - t0
    if rank == 0:
        c_one = rq.kmeans(torch.cat(gathered), K, 10, init=init, tol=0.0)
        out["sharded_kmeans"] = {"samples": per * world, "e": e, "K": K, "lloyd_iterations": 10, "seconds": t_km,
                                 "max_abs_diff_vs_single_gpu": float((c_sh - c_one).abs().max()),
                                 "bit_equal_single_gpu": bool(torch.equal(c_sh, c_one)),
                                 "what": "rq.kmeans(group=WORLD): fp64 per-cluster sums + counts all-reduced (NCCL) per iteration"}
    del gathered
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("RQB200_MODE", "auto"), choices=["auto", "exact", "fast"])
    ap.add_argument("--items", type=int, default=N_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--screen", default="auto", choices=["auto", "off", "tf32", "fp16"],
                    help="screening tier of the tensor-core route: auto = the library's rule (TF32 screen where it pays)")
    ap.add_argument("--no-extra", action="store_true", help="N > 1: skip the configs[2] / configs[4] runs under config.extra")
    ap.add_argument("--no-full-driver", action="store_true", help="skip the full_driver leg (passes 1-3 with Sinkhorn rounds)")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--e2e-items", type=int, default=0,
                    help="rows per rank of the host-buffer (e2e) leg; 0 = the same rows as the device leg, capped at 2M "
                         "(catalogue-scale runs would otherwise pin tens of GB of host memory per rank)")
    args = ap.parse_args()
    global GOLDEN, WORKLOAD
    GOLDEN, WORKLOAD = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import ai_education_generative_recommendation_b200 as rq
    from ai_education_generative_recommendation_b200 import _cabi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        group = dist.group.WORLD
    lib = _cabi.lib()
    model, cfg, cbs = golden_model(dev)
    fast_ok = True
    if args.mode in ("auto", "fast"):
        model.encode_mode = _cabi.ENCODE_FAST
        try:
            model.get_indices(torch.zeros((256, cfg["in_dim"]), device=dev))
        except _cabi.RQB200Error as exc:
            if args.mode == "fast":
                raise
            fast_ok = False
            model.encode_mode = _cabi.ENCODE_EXACT
    else:
        fast_ok = False
    if args.screen != "auto":
        model.set_screen({"off": 0, "tf32": "tf32", "fp16": 1}[args.screen])
    mode_name = ("fast(TMA-fed TF32 screening pass + tcgen05 split-fp16 three-pass tier + margin gates + exact rescue)"
                 if fast_ok else "exact(SIMT fp32, reference order)")

    n = args.items
    n_total = n * world
    lo = rank * n
    in_dim, n_levels, first_out = cfg["in_dim"], len(cfg["num_emb_list"]), cfg["layers"][0]
    x = torch.empty((n, in_dim), dtype=torch.float32, device=dev)
    _cabi.check(lib.rqb200_synth_items(SEED, lo, n, in_dim, n_total, x.data_ptr(), _cabi.stream_ptr()))
    Ks = cfg["num_emb_list"]
    # N > 1: the suffix column is global — keys travel to their owner rank through NVLink peer memory
    # (rqb200_shard_suffix_dedup: the library's own kernels store into the peers' buffers, no NCCL call per step)
    peer_dedup = sharding.PeerShardDedup(model, group, max_local_items=n) if world > 1 else None

    def step():
        codes = model.get_indices(x, use_sk=False)
        if world == 1:
            out, _ = rq.suffix_dedup(model, codes, want_stats=False)      # enqueue only: the collision statistics are read once, after the timed steps
        else:
            out = peer_dedup(codes, Ks)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~100 ms to start: begin during the warm-up, keep sampling through the timed steps
    multi_check = None
    if world > 1:
        # one-off verification: the N-GPU ids equal a single-GPU dedup of the gathered catalogue
        codes = model.get_indices(x, use_sk=False)
        mine = peer_dedup(codes, Ks)
        gathered = [torch.empty_like(codes) for _ in range(world)]
        gathered_ids = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, codes)
        dist.all_gather(gathered_ids, mine)
        if rank == 0:
            ref, _ = rq.suffix_dedup(model, torch.cat(gathered))
            multi_check = bool(torch.equal(ref, torch.cat(gathered_ids)))
            assert multi_check, "multi-GPU semantic ids differ from the single-GPU result"
        del gathered, gathered_ids
    for _ in range(max(args.warmup, 3)):
        out = step()
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.6:      # same workload, untimed: keeps the GPU under load until the sampler is running
        out = step()
    barrier()
    lib.rqb200_profile_enable(1)
    launches0 = lib.rqb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    launches = lib.rqb200_launch_count() - launches0
    prof_ms = (ctypes.c_double * 12)()
    prof_cnt = (ctypes.c_longlong * 12)()
    lib.rqb200_profile_read(prof_ms, prof_cnt, 12)
    lib.rqb200_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    tiers = model.last_stats if fast_ok else {}          # tier row counts of the LAST TIMED step (the e2e leg below runs in chunks)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / args.steps
    value = n_total / (ms_step * 1e-3)

    # ---- e2e through the host-buffer C-ABI call (pinned host in, host out), same workload per rank
    ne = args.e2e_items if args.e2e_items > 0 else min(n, 2_000_000)
    ne = min(ne, n)
    xh = torch.empty((ne, in_dim), dtype=torch.float32).pin_memory()
    xh.copy_(x[:ne])
    torch.cuda.synchronize()
    ids_host = torch.empty((ne, len(Ks) + 1), dtype=torch.int64).pin_memory()
    stats = (ctypes.c_int64 * 4)()
    mode_id = _cabi.ENCODE_FAST if fast_ok else _cabi.ENCODE_EXACT

    def e2e_step():
        _cabi.check(lib.rqb200_generate_codes_host(model._handle, mode_id, xh.data_ptr(), ne, 131072, ids_host.data_ptr(), stats,
                                                   _cabi.stream_ptr()))

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = ne * world / float(e2e_s.item())
    os.sched_setaffinity(0, all_cpus)          # the CPU baseline below gets every host core again
    if world == 1 and ne == n:
        assert np.array_equal(ids_host.numpy(), out.cpu().numpy()), "host-buffer path and device path disagree"

    # ---- the other BASELINE configurations, strong-scaled over this run's GPUs (N > 1 only; reported under config.extra)
    extras = {}
    if world > 1 and not args.no_extra:
        del xh, ids_host
        x = None
        out = None
        barrier()
        torch.cuda.empty_cache()
        plan = [("c3", 10_000_000)] + ([("c5", 100_000_000)] if world >= 8 else [])
        for key, total in plan:
            extras[key] = run_extra(key, total, rank, world, dev, group, args)
        extras.update(run_sharded_checks(rank, world, dev, group) or {})
    if rank == 0:
        peaks, peak_src = load_peaks()
        # dominant kernel: the encoder's first Linear — linear_tc2_kernel (fast route) / linear_exact_kernel (exact route).
        # Algorithmic work of that kernel per item: 4*in_dim bytes read, 2*in_dim*256 flop.
        slot = 4 if (fast_ok and prof_cnt[4] > 0) else 0
        k_ms = prof_ms[slot] / max(prof_cnt[slot], 1)
        alg_bytes = 4.0 * in_dim * n
        alg_flops = 2.0 * in_dim * first_out * n
        screen_active = fast_ok and tiers.get("three_pass_rows", 0) > 0     # the TF32 screening tier ran
        l1_kernel = "linear_tf32_kernel" if screen_active else "linear_tc2_kernel"
        l1_desc = ("encoder layer 1 over every row: ONE tcgen05 kind::tf32 pass, cta_group::2, operands delivered by TMA "
                   "(cp.async.bulk.tensor) straight from the fp32 rows" if screen_active
                   else "encoder layer 1: tcgen05 cta_group::2, split-fp16, 3 MMA passes")
        kname = (f"{l1_kernel} ({l1_desc})" if slot == 4
                 else "linear_exact_kernel<128,128,8,8> (encoder layer 1, fp32 FMA chains)")
        traffic = None
        for tname in ("r2_traffic.json", "r1_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):       # dram bytes per launch from the committed ncu --set full capture (null if not captured)
                traffic = json.load(open(tpath)).get(l1_kernel if slot == 4 else "linear_exact_kernel")
                if traffic is not None:
                    break
        gbs = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        tfl = alg_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        hbm_peak = float(peaks["hbm_gbs"])
        tc_peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        stage = {"tc_linear0": prof_ms[4] / args.steps, "tc_linear_rest": prof_ms[6] / args.steps,
                 "exact_linear0": prof_ms[0] / args.steps, "exact_linear_rest": prof_ms[1] / args.steps,
                 "quantize_incl_rescue_quantizer": prof_ms[2] / args.steps, "dedup": prof_ms[3] / args.steps,
                 "three_pass_rerun_tier": prof_ms[7] / args.steps, "exact_rescue_tier": prof_ms[8] / args.steps}
        def _st(name, bound, ms, work, peak, unit, note):
            ach = work / (ms * 1e-3) / (1e9 if unit == "GB/s" else 1e12) if ms > 0 else 0.0
            return {"stage": name, "bound": bound, "ms": ms, "achieved": ach, "peak": peak, "unit": unit,
                    "frac": ach / peak if peak else None, "algorithmic_work_per_item": work / n, "note": note}
        e_dim = int(model.e_dim)
        stages = []
        if slot == 4:           # per-stage view of the fast route (algorithmic work only, per item)
            stages = [
                _st(f"encoder layer 1 ({l1_kernel})", "hbm", stage["tc_linear0"], 4.0 * in_dim * n, hbm_peak, "GB/s",
                    "reads X once: 4*in_dim B"),
                _st("encoder layers 2+3 fused (mlp23_tc_kernel)", "hbm", stage["tc_linear_rest"],
                    (4.0 * first_out + 4.0 * e_dim) * n, hbm_peak, "GB/s", "reads H1 (4 B per feature as fp16 hi+lo), writes z"),
                _st("residual quantizer (quantize_tc_kernel + the rescue tier's exact quantizer)", "tensor",
                    stage["quantize_incl_rescue_quantizer"], 2.0 * sum(Ks) * e_dim * n, tc_peak, "TFLOP/s",
                    "2*sum(K)*e flop of ONE pass; the kernel issues 3 fp16 passes and is bound by its per-level latency chain"),
                _st("suffix dedup (radix sort + segmented rank)", "hbm", stage["dedup"], (8.0 * n_levels + 8.0 * (n_levels + 1)) * n,
                    hbm_peak, "GB/s", "reads codes, writes ids; 5 launches (digit histograms, three one-sweep digit passes, rank + output rows), working set in L2"),
            ]
        roofline = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak if hbm_peak else None,
                    "traffic": traffic, "kernel": kname, "kernel_ms": k_ms,
                    "peak_source": f"{peak_src} hbm_gbs; SURVEY §8d: a bf16-rate pass over this shape is HBM-bound (AI 96 flop/B < ridge 207)",
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "tensor_view": {"achieved_tflops": tfl, "peak_tflops": tc_peak, "frac": tfl / tc_peak if tc_peak else None,
                                    "note": ("algorithmic flops of ONE pass = what the TF32 kernel issues; the TF32 rate is half the bf16 rate the peak is quoted at"
                                             if screen_active else "algorithmic flops of ONE pass; the kernel issues 3 fp16 passes for fp32-class accuracy")},
                    "step_share": (prof_ms[slot] / args.steps) / ms_step if ms_step else None,
                    "whole_step_hbm_frac": (value / world) * (4 * in_dim + 8 * n_levels) / (hbm_peak * 1e9),
                    "stage_ms_per_step": stage, "stages": stages,
                    "tier_rows_last_step": ({"rows": n, "three_pass_rerun_rows": tiers.get("three_pass_rows", 0),
                                             "exact_rescue_rows": tiers.get("rescued_rows", 0),
                                             "exact_rescue_fraction": tiers.get("rescued_rows", 0) / n if n else 0.0,
                                             "note": "rows this rank re-ran on the three-pass tier (behind the TF32 screen) and on the "
                                                     "exact SIMT tier in the last step; every other row was certified by a margin gate"}
                                            if fast_ok else None)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "items_per_gpu": n, "global_items": n_total, "encode_mode": mode_name,
                           "step": "get_indices(use_sk=False) + suffix dedup" + (" (global: keys routed to owner ranks over NVLink peer memory)" if world > 1 else ""),
                           "l2": f"inputs {n * in_dim * 4 / 1e9:.2f} GB per GPU per step, larger than the 126 MB L2 (no flush needed)",
                           "parallelism": f"items sharded x{world}, codebooks replicated",
                           "multi_gpu_ids_equal_single_gpu": multi_check, "extra": extras or None},
                "roofline": roofline,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ne * in_dim * 4),
                        "d2h_bytes_per_step": int(ne * (len(Ks) + 1) * 8), "items_per_gpu": ne,
                        "api": "rqb200_generate_codes_host (pinned host buffers)", "host_placement_rank0": numa},
                "gpu_launches": int(launches), "clocks": clocks}
        if world == 1 and not args.no_full_driver:
            # the WHOLE driver (generate_code.py / infer.py semantics): pass 1 + the <= 30 Sinkhorn re-encode rounds
            # (eps on the last level as in the fixture's config) + suffix column, catalogue resident in HBM
            model_sk, cfg_sk, _ = golden_model(dev, with_sk=True)
            ids_sk, st_sk = rq.generate_codes(model_sk, x)                  # warm-up (workspaces, attributes)
            torch.cuda.synchronize()
            lib.rqb200_profile_enable(1)
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                ids_sk, st_sk = rq.generate_codes(model_sk, x)
            torch.cuda.synchronize()
            fd_s = (time.perf_counter() - t0) / reps
            fd_ms = (ctypes.c_double * 12)()
            fd_cnt = (ctypes.c_longlong * 12)()
            lib.rqb200_profile_read(fd_ms, fd_cnt, 12)
            lib.rqb200_profile_enable(0)
            full = {"value": n / fd_s, "unit": UNIT, "seconds": fd_s, "items": n, "sk_epsilons": cfg_sk["sk_epsilons"],
                    "rounds": st_sk["rounds"], "groups_reencoded_per_round": st_sk.get("groups_reencoded_per_round"),
                    "collision_rate_after": st_sk["collision_rate"], "max_conflicts": st_sk["max_conflicts"],
                    "stage_ms": {"group_reencode (small-batch encoder + arg-min levels)": fd_ms[9] / reps,
                                 "sinkhorn (one CTA per group, fp64)": fd_ms[5] / reps,
                                 "pass1_layer1": fd_ms[4] / reps, "pass1_quantize": fd_ms[2] / reps,
                                 "sort / group extraction / suffix (dedup slot)": fd_ms[3] / reps},
                    "what": "rq.generate_codes: pass 1 (tensor-core route) + re-encode rounds (every group through the whole model "
                            "as its own batch, reference arithmetic; fixed-point groups skipped after their first round) + suffix dedup"}
            if not args.no_cpu_baseline:
                from oracle import oracle as O
                from ai_education_generative_recommendation_b200.fixtures import synth_weights
                ns = min(n, 20000)
                _, (ew, eb), _ = synth_weights(cfg_sk)
                xs = x[:ns].cpu().numpy()
                t0 = time.perf_counter()
                try:
                    ref_ids, ref_st = O.generate_codes(xs, ew, eb, cbs, cfg_sk["sk_epsilons"], cfg_sk["sk_iters"], group_order=True)
                    cs = time.perf_counter() - t0
                    sub_ids, _ = rq.generate_codes(model_sk, x[:ns].contiguous())
                    full["cpu_port"] = {"value": ns / cs, "unit": UNIT, "seconds": cs, "items": ns, "rounds": ref_st["rounds"],
                                        "ids_equal_gpu_on_the_same_items": bool(np.array_equal(ref_ids, sub_ids.cpu().numpy())),
                                        "what": "oracle.generate_codes(group_order=True), one host process"}
                except KeyError as exc:
                    # a group size whose reference arithmetic is not restated by the oracle (1024 -> 256 with 16..175 rows: the
                    # reference's own result depends on its thread count there): no CPU twin for this catalogue
                    full["cpu_port"] = {"unavailable": str(exc)}
            line["full_driver"] = full
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = min(ne, 1_000_000)
            v, dt, ids, kind = cpu_reference_leg(sample, threads, x_host=xh.numpy())
            same = bool(np.array_equal(ids[:, :n_levels], ids_host.numpy()[:sample, :n_levels]))
            what = ("the reference's own PyTorch RQVAE.get_indices (oracle/_ref, unmodified; batch 65536, all host threads) + "
                    "suffix column (oracle C restatement)" if kind == "reference"
                    else "oracle C restatement of get_indices + suffix dedup (pthreads)")
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"{sample} rows of the same catalogue, {dt:.1f} s: {what}; codes equal to GPU: {same}"}
            if kind == "reference":      # the C port beside it (what round 1 reported)
                pv, pdt, pids, _ = cpu_reference_leg(sample, threads, x_host=xh.numpy(), kind="port")
                line["cpu_baseline"]["port_value"] = pv
                line["cpu_baseline"]["port_codes_equal_reference"] = bool(np.array_equal(pids, ids))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
