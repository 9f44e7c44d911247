"""Share of every kernel in the TIMED steps of a bench run, from the ncu launch list of the same command
(ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python bench.py --steps 2 --warmup 1 ...).

    python tools/launch_shares.py gpurun_out/launches.csv [warmup=1] [steps=2] > profiles/..._launch_shares.txt

A step starts at a full-catalogue launch of the first-layer kernel (the longest launches of the list); ncu's per-launch times
are cold-cache and serialised, so the SHARE per kernel is what has to agree with the bench line, not the absolute time.
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    warmup = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        name = re.sub(r"^.*?::(\(anonymous namespace\)|<unnamed>)::", "", r["Kernel Name"])
        name = re.sub(r"\(.*$", "", name).replace("void ", "")
        rows.append((name, us))
    first = [n for n in ("linear_tf32_kernel", "linear_tc2_kernel") if any(r[0].startswith(n) for r in rows)][0]
    longest = max(us for n, us in rows if n.startswith(first))
    starts = [i for i, (n, us) in enumerate(rows) if n.startswith(first) and us > 0.6 * longest]
    assert len(starts) >= warmup + steps, (len(starts), warmup, steps)
    per_step = starts[warmup + 1] - starts[warmup] if steps > 1 else starts[warmup] - starts[warmup - 1]
    lo, hi = starts[warmup], starts[warmup] + per_step * steps
    agg = OrderedDict()
    for n, us in rows[lo:hi]:
        a = agg.setdefault(n, [0.0, 0])
        a[0] += us
        a[1] += 1
    total = sum(a[0] for a in agg.values())
    print(f"launch list: {path}; the {steps} TIMED steps = launches {lo} .. {hi - 1} ({per_step} launches per step)\n")
    for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{100 * us / total:6.2f} %  {us / steps:10.1f} us per step  x{c // steps:<3d} {n}")
    print(f"\nsum {total / steps:.1f} us per step")


if __name__ == "__main__":
    main()
