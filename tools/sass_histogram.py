"""Instruction histogram per kernel of the shipped library (cuobjdump -sass), written to profiles/<tag>_sass_histogram.txt:
for every kernel the instruction count and the counts of the mnemonics that prove which hardware paths it uses
(UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor copies, UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, DFMA / DMUL / DADD = fp64 pipe, FFMA = fp32 pipe, HMMA = legacy mma.sync).
   python tools/sass_histogram.py [tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "ai_education_generative_recommendation_b200", "librqvae_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "DFMA", "DMUL", "DADD",
        "FFMA", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,8}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern]["_total"] += 1
        op = m.group(1)
        hist[kern][op] += 1
        if op in ("UTCHMMA", "UTMALDG") and ".2CTA" in m.group(2):
            hist[kern][op + ".2CTA"] += 1
path = os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.txt")
with open(path, "w") as f:
    f.write(f"cuobjdump -sass {os.path.relpath(SO, ROOT)} (sm_100a) — instructions per kernel; columns: total, then the mnemonics present\n\n")
    for k, c in hist.items():
        cols = [f"{key}={c[key]}" for key in KEYS + ["UTCHMMA.2CTA", "UTMALDG.2CTA"] if c[key]]
        f.write(f"{k}\n    total={c['_total']}  " + "  ".join(cols) + "\n")
    tot = collections.Counter()
    for c in hist.values():
        tot.update(c)
    f.write("\nwhole library: " + "  ".join(f"{key}={tot[key]}" for key in KEYS + ["UTCHMMA.2CTA", "UTMALDG.2CTA"] if tot[key]) + "\n")
print(path)
