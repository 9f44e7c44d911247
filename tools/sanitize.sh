#!/bin/bash
# compute-sanitizer passes over the whole hot path on small shapes (the reference has no sanitizer runs; the
# new build should).  One small end-to-end invocation (smoke(): 4096 items, tensor-core route + exact rescue + Sinkhorn
# rounds + dedup, checked against the oracle) under memcheck, racecheck and synccheck.  GPU box only; not run in round 1
# (budget).   gpurun --timeout 1800 -- 'bash tools/sanitize.sh'
mkdir -p gpurun_out
# ONE tool per gpurun call (B200_PROFILING.md): bash tools/sanitize.sh memcheck | racecheck | synccheck
for tool in ${1:-memcheck}; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 \
        python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_sanitizer_$tool.txt 2>&1
    echo "$tool exit $?" >> gpurun_out/r2_sanitizer_$tool.txt
    tail -n 4 gpurun_out/r2_sanitizer_$tool.txt
done
