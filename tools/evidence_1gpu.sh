#!/bin/bash
# One-GPU evidence run (under gpurun): GPU test suite, smoke, bench lines for the three BASELINE shapes, step timeline,
# then the ncu launch list of the same bench command.  Everything lands in gpurun_out/; copy what is to be kept to profiles/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
for c in c2 c3 c5; do
  python bench.py --steps 10 --warmup 3 --config $c > gpurun_out/bench_${c}_n1.json 2> gpurun_out/bench_${c}_n1.err || tail -n 3 gpurun_out/bench_${c}_n1.err
done
python - <<'PY'
import json
for c in ("c2", "c3", "c5"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_{c}_n1.json") if l.startswith("{")][-1])
        print(c, "value", round(d["value"] / 1e6, 1), "M/s  ms", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"] / 1e6, 2),
              " full", round(d["full_driver"]["value"]), " frac", round(d["roofline"]["frac"], 3),
              {k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()}, d["clocks"])
    except Exception as e:
        print(c, "no line:", e)
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference_arm.json
timeout 300 python tools/step_timeline.py c2_slice > gpurun_out/step_timeline_c2.txt 2>&1; tail -n 3 gpurun_out/step_timeline_c2.txt | cut -c1-140
timeout 300 python tools/time_dedup.py c2_slice > gpurun_out/time_dedup_c2.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-full-driver > gpurun_out/ncu_bench.log 2>&1; tail -n 2 gpurun_out/ncu_bench.log | cut -c1-200
# phase trace of the tensor-core quantizer (needs tools/build_variant.sh qt quantize_tc.cu -DRQB_QTC_TRACE beforehand)
QT=$PWD/ai_education_generative_recommendation_b200/librqvae_b200_qt.so
if [ -f "$QT" ]; then
  (RQB200_LIB=$QT timeout 200 python tools/trace_qtc.py c2_slice; RQB200_LIB=$QT timeout 200 python tools/trace_qtc.py c5_slice 400000) > gpurun_out/qtc_phase_trace.txt 2>&1
fi
