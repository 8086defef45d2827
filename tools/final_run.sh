#!/bin/bash
# Final single-GPU validation: both GPU test suites, smoke, full default bench with a compact summary, then the ncu evidence
# for the loss step (launch list of graph-free steps + one full capture of the fused kernel).
bash tools/gpu_check.sh nobench > /dev/null 2>&1
tail -n 4 gpurun_out/summary.txt
grep -h "passed\|failed\|error" gpurun_out/pytest_simt.log gpurun_out/pytest_tc.log | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
tail -c 300 gpurun_out/bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_full.log").read().strip().splitlines()[-1])
print("value %.0f ms %.2f e2e %.0f kernel_ms %.2f frac %.3f launches %s clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k, v in d["secondary"].items():
    if isinstance(v, dict):
        print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "ms_best", "ms", "unit")},
              v.get("roofline", {}).get("frac") if isinstance(v.get("roofline"), dict) else None)
PY
if [ "${1:-}" != "noncu" ]; then
  TRB_LOSS_PRECISION=bf16 TRB_LOSS_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file gpurun_out/launches_loss_fused.csv python tools/loss_probe.py > gpurun_out/launches_loss_fused.log 2>&1
  echo "loss launch list exit $?"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:fused_loss_kernel --launch-skip 6 --launch-count 1 -f \
      -o gpurun_out/fused_r1_final python tools/fused_probe.py child 7 0 > gpurun_out/ncu_fused.log 2>&1
  echo "fused full capture exit $?"
fi
