#!/bin/bash
# Final single-GPU validation: both GPU test suites, smoke, full default bench with a compact summary.
bash tools/gpu_check.sh nobench > /dev/null 2>&1
tail -n 4 gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
tail -c 300 gpurun_out/bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_full.log").read().strip().splitlines()[-1])
print("value %.0f ms %.2f e2e %.0f kernel_ms %.2f frac %.3f launches %s clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k, v in d["secondary"].items():
    print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "ms", "unit")},
          v.get("roofline", {}).get("frac"))
PY
