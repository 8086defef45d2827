set -u
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2.log 2> gpurun_out/bench_r2.err
echo "bench exit $?"
tail -c 1500 gpurun_out/bench_r2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2.log").read().strip().splitlines()[-1])
print("value %.0f ms %.2f e2e %.0f (%.2f ms) kernel_ms %.2f frac %.3f launches %s clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]))
print("result", d["result"])
print("cpu", d.get("cpu_baseline"))
print("loss_step", d.get("loss_step"))
for k, v in d["secondary"].items():
    if isinstance(v, dict):
        print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "ms_best", "ms", "unit", "ms_per_call", "first_evaluation_ms", "tflops", "library_launches", "error", "frac_of_bf16_sustained")},
              {kk: round(vv, 4) for kk, vv in v.get("roofline", {}).items() if isinstance(vv, float)} if isinstance(v.get("roofline"), dict) else None)
PY
