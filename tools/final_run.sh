#!/bin/bash
# Round-2 single-GPU validation: both GPU test suites, smoke, the full default bench with a compact summary, then the ncu evidence
# (launch lists of the retrieval step and of the loss step, full captures of the stream kernel and of the fused loss kernel).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -n 1 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
echo "bench exit $?"; tail -c 300 gpurun_out/bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_full.log").read().strip().splitlines()[-1])
print("value %.0f ms %.2f e2e %.0f kernel_ms %.2f frac %.3f launches %s clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]))
print("result", {k: v for k, v in d["result"].items() if k != "parity_sample_detail"})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print("loss_step", d.get("loss_step"))
for k, v in d["secondary"].items():
    if isinstance(v, dict):
        print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "ms_best", "ms", "unit", "ms_per_call", "first_evaluation_ms", "tflops", "library_launches", "error", "frac_of_bf16_sustained", "speedup", "launches_removed_per_step")},
              v.get("roofline", {}).get("frac") if isinstance(v.get("roofline"), dict) else None)
PY
if [ "${1:-}" != "noncu" ]; then
  TRB_BENCH_CUDA_PROFILER=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/r02_launches_retrieval_1m_step.csv \
      python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
  echo "retrieval launch list exit $?"
  TRB_LOSS_PRECISION=bf16 TRB_LOSS_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv \
      --log-file gpurun_out/r02_launches_loss_fused.csv python tools/loss_probe.py > gpurun_out/launches_loss_fused.log 2>&1
  echo "loss launch list exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:retrieval_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_stream \
      python tools/tc_probe.py one > gpurun_out/ncu_stream.log 2>&1
  echo "stream full capture exit $?"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_loss_kernel --launch-skip 6 --launch-count 1 -f \
      -o gpurun_out/r02_loss_fused python tools/fused_probe.py child 7 0 > gpurun_out/ncu_fused.log 2>&1
  echo "fused full capture exit $?"
  TRB_FUSED_DEBUG=1 timeout 300 python tools/fused_probe.py child 7 0 2>&1 | grep "STAMPS\|TIME" > gpurun_out/fused_stamps_r2.log
fi
