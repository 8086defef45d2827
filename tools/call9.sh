set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loss.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_loss.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_loss.log; grep -n "AssertionError\|Error" gpurun_out/pytest_loss.log | head -5
for pdl in 0 1; do
  if [ $pdl = 0 ]; then export TRB_FUSED_NO_PDL=1; else unset TRB_FUSED_NO_PDL; fi
  echo "== PDL=$pdl"
  timeout 300 python tools/fused_probe.py child 7 0 2>&1 | grep "TIME\|FAIL"
  python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench, textreid_b200 as trb
pk = bench.peaks(); dev = torch.device("cuda", 0)
fb = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for key in ("n128_k2048", "n256_k4096"):
    r = bench.loss_step_line(trb, pk, dev, lambda: fb.fill_(1), key, "bf16", "stepgraph", iters=30)
    print(key, "stepgraph %.1f us (best %.1f) launches %d" % (r["ms_per_step"] * 1e3, r["ms_best"] * 1e3, r["library_launches"]))
PY
done
