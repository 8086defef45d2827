#!/bin/bash
# One gpurun trip: GPU parity tests (SIMT and tensor-core suites in separate processes so that a trap in one
# cannot poison the other), smoke, and a short bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "not bf16 and not sharded" > gpurun_out/pytest_simt.log 2>&1
echo "simt exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "bf16 or sharded" > gpurun_out/pytest_tc.log 2>&1
echo "tc exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
if [ "${1:-}" != "nobench" ]; then
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
  echo "bench exit $?" | tee -a gpurun_out/summary.txt
fi
for f in gpurun_out/pytest_simt.log gpurun_out/pytest_tc.log gpurun_out/smoke.log; do tail -n 3 $f; done
tail -c 3000 gpurun_out/bench.log 2>/dev/null; true
