set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?"
tail -n 25 gpurun_out/pytest_all.log
for enq in 0 1; do
echo "== TRB_PROBE_ENQUEUE=$enq"
TRB_PROBE_ENQUEUE=$enq timeout 300 python tools/fused_probe.py child 7 0 2>&1 | grep TIME
done
TRB_FUSED_DEBUG=1 timeout 300 python tools/fused_probe.py child 7 0 2>&1 | grep "STAMPS" | tail -6
