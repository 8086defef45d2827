set -u
mkdir -p gpurun_out
rm -f gpurun_out/tc_ab.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "bf16 or sharded or retriev" -x > gpurun_out/pytest_tc.log 2>&1
echo "tc exit $?"
tail -n 5 gpurun_out/pytest_tc.log
bash tools/tc_ab.sh > /dev/null 2>&1
cat gpurun_out/tc_ab.log
