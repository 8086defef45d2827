set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?"
tail -n 30 gpurun_out/pytest_all.log
