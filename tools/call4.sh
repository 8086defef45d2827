set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --timeout 900 -x -s -k "config4" > gpurun_out/pytest_cfg.log 2>&1
echo "pytest exit $?"
tail -n 30 gpurun_out/pytest_cfg.log
