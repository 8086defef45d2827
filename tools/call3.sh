set -u
mkdir -p gpurun_out
timeout 300 python tools/tc_probe.py quick 2>&1 | tee gpurun_out/tc_quick.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:retrieval_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_stream python tools/tc_probe.py one > gpurun_out/ncu_stream.log 2>&1
echo "ncu exit $?"
