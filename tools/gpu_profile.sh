#!/bin/bash
# ncu evidence for the headline kernel (1 GPU): launch list of a bench run + one full capture of the stream kernel.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_bench.log 2>&1
echo "launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:retrieval_tc_kernel -s 6 -c 2 -f -o gpurun_out/prof_retrieval $B > gpurun_out/prof_bench.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
