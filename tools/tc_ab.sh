#!/bin/bash
# A/B timing of stream-kernel build variants (TRB_LIB selects the library): python tools/tc_probe.py per variant
# usage: tools/tc_ab.sh [variant library names without .so ...]   ("" = the shipped library is always run first)
set -u
mkdir -p gpurun_out
for lib in "" "$@"; do
  if [ -n "$lib" ]; then export TRB_LIB=$PWD/textreid_b200/$lib.so; else unset TRB_LIB; fi
  echo "== variant ${lib:-default}" | tee -a gpurun_out/tc_ab.log
  timeout 300 python tools/tc_probe.py quick 2>&1 | tee -a gpurun_out/tc_ab.log
done
unset TRB_LIB
