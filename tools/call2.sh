set -u
mkdir -p gpurun_out
rm -f gpurun_out/tc_dbg.log
for lib in libtrb_probe16 libtrb_probe8; do
export TRB_LIB=$PWD/textreid_b200/$lib.so
for dbg in 0 4 5; do
  echo "== $lib TRB_TC_DEBUG=$dbg" | tee -a gpurun_out/tc_dbg.log
  TRB_TC_DEBUG=$dbg timeout 300 python tools/tc_probe.py quick 2>&1 | tee -a gpurun_out/tc_dbg.log
done
done
