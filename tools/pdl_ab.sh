for v in 1 0; do
  echo "== TRB_FUSED_PDL=$v"; TRB_FUSED_PDL=$v python tools/fused_probe.py child 7 0 2>&1 | grep "PASS\|FAIL\|TIME\|rror"
done
TRB_FUSED_DEBUG=1 python tools/fused_probe.py child 7 0 2>&1 | grep "STAMPS"
