# A/B of launch-level switches of the fused loss step: L2 prefetch of the projection by the prologue, programmatic dependent launch
for v in "" "TRB_FUSED_NO_PREFETCH=1" "TRB_FUSED_PDL=0"; do
  echo "== $v"; env $v python tools/fused_probe.py child 7 0 2>&1 | grep "FAIL\|TIME\|rror"
done
TRB_FUSED_DEBUG=1 python tools/fused_probe.py child 7 0 2>&1 | grep "STAMPS.*inst"
