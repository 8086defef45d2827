for v in "" libtrb_p8 libtrb_u1 libtrb_noalu libtrb_p2; do
  if [ -n "$v" ]; then export TRB_LIB=$PWD/textreid_b200/$v.so; fi
  echo "== $v"; TRB_FUSED_DEBUG=1 python tools/fused_probe.py child 7 0 2>&1 | grep "STAMPS.hot L2. nce\|graph"
done
