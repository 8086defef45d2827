set -u
mkdir -p gpurun_out
TRB_PROBE_SHAPE=n256 TRB_LOSS_PRECISION=bf16 TRB_LOSS_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 40 --csv --log-file gpurun_out/launches_loss_n256.csv python tools/loss_probe.py > gpurun_out/launches_loss_n256.log 2>&1
echo "exit $?"
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/launches_loss_n256.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
for r in rows[hdr + 1:]:
    name = r[4].split("(")[0][-40:]
    print("%-42s grid %-14s %8.1f us" % (name, r[8], float(r[-1]) / 1e3))
PY
