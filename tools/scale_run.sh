#!/bin/bash
# Scaling sweep on one box: bench.py at N = 1, 2, 4, 8 (whatever the box has).
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for n in ${SCALE_NS:-1 2 4 8}; do
  if [ $n -le $NG ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_n$n.log").read().strip().splitlines()[-1])
    print("N=$n ms/step %.2f value %.0f q/s  e2e %.0f q/s kernel_ms %.2f frac %.3f R@1 %.3f mAP %.4f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["result"]["R@1"], d["result"]["mAP"]))
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
  fi
done
