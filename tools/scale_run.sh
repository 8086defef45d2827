#!/bin/bash
# Round-2 scaling sweep on one 8-GPU box: N = 1, 2, 4, 8 back to back (strong scaling of BASELINE configs[3]); the multi-GPU
# parity tests (exact tie-heavy fixture under NCCL, inference() under a process group, cross-rank queue) run first on 4 GPUs.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded.py -m gpu -q --timeout 600 > gpurun_out/pytest_nccl.log 2>&1
echo "nccl tests exit $?"; tail -n 2 gpurun_out/pytest_nccl.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/scale_n1.log 2> gpurun_out/scale_n1.err
for n in 2 4 8; do
  TRB_PROFILE_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
  echo "n$n exit $?"; grep PHASES gpurun_out/scale_n$n.err | head -1
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open("gpurun_out/scale_n%d.log" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "ERR", e); continue
    base = base or d
    print("N=%d value %.0f q/s (%.2fx) ms %.3f | e2e %.0f q/s (%.2fx) ms %.2f | kernel_ms %.2f frac %.3f | launches %s | stored-result match %s" % (
        n, d["value"], d["value"] / base["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["value"] / base["e2e"]["value"], d["e2e"]["ms_per_step"],
        d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["result"].get("matches_stored_single_gpu_result")))
PY
