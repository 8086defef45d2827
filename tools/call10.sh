set -u
mkdir -p gpurun_out
TRB_PROFILE_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err
echo "n8 exit $?"; grep PHASES gpurun_out/bench_n8.err | head -1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n8.log").read().strip().splitlines()[-1])
print("value %.0f ms %.3f e2e %.0f (%.2f ms) kernel_ms %.2f frac %.3f launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"]))
print(" result", d["result"])
PY
