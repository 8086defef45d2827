set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err
echo "n1 exit $?"; tail -c 600 gpurun_out/bench_n1.err
TRB_PROFILE_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "n2 exit $?"; grep PHASES gpurun_out/bench_n2.err; tail -c 800 gpurun_out/bench_n2.err | tail -5
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.log", "gpurun_out/bench_n2.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms %.2f e2e %.0f (%.2f ms) kernel_ms %.2f frac %.3f launches %s" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"]))
        print(" result", {k: v for k, v in d["result"].items() if k != "parity_sample_detail"})
    except Exception as e:
        print(f, "ERR", e)
PY
