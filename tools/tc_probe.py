"""Probe of the tensor-core stream kernel: where does the time go? (run on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textreid_b200.synthetic import eval_data
from textreid_b200.sharded import ShardWorker, CudaBackend, _local_plans, _as_pid

def run(Q, G, D, get_map, n_ids, iters=3, nsplit=None):
    text, q_pid, image, g_pid = eval_data(Q, G, D, n_ids, 0, G, "cuda", torch.bfloat16)
    plan = _local_plans(_as_pid(q_pid.long()), [_as_pid(g_pid.long())], [0], get_map, "bf16")[0]
    w = ShardWorker(text, image, plan, CudaBackend())
    if get_map:
        w.set_thresholds(w.local_thresholds())
    w.stream(nsplit)
    w.record_events = True
    ts = []
    for _ in range(iters):
        w.stream(nsplit)
        torch.cuda.synchronize()
        a, b = w.stream_events
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    print("Q=%d G=%d D=%d mAP=%s max_rel=%d nsplit=%s: %.2f ms  %.1f TFLOP/s" % (Q, G, D, get_map, plan.max_rel, nsplit, ms, 2.0 * Q * G * D / ms / 1e9), flush=True)

if __name__ == "__main__":
    Q, G = 100000, 1000000
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run(Q, G, 256, True, 250000, iters=1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "r4":
        run(Q, G, 256, True, 250000, iters=3)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        run(Q, G, 256, True, 250000, iters=5)
        run(Q, G, 256, False, 250000)
        run(Q, G, 256, True, 125000)
        sys.exit(0)
    run(Q, G, 256, True, 250000)
    run(Q, G, 256, False, 250000)
    run(Q, G, 128, True, 250000)
    run(Q, G, 128, False, 250000)
    run(Q, G, 64, False, 250000)
    run(Q, G, 256, True, 125000)      # 8 relevant per query -> RT=8 variant
    run(Q, G, 256, False, 250000, nsplit=1)
    run(Q, G, 256, True, 250000, nsplit=1)
    run(Q, G, 256, True, 250000, nsplit=14)
