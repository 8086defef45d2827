set -u
mkdir -p gpurun_out
python - <<'PY'
import sys, os, torch
sys.path.insert(0, ".")
from textreid_b200.synthetic import eval_data
from textreid_b200.sharded import ShardWorker, CudaBackend, _local_plans, _as_pid
Q, G = 100000, 1000000
text, q_pid, image, g_pid = eval_data(Q, G, 256, 250000, 0, G, "cuda", torch.bfloat16)
plan = _local_plans(_as_pid(q_pid.long()), [_as_pid(g_pid.long())], [0], True, "bf16")[0]
w = ShardWorker(text, image, plan, CudaBackend())
w.set_thresholds(w.local_thresholds())
for env, ns in ((None, None), ("1", 4), ("1", 8), ("1", 16), ("1", 32)):
    if env: os.environ["TRB_TC_SPLIT_ALL"] = env
    w.stream(ns); w.record_events = True
    ts = []
    for _ in range(4):
        w.stream(ns); torch.cuda.synchronize(); a, b = w.stream_events; ts.append(a.elapsed_time(b))
    print("split_all=%s nsplit=%s: %.2f ms" % (env, ns, sorted(ts)[1]), flush=True)
PY
