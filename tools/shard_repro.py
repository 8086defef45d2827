"""Single-GPU reproduction of the 3-rank NCCL fixture of tests/test_sharded.py (tie-heavy, uneven shards, D = 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import textreid_oracle as O
from tests.sharded_worker import make_case, shard_slices
from textreid_b200.sharded import retrieve_sharded_local
import textreid_b200 as trb

def T(x): return x.cuda()
Q, G, D = 151, 700, 64
text, image, tpid, ipid = make_case(Q, G, D, max(G // 5, 1), seed=5, exact=True)
sim = (O.normalize_rows(text).bfloat16().double() @ O.normalize_rows(image).bfloat16().double().t()).float()
ranks = O.hit_ranks(sim, tpid, ipid)
for world in (1, 2, 3, 4):
    sl = shard_slices(G, world)
    for prec in ("fp32", "bf16"):
        res = retrieve_sharded_local(T(text), [T(image[a:b]) for a, b in sl], T(tpid), [T(ipid[a:b]) for a, b in sl], (1, 5, 10), True, prec)
        hr, rp = res.hit_ranks.cpu().long(), res.rel_ptr.cpu()
        bad = []
        for q in range(Q):
            got = hr[rp[q]:rp[q + 1]]
            if not torch.equal(got, ranks[q]):
                bad.append(q)
        print("world", world, prec, "shards", sl, "bad queries", bad[:10], flush=True)
        for q in bad[:3]:
            rel = (ipid == tpid[q]).nonzero().flatten()
            print("   q", q, "relevant items", rel.tolist(), "sims", sim[q, rel].tolist(), "got", hr[rp[q]:rp[q + 1]].tolist(), "want", ranks[q].tolist())
            for r in rel.tolist():
                ties = (sim[q] == sim[q, r]).nonzero().flatten().tolist()
                print("      item", r, "thr", float(sim[q, r]), "#greater", int((sim[q] > sim[q, r]).sum()), "ties at", ties[:40])

# ---- per-shard contributions of one failing configuration (world 3, bf16) ----
from textreid_b200.sharded import _local_plans, ShardWorker, CudaBackend, _as_pid
from textreid_b200 import sharded
sharded.clear_plan_cache()
world = 3
sl = shard_slices(G, world)
q_pids = _as_pid(T(tpid)); pids = [_as_pid(T(ipid[a:b])) for a, b in sl]
bases = [a for a, _ in sl]
plans = _local_plans(q_pids, pids, bases, True, "bf16")
workers = [ShardWorker(T(text), T(image[a:b]), pl, CudaBackend(), False) for (a, b), pl in zip(sl, plans)]
parts = [w.local_thresholds() for w in workers]
thr = torch.stack(parts).sum(0)
for w in workers: w.set_thresholds(thr)
outs = [w.stream(None) for w in workers]
rel_ptr = plans[0].rel.rel_ptr.cpu(); gidx = plans[0].gidx_store.cpu(); thr_c = thr.cpu()
for q in (58, 85):
    for s in range(int(rel_ptr[q]), int(rel_ptr[q + 1])):
        r = int(gidx[s]); t = float(thr_c[s])
        line = "q %d slot %d item %d thr %g (oracle sim %g):" % (q, s - int(rel_ptr[q]), r, t, float(sim[q, r]))
        for k, (a, b) in enumerate(sl):
            seg = sim[q, a:b]
            idx = torch.arange(a, b)
            want = int(((seg > t) | ((seg == t) & (idx < r))).sum())
            got = int(outs[k][2][s])
            line += "  shard%d got %d want %d%s" % (k, got, want, "" if got == want else " <<<")
        print(line)
