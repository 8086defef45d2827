"""Rank worker for the sharded-evaluation tests (launched by torch.multiprocessing / torchrun).

backend == "oracle": CPU stand-in for the device kernels (gloo) -- exercises only the HOST logic of
textreid_b200.sharded (CSR bookkeeping, slot offsets, collectives, merge orchestration).
backend == "cuda": the real library over NCCL, one GPU per rank.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import textreid_oracle as O
from textreid_b200.evaluation import RetrievalResult


class OracleBackend:
    """Same contracts as the CUDA kernels (include/textreid_b200.h), computed with torch on the CPU."""
    name = "oracle"

    def normalize(self, x):
        return O.normalize_rows(x.float())

    def thresholds_fp32(self, qn, gn, rel_ptr, rel_row, thr):
        counts = rel_ptr[1:] - rel_ptr[:-1]
        qidx = torch.repeat_interleave(torch.arange(qn.shape[0]), counts)
        rows = rel_row[:qidx.numel()]
        ok = rows >= 0
        thr[:qidx.numel()][ok] = (qn[qidx[ok]] * gn[rows[ok]]).sum(1)

    def nsplit(self, Q, G, device):
        return 2 if G >= 4 else 1

    def merge_lists(self, cand_sim, cand_idx):
        Q = cand_sim.shape[0]
        cs, ci = cand_sim.reshape(Q, -1), cand_idx.reshape(Q, -1)
        key = torch.argsort(ci, dim=1, stable=True)
        cs, ci = torch.gather(cs, 1, key), torch.gather(ci, 1, key)
        order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :10]
        ts, ti = torch.gather(cs, 1, order), torch.gather(ci, 1, order)
        ti = torch.where(ts == float("-inf"), torch.full_like(ti, -1), ti)
        return ts.unsqueeze(1).contiguous(), ti.unsqueeze(1).contiguous()

    def finish_partial(self, sim_my, idx_my, my_ptr, cnt, total):
        Qc = sim_my.shape[0]
        idx_my = torch.where(idx_my < 0, torch.full_like(idx_my, torch.iinfo(torch.int64).max), idx_my)
        ts, ti = self.merge_lists(sim_my, idx_my)
        ts, ti = ts[:, 0], ti[:, 0]
        ti = torch.where(ti == torch.iinfo(torch.int64).max, torch.full_like(ti, -1), ti)
        hit_ranks = torch.zeros(max(total, 1), dtype=torch.int32)
        first, ap = [], []
        for q in range(Qc):
            lo, hi = int(my_ptr[q]), int(my_ptr[q + 1])
            r = torch.sort(cnt[lo:hi])[0]
            hit_ranks[lo:hi] = r
            first.append(int(r[0]) if r.numel() else 2 ** 31 - 1)
            s = torch.tensor(0.0)
            for j, x in enumerate(r.tolist()):
                s = s + torch.tensor(float(j + 1)) / torch.tensor(float(x + 1))
            ap.append(s / torch.tensor(float(r.numel())) if r.numel() else torch.tensor(float("nan")))
        return ts, ti, torch.tensor(first, dtype=torch.int32), torch.stack(ap), hit_ranks

    def metrics(self, first_hit, ap, topk):
        Q = first_hit.numel()
        cmc = torch.stack([(first_hit < k).float().sum() / Q * 100 for k in topk])
        return cmc, ap.double().mean().float() * 100

    def stream_fp32(self, qn, gn, g_base, rel_ptr, thr, thr_gidx, cnt, nsplit):
        Q, G = qn.shape[0], gn.shape[0]
        sim = qn @ gn.t()
        gidx = torch.arange(G) + g_base
        cand_sim = torch.full((Q, nsplit, 10), float("-inf"))
        cand_idx = torch.full((Q, nsplit, 10), torch.iinfo(torch.int64).max, dtype=torch.int64)
        for s in range(nsplit):
            lo, hi = G * s // nsplit, G * (s + 1) // nsplit
            if hi > lo:
                order = torch.argsort(sim[:, lo:hi], dim=1, descending=True, stable=True)[:, :10]
                k = order.shape[1]
                cand_sim[:, s, :k] = torch.gather(sim[:, lo:hi], 1, order)
                cand_idx[:, s, :k] = gidx[lo:hi][order]
        if rel_ptr is not None:
            counts = rel_ptr[1:] - rel_ptr[:-1]
            qidx = torch.repeat_interleave(torch.arange(Q), counts)
            n = qidx.numel()
            s_rows = sim[qidx]                                   # [n, G]
            t, ti = thr[:n].unsqueeze(1), thr_gidx[:n].unsqueeze(1)
            before = (s_rows > t) | ((s_rows == t) & (gidx.unsqueeze(0) < ti))
            cnt[:n] += before.sum(1).to(torch.int32)
        return cand_sim, cand_idx

    def finish(self, cand_sim, cand_idx, nlists, q_pids, g_pids, rel, cnt, topk):
        Q = q_pids.numel()
        cs, ci = cand_sim.reshape(Q, -1), cand_idx.reshape(Q, -1)
        key = torch.argsort(ci, dim=1, stable=True)              # ties by index, then stable by similarity
        cs, ci = torch.gather(cs, 1, key), torch.gather(ci, 1, key)
        order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :10]
        top_sim, top_idx = torch.gather(cs, 1, order), torch.gather(ci, 1, order)
        top_idx = torch.where(top_sim == float("-inf"), torch.full_like(top_idx, -1), top_idx)
        topk = list(topk)
        if rel is None:
            hit = (g_pids[top_idx.clamp(min=0)] == q_pids.view(-1, 1)) & (top_idx >= 0)
            first = torch.where(hit.any(1), hit.float().argmax(1), torch.full((Q,), 2 ** 31 - 1)).to(torch.int32)
            cmc = torch.stack([(first < k).float().mean() * 100 for k in topk])
            return RetrievalResult(cmc, None, top_idx, top_sim, first, None, None, None)
        ranks, ap, first = [], [], []
        for q in range(Q):
            r = torch.sort(cnt[rel.rel_ptr[q]:rel.rel_ptr[q + 1]])[0]
            ranks.append(r)
            first.append(int(r[0]) if r.numel() else 2 ** 31 - 1)
            s = torch.tensor(0.0)
            for j, x in enumerate(r.tolist()):
                s = s + torch.tensor(float(j + 1)) / torch.tensor(float(x + 1))
            ap.append(s / torch.tensor(float(r.numel())) if r.numel() else torch.tensor(float("nan")))
        first = torch.tensor(first, dtype=torch.int32)
        ap = torch.stack(ap)
        cmc = torch.stack([(first < k).float().sum() / Q * 100 for k in topk])
        return RetrievalResult(cmc, ap.double().mean().float() * 100, top_idx, top_sim, first, ap,
                               torch.cat(ranks).to(torch.int32) if ranks else None, rel.rel_ptr)


def make_case(Q, G, D, n_ids, seed, exact=False):
    """exact=True: +-1/8 Rademacher rows with D=64 (unit norm, every dot product exact in fp32 under any
    summation order, many ties) so that shard-local and global arithmetic agree bit for bit."""
    g = torch.Generator().manual_seed(seed)
    ipid = torch.randint(0, n_ids, (G,), generator=g)
    src = torch.randint(0, G, (Q,), generator=g)
    tpid = ipid[src].clone()
    if exact:
        assert D == 64
        image = (torch.randint(0, 2, (G, D), generator=g).float() * 2 - 1) / 8.0
        text = (torch.randint(0, 2, (Q, D), generator=g).float() * 2 - 1) / 8.0
    else:
        image = torch.randn(G, D, generator=g)
        text = 0.5 * image[src] + torch.randn(Q, D, generator=g)
    return text, image, tpid, ipid


def shard_slices(G, world, uneven=True):
    cuts = [0]
    for r in range(world):
        share = G // world + (37 if (uneven and r == 0) else 0)
        cuts.append(min(G, cuts[-1] + share))
    cuts[-1] = G
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class _GoldenSplit:
    """Dataset / loader / model stand-ins that replay the embeddings of tests/golden/evaluation_small.npz (recorded from the
    unmodified reference); rank r encodes the items idx % world == r, and item 0 twice (a sampler that pads)."""

    def __init__(self, rank, world):
        import numpy as np
        g = np.load(os.path.join(ROOT, "tests", "golden", "evaluation_small.npz"))
        self.v, self.t = torch.from_numpy(g["v"]), torch.from_numpy(g["t"])
        self.image_ids, self.pids = [int(x) for x in g["image_ids"]], [int(x) for x in g["pids"]]
        self.mine = [i for i in range(self.v.shape[0]) if i % world == rank] + ([0] if rank == world - 1 else [])
        self.r1 = torch.from_numpy(g["plain.r1"])
        outer = self

        class DS:
            def __len__(self): return len(outer.image_ids)
            def get_id_info(self, idx): return outer.image_ids[idx], outer.pids[idx]

        class Cap:
            def to(self, device): return self

        class Loader:
            dataset = DS()

            def __iter__(self):
                for b0 in range(0, len(outer.mine), 16):
                    idx = outer.mine[b0:b0 + 16]
                    yield torch.tensor(idx, dtype=torch.float32).unsqueeze(1), [Cap() for _ in idx], tuple(idx)

        class Model(torch.nn.Module):
            pad = 0         # zero columns appended to the embeddings (cosine similarities unchanged; bf16 path needs D % 64 == 0)

            def forward(self, images, captions):
                idx = images[:, 0].long().cpu()
                f = torch.nn.functional.pad
                return [f(outer.v[idx], (0, self.pad)).to(images.device), f(outer.t[idx], (0, self.pad)).to(images.device)]

        self.loader, self.model = Loader(), Model()


def worker(rank, world, backend_name, port, out_dir, precision="fp32", Q=151, G=700, D=64, exact=True):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if backend_name == "cuda":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        dev = torch.device("cuda", rank)
        backend = None
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = torch.device("cpu")
        backend = OracleBackend()
    from textreid_b200 import sharded
    from textreid_b200.sharded import retrieve_sharded
    text, image, tpid, ipid = make_case(Q, G, D, max(G // 5, 1), seed=5, exact=exact)
    lo, hi = shard_slices(G, world)[rank]
    tpid_d, ipid_d = tpid.to(dev), ipid[lo:hi].to(dev).contiguous()
    res = retrieve_sharded(text.to(dev), image[lo:hi].to(dev), tpid_d, ipid_d, (1, 5, 10), True,
                           precision, backend=backend, return_hit_ranks=True)
    res_topk = retrieve_sharded(text.to(dev), image[lo:hi].to(dev), tpid_d, ipid_d, (1, 5, 10), False,
                                precision, backend=backend)
    # second evaluation of the same split (same pid tensors, new embeddings objects): served from the cached plan, same result
    before = dict(sharded.plan_cache_stats)
    res2 = retrieve_sharded(text.to(dev).clone(), image[lo:hi].to(dev).clone(), tpid_d, ipid_d, (1, 5, 10), True,
                            precision, backend=backend, return_hit_ranks=True)
    assert sharded.plan_cache_stats["hits"] == before["hits"] + 1 and sharded.plan_cache_stats["misses"] == before["misses"]
    assert torch.equal(res2.top_idx, res.top_idx) and torch.equal(res2.cmc, res.cmc) and torch.equal(res2.hit_ranks, res.hit_ranks)
    out = {"cmc": res.cmc.cpu(), "mAP": res.mAP.cpu(), "top_idx": res.top_idx.cpu(), "top_sim": res.top_sim.cpu(),
           "hit_ranks": res.hit_ranks.cpu(), "rel_ptr": res.rel_ptr.cpu(), "ap": res.ap.cpu(),
           "cmc_topk": res_topk.cmc.cpu(), "top_idx_topk": res_topk.top_idx.cpu()}
    # the embedding hand-off of inference(): every rank encodes a part of the split, device all-gather, rows by dataset index
    from textreid_b200.evaluation import compute_on_dataset_tensors, gather_embeddings
    split = _GoldenSplit(rank, world)
    idx, v_loc, t_loc = compute_on_dataset_tensors(split.model, split.loader, dev)
    idx_all, v_all, t_all = gather_embeddings(idx, v_loc, t_loc)
    n = split.v.shape[0]
    assert idx_all.tolist() == list(range(n)) and torch.equal(v_all.cpu(), split.v) and torch.equal(t_all.cpu(), split.t)
    if backend_name == "cuda":
        # the reference entry under a process group: sharded evaluation, value on the main process only (inference.py:86-87)
        from textreid_b200.evaluation import inference
        for prec in ("fp32", "bf16"):
            split.model.pad = 32 if prec == "bf16" else 0
            r1 = inference(split.model, split.loader, device=str(dev), output_folder=out_dir, save_data=False, rerank=False,
                           precision=prec)
            assert (r1 is None) == (rank != 0)
            if rank == 0:
                out["inference_r1_" + prec] = r1.cpu()
        out["inference_r1_golden"] = split.r1
        # cross-rank MoCo queue: every rank steps on its own batch, the keys of all ranks enter every queue
        import textreid_b200 as trb
        from textreid_b200.synthetic import loss_inputs
        N, D, C = 8, 64, 101
        K = 4 * N * world      # the global batch N * world must divide the queue (head.py:101), also at world 3
        base = loss_inputs(N, D, K, C, seed=77)
        vq, tq, idq = base["v_queue"].to(dev), base["t_queue"].to(dev), base["id_queue"].to(dev)
        ptr = torch.zeros(1, dtype=torch.int64, device=dev)
        exp_v, exp_t, exp_id, exp_ptr = base["v_queue"].clone(), base["t_queue"].clone(), base["id_queue"].clone(), torch.zeros(1, dtype=torch.int64)
        pr = base["projection"].to(dev)
        for step in range(3):
            per_rank = [loss_inputs(N, D, K, C, seed=1000 * step + r) for r in range(world)]
            mine = {k: v.to(dev) for k, v in per_rank[rank].items()}
            d = trb.moco_loss_dict(mine["v_embed"], mine["t_embed"], mine["v_key"], mine["t_key"], mine["labels"], vq, tq, idq, ptr, pr,
                                   epsilon=0.1, precision="fp32", gather_group=True)
            ref_l, _, _, _ = O.moco_loss_dict_with_grads(*[per_rank[rank][k] for k in ("v_embed", "t_embed", "v_key", "t_key", "labels")],
                                                         exp_v, exp_t, exp_id, base["projection"], epsilon=0.1)
            for k in ref_l:       # logits from the queue as it was BEFORE this step's global enqueue
                assert abs(float(d[k]) - float(ref_l[k])) <= 2e-5 * max(1.0, abs(float(ref_l[k]))), (step, k)
            O.enqueue(exp_v, exp_t, exp_id, exp_ptr, torch.cat([p["v_key"] for p in per_rank]), torch.cat([p["t_key"] for p in per_rank]),
                      torch.cat([p["labels"] for p in per_rank]))
            assert torch.equal(vq.cpu(), exp_v) and torch.equal(tq.cpu(), exp_t) and torch.equal(idq.cpu(), exp_id) and int(ptr) == int(exp_ptr)
        out["cross_rank_queue_ok"] = torch.tensor(1)
    torch.save(out, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":      # torchrun entry: python tests/sharded_worker.py <backend> <out_dir> <precision> <exact 0|1>
    exact = len(sys.argv) > 4 and sys.argv[4] == "1"
    worker(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), sys.argv[1], int(os.environ.get("MASTER_PORT", "29511")),
           sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "fp32", Q=300 if not exact else 151, G=3000 if not exact else 700, D=64,
           exact=exact)
