"""Sampled fp64 re-evaluation of a retrieval result at sizes where no [Q, G] matrix fits (BASELINE configs[3]: 100k x 1M).

A diagnostic, not a compute path: it re-derives, for a random sample of queries, the similarities to the WHOLE gallery in
float64 from the very bf16 operands the tensor-core stream consumed (read back out of the packed images), and checks the
kernel's top-10 lists and hit ranks against the order those float64 values imply.  The only difference between the two sides
is the fp32 accumulation order inside the tensor core, so every comparison that is decided by more than ``margin`` must
agree exactly; the rest must stay inside the interval the margin allows.  bench.py reports the outcome as
``result.parity_sample``; tests/test_gpu_retrieval.py cross-checks this module against the CPU restatement of the reference.
"""
from __future__ import annotations

from typing import Dict

import torch

from .retrieval_tc import pack_rows, unpack_rows


@torch.no_grad()
def sampled_similarity_fp64(text_embed, image_embed, sample: torch.Tensor, normalized: bool = False, block: int = 1 << 17):
    """[len(sample), G] float64 similarities of the sampled queries from the kernel's own bf16 operands."""
    D = text_embed.shape[1]
    G = image_embed.shape[0]
    dev = text_embed.device
    qp = pack_rows(text_embed[sample].contiguous(), None, normalize=not normalized)
    qb = unpack_rows(qp, torch.arange(sample.numel(), device=dev), D).double()
    sim = torch.empty(sample.numel(), G, dtype=torch.float64, device=dev)
    for g0 in range(0, G, block):
        g1 = min(G, g0 + block)
        gp = pack_rows(image_embed[g0:g1].contiguous(), None, normalize=not normalized)
        gb = unpack_rows(gp, torch.arange(g1 - g0, device=dev), D).double()
        sim[:, g0:g1] = qb @ gb.t()
    return sim


@torch.no_grad()
def sampled_check(text_embed, image_embed, q_pid, g_pid, res, n_sample: int = 64, seed: int = 0, margin: float = 2e-6,
                  normalized: bool = False) -> Dict:
    """Check ``res`` (a RetrievalResult of the bf16 path with hit ranks) on ``n_sample`` random queries.  Returns a dict with
    the counts and ``status`` = "ok" | "FAIL"."""
    dev = text_embed.device
    Q, G = text_embed.shape[0], image_embed.shape[0]
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sample = torch.randperm(Q, generator=gen)[:min(n_sample, Q)].to(dev)
    sim = sampled_similarity_fp64(text_embed, image_embed, sample, normalized)
    q_pid, g_pid = q_pid.reshape(-1).to(dev), g_pid.reshape(-1).to(dev)
    k = min(10, G)
    out = dict(n_queries=int(sample.numel()), gallery=G, margin=margin, top10_decided=0, top10_match=0, top10_undecided=0,
               slots=0, slots_decided=0, slots_decided_exact=0, slots_outside_interval=0, max_sim_err=0.0)
    top_v, top_i = torch.topk(sim, min(k + 1, G), dim=1, largest=True, sorted=True)
    got_i, got_v = res.top_idx[sample][:, :k], res.top_sim[sample][:, :k].double()
    out["max_sim_err"] = float((got_v - torch.gather(sim, 1, got_i.clamp(min=0))).abs().max())
    gaps = top_v[:, :-1] - top_v[:, 1:] if top_v.shape[1] > k else None
    for j in range(sample.numel()):
        decided = gaps is None or bool((gaps[j] > margin).all())
        if decided:
            out["top10_decided"] += 1
            out["top10_match"] += int(torch.equal(got_i[j], top_i[j, :k]))
        else:
            out["top10_undecided"] += 1
            # the SET must still agree up to entries within the margin of the 10th value
            kth = top_v[j, k - 1]
            sure = top_i[j, :k][top_v[j, :k] > kth + margin]
            out["top10_match"] += int(bool(torch.isin(sure, got_i[j]).all()))
            out["top10_decided"] += 1
    rel_ptr = res.rel_ptr
    idx = torch.arange(G, device=dev)
    for j, q in enumerate(sample.tolist()):
        rel = (g_pid == q_pid[q]).nonzero().reshape(-1)
        if rel.numel() == 0:
            continue
        s = sim[j]
        t = s[rel]                                                                         # [R]
        exact = (s.unsqueeze(0) > t.unsqueeze(1)).sum(1) + ((s.unsqueeze(0) == t.unsqueeze(1)) & (idx.unsqueeze(0) < rel.unsqueeze(1))).sum(1)
        lo = (s.unsqueeze(0) > (t + margin).unsqueeze(1)).sum(1)
        hi = (s.unsqueeze(0) >= (t - margin).unsqueeze(1)).sum(1) - 1
        got = res.hit_ranks[int(rel_ptr[q]):int(rel_ptr[q + 1])].to(torch.int64)          # ascending
        lo_s, hi_s, ex_s = torch.sort(lo)[0], torch.sort(hi)[0], torch.sort(exact)[0]
        out["slots"] += int(rel.numel())
        out["slots_outside_interval"] += int(((got < lo_s) | (got > hi_s)).sum())
        dec = lo_s == hi_s
        out["slots_decided"] += int(dec.sum())
        out["slots_decided_exact"] += int((got[dec] == ex_s[dec]).sum())
    ok = (out["top10_match"] == out["top10_decided"] and out["slots_outside_interval"] == 0
          and out["slots_decided_exact"] == out["slots_decided"] and out["max_sim_err"] < 1e-5)
    out["status"] = "ok" if ok else "FAIL"
    return out
