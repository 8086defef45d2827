"""CPU oracle for the TextReID hot path.  TEST INFRASTRUCTURE ONLY.

This module is a from-scratch CPU restatement (torch CPU tensors, fp32 by
default, fp64 on request) of the reference algorithms on the north-star path:

    MoCo loss dict   lib/models/embeddings/moco_head/head.py:111-176
                     lib/models/embeddings/moco_head/loss.py:21-39
                     lib/models/losses.py:6-62, 102-128, 206-217
    EMA / enqueue    lib/models/embeddings/moco_head/head.py:73-109
    retrieval eval   lib/data/metrics/evaluation.py:11-37, 68-73, 117-120

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; the product package ``textreid_b200``
never does (it fails loudly when the CUDA library is missing).

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, imported from /root/reference in the build container by
``tools/make_golden.py``; the resulting fixtures live in ``tests/golden`` and
``tests/test_oracle_golden.py`` replays them without the reference present.

Tie rule: the reference's ``argsort``/``topk`` leave tie order undefined;
the north star pins (similarity desc, gallery index asc), i.e. a stable
descending sort.  ``rank(..., stable=True)`` is therefore the default.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

__all__ = [
    "normalize_rows",
    "queue_positive_mask",
    "moco_logits",
    "infonce_loss",
    "instance_loss",
    "global_align_loss",
    "moco_loss_dict",
    "moco_loss_dict_with_grads",
    "ema_update",
    "enqueue",
    "rank",
    "first_occurrence",
    "similarity_matrix",
    "retrieve",
    "hit_ranks",
    "jaccard_rerank_matrix",
]


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def normalize_rows(x: torch.Tensor, dim: int = 1, eps: float = 1e-12) -> torch.Tensor:
    """x / max(||x||_2, eps) along ``dim`` (what F.normalize computes;
    head.py:128-129,139,145; losses.py:51,112-113; evaluation.py:117-118)."""
    nrm = x.pow(2).sum(dim=dim, keepdim=True).sqrt()
    return x / nrm.clamp_min(eps)


def queue_positive_mask(id_queue: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """mask[k] = True when queue slot k holds an id present in the batch.

    head.py:148-157 builds ``neg_idx`` = slots whose id matches *no* batch id
    (one global column set for all rows, via eq/nonzero/unique/counts==1).
    The complement of that set is returned here.
    """
    ids = id_queue.reshape(-1)
    return (ids.unsqueeze(0) == labels.reshape(-1, 1)).any(dim=0)


def moco_logits(v_q, t_q, v_k, t_k, v_queue, t_queue, mask):
    """Positive and (gathered) negative logits, head.py:160-170.

    v_q/t_q are the L2-normalised query embeddings, v_k/t_k the normalised key
    embeddings, queues are [D, K].  Returns v_pos [N,1], v_neg [N,K'],
    t_pos [N,1], t_neg [N,K'] with K' = number of unmasked slots.
    """
    keep = (~mask).nonzero(as_tuple=False).reshape(-1)
    v_pos = (v_q * t_k).sum(dim=1, keepdim=True)
    t_pos = (t_q * v_k).sum(dim=1, keepdim=True)
    v_neg = v_q @ t_queue.detach()[:, keep]
    t_neg = t_q @ v_queue.detach()[:, keep]
    return v_pos, v_neg, t_pos, t_neg


def _ce_target0(logits: torch.Tensor) -> torch.Tensor:
    """mean_n(logsumexp(row) - row[0]) == F.cross_entropy(logits, zeros)."""
    return (torch.logsumexp(logits, dim=1) - logits[:, 0]).mean()


def infonce_loss(v_pos, v_neg, t_pos, t_neg, T: float = 0.07) -> torch.Tensor:
    """losses.py:206-217: CE over [pos | neg] / T with target column 0, v + t."""
    v_logits = torch.cat([v_pos, v_neg], dim=1) / T
    t_logits = torch.cat([t_pos, t_neg], dim=1) / T
    return _ce_target0(v_logits) + _ce_target0(t_logits)


def _smoothed_ce(logits: torch.Tensor, labels: torch.Tensor, epsilon: float) -> torch.Tensor:
    """CrossEntropyLabelSmooth.forward, losses.py:26-39:
    target = (1-eps)*onehot + eps/C ; loss = (-target*log_softmax).mean(0).sum()."""
    n, c = logits.shape
    logp = torch.log_softmax(logits, dim=1)
    target = torch.zeros_like(logp)
    target[torch.arange(n, device=logits.device), labels.reshape(-1).long()] = 1.0
    target = (1 - epsilon) * target + epsilon / c
    return (-target * logp).mean(0).sum()


REFERENCE_SMOOTHING = 0.1      # CrossEntropyLabelSmooth's default epsilon, losses.py:18


def instance_loss(projection, v_embed, t_embed, labels, epsilon: float = 0.0,
                  scale: float = 1.0) -> torch.Tensor:
    """losses.py:42-62 with norm=False: logits = scale * embed @ (W / ||W||_col).

    ``epsilon`` is only an on/off switch in the reference: ``if epsilon > 0`` builds ``CrossEntropyLabelSmooth(num_classes=...)``
    WITHOUT forwarding it (losses.py:56-57), so the smoothing weight that is applied is always the class default 0.1
    (losses.py:18).  Restated as is; pinned by the golden case ``loss_fn_eps02`` (EPSILON = 0.2 gives the 0.1 result)."""
    w_hat = normalize_rows(projection, dim=0)
    zv = scale * (v_embed @ w_hat)
    zt = scale * (t_embed @ w_hat)
    if epsilon > 0:
        return _smoothed_ce(zv, labels, REFERENCE_SMOOTHING) + _smoothed_ce(zt, labels, REFERENCE_SMOOTHING)
    lab = labels.reshape(-1).long()
    idx = torch.arange(zv.shape[0], device=zv.device)
    ce = lambda z: (torch.logsumexp(z, dim=1) - z[idx, lab]).mean()
    return ce(zv) + ce(zt)


def global_align_loss(v_embed, t_embed, labels, alpha=0.6, beta=0.4,
                      scale_pos=10, scale_neg=40) -> torch.Tensor:
    """losses.py:102-128: literal log(1+exp(.)) over same-id / different-id pairs."""
    n = labels.shape[0]
    s = normalize_rows(v_embed) @ normalize_rows(t_embed).t()
    lab = labels.reshape(-1)
    same = lab.unsqueeze(0) == lab.unsqueeze(1)
    pos = torch.log(1 + torch.exp(-scale_pos * (s[same] - alpha)))
    neg = torch.log(1 + torch.exp(scale_neg * (s[~same] - beta)))
    return (pos.sum() + neg.sum()) * 2.0 / n


def moco_loss_dict(v_embed, t_embed, v_key, t_key, labels, v_queue, t_queue, id_queue,
                   projection, *, T: float = 0.07, epsilon: float = 0.1,
                   v_embed_q: Optional[torch.Tensor] = None,
                   t_embed_q: Optional[torch.Tensor] = None,
                   alpha=0.6, beta=0.4, scale_pos=10, scale_neg=40) -> Dict[str, torch.Tensor]:
    """The loss dict of MoCoHead.forward (train branch, head.py:126-172) given
    the post-Linear embeddings.  ``v_key``/``t_key`` are the *normalised* key
    embeddings.  With FC=True the InfoNCE queries come from a separate head
    (head.py:118-124): pass them un-normalised as ``v_embed_q``/``t_embed_q``.
    """
    vq = normalize_rows(v_embed if v_embed_q is None else v_embed_q)
    tq = normalize_rows(t_embed if t_embed_q is None else t_embed_q)
    mask = queue_positive_mask(id_queue, labels)
    v_pos, v_neg, t_pos, t_neg = moco_logits(vq, tq, v_key, t_key, v_queue, t_queue, mask)
    return {
        "instance_loss": instance_loss(projection, v_embed, t_embed, labels, epsilon=epsilon),
        "infonce_loss": infonce_loss(v_pos, v_neg, t_pos, t_neg, T),
        "global_align_loss": global_align_loss(v_embed, t_embed, labels, alpha, beta,
                                               scale_pos, scale_neg),
    }


def moco_loss_dict_with_grads(v_embed, t_embed, v_key, t_key, labels, v_queue, t_queue,
                              id_queue, projection, *, weights=(1.0, 1.0, 1.0), **kw):
    """Loss dict plus d(sum_i w_i loss_i)/d(v_embed, t_embed, projection), the way
    trainer.py:82,90 consumes it (weights all 1)."""
    ve = v_embed.detach().clone().requires_grad_(True)
    te = t_embed.detach().clone().requires_grad_(True)
    pr = projection.detach().clone().requires_grad_(True)
    d = moco_loss_dict(ve, te, v_key, t_key, labels, v_queue, t_queue, id_queue, pr, **kw)
    total = (weights[0] * d["instance_loss"] + weights[1] * d["infonce_loss"]
             + weights[2] * d["global_align_loss"])
    total.backward()
    losses = {k: v.detach() for k, v in d.items()}
    return losses, ve.grad, te.grad, pr.grad


# --------------------------------------------------------------------------
# momentum update and queue maintenance
# --------------------------------------------------------------------------
def ema_update(params_k: Sequence[torch.Tensor], params_q: Sequence[torch.Tensor], m: float) -> None:
    """head.py:73-94: p_k <- p_k*m + p_q*(1-m), two products then one add,
    with m and (1-m) as python floats (so 1-m is formed in double)."""
    one_minus = 1.0 - m
    for pk, pq in zip(params_k, params_q):
        pk.copy_(pk * m + pq * one_minus)


def enqueue(v_queue, t_queue, id_queue, queue_ptr, v_keys, t_keys, ids) -> None:
    """head.py:96-109: write keys^T into columns [ptr, ptr+N), advance ptr mod K."""
    n = v_keys.shape[0]
    k = v_queue.shape[1]
    if k % n != 0:
        raise AssertionError("queue length must be a multiple of the batch size")
    p = int(queue_ptr.reshape(-1)[0])
    v_queue[:, p:p + n] = v_keys.t()
    t_queue[:, p:p + n] = t_keys.t()
    id_queue[:, p:p + n] = ids.reshape(1, -1)
    queue_ptr[0] = (p + n) % k


# --------------------------------------------------------------------------
# retrieval evaluation
# --------------------------------------------------------------------------
def rank(similarity: torch.Tensor, q_pids: torch.Tensor, g_pids: torch.Tensor,
         topk=(1, 5, 10), get_mAP: bool = True, stable: bool = True,
         per_column_loop: bool = True):
    """evaluation.py:11-37.

    get_mAP=True : full descending sort of every row, CMC@topk and mAP.
    get_mAP=False: only the best max(topk) columns per row, CMC@topk.
    ``per_column_loop`` keeps the reference's G-iteration python loop for the
    precision terms (evaluation.py:33); the vectorised form is arithmetically
    identical (same fp32 division per element) and exists for large cases.
    Returns (cmc, mAP, indices) or (cmc, indices) like the reference.
    """
    topk_t = torch.as_tensor(list(topk) if not torch.is_tensor(topk) else topk).to(similarity.device)
    depth = int(topk_t.max())
    if get_mAP:
        order = torch.argsort(similarity, dim=1, descending=True, stable=stable)
    else:
        if stable:
            order = torch.argsort(similarity, dim=1, descending=True, stable=True)[:, :depth]
        else:
            order = torch.topk(similarity, k=depth, dim=1, largest=True, sorted=True)[1]
    hit = g_pids[order] == q_pids.reshape(-1, 1)

    depth = min(depth, hit.shape[1])
    reached = hit[:, :depth].cumsum(1).clamp(max=1)
    cmc = reached.float().mean(0) * 100
    cmc = cmc[topk_t - 1]
    if not get_mAP:
        return cmc, order

    n_rel = hit.sum(1)
    running = hit.cumsum(1)
    if per_column_loop:
        cols = [running[:, i] / (i + 1.0) for i in range(running.shape[1])]
        prec = torch.stack(cols, 1) * hit
    else:
        denom = torch.arange(1, running.shape[1] + 1, dtype=torch.float32, device=similarity.device)
        prec = (running.to(torch.float32) / denom) * hit
    ap = prec.sum(1) / n_rel
    return cmc, ap.mean() * 100, order


def first_occurrence(keys: Sequence) -> torch.Tensor:
    """evaluation.py:68-73 (get_unique): index of the first appearance of each key,
    in order of first appearance."""
    seen = {}
    for pos, key in enumerate(keys):
        key = key.item() if torch.is_tensor(key) else key
        seen.setdefault(key, pos)
    return torch.tensor(list(seen.values()), dtype=torch.long)


def similarity_matrix(text_embed: torch.Tensor, image_embed: torch.Tensor) -> torch.Tensor:
    """evaluation.py:117-120: cosine similarity, queries (text) by gallery (image)."""
    return normalize_rows(text_embed) @ normalize_rows(image_embed).t()


def hit_ranks(similarity: torch.Tensor, q_pids: torch.Tensor, g_pids: torch.Tensor):
    """Integer artefacts behind rank(): for every query the 0-based ranks (stable
    descending order) of its relevant gallery items, ascending.  Returns a list of
    1-D int64 tensors.  Derived from evaluation.py:14,20-21."""
    order = torch.argsort(similarity, dim=1, descending=True, stable=True)
    hit = g_pids[order] == q_pids.reshape(-1, 1)
    return [row.nonzero(as_tuple=False).reshape(-1) for row in hit]


def hit_rank_bounds(similarity: torch.Tensor, q_pids: torch.Tensor, g_pids: torch.Tensor, margin: float = 0.0):
    """The same integer artefacts as ``hit_ranks`` by COUNTING instead of sorting (usable on [few queries, 10^6] matrices):
    for every (query, relevant item r) the 0-based rank  #{g : s_g > s_r} + #{g : s_g == s_r and g < r}  under the pinned
    order (similarity descending, index ascending) -- equivalent to the position of r in the stable descending argsort of
    evaluation.py:14.  With ``margin`` > 0 it returns the interval [lo, hi] the rank can lie in when similarities closer
    than ``margin`` to s_r are treated as undecided (accumulation noise of a lower-precision implementation); margin = 0
    gives lo == hi == the exact rank.  Returns (rel_ptr [Q+1], rel_col [total], lo [total], hi [total]), slots per query in
    ascending gallery index."""
    Q, G = similarity.shape
    idx = torch.arange(G)
    ptr, cols, los, his = [0], [], [], []
    for q in range(Q):
        rel = (g_pids == q_pids[q]).nonzero().reshape(-1)
        s = similarity[q]
        for r in rel.tolist():
            t = s[r]
            if margin > 0:
                lo = int((s > t + margin).sum())
                hi = int((s >= t - margin).sum()) - 1                      # everything that may precede r, minus r itself
            else:
                lo = hi = int((s > t).sum()) + int(((s == t) & (idx < r)).sum())
            cols.append(r); los.append(lo); his.append(hi)
        ptr.append(len(cols))
    return (torch.tensor(ptr, dtype=torch.int64), torch.tensor(cols, dtype=torch.int64), torch.tensor(los, dtype=torch.int64),
            torch.tensor(his, dtype=torch.int64))


def retrieve(text_embed, image_embed, text_pid, image_pid, topk=(1, 5, 10), get_mAP=True,
             per_column_loop: bool = False):
    """Embedding-level entry: normalise, similarity, rank (evaluation.py:117-120 + rank)."""
    sim = similarity_matrix(text_embed, image_embed)
    return rank(sim, text_pid, image_pid, topk, get_mAP, per_column_loop=per_column_loop)


def jaccard_rerank_matrix(q_feats: torch.Tensor, g_feats: torch.Tensor, neighbor_num: int = 5,
                          alpha: float = 0.05) -> torch.Tensor:
    """evaluation.py:40-65 (k_reciprocal): alpha * Jaccard(top-n(q->g), top-n(g->g))
    as a float64 [Q, G] matrix.  |A|=|B|=n so J = i / (2n - i), i = |A & B|."""
    qg = torch.argsort(q_feats @ g_feats.t(), dim=1, descending=True, stable=True)[:, :neighbor_num]
    gg = torch.argsort(g_feats @ g_feats.t(), dim=1, descending=True, stable=True)[:, :neighbor_num]
    inter = (qg[:, None, :, None] == gg[None, :, None, :]).sum(dim=(2, 3)).to(torch.float64)
    return alpha * inter / (2 * neighbor_num - inter)
