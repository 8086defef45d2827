"""Materialised similarity and k-reciprocal Jaccard re-ranking (lib/data/metrics/evaluation.py:40-65,
120, 151-156).  The re-ranked scores are float64 like the reference (np.zeros -> from_numpy)."""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from .evaluation import TOPK_DEPTH, RetrievalResult, _finish_and_metrics, build_relevance, retrieve


def similarity_matrix(qn: torch.Tensor, gn: torch.Tensor) -> torch.Tensor:
    """sim = qn @ gn^T for already-normalised rows, written to HBM (compatibility paths only)."""
    _lib.require_cuda(qn, gn)
    qn, gn = qn.contiguous().float(), gn.contiguous().float()
    Q, D = qn.shape
    G = gn.shape[0]
    sim = torch.empty(Q, G, dtype=torch.float32, device=qn.device)
    _lib.check(_lib.load().trb_similarity_f32(_lib.ptr(qn), _lib.ptr(gn), _lib.ptr(sim), Q, G, D,
                                              _lib.stream_ptr(qn.device)), "trb_similarity_f32")
    return sim


def neighbor_lists(q_feats: torch.Tensor, g_feats: torch.Tensor, neighbor_num: int = 5) -> Tuple[torch.Tensor, torch.Tensor]:
    """qg_nn [Q, n]: best n gallery items of every query; gg_nn [G, n]: best n gallery items of every gallery item
    (evaluation.py:44-51, with the pinned tie order).  Features are already L2-normalised.  Uses the fused fp32
    top-10 stream; no similarity matrix is written."""
    if not 1 <= neighbor_num <= TOPK_DEPTH:
        raise ValueError("neighbor_num must be in [1, %d]" % TOPK_DEPTH)
    dev = q_feats.device
    zq = torch.zeros(q_feats.shape[0], dtype=torch.int64, device=dev)
    zg = torch.zeros(g_feats.shape[0], dtype=torch.int64, device=dev)
    k = (min(neighbor_num, g_feats.shape[0]),)
    qg = retrieve(q_feats, g_feats, zq, zg, k, get_mAP=False, precision="fp32", normalized=True).top_idx[:, :neighbor_num]
    gg = retrieve(g_feats, g_feats, zg, zg, k, get_mAP=False, precision="fp32", normalized=True).top_idx[:, :neighbor_num]
    return qg.contiguous(), gg.contiguous()


def jaccard_rerank_matrix(q_feats: torch.Tensor, g_feats: torch.Tensor, neighbor_num: int = 5, alpha: float = 0.05, nn=None):
    """Drop-in for k_reciprocal (evaluation.py:40-65): alpha * Jaccard as a float64 [Q, G] matrix.  ``nn`` = the neighbour
    lists ``neighbor_lists(q_feats, g_feats, neighbor_num)`` when the caller already has them."""
    qg, gg = nn if nn is not None else neighbor_lists(q_feats, g_feats, neighbor_num)
    Q, G = q_feats.shape[0], g_feats.shape[0]
    out = torch.empty(Q, G, dtype=torch.float64, device=q_feats.device)
    _lib.check(_lib.load().trb_jaccard_f64(_lib.ptr(qg), _lib.ptr(gg), neighbor_num, float(alpha), _lib.ptr(out), Q, G,
                                           _lib.stream_ptr(out.device)), "trb_jaccard_f64")
    return out


def rerank_rank(similarity: torch.Tensor, qg_nn: torch.Tensor, gg_nn: torch.Tensor, q_pids, g_pids, topk=(1, 5, 10),
                alpha: float = 0.05) -> RetrievalResult:
    """rank(alpha * jaccard + similarity, ...) with the float64 scores formed on the fly (evaluation.py:151-156)."""
    _lib.require_cuda(similarity, qg_nn, gg_nn)
    dev = similarity.device
    Q, G = similarity.shape
    q_pids = q_pids.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    g_pids = g_pids.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    rel = build_relevance(q_pids, g_pids)
    cand_sim = torch.empty(Q, TOPK_DEPTH, dtype=torch.float32, device=dev)
    cand_idx = torch.empty(Q, TOPK_DEPTH, dtype=torch.int64, device=dev)
    cnt = torch.zeros(max(rel.total, 1), dtype=torch.int32, device=dev)
    n = qg_nn.shape[1]
    _lib.check(_lib.load().trb_rank_rerank_f64(
        _lib.ptr(similarity), similarity.stride(0), similarity.stride(1), Q, G, _lib.ptr(qg_nn.contiguous()),
        _lib.ptr(gg_nn.contiguous()), n, float(alpha), _lib.ptr(rel.rel_ptr), rel.col_ptr(), _lib.ptr(cand_sim), _lib.ptr(cand_idx),
        _lib.ptr(cnt), _lib.stream_ptr(dev)), "trb_rank_rerank_f64")
    return _finish_and_metrics(cand_sim, cand_idx, 1, q_pids, g_pids, rel, cnt, topk)
