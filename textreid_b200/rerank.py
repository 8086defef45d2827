"""Materialised similarity and k-reciprocal Jaccard re-ranking (lib/data/metrics/evaluation.py:40-65,
120, 151-156).  The re-ranked scores are float64 like the reference (np.zeros -> from_numpy)."""
from __future__ import annotations

import torch

from . import _lib


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


def jaccard_rerank_matrix(q_feats: torch.Tensor, g_feats: torch.Tensor, neighbor_num: int = 5, alpha: float = 0.05):
    raise NotImplementedError("k-reciprocal re-ranking (SURVEY section 8f, row 1) is not built yet")


def jaccard_rerank_rank(jac, sim, q_pids, g_pids, topk):
    raise NotImplementedError("k-reciprocal re-ranking (SURVEY section 8f, row 1) is not built yet")
