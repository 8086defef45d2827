"""Text-to-image retrieval evaluation on B200: host side.

Mirrors the reference's call shapes

    rank(similarity, q_pids, g_pids, topk, get_mAP)           lib/data/metrics/evaluation.py:11-37
    evaluation(dataset, predictions, output_folder, topk,     lib/data/metrics/evaluation.py:76-173
               save_data, rerank)
    inference(model, data_loader, dataset_name, device,       lib/engine/inference.py:48-96
              output_folder, save_data, rerank)

and adds the embedding-level entry ``retrieve`` that never materialises the [Q, G] similarity
matrix.  All arithmetic on the hot path runs in libtextreid_b200.so (hand-written sm_100a CUDA);
torch is used for memory, streams, index bookkeeping (sorting pids, prefix sums) and collectives.
There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

TOPK_DEPTH = _lib.TOPK_DEPTH
_SM_COUNT_FALLBACK = 148


# --------------------------------------------------------------------------------------------
# relevance index: which gallery rows share the query's pid (evaluation.py:20-21)
# --------------------------------------------------------------------------------------------
@dataclass
class RelevanceIndex:
    rel_ptr: torch.Tensor      # [Q+1] int64, CSR offsets (one slot per (query, relevant gallery item))
    rel_col: torch.Tensor      # [total] int64, global gallery index of every slot, ascending per query
    total: int
    store: Optional[torch.Tensor] = None   # backing buffer of rel_col with >= 1 element (never a NULL device pointer)

    def col_ptr(self):
        return _lib.ptr(self.store if self.store is not None else self.rel_col)


def build_relevance(q_pids: torch.Tensor, g_pids: torch.Tensor) -> RelevanceIndex:
    """CSR of relevant gallery items per query from the pid vectors (index bookkeeping in torch;
    one host sync for the total slot count)."""
    q_pids = q_pids.reshape(-1).to(torch.int64)
    g_pids = g_pids.reshape(-1).to(torch.int64)
    dev = q_pids.device
    if g_pids.numel() == 0:
        z = torch.zeros(q_pids.numel() + 1, dtype=torch.int64, device=dev)
        st = torch.zeros(1, dtype=torch.int64, device=dev)
        return RelevanceIndex(z, st[:0], 0, st)
    sorted_pid, order = torch.sort(g_pids, stable=True)
    lo = torch.searchsorted(sorted_pid, q_pids, right=False)
    hi = torch.searchsorted(sorted_pid, q_pids, right=True)
    counts = hi - lo
    rel_ptr = torch.zeros(q_pids.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=rel_ptr[1:])
    total = int(rel_ptr[-1].item())
    slot_q = torch.repeat_interleave(torch.arange(q_pids.numel(), device=dev), counts, output_size=total)
    within = torch.arange(total, device=dev) - rel_ptr[:-1][slot_q]
    store = torch.zeros(max(total, 1), dtype=torch.int64, device=dev)   # never a NULL device pointer
    store[:total] = order[lo[slot_q] + within]   # stable sort => ascending gallery index inside a pid group
    return RelevanceIndex(rel_ptr, store[:total], total, store)


# --------------------------------------------------------------------------------------------
# result container
# --------------------------------------------------------------------------------------------
@dataclass
class RetrievalResult:
    cmc: torch.Tensor                     # [len(topk)] fp32 percent (R@k)
    mAP: Optional[torch.Tensor]           # 0-d fp32 percent, None in top-k-only mode
    top_idx: torch.Tensor                 # [Q, 10] int64 global gallery indices, best first (-1 padding)
    top_sim: torch.Tensor                 # [Q, 10] fp32
    first_hit: torch.Tensor               # [Q] int32
    ap: Optional[torch.Tensor]            # [Q] fp32
    hit_ranks: Optional[torch.Tensor]     # [total] int32, ascending per query
    rel_ptr: Optional[torch.Tensor]       # [Q+1] int64
    thresholds: Optional[torch.Tensor] = None   # [total] fp32 similarity of every relevant pair (debug / tests)


def _topk_host_array(topk) -> Tuple[C.Array, int, List[int]]:
    vals = [int(k) for k in (topk.tolist() if torch.is_tensor(topk) else list(topk))]
    if not 1 <= len(vals) <= 8:
        raise ValueError("topk must hold between 1 and 8 cut-offs")
    if max(vals) > TOPK_DEPTH or min(vals) < 1:
        raise ValueError("topk cut-offs must lie in [1, %d] (the reference uses [1, 5, 10])" % TOPK_DEPTH)
    return (C.c_int32 * len(vals))(*vals), len(vals), vals


def _sm_count(device) -> int:
    try:
        return torch.cuda.get_device_properties(device).multi_processor_count
    except Exception:  # pragma: no cover
        return _SM_COUNT_FALLBACK


def _finish_and_metrics(cand_sim, cand_idx, nlists, q_pids, g_pids, rel: Optional[RelevanceIndex], cnt, topk,
                        want_hit_ranks=True) -> RetrievalResult:
    lib = _lib.load()
    dev = cand_sim.device
    Q = q_pids.numel()
    st = _lib.stream_ptr(dev)
    top_sim = torch.empty(Q, TOPK_DEPTH, dtype=torch.float32, device=dev)
    top_idx = torch.empty(Q, TOPK_DEPTH, dtype=torch.int64, device=dev)
    first_hit = torch.empty(Q, dtype=torch.int32, device=dev)
    get_map = rel is not None
    ap = torch.empty(Q, dtype=torch.float32, device=dev) if get_map else None
    hit_ranks = torch.empty(max(rel.total, 1), dtype=torch.int32, device=dev) if (get_map and want_hit_ranks) else None
    _lib.check(lib.trb_retrieval_finish(
        _lib.ptr(cand_sim), _lib.ptr(cand_idx), nlists, Q, _lib.ptr(q_pids), _lib.ptr(g_pids), g_pids.numel(),
        _lib.ptr(rel.rel_ptr) if get_map else None, _lib.ptr(cnt) if get_map else None,
        _lib.ptr(top_sim), _lib.ptr(top_idx), _lib.ptr(first_hit), _lib.ptr(hit_ranks), _lib.ptr(ap), st),
        "trb_retrieval_finish")
    arr, n, _ = _topk_host_array(topk)
    cmc = torch.empty(n, dtype=torch.float32, device=dev)
    mAP = torch.empty((), dtype=torch.float32, device=dev) if get_map else None
    _lib.check(lib.trb_retrieval_metrics(_lib.ptr(first_hit), _lib.ptr(ap), Q, arr, n, _lib.ptr(cmc),
                                         _lib.ptr(mAP), st), "trb_retrieval_metrics")
    if hit_ranks is not None:
        hit_ranks = hit_ranks[:rel.total]
    return RetrievalResult(cmc, mAP, top_idx, top_sim, first_hit, ap, hit_ranks,
                           rel.rel_ptr if get_map else None)


def reference_tail_from_hit_ranks(res: RetrievalResult, G: int, topk) -> Tuple[torch.Tensor, torch.Tensor]:
    """Parity mode: evaluate the reference's float reductions (evaluation.py:23-26,31-36), in the
    reference's own summation order on the CPU, on the integer artefacts the kernels produced
    (hit ranks).  Bit-exact to the reference run on CPU; O(Q*G) memory, for CUHK-PEDES-sized cases."""
    Q = res.first_hit.numel()
    rel_ptr = res.rel_ptr.cpu()
    ranks = res.hit_ranks.cpu().to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(Q), rel_ptr[1:] - rel_ptr[:-1])
    matches = torch.zeros(Q, G, dtype=torch.bool)
    matches[rows, ranks] = True
    topk_t = torch.as_tensor([int(k) for k in (topk.tolist() if torch.is_tensor(topk) else topk)])
    depth = int(topk_t.max())
    all_cmc = matches[:, :depth].cumsum(1)
    all_cmc[all_cmc > 1] = 1
    all_cmc = all_cmc.float().mean(0) * 100
    all_cmc = all_cmc[topk_t - 1]
    num_rel = matches.sum(1)
    tmp_cmc = matches.cumsum(1)
    denom = torch.arange(1, G + 1, dtype=torch.float32)
    tmp_cmc = (tmp_cmc.to(torch.float32) / denom) * matches
    AP = tmp_cmc.sum(1) / num_rel
    return all_cmc, AP.mean() * 100


# --------------------------------------------------------------------------------------------
# rank(): drop-in for evaluation.py:11-37 on a materialised similarity matrix
# --------------------------------------------------------------------------------------------
def rank_artifacts(similarity: torch.Tensor, q_pids: torch.Tensor, g_pids: torch.Tensor, topk=(1, 5, 10),
                   get_mAP: bool = True) -> RetrievalResult:
    _lib.require_cuda(similarity, q_pids, g_pids)
    if similarity.dim() != 2 or similarity.dtype != torch.float32:
        raise ValueError("similarity must be a 2-D float32 tensor")
    lib = _lib.load()
    dev = similarity.device
    Q, G = similarity.shape
    q_pids = q_pids.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    g_pids = g_pids.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    if q_pids.numel() != Q or g_pids.numel() != G:
        raise ValueError("pid vectors do not match the similarity shape")
    rel = build_relevance(q_pids, g_pids) if get_mAP else None
    cand_sim = torch.empty(Q, TOPK_DEPTH, dtype=torch.float32, device=dev)
    cand_idx = torch.empty(Q, TOPK_DEPTH, dtype=torch.int64, device=dev)
    cnt = torch.zeros(max(rel.total, 1), dtype=torch.int32, device=dev) if get_mAP else None
    _lib.check(lib.trb_rank_similarity_f32(
        _lib.ptr(similarity), similarity.stride(0), similarity.stride(1), Q, G,
        _lib.ptr(rel.rel_ptr) if get_mAP else None, rel.col_ptr() if get_mAP else None,
        _lib.ptr(cand_sim), _lib.ptr(cand_idx), _lib.ptr(cnt), _lib.stream_ptr(dev)), "trb_rank_similarity_f32")
    return _finish_and_metrics(cand_sim, cand_idx, 1, q_pids, g_pids, rel, cnt, topk)


def rank(similarity, q_pids, g_pids, topk=[1, 5, 10], get_mAP=True, full_indices=False, parity=False):
    """Same signature and return tuple as the reference's ``rank``: ``(all_cmc, mAP, indices)`` when
    ``get_mAP`` else ``(all_cmc, indices)``; ties are ordered by gallery index (stable sort).

    ``indices`` holds the best max(topk)... 10 gallery indices per query ([Q, 10]); every caller in the
    reference discards it (evaluation.py:147-170).  ``full_indices=True`` returns the full [Q, G]
    permutation for compatibility (computed with torch.sort, outside the hot path).
    ``parity=True`` re-evaluates the final float reductions in the reference's CPU summation order.
    """
    res = rank_artifacts(similarity, q_pids, g_pids, topk, get_mAP)
    cmc, mAP = res.cmc, res.mAP
    if parity and get_mAP:
        cmc, mAP = reference_tail_from_hit_ranks(res, similarity.shape[1], topk)
        cmc, mAP = cmc.to(similarity.device), mAP.to(similarity.device)
    if full_indices and get_mAP:
        indices = torch.argsort(similarity, dim=1, descending=True, stable=True)
    else:
        depth = max(int(k) for k in (topk.tolist() if torch.is_tensor(topk) else topk))
        indices = res.top_idx[:, :depth] if not get_mAP else res.top_idx
    if not get_mAP:
        return cmc, indices
    return cmc, mAP, indices


# --------------------------------------------------------------------------------------------
# retrieve(): fused normalise + similarity + top-k + ranks, no [Q, G] matrix
# --------------------------------------------------------------------------------------------
def l2_normalize_rows(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    _lib.require_cuda(x)
    x = x.contiguous().float()
    y = torch.empty_like(x)
    _lib.check(_lib.load().trb_l2_normalize_rows_f32(_lib.ptr(x), _lib.ptr(y), None, x.shape[0], x.shape[1],
                                                      eps, _lib.stream_ptr(x.device)), "trb_l2_normalize_rows_f32")
    return y


def _choose_nsplit(Q: int, G: int, tile_m: int, tile_n: int, sms: int) -> int:
    q_tiles = max(1, -(-Q // tile_m))
    g_tiles = max(1, -(-G // tile_n))
    want = -(-2 * sms // q_tiles)
    return int(max(1, min(want, g_tiles, 1024)))


def _stream_fp32(qn, gn, g_base, rel_ptr, thr, thr_gidx, cnt, nsplit):
    lib = _lib.load()
    dev = qn.device
    Q, D = qn.shape
    G = gn.shape[0]
    cand_sim = torch.empty(Q, nsplit, TOPK_DEPTH, dtype=torch.float32, device=dev)
    cand_idx = torch.empty(Q, nsplit, TOPK_DEPTH, dtype=torch.int64, device=dev)
    _lib.check(lib.trb_retrieval_stream_f32(
        _lib.ptr(qn), _lib.ptr(gn), Q, G, D, g_base, _lib.ptr(rel_ptr), _lib.ptr(thr), _lib.ptr(thr_gidx), nsplit,
        _lib.ptr(cand_sim), _lib.ptr(cand_idx), _lib.ptr(cnt), _lib.stream_ptr(dev)), "trb_retrieval_stream_f32")
    return cand_sim, cand_idx


def retrieve(text_embed: torch.Tensor, image_embed: torch.Tensor, text_pid: torch.Tensor, image_pid: torch.Tensor,
             topk=(1, 5, 10), get_mAP: bool = True, precision: str = "fp32", normalized: bool = False,
             nsplit: Optional[int] = None) -> RetrievalResult:
    """Embedding-level retrieval evaluation: queries = text, gallery = image (evaluation.py:117-120 + rank).

    precision "fp32": FFMA path, indices / R@k / mAP bit-exact with the stable-sort reference on
    exactly representable inputs, similarities within 1e-5.  "bf16": tcgen05 path (1e-3).
    """
    _lib.require_cuda(text_embed, image_embed, text_pid, image_pid)
    if text_embed.dim() != 2 or image_embed.dim() != 2 or text_embed.shape[1] != image_embed.shape[1]:
        raise ValueError("embeddings must be [Q, D] and [G, D]")
    if precision not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    dev = text_embed.device
    from .sharded import _as_pid, retrieve_sharded_local
    q_pids = _as_pid(text_pid if text_pid.device == dev else text_pid.to(dev))
    g_pids = _as_pid(image_pid if image_pid.device == dev else image_pid.to(dev))
    # one shard, no collectives; the pid bookkeeping (sort orders, relevance CSR, band arrays) is cached in a ShardPlan keyed
    # by the pid tensors, so evaluating the same split again (trainer.py:124: every EVALUATE_PERIOD epochs) only packs,
    # captures thresholds, streams and merges
    return retrieve_sharded_local(text_embed, [image_embed], q_pids, [g_pids], topk, get_mAP, precision, None, nsplit, normalized)


# --------------------------------------------------------------------------------------------
# evaluation() / inference(): the reference's entry points
# --------------------------------------------------------------------------------------------
def first_occurrence_index(image_ids: Sequence[int], device) -> torch.Tensor:
    """evaluation.py:68-73 (get_unique), vectorised: index of the first appearance of each image id,
    in order of first appearance."""
    ids = torch.as_tensor(np.asarray(image_ids), device=device)
    uniq, inverse = torch.unique(ids, return_inverse=True)
    first = torch.full((uniq.numel(),), ids.numel(), dtype=torch.int64, device=device)
    first.scatter_reduce_(0, inverse, torch.arange(ids.numel(), device=device), reduce="amin")
    return torch.sort(first)[0]


def _table(rows, headers):
    try:
        from tabulate import tabulate
        return tabulate(rows, tablefmt="psql", headers=headers, numalign="left")
    except Exception:  # pragma: no cover
        return "\n".join(str(r) for r in [headers] + list(rows))


def eval_precision(precision=None) -> str:
    """Arithmetic path of the evaluation entries: explicit argument, else the environment variable TRB_EVAL_PRECISION, else
    "fp32" -- the path on which indices, R@k and mAP are bit-exact with the stable-sort reference.  "bf16" selects the tcgen05
    stream (the path the 1/2/4/8-GPU throughput numbers are quoted on)."""
    p = precision or os.environ.get("TRB_EVAL_PRECISION", "fp32")
    if p not in ("fp32", "bf16"):
        raise ValueError("evaluation precision must be 'fp32' or 'bf16' (got %r)" % (p,))
    return p


# pid / image-id bookkeeping of a dataset split: pure metadata, identical at every evaluation of the split.  Cached per
# (dataset object, index list) so that (a) the per-item Python loop of evaluation.py:101-106 runs once and (b) the SAME pid
# tensors are handed to retrieve() every time, which is what lets the ShardPlan cache recognise the split.
_ID_CACHE: dict = {}


class SplitInfo:
    def __init__(self, dataset, idx: np.ndarray, dev):
        image_ids, pids = [], []
        for i in idx.tolist():
            image_id, pid = dataset.get_id_info(i)
            image_ids.append(image_id)
            pids.append(pid)
        self.text_pid = torch.as_tensor(np.asarray(pids), device=dev).to(torch.int64).contiguous()
        self.keep = first_occurrence_index(image_ids, dev)
        self.image_pid = self.text_pid[self.keep].contiguous()
        self.shards: dict = {}

    def shard(self, which: str, lo: int, hi: int) -> torch.Tensor:
        key = (which, lo, hi)
        if key not in self.shards:
            src = self.image_pid if which == "image" else self.text_pid
            self.shards[key] = src[lo:hi].contiguous()
        return self.shards[key]


def split_info(dataset, idx, dev) -> SplitInfo:
    idx = np.asarray(idx, dtype=np.int64)
    key = (id(dataset), str(dev), idx.shape[0], hash(idx.tobytes()))
    hit = _ID_CACHE.get(key)
    if hit is None or hit[0] is not dataset:
        if len(_ID_CACHE) >= 4:
            _ID_CACHE.pop(next(iter(_ID_CACHE)))
        hit = (dataset, SplitInfo(dataset, idx, dev))
        _ID_CACHE[key] = hit
    return hit[1]


def _even_bounds(n: int, world: int, rank: int):
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def evaluate_embeddings(dataset, idx, image_all, text_all, output_folder, topk, save_data=True, rerank=True, precision=None,
                        group=None):
    """Core of ``evaluation`` on contiguous tensors: ``idx`` [n] dataset indices (host), ``image_all`` / ``text_all`` [n, D]
    device tensors in the same order (row i = the model's outputs for dataset item idx[i]).  With an initialised process
    group and the trainer's flags (save_data=False, rerank=False) every rank must call it with the SAME full tensors; each
    rank then scores its contiguous slice of the gallery (textreid_b200.sharded) and all ranks return the same value."""
    logger = logging.getLogger("PersonSearch.inference")
    precision = eval_precision(precision)
    dev = text_all.device
    topk_list = [int(k) for k in (topk.tolist() if torch.is_tensor(topk) else topk)]
    info = split_info(dataset, idx, dev)
    image_n = l2_normalize_rows(image_all[info.keep])
    text_n = l2_normalize_rows(text_all)
    results = {}
    from .rerank import jaccard_rerank_matrix, neighbor_lists, rerank_rank, similarity_matrix
    single = isinstance(group, str)         # "single": evaluate in this process even though a process group exists
    world = 1
    if not single and torch.distributed.is_available() and torch.distributed.is_initialized():
        world = torch.distributed.get_world_size(group)
    if not save_data and not rerank:
        # fast path (trainer.py:124): nothing needs the [Q, G] matrix
        if world > 1:
            from .sharded import retrieve_sharded
            rank_ = torch.distributed.get_rank(group)
            G, Qn = image_n.shape[0], text_n.shape[0]
            g_sizes = [_even_bounds(G, world, r)[1] - _even_bounds(G, world, r)[0] for r in range(world)]
            t_sizes = [_even_bounds(Qn, world, r)[1] - _even_bounds(Qn, world, r)[0] for r in range(world)]
            g_lo, g_hi = _even_bounds(G, world, rank_)
            t_lo, t_hi = _even_bounds(Qn, world, rank_)
            t2i = retrieve_sharded(text_n, image_n[g_lo:g_hi], info.text_pid, info.shard("image", g_lo, g_hi), topk_list, False,
                                   precision, group=group, shard_sizes=g_sizes, normalized=True)
            i2t = retrieve_sharded(image_n, text_n[t_lo:t_hi], info.image_pid, info.shard("text", t_lo, t_hi), topk_list, False,
                                   precision, group=group, shard_sizes=t_sizes, normalized=True)
        else:
            t2i = retrieve(text_n, image_n, info.text_pid, info.image_pid, topk_list, get_mAP=False, precision=precision, normalized=True)
            i2t = retrieve(image_n, text_n, info.image_pid, info.text_pid, topk_list, get_mAP=False, precision=precision, normalized=True)
        results["t2i"], results["i2t"] = t2i.cmc, i2t.cmc
        rows = [[k, float(results["t2i"][j]), float(results["i2t"][j])] for j, k in enumerate(topk_list)]
        logger.info("\n" + _table(rows, ["topk", "t2i", "i2t"]))
        evaluation.last_results = results
        return results["t2i"][0]
    return _evaluate_materialised(logger, output_folder, topk_list, save_data, rerank, info.image_pid, info.text_pid, None,
                                  image_n, text_n, None, None)


def evaluation(dataset, predictions, output_folder, topk, save_data=True, rerank=True, precision=None):
    """Drop-in for lib/data/metrics/evaluation.py:76-173.  Returns t2i R@1 (0-d fp32 tensor, percent).

    ``predictions``: {dataset_index: [image_embed(D), text_embed(D)]} or None to read
    ``inference_data.npz`` from ``output_folder`` like the reference.  Re-ranking (k-reciprocal,
    evaluation.py:40-65) is applied when ``rerank`` is True.  ``precision``: see ``eval_precision``.
    """
    logger = logging.getLogger("PersonSearch.inference")
    data_dir = os.path.join(output_folder, "inference_data.npz")
    dev = torch.device("cuda", torch.cuda.current_device())
    topk_list = [int(k) for k in (topk.tolist() if torch.is_tensor(topk) else topk)]
    if predictions is not None:
        keys = list(predictions.keys())
        image_all = torch.stack([predictions[i][0] for i in keys], dim=0).to(dev)
        text_all = torch.stack([predictions[i][1] for i in keys], dim=0).to(dev)
        return evaluate_embeddings(dataset, keys, image_all, text_all, output_folder, topk_list, save_data, rerank, precision)
    data = np.load(data_dir)
    logger.info("Load inference data from {}".format(data_dir))
    image_pid = torch.as_tensor(data["image_pid"], device=dev)
    text_pid = torch.as_tensor(data["text_pid"], device=dev)
    similarity = torch.as_tensor(data["similarity"], device=dev)
    rvn_mat = torch.as_tensor(data["rvn_mat"], device=dev) if rerank else None
    rtn_mat = torch.as_tensor(data["rtn_mat"], device=dev) if rerank else None
    return _evaluate_materialised(logger, output_folder, topk_list, save_data, rerank, image_pid, text_pid, similarity, None, None,
                                  rvn_mat, rtn_mat)


def _evaluate_materialised(logger, output_folder, topk_list, save_data, rerank, image_pid, text_pid, similarity, image_n, text_n,
                           rvn_mat, rtn_mat):
    """The compatibility paths of evaluation.py:120-173 that need the [Q, G] matrix: npz cache, k-reciprocal re-ranking, and the
    cached-npz replay.  fp32 FFMA arithmetic throughout."""
    from .rerank import jaccard_rerank_matrix, neighbor_lists, rerank_rank, similarity_matrix
    data_dir = os.path.join(output_folder, "inference_data.npz")
    results = {}
    nn_t2i = nn_i2t = None
    if similarity is None:
        similarity = similarity_matrix(text_n, image_n)
        if rerank:
            nn_t2i = neighbor_lists(text_n, image_n)      # (text -> image, image -> image)
            nn_i2t = neighbor_lists(image_n, text_n)      # (image -> text, text -> text)
        if save_data:
            payload = dict(image_pid=image_pid.cpu().numpy(), text_pid=text_pid.cpu().numpy(),
                           similarity=similarity.cpu().numpy())
            if rerank:
                payload.update(rvn_mat=jaccard_rerank_matrix(text_n, image_n, nn=nn_t2i).cpu().numpy(),
                               rtn_mat=jaccard_rerank_matrix(image_n, text_n, nn=nn_i2t).cpu().numpy())
            np.savez(data_dir, **payload)
    if rerank:
        sim_t = similarity.t()
        c, m, _ = rank(sim_t, image_pid, text_pid, topk_list, get_mAP=True)
        results["i2t"], results["i2t_mAP"] = c, m
        c, m, _ = rank(similarity, text_pid, image_pid, topk_list, get_mAP=True)
        results["t2i"], results["t2i_mAP"] = c, m
        if nn_t2i is not None:
            r = rerank_rank(sim_t, nn_i2t[0], nn_i2t[1], image_pid, text_pid, topk_list)
            results["re_i2t"], results["re_i2t_mAP"] = r.cmc, r.mAP
            r = rerank_rank(similarity, nn_t2i[0], nn_t2i[1], text_pid, image_pid, topk_list)
            results["re_t2i"], results["re_t2i_mAP"] = r.cmc, r.mAP
        else:   # cached npz: the float64 matrices were loaded; their sum with the similarity is ranked as float64 scores
            for key, mat, s_, qp, gp in (("re_i2t", rtn_mat, sim_t, image_pid, text_pid), ("re_t2i", rvn_mat, similarity, text_pid, image_pid)):
                c, m = _rank_float64_matrix(mat + s_, qp, gp, topk_list)
                results[key], results[key + "_mAP"] = c, m
        cols = ["t2i", "re_t2i", "i2t", "re_i2t"]
        rows = [[k] + [float(results[c][j]) for c in cols] for j, k in enumerate(topk_list)]
        rows.append(["mAP"] + [float(results[c + "_mAP"]) for c in cols])
        logger.info("\n" + _table(rows, ["topk", "t2i", "re-t2i", "i2t", "re-i2t"]))
    else:
        results["t2i"], _ = rank(similarity, text_pid, image_pid, topk_list, get_mAP=False)
        results["i2t"], _ = rank(similarity.t(), image_pid, text_pid, topk_list, get_mAP=False)
        rows = [[k, float(results["t2i"][j]), float(results["i2t"][j])] for j, k in enumerate(topk_list)]
        logger.info("\n" + _table(rows, ["topk", "t2i", "i2t"]))
    evaluation.last_results = results
    return results["t2i"][0]


def _rank_float64_matrix(scores: torch.Tensor, q_pids, g_pids, topk):
    """Cached-npz path: rank an already materialised float64 score matrix (evaluation.py:85-95, 151-156)."""
    dev = scores.device
    Q, G = scores.shape
    q_pids = q_pids.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    g_pids = g_pids.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    rel = build_relevance(q_pids, g_pids)
    cand_sim = torch.empty(Q, TOPK_DEPTH, dtype=torch.float32, device=dev)
    cand_idx = torch.empty(Q, TOPK_DEPTH, dtype=torch.int64, device=dev)
    cnt = torch.zeros(max(rel.total, 1), dtype=torch.int32, device=dev)
    _lib.check(_lib.load().trb_rank_scores_f64(_lib.ptr(scores), scores.stride(0), scores.stride(1), Q, G, _lib.ptr(rel.rel_ptr),
                                               rel.col_ptr(), _lib.ptr(cand_sim), _lib.ptr(cand_idx), _lib.ptr(cnt),
                                               _lib.stream_ptr(dev)), "trb_rank_scores_f64")
    res = _finish_and_metrics(cand_sim, cand_idx, 1, q_pids, g_pids, rel, cnt, topk)
    return res.cmc, res.mAP


def compute_on_dataset(model, data_loader, device):
    """lib/engine/inference.py:14-26, with the per-item dict kept for signature compatibility (callers that want the
    reference's ``{idx: [v, t]}``); ``inference`` itself uses ``compute_on_dataset_tensors``."""
    model.eval()
    results: Dict[int, list] = {}
    for batch in data_loader:
        images, captions, image_ids = batch
        images = images.to(device)
        captions = [c.to(device) for c in captions]
        with torch.no_grad():
            output = model(images, captions)
        for result in output:
            for img_id, pred in zip(image_ids, result):
                results.setdefault(int(img_id), []).append(pred)
    return results


def compute_on_dataset_tensors(model, data_loader, device):
    """The embedding hand-off without the dict: dataset indices [n] (host int64 array) and the model's two outputs as
    contiguous [n, D] device tensors, rows in loader order.  An index the sampler yields twice keeps its first outputs, like
    ``prediction[0], prediction[1]`` of the reference's per-index list (evaluation.py:105-106)."""
    model.eval()
    idx, vs, ts = [], [], []
    for images, captions, image_ids in data_loader:
        images = images.to(device)
        captions = [c.to(device) for c in captions]
        with torch.no_grad():
            v, t = model(images, captions)
        idx.extend(int(i) for i in image_ids)
        vs.append(v)
        ts.append(t)
    idx = np.asarray(idx, dtype=np.int64)
    v_all, t_all = torch.cat(vs, dim=0), torch.cat(ts, dim=0)
    _, first = np.unique(idx, return_index=True)
    if first.shape[0] != idx.shape[0]:
        first = np.sort(first)
        sel = torch.as_tensor(first, device=v_all.device)
        idx, v_all, t_all = idx[first], v_all[sel], t_all[sel]
    return idx, v_all.contiguous(), t_all.contiguous()


def gather_embeddings(idx, v_local, t_local, group=None):
    """Replacement of _accumulate_predictions_from_multiple_gpus (inference.py:29-45, comm.py:47-87: pickle + CPU tensors):
    all-gather of the index vectors and of the [n_r, D] embedding blocks on the device (NCCL), rows then ordered by dataset
    index on EVERY rank (the sharded evaluation needs the full query set everywhere)."""
    from .sharded import _all_gather_varlen
    dev = v_local.device
    idx_t = torch.as_tensor(idx, device=dev)
    idx_all = torch.cat(_all_gather_varlen(idx_t, group))
    v_all = torch.cat(_all_gather_varlen(v_local.contiguous(), group))
    t_all = torch.cat(_all_gather_varlen(t_local.contiguous(), group))
    idx_host = idx_all.cpu().numpy()
    uniq, first = np.unique(idx_host, return_index=True)          # sorted by dataset index; first occurrence wins
    sel = torch.as_tensor(first, device=dev)
    if uniq.shape[0] and uniq.shape[0] != int(uniq[-1]) + 1:
        logging.getLogger("PersonSearch.inference").warning(
            "Number of images that were gathered from multiple processes is not a contiguous set. "
            "Some images might be missing from the evaluation")
    return uniq, v_all[sel].contiguous(), t_all[sel].contiguous()


def inference(model, data_loader, dataset_name="cuhkpedes-test", device="cuda", output_folder="", save_data=True,
              rerank=True, precision=None):
    """Drop-in for lib/engine/inference.py:48-96.  Embeddings stay on the device as contiguous [n, D] tensors from the
    encoder loop to the kernels (no per-item dict, no pickle).  Under an initialised process group every rank encodes its
    part of the split; the blocks are all-gathered over NCCL and -- with the trainer's flags save_data=False, rerank=False
    (trainer.py:124) -- every rank scores its slice of the gallery (``textreid_b200.sharded``).  Like the reference, only the
    main process returns the value (others return None); the compatibility modes that need the [Q, G] matrix (npz cache,
    re-ranking) run on the main process alone."""
    logger = logging.getLogger("PersonSearch.inference")
    dataset = data_loader.dataset
    logger.info("Start evaluation on {} dataset({} images).".format(dataset_name, len(dataset)))
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
    main = (not dist_on) or torch.distributed.get_rank() == 0
    if os.path.exists(os.path.join(output_folder, "inference_data.npz")):
        if not main:
            return None
        return evaluation(dataset=dataset, predictions=None, output_folder=output_folder, save_data=save_data, rerank=rerank,
                          topk=[1, 5, 10], precision=precision)
    idx, v_all, t_all = compute_on_dataset_tensors(model, data_loader, torch.device(device))
    if dist_on:
        idx, v_all, t_all = gather_embeddings(idx, v_all, t_all)
        if save_data or rerank:
            if not main:
                return None
            r = evaluate_embeddings(dataset, idx, v_all, t_all, output_folder, [1, 5, 10], save_data, rerank, precision,
                                    group="single")
            return r
    r = evaluate_embeddings(dataset, idx, v_all, t_all, output_folder, [1, 5, 10], save_data, rerank, precision)
    return r if main else None
