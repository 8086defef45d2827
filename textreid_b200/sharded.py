"""Gallery-sharded retrieval evaluation across GPUs (one process per GPU, torch.distributed / NCCL).

Only the gallery shards: every rank holds all Q queries and a contiguous slice of
the gallery.  Exchange steps, all tiny next to the per-rank GEMM stream:

  0. all-gather the gallery pids                         -> every rank builds the same relevance CSR
  1. per rank: similarities of ITS relevant items        -> all-reduce(sum) of the threshold vector
     (each slot is written by exactly one rank, so the sum with zeros is exact)
  2. per rank: one stream over its slice                 -> all-gather of per-query top-10 candidate
     lists and all-reduce(sum) of the integer rank counts
  3. every rank merges P x lists -> top-10, hit ranks, AP, R@k, mAP (identical results on all ranks)

The reference has no counterpart (it gathers pickled embeddings to rank 0, lib/engine/inference.py:29-45,
lib/utils/comm.py:47-87); the oracle for this path is single-process evaluation on the concatenated gallery.

The per-shard device work goes through a small backend object so that the host logic (CSR bookkeeping,
slot offsets, collectives, merge) can be exercised with gloo on CPU by the tests with a stand-in backend;
the product backend is the CUDA library and nothing else.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from .evaluation import (RelevanceIndex, RetrievalResult, TOPK_DEPTH, _choose_nsplit, _finish_and_metrics, _sm_count,
                         _stream_fp32, build_relevance, l2_normalize_rows)


class CudaBackend:
    """Device work of one shard, served by libtextreid_b200.so."""

    name = "cuda"

    def normalize(self, x):
        return l2_normalize_rows(x)

    def thresholds_fp32(self, qn, gn, rel_ptr, rel_row, thr):
        _lib.check(_lib.load().trb_retrieval_thresholds_f32(_lib.ptr(qn), _lib.ptr(gn), _lib.ptr(rel_ptr), _lib.ptr(rel_row),
                                                            _lib.ptr(thr), qn.shape[0], qn.shape[1],
                                                            _lib.stream_ptr(qn.device)), "trb_retrieval_thresholds_f32")

    def stream_fp32(self, qn, gn, g_base, rel_ptr, thr, thr_gidx, cnt, nsplit):
        return _stream_fp32(qn, gn, g_base, rel_ptr, thr, thr_gidx, cnt, nsplit)

    def nsplit(self, Q, G, device):
        return _choose_nsplit(Q, G, 128, 128, _sm_count(device))

    def finish(self, cand_sim, cand_idx, nlists, q_pids, g_pids, rel, cnt, topk):
        return _finish_and_metrics(cand_sim, cand_idx, nlists, q_pids, g_pids, rel, cnt, topk)

    def merge_lists(self, cand_sim, cand_idx, q_pids, g_pids):
        """[Q, L, 10] candidate lists -> [Q, 1, 10]: the rank-local merge that keeps the all-gather at 120 B/query."""
        Q, L, K = cand_sim.shape
        if L == 1:
            return cand_sim, cand_idx
        dev = cand_sim.device
        top_sim = torch.empty(Q, 1, K, dtype=torch.float32, device=dev)
        top_idx = torch.empty(Q, 1, K, dtype=torch.int64, device=dev)
        _lib.check(_lib.load().trb_retrieval_finish(
            _lib.ptr(cand_sim), _lib.ptr(cand_idx), L, Q, _lib.ptr(q_pids), _lib.ptr(g_pids), g_pids.numel(), None, None,
            _lib.ptr(top_sim), _lib.ptr(top_idx), None, None, None, _lib.stream_ptr(dev)), "trb_retrieval_finish")
        return top_sim, top_idx


class ShardWorker:
    """State of one gallery shard between the exchange steps."""

    def __init__(self, text_embed, image_shard, q_pids, g_pids_all, g_base, get_mAP, precision, backend, normalized=False):
        self.backend = backend
        self.precision = precision
        self.get_mAP = get_mAP
        self.g_base = int(g_base)
        self.q_pids = q_pids
        self.g_pids_all = g_pids_all
        self.Gs = image_shard.shape[0]
        self.dev = text_embed.device
        self.record_events = False       # bench.py: CUDA events around the stream kernel alone
        self.stream_events = None
        self.rel: Optional[RelevanceIndex] = build_relevance(q_pids, g_pids_all) if get_mAP else None
        self.max_rel = 0
        if get_mAP and self.rel.total > 0:
            # one more host read next to build_relevance's: picks the 4- or 8-threshold kernel variant
            self.max_rel = int((self.rel.rel_ptr[1:] - self.rel.rel_ptr[:-1]).max().item())
        if precision == "fp32":
            self.qn = text_embed.contiguous().float() if normalized else backend.normalize(text_embed)
            self.gn = image_shard.contiguous().float() if normalized else backend.normalize(image_shard)
        elif precision == "bf16":
            self._prepare_tc(text_embed, image_shard, normalized)
        else:
            raise ValueError("precision must be 'fp32' or 'bf16'")

    # ---- bf16 tensor-core preparation: packed operands and band bookkeeping ----
    def _prepare_tc(self, text_embed, image_shard, normalized):
        """Queries are packed once, in pid order.  The gallery is packed in INDEX order for the stream (ties are then
        resolved by position) and, when ranks are wanted, a second time in pid order for the threshold capture, where the
        relevant items of a 128-query tile form one contiguous band."""
        from .retrieval_tc import pack_rows
        lib = _lib.load()
        dev = self.dev
        Q = text_embed.shape[0]
        self.Qp, self.Gp = int(lib.trb_packed_rows(Q)), int(lib.trb_packed_rows(self.Gs))
        q_sorted, q_order = torch.sort(self.q_pids, stable=True)
        self.q_packed = pack_rows(text_embed, q_order, normalize=not normalized)
        self.q_row_id = torch.full((self.Qp,), -1, dtype=torch.int64, device=dev)
        self.q_row_id[:Q] = q_order
        self.g_packed = pack_rows(image_shard, None, normalize=not normalized)
        if self.get_mAP:
            rel = self.rel
            g_pids_local = self.g_pids_all[self.g_base:self.g_base + self.Gs]
            g_sorted, g_order = torch.sort(g_pids_local, stable=True)
            self.g_packed_pid = pack_rows(image_shard, g_order, normalize=not normalized)
            self.g_row_id = torch.full((self.Gp,), -1, dtype=torch.int64, device=dev)
            self.g_row_id[:self.Gs] = g_order + self.g_base
            self.band_lo = torch.zeros(self.Qp, dtype=torch.int32, device=dev)
            self.band_hi = torch.zeros(self.Qp, dtype=torch.int32, device=dev)
            self.band_lo[:Q] = torch.searchsorted(g_sorted, q_sorted, right=False).to(torch.int32)
            self.band_hi[:Q] = torch.searchsorted(g_sorted, q_sorted, right=True).to(torch.int32)
            # slots of a query are ordered by global gallery index: this shard's items start after the
            # items that live on lower shards
            below = torch.zeros(rel.total + 1, dtype=torch.int64, device=dev)
            torch.cumsum((rel.rel_col < self.g_base).to(torch.int64), 0, out=below[1:])
            off = below[rel.rel_ptr[1:]] - below[rel.rel_ptr[:-1]]
            self.rel_off = torch.zeros(self.Qp, dtype=torch.int32, device=dev)
            self.rel_off[:Q] = off[q_order].to(torch.int32)

    # ---- step 1 ----
    def local_thresholds(self) -> torch.Tensor:
        rel = self.rel
        thr = torch.zeros(max(rel.total, 1), dtype=torch.float32, device=self.dev)
        if self.precision == "fp32":
            local = rel.rel_col - self.g_base
            rel_row = torch.where((local >= 0) & (local < self.Gs), local, torch.full_like(local, -1)).contiguous()
            if rel_row.numel() == 0:
                rel_row = torch.full((1,), -1, dtype=torch.int64, device=self.dev)
            self.backend.thresholds_fp32(self.qn, self.gn, rel.rel_ptr, rel_row, thr)
        else:
            scratch_gidx = torch.zeros(max(rel.total, 1), dtype=torch.int64, device=self.dev)
            Q = self.q_pids.numel()
            D = self.q_packed.numel() // (2 * self.Qp)
            _lib.check(_lib.load().trb_retrieval_stream_tc(
                _lib.ptr(self.q_packed), _lib.ptr(self.g_packed_pid), Q, self.Gs, D, _lib.ptr(self.q_row_id),
                _lib.ptr(self.g_row_id), self.g_base, _lib.ptr(rel.rel_ptr), _lib.ptr(thr), _lib.ptr(scratch_gidx), _lib.ptr(self.band_lo),
                _lib.ptr(self.band_hi), _lib.ptr(self.rel_off), 1, 1, self.max_rel, None, None, None, _lib.stream_ptr(self.dev)),
                "trb_retrieval_stream_tc(mode=1)")
        return thr

    # ---- step 2 ----
    def stream(self, thr: Optional[torch.Tensor], nsplit: Optional[int] = None):
        rel = self.rel
        cnt = torch.zeros(max(rel.total, 1), dtype=torch.int32, device=self.dev) if self.get_mAP else None
        Q = self.q_pids.numel()
        if self.precision == "fp32":
            ns = nsplit or self.backend.nsplit(Q, self.Gs, self.dev)
            gidx = rel.store if self.get_mAP else None
            ev = self._events()
            cand_sim, cand_idx = self.backend.stream_fp32(self.qn, self.gn, self.g_base, rel.rel_ptr if self.get_mAP else None,
                                                          thr, gidx, cnt, ns)
            self._events(ev)
            return cand_sim, cand_idx, cnt
        from .retrieval_tc import choose_nsplit_tc
        lib = _lib.load()
        D = self.q_packed.numel() // (2 * self.Qp)
        num_gtiles = -(-self.Gs // 256)
        ns = nsplit or choose_nsplit_tc(-(-Q // 128), num_gtiles, _sm_count(self.dev))
        ns = max(1, min(ns, num_gtiles))
        lists = int(lib.trb_retrieval_tc_lists_per_split()) * ns
        cand_sim = torch.empty(Q, lists, TOPK_DEPTH, dtype=torch.float32, device=self.dev)
        cand_idx = torch.empty(Q, lists, TOPK_DEPTH, dtype=torch.int64, device=self.dev)
        gidx = None
        if self.get_mAP:
            gidx = rel.store
        ev = self._events()
        _lib.check(lib.trb_retrieval_stream_tc(
            _lib.ptr(self.q_packed), _lib.ptr(self.g_packed), Q, self.Gs, D, _lib.ptr(self.q_row_id), None, self.g_base,
            _lib.ptr(rel.rel_ptr) if self.get_mAP else None, _lib.ptr(thr), _lib.ptr(gidx), None, None, None, 0, ns,
            self.max_rel, _lib.ptr(cand_sim), _lib.ptr(cand_idx), _lib.ptr(cnt), _lib.stream_ptr(self.dev)), "trb_retrieval_stream_tc(mode=0)")
        self._events(ev)
        return cand_sim, cand_idx, cnt

    def _events(self, started=None):
        if not self.record_events:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        if started is None:
            return e
        self.stream_events = (started, e)
        return None


def _finish(backend, cand_sims, cand_idxs, q_pids, g_pids_all, rel, cnt, topk):
    cand_sim = torch.cat(cand_sims, dim=1).contiguous()
    cand_idx = torch.cat(cand_idxs, dim=1).contiguous()
    return backend.finish(cand_sim, cand_idx, cand_sim.shape[1], q_pids, g_pids_all, rel, cnt, topk)


def retrieve_sharded_local(text_embed, image_shards: Sequence[torch.Tensor], text_pid, image_pid_shards, topk=(1, 5, 10),
                           get_mAP=True, precision="fp32", backend=None, nsplit=None) -> RetrievalResult:
    """All shards processed by ONE process, with the collectives replaced by local sums / concatenation.
    Same code path per shard as the distributed driver; used to validate the exchange protocol."""
    backend = backend or CudaBackend()
    q_pids = text_pid.reshape(-1).to(torch.int64).contiguous()
    g_pids_all = torch.cat([p.reshape(-1).to(torch.int64) for p in image_pid_shards]).contiguous()
    bases, b = [], 0
    for s in image_shards:
        bases.append(b)
        b += s.shape[0]
    workers = [ShardWorker(text_embed, s, q_pids, g_pids_all, base, get_mAP, precision, backend)
               for s, base in zip(image_shards, bases)]
    thr = None
    if get_mAP:
        thr = torch.stack([w.local_thresholds() for w in workers]).sum(0)
    outs = [w.stream(thr, nsplit) for w in workers]
    cnt = torch.stack([o[2] for o in outs]).sum(0).to(torch.int32) if get_mAP else None
    res = _finish(backend, [o[0] for o in outs], [o[1] for o in outs], q_pids, g_pids_all, workers[0].rel, cnt, topk)
    res.thresholds = thr[:workers[0].rel.total] if get_mAP else None
    return res


def _all_gather_varlen(t: torch.Tensor, group=None) -> List[torch.Tensor]:
    """all_gather of 1-D or [n, ...] tensors whose first dimension differs per rank (pad to the max)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:s] for b, s in zip(bufs, sizes)]


def retrieve_sharded(text_embed, image_shard, text_pid, image_pid_shard, topk=(1, 5, 10), get_mAP=True, precision="fp32",
                     group=None, backend=None, nsplit=None) -> RetrievalResult:
    """Distributed driver: call on every rank with the full query set and this rank's gallery slice
    (slices are contiguous and ordered by rank).  Returns the same RetrievalResult on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("retrieve_sharded needs an initialised torch.distributed process group")
    backend = backend or CudaBackend()
    rank = dist.get_rank(group)
    q_pids = text_pid.reshape(-1).to(torch.int64).contiguous()
    pid_parts = _all_gather_varlen(image_pid_shard.reshape(-1).to(torch.int64).contiguous(), group)
    g_pids_all = torch.cat(pid_parts).contiguous()
    g_base = sum(p.numel() for p in pid_parts[:rank])
    w = ShardWorker(text_embed, image_shard, q_pids, g_pids_all, g_base, get_mAP, precision, backend)
    thr = None
    if get_mAP:
        thr = w.local_thresholds()
        dist.all_reduce(thr, op=dist.ReduceOp.SUM, group=group)      # one non-zero contributor per slot: exact
    cand_sim, cand_idx, cnt = w.stream(thr, nsplit)
    if get_mAP:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    # candidate lists: merge this rank's L lists to one per query, then all-gather [Q, 10] x (fp32, int64)
    world = dist.get_world_size(group)
    if hasattr(backend, "merge_lists"):
        cand_sim, cand_idx = backend.merge_lists(cand_sim, cand_idx, q_pids, g_pids_all)
        sims = torch.empty((world,) + tuple(cand_sim.shape), dtype=cand_sim.dtype, device=cand_sim.device)
        idxs = torch.empty((world,) + tuple(cand_idx.shape), dtype=cand_idx.dtype, device=cand_idx.device)
        dist.all_gather_into_tensor(sims, cand_sim.contiguous(), group=group)
        dist.all_gather_into_tensor(idxs, cand_idx.contiguous(), group=group)
        sim_parts, idx_parts = [sims[r] for r in range(world)], [idxs[r] for r in range(world)]
    else:   # stand-in backends (tests): list counts may differ per rank -> gather along a leading list axis
        sim_parts = [x.permute(1, 0, 2) for x in _all_gather_varlen(cand_sim.permute(1, 0, 2).contiguous(), group)]
        idx_parts = [x.permute(1, 0, 2) for x in _all_gather_varlen(cand_idx.permute(1, 0, 2).contiguous(), group)]
    res = _finish(backend, sim_parts, idx_parts, q_pids, g_pids_all, w.rel, cnt, topk)
    res.thresholds = thr[:w.rel.total] if get_mAP else None
    return res
