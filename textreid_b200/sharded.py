"""Gallery-sharded retrieval evaluation across GPUs (one process per GPU, torch.distributed / NCCL).

Only the gallery shards: every rank holds all Q queries and a contiguous slice of
the gallery.  Exchange steps, all tiny next to the per-rank GEMM stream:

  0. once per (dataset, sharding), cached in a ShardPlan: all-gather of the per-query COUNTS of relevant items and
     all-reduce of the slot -> global gallery index table  -> every rank builds the same relevance CSR
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
                         _stream_fp32, _topk_host_array, l2_normalize_rows)


class PhaseTimer:
    """Optional CUDA-event phase marks (TRB_PROFILE_PHASES=1): where a step's time goes besides the stream kernel."""
    enabled = bool(int(__import__("os").environ.get("TRB_PROFILE_PHASES", "0")))
    marks: list = []

    @classmethod
    def mark(cls, name):
        if cls.enabled and torch.cuda.is_available():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            cls.marks.append((name, e, __import__("time").perf_counter()))

    @classmethod
    def report(cls):
        if not cls.marks:
            return ""
        torch.cuda.synchronize()
        out = []
        for (n0, e0, t0), (n1, e1, t1) in zip(cls.marks[:-1], cls.marks[1:]):
            out.append("%s: gpu %.3f ms host %.3f ms" % (n1, e0.elapsed_time(e1), (t1 - t0) * 1e3))
        cls.marks = []
        return "; ".join(out)


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

    def finish_partial(self, sim_my, idx_my, my_ptr, cnt, total):
        """top-10 / first hit / AP (and the sorted hit ranks, written into a zeroed full-size buffer) for a block of
        queries whose candidate lists are [Qc, L, 10] and whose CSR slice is my_ptr [Qc+1] (global slot numbers)."""
        lib = _lib.load()
        dev = sim_my.device
        Qc, L, K = sim_my.shape
        top_sim = torch.empty(Qc, K, dtype=torch.float32, device=dev)
        top_idx = torch.empty(Qc, K, dtype=torch.int64, device=dev)
        first_hit = torch.empty(Qc, dtype=torch.int32, device=dev)
        ap = torch.empty(Qc, dtype=torch.float32, device=dev)
        hit_ranks = torch.zeros(max(total, 1), dtype=torch.int32, device=dev)
        _lib.check(lib.trb_retrieval_finish(
            _lib.ptr(sim_my), _lib.ptr(idx_my), L, Qc, None, None, 0, _lib.ptr(my_ptr), _lib.ptr(cnt),
            _lib.ptr(top_sim), _lib.ptr(top_idx), _lib.ptr(first_hit), _lib.ptr(hit_ranks), _lib.ptr(ap),
            _lib.stream_ptr(dev)), "trb_retrieval_finish")
        return top_sim, top_idx, first_hit, ap, hit_ranks

    def metrics(self, first_hit, ap, topk):
        arr, n, _ = _topk_host_array(topk)
        dev = first_hit.device
        cmc = torch.empty(n, dtype=torch.float32, device=dev)
        mAP = torch.empty((), dtype=torch.float32, device=dev)
        _lib.check(_lib.load().trb_retrieval_metrics(_lib.ptr(first_hit), _lib.ptr(ap), first_hit.numel(), arr, n, _lib.ptr(cmc),
                                                     _lib.ptr(mAP), _lib.stream_ptr(dev)), "trb_retrieval_metrics")
        return cmc, mAP

    def merge_lists(self, cand_sim, cand_idx):
        """[Q, L, 10] candidate lists -> [Q, 1, 10]: the rank-local merge that keeps the all-gather at 120 B/query."""
        Q, L, K = cand_sim.shape
        if L == 1:
            return cand_sim, cand_idx
        dev = cand_sim.device
        top_sim = torch.empty(Q, 1, K, dtype=torch.float32, device=dev)
        top_idx = torch.empty(Q, 1, K, dtype=torch.int64, device=dev)
        _lib.check(_lib.load().trb_retrieval_finish(
            _lib.ptr(cand_sim), _lib.ptr(cand_idx), L, Q, None, None, 0, None, None,
            _lib.ptr(top_sim), _lib.ptr(top_idx), None, None, None, _lib.stream_ptr(dev)), "trb_retrieval_finish")
        return top_sim, top_idx


class ShardPlan:
    """Everything about one gallery shard that depends on the PIDS only (not on the embeddings): sort orders, the relevance
    CSR, this rank's slot offsets, the global gallery index of every slot, the band arrays of the tensor-core threshold
    capture.  The reference evaluates the same test split every EVALUATE_PERIOD epochs (trainer.py:124) with a new model each
    time, so a plan is built once per (dataset, sharding) and reused: the per-evaluation work is then pack + thresholds +
    stream + merge, with no sort, no host read and one collective less.

    Slot layout of the relevance CSR (one slot per (query, relevant gallery item)): the slots of query q are ordered by
    global gallery index, i.e. by (owning rank, position in the rank's pid-sorted shard).  A rank therefore only needs
    every rank's per-query COUNT of relevant items to place its own slots: no global pid gather, no global sort."""

    def __init__(self, q_pids, g_pids_local, g_base, get_mAP, precision):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision, self.get_mAP, self.g_base = precision, bool(get_mAP), int(g_base)
        self.q_pids, self.g_pids_local = q_pids, g_pids_local
        self.Q, self.Gs, self.dev = q_pids.numel(), g_pids_local.numel(), q_pids.device
        self.rel: Optional[RelevanceIndex] = None
        self.max_rel = self.total = self.total_local = 0
        self.g_pids_all = None            # top-k-only mode: pids of every shard (first hit inside the top-10)
        self.shard_sizes = None
        if precision == "bf16" or get_mAP:
            self.q_sorted, self.q_order = torch.sort(q_pids, stable=True)
        if get_mAP:
            self.g_sorted, self.g_order = torch.sort(g_pids_local, stable=True)   # stable: ascending index inside a pid
            self.lo_s = torch.searchsorted(self.g_sorted, self.q_sorted, right=False)   # per pid-sorted query row
            self.hi_s = torch.searchsorted(self.g_sorted, self.q_sorted, right=True)

    # ---- step 0: how many relevant items of every query live on this shard ----
    def local_counts(self) -> torch.Tensor:
        counts = torch.zeros(self.Q, dtype=torch.int32, device=self.dev)
        counts[self.q_order] = (self.hi_s - self.lo_s).to(torch.int32)
        return counts

    def set_layout(self, counts_all: torch.Tensor, rank: int) -> None:
        """counts_all [P, Q] int32 (rank-major).  Fixes the CSR, this rank's offsets and the local slot arithmetic; ONE host
        read, once per plan."""
        dev = self.dev
        c64 = counts_all.to(torch.int64)
        per_q = c64.sum(0)
        rel_ptr = torch.zeros(self.Q + 1, dtype=torch.int64, device=dev)
        torch.cumsum(per_q, 0, out=rel_ptr[1:])
        off = c64[:rank].sum(0) if rank > 0 else torch.zeros(self.Q, dtype=torch.int64, device=dev)
        local = c64[rank]
        host = torch.stack([rel_ptr[-1], per_q.max() if self.Q else rel_ptr[-1], local.sum()]).cpu()
        total, self.max_rel, self.total_local = int(host[0]), int(host[1]), int(host[2])
        self.total = total
        self.gidx_store = torch.zeros(max(total, 1), dtype=torch.int64, device=dev)
        self.rel = RelevanceIndex(rel_ptr, self.gidx_store[:total], total, self.gidx_store)
        self.off_q = off                                   # [Q] slots of lower ranks, original query order
        # local slots: query (pid-sorted row i) owns sorted-gallery rows [lo_s[i], hi_s[i])
        cnt_s = self.hi_s - self.lo_s
        row_i = torch.repeat_interleave(torch.arange(self.Q, device=dev), cnt_s, output_size=self.total_local)
        start = torch.cumsum(cnt_s, 0) - cnt_s
        within = torch.arange(self.total_local, device=dev) - start[row_i]
        q = self.q_order[row_i]
        self.slot = rel_ptr[q] + off[q] + within                         # [total_local] global slot of every local slot
        self.slot_row = self.g_order[self.lo_s[row_i] + within]          # its LOCAL gallery row
        # global gallery index of this shard's slots (zeros elsewhere); summed over the ranks by set_gidx
        self.gidx_store[self.slot] = self.slot_row + self.g_base
        if self.precision == "fp32":
            self.rel_row = torch.full((max(total, 1),), -1, dtype=torch.int64, device=dev)
            self.rel_row[self.slot] = self.slot_row
        else:
            lib = _lib.load()
            self.Qp, self.Gp = int(lib.trb_packed_rows(self.Q)), int(lib.trb_packed_rows(self.Gs))
            self.rel_off = torch.zeros(self.Qp, dtype=torch.int32, device=dev)
            self.rel_off[:self.Q] = off[self.q_order].to(torch.int32)
            self.g_row_id = torch.full((self.Gp,), -1, dtype=torch.int64, device=dev)
            self.g_row_id[:self.Gs] = self.g_order + self.g_base
            self.band_lo = torch.zeros(self.Qp, dtype=torch.int32, device=dev)
            self.band_hi = torch.zeros(self.Qp, dtype=torch.int32, device=dev)
            self.band_lo[:self.Q] = self.lo_s.to(torch.int32)
            self.band_hi[:self.Q] = self.hi_s.to(torch.int32)
            self.gidx_scratch = torch.zeros(max(total, 1), dtype=torch.int64, device=dev)   # the capture kernel's own copy

    def finish_common(self):
        if self.precision == "bf16":
            lib = _lib.load()
            if not hasattr(self, "Qp"):
                self.Qp, self.Gp = int(lib.trb_packed_rows(self.Q)), int(lib.trb_packed_rows(self.Gs))
            self.q_row_id = torch.full((self.Qp,), -1, dtype=torch.int64, device=self.dev)
            self.q_row_id[:self.Q] = self.q_order


class ShardWorker:
    """Per-evaluation state of one gallery shard between the exchange steps: the normalised / packed embeddings and the
    thresholds.  Everything pid-derived comes from the ShardPlan."""

    def __init__(self, text_embed, image_shard, plan: ShardPlan, backend, normalized=False):
        self.backend, self.plan = backend, plan
        self.precision, self.get_mAP, self.g_base = plan.precision, plan.get_mAP, plan.g_base
        self.Q, self.Gs, self.dev = plan.Q, plan.Gs, text_embed.device
        if text_embed.shape[0] != plan.Q or image_shard.shape[0] != plan.Gs:
            raise ValueError("embeddings do not match the plan (%d x %d expected)" % (plan.Q, plan.Gs))
        self.record_events = False       # bench.py: CUDA events around the stream kernel alone
        self.stream_events = None
        self.rel, self.max_rel = plan.rel, plan.max_rel
        if self.precision == "fp32":
            self.qn = text_embed.contiguous().float() if normalized else backend.normalize(text_embed)
            self.gn = image_shard.contiguous().float() if normalized else backend.normalize(image_shard)
        else:
            from .retrieval_tc import pack_rows
            self.D = text_embed.shape[1]
            # queries once, in pid order; gallery in INDEX order for the stream (ties resolved by position) and, when
            # ranks are wanted, again in pid order for the threshold capture (relevant items form a contiguous band)
            self.q_packed = pack_rows(text_embed, plan.q_order, normalize=not normalized)
            self.g_packed = pack_rows(image_shard, None, normalize=not normalized)
            if self.get_mAP:
                self.g_packed_pid = pack_rows(image_shard, plan.g_order, normalize=not normalized)

    # ---- step 1: similarities of this shard's relevant items; zeros elsewhere ----
    def local_thresholds(self):
        plan, rel, dev = self.plan, self.plan.rel, self.dev
        thr = torch.zeros(max(rel.total, 1), dtype=torch.float32, device=dev)
        if self.precision == "fp32":
            self.backend.thresholds_fp32(self.qn, self.gn, rel.rel_ptr, plan.rel_row, thr)
        else:
            _lib.check(_lib.load().trb_retrieval_stream_tc(
                _lib.ptr(self.q_packed), _lib.ptr(self.g_packed_pid), self.Q, self.Gs, self.D, _lib.ptr(plan.q_row_id),
                _lib.ptr(plan.g_row_id), self.g_base, _lib.ptr(rel.rel_ptr), _lib.ptr(thr), _lib.ptr(plan.gidx_scratch),
                _lib.ptr(plan.band_lo), _lib.ptr(plan.band_hi), _lib.ptr(plan.rel_off), 1, 1, self.max_rel, None, None, None,
                _lib.stream_ptr(dev)), "trb_retrieval_stream_tc(mode=1)")
        return thr

    def set_thresholds(self, thr: torch.Tensor) -> None:
        """Thresholds of ALL shards (after the exchange)."""
        self.thr = thr

    # ---- step 2 ----
    def stream(self, nsplit: Optional[int] = None):
        plan, rel = self.plan, self.plan.rel
        cnt = torch.zeros(max(rel.total, 1), dtype=torch.int32, device=self.dev) if self.get_mAP else None
        thr = self.thr if self.get_mAP else None
        gidx = plan.gidx_store if self.get_mAP else None
        Q = self.Q
        if self.precision == "fp32":
            ns = nsplit or self.backend.nsplit(Q, self.Gs, self.dev)
            ev = self._events()
            cand_sim, cand_idx = self.backend.stream_fp32(self.qn, self.gn, self.g_base, rel.rel_ptr if self.get_mAP else None,
                                                          thr, gidx, cnt, ns)
            self._events(ev)
            return cand_sim, cand_idx, cnt
        from .retrieval_tc import choose_nsplit_tc
        lib = _lib.load()
        num_gtiles = -(-self.Gs // 256)
        ns = nsplit or choose_nsplit_tc(-(-Q // 128), num_gtiles, _sm_count(self.dev))
        ns = max(1, min(ns, num_gtiles))
        lists = int(lib.trb_retrieval_tc_lists_per_split()) * ns
        cand_sim = torch.empty(Q, lists, TOPK_DEPTH, dtype=torch.float32, device=self.dev)
        cand_idx = torch.empty(Q, lists, TOPK_DEPTH, dtype=torch.int64, device=self.dev)
        ev = self._events()
        _lib.check(lib.trb_retrieval_stream_tc(
            _lib.ptr(self.q_packed), _lib.ptr(self.g_packed), Q, self.Gs, self.D, _lib.ptr(plan.q_row_id), None, self.g_base,
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


# ---- plan cache: keyed by the identity (address, version, size) of the pid tensors, which the entries keep alive ----
_PLAN_CACHE: dict = {}
_PLAN_CACHE_CAP = 8
plan_cache_stats = {"hits": 0, "misses": 0}


def _plan_key(q_pids, g_pids_list, get_mAP, precision, tag):
    k = [q_pids.data_ptr(), q_pids._version, q_pids.numel(), str(q_pids.device), bool(get_mAP), precision, tag]
    for g in g_pids_list:
        k += [g.data_ptr(), g._version, g.numel()]
    return tuple(k)


def _cache_get(key, q_pids=None, g_pids_list=None, group=None, distributed=False):
    """Look a plan up by the identity of its pid tensors; on a miss, fall back to CONTENT equality with the cached entries of
    the same shapes (an evaluation loop that re-uploads the split's pid vectors every time creates new tensors with the same
    values).  The content check is one device-side comparison and a host read -- far cheaper than re-planning (two sorts, the
    slot arithmetic, two collectives); under a process group the verdict is min-reduced so that every rank takes the same
    branch."""
    e = _PLAN_CACHE.get(key)
    if e is None and q_pids is not None:
        for k2, cand in list(_PLAN_CACHE.items()):
            cq, cg = cand[1], cand[2]
            cg = cg if isinstance(cg, (list, tuple)) else [cg]
            if k2[3:7] != key[3:7] or cq.numel() != q_pids.numel() or len(cg) != len(g_pids_list):
                continue
            if any(a.numel() != b.numel() for a, b in zip(cg, g_pids_list)):
                continue
            same = torch.equal(cq, q_pids) and all(torch.equal(a, b) for a, b in zip(cg, g_pids_list))
            if distributed:
                flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=q_pids.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                same = bool(int(flag.item()))
            if same:
                e = cand
            break
    plan_cache_stats["hits" if e is not None else "misses"] += 1
    return e


def _cache_put(key, value):
    if len(_PLAN_CACHE) >= _PLAN_CACHE_CAP:
        _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
    _PLAN_CACHE[key] = value


def clear_plan_cache():
    _PLAN_CACHE.clear()


def _as_pid(t: torch.Tensor) -> torch.Tensor:
    """int64 contiguous 1-D view; returns the SAME tensor object when it already is one, so that the plan cache can recognise
    the pid vectors of the previous evaluation."""
    if t.dtype is torch.int64 and t.dim() == 1 and t.is_contiguous():
        return t
    return t.reshape(-1).to(torch.int64).contiguous()


def _finish(backend, cand_sims, cand_idxs, q_pids, g_pids_all, rel, cnt, topk):
    cand_sim = torch.cat(cand_sims, dim=1).contiguous() if len(cand_sims) > 1 else cand_sims[0].contiguous()
    cand_idx = torch.cat(cand_idxs, dim=1).contiguous() if len(cand_idxs) > 1 else cand_idxs[0].contiguous()
    if g_pids_all is None:      # ranks come from the counts; candidate pids are not needed
        g_pids_all = torch.zeros(1, dtype=torch.int64, device=q_pids.device)[:0]
    return backend.finish(cand_sim, cand_idx, cand_sim.shape[1], q_pids, g_pids_all, rel, cnt, topk)


def _local_plans(q_pids, pids, bases, get_mAP, precision):
    """Plans of all shards of a single-process (multi-shard) evaluation, cached on the identity of the pid tensors."""
    key = _plan_key(q_pids, pids, get_mAP, precision, ("local",) + tuple(bases))
    hit = _cache_get(key, q_pids, pids)
    if hit is not None:
        return hit[0]
    plans = [ShardPlan(q_pids, p, base, get_mAP, precision) for p, base in zip(pids, bases)]
    if get_mAP:
        counts_all = torch.stack([pl.local_counts() for pl in plans])
        for r, pl in enumerate(plans):
            pl.set_layout(counts_all, r)
        gidx = torch.stack([pl.gidx_store for pl in plans]).sum(0)          # one contributor per slot
        for pl in plans:
            pl.gidx_store.copy_(gidx)
    else:
        g_all = torch.cat(pids).contiguous()
        for pl in plans:
            pl.g_pids_all = g_all
    for pl in plans:
        pl.finish_common()
    _cache_put(key, (plans, q_pids, pids))      # the entry keeps the pid tensors (and so their addresses) alive
    return plans


def retrieve_sharded_local(text_embed, image_shards: Sequence[torch.Tensor], text_pid, image_pid_shards, topk=(1, 5, 10),
                           get_mAP=True, precision="fp32", backend=None, nsplit=None, normalized=False) -> RetrievalResult:
    """All shards processed by ONE process, with the collectives replaced by local sums / concatenation.
    Same code path per shard as the distributed driver; used to validate the exchange protocol (and, with a single
    shard, as the single-GPU tensor-core entry)."""
    backend = backend or CudaBackend()
    q_pids = _as_pid(text_pid)
    pids = [_as_pid(p) for p in image_pid_shards]
    bases, b = [], 0
    for s in image_shards:
        bases.append(b)
        b += s.shape[0]
    plans = _local_plans(q_pids, pids, bases, get_mAP, precision)
    workers = [ShardWorker(text_embed, s, pl, backend, normalized) for s, pl in zip(image_shards, plans)]
    thr = None
    if get_mAP:
        parts = [w.local_thresholds() for w in workers]
        thr = torch.stack(parts).sum(0) if len(parts) > 1 else parts[0]
        for w in workers:
            w.set_thresholds(thr)
    outs = [w.stream(nsplit) for w in workers]
    cnt = None
    if get_mAP:
        cnt = torch.stack([o[2] for o in outs]).sum(0).to(torch.int32) if len(outs) > 1 else outs[0][2]
    res = _finish(backend, [o[0] for o in outs], [o[1] for o in outs], q_pids, plans[0].g_pids_all, plans[0].rel, cnt, topk)
    res.thresholds = thr[:plans[0].rel.total] if get_mAP else None
    return res


def _pack_f32_i32(f: torch.Tensor, i: torch.Tensor) -> torch.Tensor:
    """(fp32, index < 2^31) -> one int64 per pair, so that one collective moves both."""
    return (i.to(torch.int64) << 32) | (f.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF)


def _unpack_f32_i32(p: torch.Tensor):
    f = (p & 0xFFFFFFFF).to(torch.int32).view(torch.float32)
    return f, p >> 32


def _finish_scattered(backend, cand_sim, cand_idx, plan, cnt, topk, group, want_hit_ranks):
    """Finish with the queries partitioned over the ranks: an all-to-all hands every rank the candidate list of ITS
    Q/P queries from every rank (8 B per candidate), each rank derives top-10 / first hit / AP for them, and the small
    per-query results are gathered back, so every rank ends with identical results.  Indices travel as 31-bit values
    packed next to the fp32 similarity (the caller checks the gallery size)."""
    dev = cand_sim.device
    rel = plan.rel
    world = dist.get_world_size(group)
    Q, K = cand_sim.shape[0], cand_sim.shape[2]
    Qc = plan.Qc
    Qp = Qc * world
    tx = torch.empty((Qp, K), dtype=torch.int64, device=dev)
    tx[:Q] = _pack_f32_i32(cand_sim[:, 0], cand_idx[:, 0].clamp(min=-1))
    if Qp > Q:
        tx[Q:] = _pack_f32_i32(torch.full((1,), float("-inf"), device=dev), torch.full((1,), -1, dtype=torch.int64, device=dev))
    rx = torch.empty_like(tx)                                              # [P, Qc, K]: lists of my queries
    dist.all_to_all_single(rx.reshape(-1), tx.reshape(-1), group=group)
    sim_my, idx_my = _unpack_f32_i32(rx.reshape(world, Qc, K).permute(1, 0, 2).contiguous())   # [Qc, P, K]
    top_sim, top_idx, first_hit, ap, hit_ranks = backend.finish_partial(sim_my.contiguous(), idx_my.contiguous(), plan.my_ptr, cnt,
                                                                         rel.total)
    # per-query results -> one int64 block: 10 x (sim, idx) pairs + (ap, first_hit)
    block = torch.cat([_pack_f32_i32(top_sim, top_idx), _pack_f32_i32(ap, first_hit.to(torch.int64)).unsqueeze(1)], dim=1)
    allb = _all_gather_stack(block, group).reshape(Qp, K + 1)[:Q]
    top_sim, top_idx = _unpack_f32_i32(allb[:, :K].contiguous())
    ap, first_hit = _unpack_f32_i32(allb[:, K].contiguous())
    first_hit = first_hit.to(torch.int32)
    if want_hit_ranks:
        dist.all_reduce(hit_ranks, op=dist.ReduceOp.SUM, group=group)      # disjoint slots per rank: exact
        hit_ranks = hit_ranks[:rel.total]
    else:
        hit_ranks = None
    cmc, mAP = backend.metrics(first_hit.contiguous(), ap.contiguous(), topk)
    return RetrievalResult(cmc, mAP, top_idx.contiguous(), top_sim.contiguous(), first_hit, ap, hit_ranks, rel.rel_ptr)


def _all_gather_stack(t: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather of equally-shaped tensors into one [world, *shape] tensor (flat buffers: works on NCCL and gloo)."""
    world = dist.get_world_size(group)
    flat = t.contiguous().reshape(-1)
    out = torch.empty(world * flat.numel(), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.reshape((world,) + tuple(t.shape))


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


def _distributed_plan(q_pids, g_pids_local, get_mAP, precision, group, shard_sizes) -> ShardPlan:
    """This rank's plan; built with two small collectives (per-query counts, slot indices) and one host read, then cached.
    Every rank sees the same hit / miss sequence as long as every rank re-uses (or re-creates) its pid tensors alike."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    key = _plan_key(q_pids, [g_pids_local], get_mAP, precision, ("dist", rank, world, id(group)))
    hit = _cache_get(key, q_pids, [g_pids_local], group, distributed=True)
    if hit is not None:
        return hit[0]
    dev = q_pids.device
    if shard_sizes is None:
        n = torch.tensor([g_pids_local.numel()], dtype=torch.int64, device=dev)
        shard_sizes = _all_gather_stack(n, group).reshape(-1).tolist()      # one host read
    shard_sizes = [int(x) for x in shard_sizes]
    plan = ShardPlan(q_pids, g_pids_local, int(sum(shard_sizes[:rank])), get_mAP, precision)
    plan.shard_sizes = shard_sizes
    if get_mAP:
        plan.set_layout(_all_gather_stack(plan.local_counts(), group), rank)
        dist.all_reduce(plan.gidx_store, op=dist.ReduceOp.SUM, group=group)    # one contributor per slot: the sum is exact
        # scattered finish: the queries are dealt to the ranks in contiguous blocks of Qc
        Qc = -(-plan.Q // world)
        rel_ptr_pad = torch.full((Qc * world + 1,), plan.total, dtype=torch.int64, device=dev)
        rel_ptr_pad[:plan.Q + 1] = plan.rel.rel_ptr
        plan.Qc = Qc
        plan.my_ptr = rel_ptr_pad[rank * Qc: (rank + 1) * Qc + 1].contiguous()
    else:
        plan.g_pids_all = torch.cat(_all_gather_varlen(g_pids_local, group)).contiguous()   # pids of the top-10 candidates
    plan.finish_common()
    _cache_put(key, (plan, q_pids, g_pids_local))
    return plan


def retrieve_sharded(text_embed, image_shard, text_pid, image_pid_shard, topk=(1, 5, 10), get_mAP=True, precision="fp32",
                     group=None, backend=None, nsplit=None, shard_sizes: Optional[Sequence[int]] = None,
                     return_hit_ranks: bool = False, normalized: bool = False) -> RetrievalResult:
    """Distributed driver: call on every rank with the full query set and this rank's gallery slice
    (slices are contiguous and ordered by rank).  Returns the same RetrievalResult on every rank.
    ``shard_sizes`` (rows per rank) saves the size exchange when the caller knows the split; ``return_hit_ranks``
    additionally gathers the per-slot hit ranks (a diagnostic; R@k / AP / mAP do not need it).

    Collectives per evaluation once the plan of this (dataset, sharding) is cached: all-reduce of the thresholds (4 B per
    relevant pair), all-reduce of the rank counts (4 B per pair), all-to-all of the per-query top-10 candidates (8 B each) and
    all-gather of the per-query results.  No collective touches similarity data."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("retrieve_sharded needs an initialised torch.distributed process group")
    backend = backend or CudaBackend()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    PhaseTimer.mark("start")
    q_pids = _as_pid(text_pid)
    g_pids_local = _as_pid(image_pid_shard)
    plan = _distributed_plan(q_pids, g_pids_local, get_mAP, precision, group, shard_sizes)
    PhaseTimer.mark("plan")
    w = ShardWorker(text_embed, image_shard, plan, backend, normalized)
    PhaseTimer.mark("prepare(pack)")
    if get_mAP:
        thr = w.local_thresholds()
        PhaseTimer.mark("thresholds")
        dist.all_reduce(thr, op=dist.ReduceOp.SUM, group=group)      # one contributor per slot, zeros elsewhere: exact
        w.set_thresholds(thr)
        PhaseTimer.mark("allreduce_thr")
    cand_sim, cand_idx, cnt = w.stream(nsplit)
    PhaseTimer.mark("stream")
    if get_mAP:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    PhaseTimer.mark("allreduce_cnt")
    # candidate lists: merge this rank's L lists to one per query, then hand every rank the lists of ITS queries
    g_total = int(sum(plan.shard_sizes))
    if get_mAP and hasattr(backend, "finish_partial") and world > 1 and g_total < (1 << 31) - 1:
        cand_sim, cand_idx = backend.merge_lists(cand_sim, cand_idx)
        PhaseTimer.mark("local_merge")
        res = _finish_scattered(backend, cand_sim, cand_idx, plan, cnt, topk, group, return_hit_ranks)
        PhaseTimer.mark("scattered_finish+metrics")
        res.thresholds = w.thr[:plan.rel.total]
        return res
    if hasattr(backend, "merge_lists"):
        cand_sim, cand_idx = backend.merge_lists(cand_sim, cand_idx)
        sims, idxs = _all_gather_stack(cand_sim, group), _all_gather_stack(cand_idx, group)
        sim_parts, idx_parts = [sims.permute(1, 0, 2, 3).reshape(cand_sim.shape[0], -1, cand_sim.shape[2])], \
                               [idxs.permute(1, 0, 2, 3).reshape(cand_idx.shape[0], -1, cand_idx.shape[2])]
    else:   # stand-in backends (tests): list counts may differ per rank -> gather along a leading list axis
        sim_parts = [x.permute(1, 0, 2) for x in _all_gather_varlen(cand_sim.permute(1, 0, 2).contiguous(), group)]
        idx_parts = [x.permute(1, 0, 2) for x in _all_gather_varlen(cand_idx.permute(1, 0, 2).contiguous(), group)]
    PhaseTimer.mark("merge+allgather_cand")
    res = _finish(backend, sim_parts, idx_parts, q_pids, plan.g_pids_all, plan.rel, cnt, topk)
    PhaseTimer.mark("finish+metrics")
    res.thresholds = w.thr[:plan.rel.total] if get_mAP else None
    return res
