"""MoCo loss dict on B200: host side of trb_moco_loss / trb_enqueue / trb_ema_update_*.

Reference surface (lib/models/embeddings/moco_head/loss.py:21-39, lib/models/losses.py):
the loss evaluator returns ``{"instance_loss", "infonce_loss", "global_align_loss"}`` of 0-d
fp32 tensors that the trainer sums and back-propagates (lib/engine/trainer.py:82,90).  Here the
three losses AND their gradients come out of one stream-ordered library call in ``forward``;
``backward`` only scales by the upstream gradients (device-side, no host sync).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, Sequence, Tuple

import numpy as np
import torch

from . import _lib

PRECISIONS = {"fp32": 0, "bf16": 1}
_GRAPHS: Dict[Tuple, tuple] = {}       # CUDA graphs of the loss step, keyed by shapes / flags / persistent state (see moco_loss_dict)
_GRAPH_CAP = 8
REFERENCE_SMOOTHING = 0.1              # CrossEntropyLabelSmooth's default epsilon (losses.py:18)


def effective_smoothing(epsilon: float, smoothing=None) -> float:
    """Label-smoothing weight the reference actually applies for a configured EPSILON: ``instance_loss`` only tests
    ``epsilon > 0`` and then builds ``CrossEntropyLabelSmooth(num_classes=...)`` without forwarding the value (losses.py:56-57),
    so every positive EPSILON smooths with the class default 0.1 (losses.py:18).  ``smoothing`` overrides it explicitly."""
    if smoothing is not None:
        return float(smoothing)
    return REFERENCE_SMOOTHING if epsilon > 0 else 0.0
LOSS_KEYS = ("instance_loss", "infonce_loss", "global_align_loss")
_workspaces: Dict[Tuple, torch.Tensor] = {}


def _alloc_workspace(nbytes: int, device) -> torch.Tensor:
    """Uninitialised: the prologue launch clears the grid-barrier words itself, so no fill kernel is spent (or captured into a
    caller's CUDA graph)."""
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def _workspace(shape: _lib.MocoShape, precision: int, device) -> torch.Tensor:
    key = (shape.N, shape.D, shape.K, shape.C, precision, str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = _lib.load().trb_moco_loss_workspace_bytes(C.byref(shape), precision)
        if nbytes < 0:
            _lib.check(int(nbytes), "trb_moco_loss_workspace_bytes")
        ws = _alloc_workspace(int(nbytes), device)
        _workspaces[key] = ws
    return ws


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """Contiguous float32 view of ``t`` for pointer extraction (no copy, no dispatcher call, in the common case)."""
    if t.dtype is torch.float32 and t.is_contiguous():
        return t
    return t.detach().contiguous().float()


class _MoCoLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v_embed, t_embed, v_qraw, t_qraw, projection, v_key, t_key, labels, v_queue, t_queue,
                id_queue, hp, precision, normalize_keys, separate_q, queue_ptr, enqueue, cuda_graph):
        lib = _lib.load()
        _lib.require_cuda(v_embed, t_embed, projection, v_key, t_key, labels, v_queue, t_queue, id_queue)
        dev = v_embed.device
        N, D = v_embed.shape
        K = v_queue.shape[1]
        Cn = projection.shape[1]
        if v_queue.shape[0] != D or projection.shape[0] != D or t_embed.shape != v_embed.shape:
            raise ValueError("inconsistent shapes for the MoCo loss")
        shape = _lib.MocoShape(N, D, K, Cn)
        ve, te = _f32c(v_embed), _f32c(t_embed)
        vq, tq = (_f32c(v_qraw), _f32c(t_qraw)) if separate_q else (ve, te)
        vk, tk = _f32c(v_key), _f32c(t_key)
        proj = _f32c(projection)
        lab = labels if (labels.dtype is torch.int64 and labels.dim() == 1 and labels.is_contiguous()) \
            else labels.detach().reshape(-1).to(torch.int64).contiguous()
        need_grad = any(ctx.needs_input_grad[:5])
        if enqueue:
            for q in (v_queue, t_queue):
                if q.dtype != torch.float32 or not q.is_contiguous():
                    raise ValueError("queues must be contiguous float32 [D, K] buffers")
            if id_queue.dtype != torch.int64 or queue_ptr.dtype != torch.int64 or not id_queue.is_contiguous():
                raise ValueError("id_queue / queue_ptr must be contiguous int64 buffers")
            if K % N != 0:
                raise AssertionError("K %% batch_size != 0 (head.py:101)")
            vqu, tqu, idq = v_queue, t_queue, id_queue.reshape(-1)
        else:
            vqu, tqu = _f32c(v_queue), _f32c(t_queue)
            idq = id_queue.detach().reshape(-1).to(torch.int64).contiguous()

        def allocate():
            out = dict(losses=torch.empty(3, dtype=torch.float32, device=dev), vkn=torch.empty_like(vk), tkn=torch.empty_like(tk))
            if need_grad:
                for name in ("d_inst", "d_nce", "d_ga"):
                    out[name] = torch.empty(2, N, D, dtype=torch.float32, device=dev)
                out["d_proj"] = torch.empty(D, Cn, dtype=torch.float32, device=dev)
            return out

        def launch(out, ws, do_enqueue, src=None):
            a_ve, a_te, a_vq, a_tq, a_vk, a_tk, a_lab = src if src is not None else (ve, te, vq, tq, vk, tk, lab)
            common_in = (_lib.ptr(a_ve), _lib.ptr(a_te), _lib.ptr(a_vq), _lib.ptr(a_tq), _lib.ptr(a_vk), _lib.ptr(a_tk),
                         int(normalize_keys), _lib.ptr(out["vkn"]), _lib.ptr(out["tkn"]), _lib.ptr(a_lab), _lib.ptr(vqu),
                         _lib.ptr(tqu), _lib.ptr(idq))
            common_out = (_lib.ptr(proj), C.byref(shape), C.byref(hp), precision, _lib.ptr(out["losses"]),
                          _lib.ptr(out.get("d_inst")), _lib.ptr(out.get("d_nce")), _lib.ptr(out.get("d_ga")),
                          _lib.ptr(out.get("d_proj")), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
            if do_enqueue:   # head.py:175 -- the queues are written after the logits were taken from their old contents
                _lib.check(lib.trb_moco_step(*common_in, _lib.ptr(queue_ptr), *common_out), "trb_moco_step")
                _lib.add_launches(lib.trb_moco_step_launches(C.byref(shape), precision))
            else:
                _lib.check(lib.trb_moco_loss(*common_in, *common_out), "trb_moco_loss")
                _lib.add_launches(lib.trb_moco_loss_launches(C.byref(shape), precision))

        if cuda_graph:
            # The whole step is a fixed launch sequence: replay it as one CUDA graph.  The per-step inputs (embeddings, keys,
            # labels) are fresh allocations in a real training loop, so the graph reads them from STATIC buffers it owns and
            # one multi-tensor copy refreshes those before each replay; the cache key holds shapes, flags and the addresses
            # of the persistent state only (queues, pointer, projection: module buffers / parameters).  The outputs of a
            # replay live in graph-owned buffers that the next replay overwrites.
            key = (vqu.data_ptr(), tqu.data_ptr(), idq.data_ptr(), proj.data_ptr(), queue_ptr.data_ptr() if enqueue else 0,
                   N, D, K, Cn, precision, bool(normalize_keys), bool(separate_q), need_grad, bool(enqueue), str(dev),
                   hp.T, hp.epsilon, hp.alpha, hp.beta, hp.scale_pos, hp.scale_neg)
            entry = _GRAPHS.get(key)
            live = (ve, te, vq, tq, vk, tk) if separate_q else (ve, te, vk, tk)
            if entry is None:
                if len(_GRAPHS) >= _GRAPH_CAP:
                    _GRAPHS.pop(next(iter(_GRAPHS)))
                out = allocate()
                nbytes = lib.trb_moco_loss_workspace_bytes(C.byref(shape), precision)
                if nbytes < 0:
                    _lib.check(int(nbytes), "trb_moco_loss_workspace_bytes")
                ws = _alloc_workspace(int(nbytes), dev)
                static = [torch.empty_like(t) for t in live]
                s_lab = torch.empty_like(lab)
                torch._foreach_copy_(static, list(live))
                s_lab.copy_(lab)
                if separate_q:
                    src = (static[0], static[1], static[2], static[3], static[4], static[5], s_lab)
                else:
                    src = (static[0], static[1], static[0], static[1], static[2], static[3], s_lab)
                launch(out, ws, False, src)              # eager warm-up (validates arguments, sets kernel attributes)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    launch(out, ws, enqueue, src)
                entry = (graph, out, ws, static, s_lab)
                _GRAPHS[key] = entry
            else:
                torch._foreach_copy_(entry[3], list(live))
                entry[4].copy_(lab)
            entry[0].replay()
            out = entry[1]
        else:
            out = allocate()
            launch(out, _workspace(shape, precision, dev), enqueue)
        ctx.separate_q = separate_q
        ctx.grads = (out.get("d_inst"), out.get("d_nce"), out.get("d_ga"), out.get("d_proj"))
        ctx.handover = not cuda_graph          # buffers belong to this call alone: backward may give d_proj away without a copy
        ctx.set_materialize_grads(False)       # no zero-filled gradients for the (non-differentiable) key outputs
        ctx.mark_non_differentiable(out["vkn"], out["tkn"])
        li, ln, lg = out["losses"].unbind(0)
        return li, ln, lg, out["vkn"], out["tkn"]

    @staticmethod
    def backward(ctx, g_inst, g_nce, g_ga, _gvk, _gtk):
        """One launch (trb_moco_grad_combine): saved per-loss gradients x the three upstream scalars, read on the device."""
        lib = _lib.load()
        if ctx.grads is None:
            raise RuntimeError("moco_loss_dict: the gradient buffers were handed over by the first backward pass; "
                               "call moco_loss_dict again (or use cuda_graph=True) for a second one")
        d_inst, d_nce, d_ga, d_proj = ctx.grads
        dev = d_inst.device
        gs = [None if g is None else (g if g.dtype == torch.float32 and g.is_cuda else g.to(device=dev, dtype=torch.float32))
              for g in (g_inst, g_nce, g_ga)]
        gv, gt = torch.empty_like(d_inst[0]), torch.empty_like(d_inst[0])
        gvq = gtq = None
        if ctx.separate_q:
            gvq, gtq = torch.empty_like(gv), torch.empty_like(gv)
        gp = None
        if ctx.needs_input_grad[4]:
            # eager / captured-by-the-caller steps own their buffers: d_proj itself becomes the gradient (scaled in place on the
            # device, untouched when the upstream gradient is 1); a library-level graph replay overwrites its buffers, so copy
            gp = d_proj if ctx.handover else torch.empty_like(d_proj)
        _lib.check(lib.trb_moco_grad_combine(_lib.ptr(d_inst), _lib.ptr(d_nce), _lib.ptr(d_ga), _lib.ptr(d_proj), _lib.ptr(gs[0]),
                                             _lib.ptr(gs[1]), _lib.ptr(gs[2]), int(ctx.separate_q), gv.numel(),
                                             d_proj.numel(), _lib.ptr(gv), _lib.ptr(gt), _lib.ptr(gvq), _lib.ptr(gtq),
                                             _lib.ptr(gp), _lib.stream_ptr(dev)), "trb_moco_grad_combine")
        if ctx.handover:
            ctx.grads = None
            del d_proj
        return (gv, gt, gvq, gtq, gp) + (None,) * 13


def moco_loss_dict(v_embed, t_embed, v_key, t_key, labels, v_queue, t_queue, id_queue, queue_ptr, projection, *,
                   T: float = 0.07, epsilon: float = 0.0, alpha: float = 0.6, beta: float = 0.4, scale_pos: float = 10,
                   scale_neg: float = 40, enqueue: bool = True, v_embed_q=None, t_embed_q=None,
                   normalize_keys: bool = False, precision: str = "fp32", cuda_graph: bool = False,
                   smoothing=None, gather_group=None) -> Dict[str, torch.Tensor]:
    """Functional core of MoCoHead.forward's train branch (head.py:126-175) + LossComputation.forward
    (moco_head/loss.py:21-39).

    v_embed, t_embed [N, D]  post-Linear embeddings (un-normalised)
    v_key, t_key     [N, D]  key embeddings; L2-normalised already unless ``normalize_keys``
    v_embed_q/t_embed_q      InfoNCE query inputs of the FC=True variant (head.py:118-124); default = embeds
    queues [D, K] fp32, id_queue [1, K] int64, queue_ptr [1] int64 are mutated in place when ``enqueue``.
    ``epsilon`` follows the reference: it is cfg.MODEL.EMBEDDING.EPSILON, and any positive value switches label smoothing on
    with the weight 0.1 (see ``effective_smoothing``); ``smoothing`` sets the weight explicitly instead.
    ``gather_group`` (a torch.distributed process group, or True for the default group): cross-rank queue -- the normalised
    keys and ids of EVERY rank are all-gathered (NCCL, rank order) and enqueued on every rank, so data-parallel training keeps
    one global queue like single-GPU training with the global batch.  The reference keeps per-rank
    queues (DDP with broadcast_buffers=False, train_net.py:50-56); this is opt-in and changes behaviour, not speed.  The
    logits of the step are taken from the queue as it was before, exactly like the local enqueue.  Needs K % (N * world) == 0.
    ``cuda_graph=True`` replays the whole step (loss, gradients, enqueue) as one CUDA graph: the per-step inputs are copied
    into graph-owned static buffers (one multi-tensor copy), so freshly allocated inputs hit the same graph every step; the
    returned losses / saved gradients live in graph-owned buffers that the next replay overwrites (fine for the reference's
    loop: ``backward()`` runs before the next ``forward()``, trainer.py:81-91).  Worth it for the unfused launch sequences
    (fp32: 23 launches, bf16 outside the fused gate: 43); the fused bf16 step is two launches and gains nothing.
    """
    if precision not in PRECISIONS:
        raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
    separate_q = v_embed_q is not None or t_embed_q is not None
    if separate_q and (v_embed_q is None or t_embed_q is None):
        raise ValueError("v_embed_q and t_embed_q must be given together")
    hp = _lib.MocoHParams(T, effective_smoothing(epsilon, smoothing), alpha, beta, scale_pos, scale_neg)
    cross_rank = bool(enqueue) and gather_group is not None and gather_group is not False
    li, ln, lg, vkn, tkn = _MoCoLossFunction.apply(
        v_embed, t_embed, v_embed_q if separate_q else v_embed, t_embed_q if separate_q else t_embed, projection,
        v_key, t_key, labels, v_queue, t_queue, id_queue, hp, PRECISIONS[precision], normalize_keys, separate_q,
        queue_ptr, bool(enqueue) and not cross_rank, bool(cuda_graph))
    if cross_rank:
        enqueue_all_ranks(v_queue, t_queue, id_queue, queue_ptr, vkn, tkn, labels, None if gather_group is True else gather_group)
    return {"instance_loss": li, "infonce_loss": ln, "global_align_loss": lg}


@torch.no_grad()
def enqueue_all_ranks(v_queue, t_queue, id_queue, queue_ptr, v_keys, t_keys, ids, group=None) -> None:
    """_dequeue_and_enqueue (head.py:96-109) with the keys of every rank: one all-gather of [v_keys | t_keys] (fp32) and one of
    the ids over NCCL, then the ordinary enqueue of world * N columns in rank order -- every rank ends with the same queues."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("cross-rank queue needs an initialised torch.distributed process group")
    world = dist.get_world_size(group)
    N, D = v_keys.shape
    both = torch.cat([_f32c(v_keys), _f32c(t_keys)], dim=1).contiguous()                 # [N, 2D]
    gathered = torch.empty(world * N, 2 * D, dtype=torch.float32, device=both.device)
    dist.all_gather_into_tensor(gathered, both, group=group)
    ids = ids.detach().reshape(-1).to(torch.int64).contiguous()
    all_ids = torch.empty(world * N, dtype=torch.int64, device=ids.device)
    dist.all_gather_into_tensor(all_ids, ids, group=group)
    dequeue_and_enqueue(v_queue, t_queue, id_queue, queue_ptr, gathered[:, :D].contiguous(), gathered[:, D:].contiguous(), all_ids)


def fused_debug_logits(shape_ndkc, precision: str = "bf16", device=None):
    """Debug: the [256, 128] logits tile dumped by the last fused launch made with TRB_FUSED_DEBUG_LOGITS=<instance tile> (and no
    ``cuda_graph``) for this shape on the current stream.  Rows = modality * 128 + batch row; columns = the tile's classes."""
    N, D, K, Cn = [int(x) for x in shape_ndkc]
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    shape = _lib.MocoShape(N, D, K, Cn)
    ws = _workspace(shape, PRECISIONS[precision], dev)
    out = np.empty((256, 128), dtype=np.float32)
    torch.cuda.synchronize(dev)
    _lib.check(_lib.load().trb_moco_loss_debug_logits(_lib.ptr(ws), C.byref(shape), out.ctypes.data_as(C.c_void_p)),
               "trb_moco_loss_debug_logits")
    return torch.from_numpy(out)


@torch.no_grad()
def dequeue_and_enqueue(v_queue, t_queue, id_queue, queue_ptr, v_keys, t_keys, ids) -> None:
    """head.py:96-109 with the pointer kept on the device (no int(queue_ptr) sync)."""
    _lib.require_cuda(v_queue, t_queue, id_queue, queue_ptr, v_keys, t_keys, ids)
    for q in (v_queue, t_queue):
        if q.dtype != torch.float32 or not q.is_contiguous():
            raise ValueError("queues must be contiguous float32 [D, K] buffers")
    if id_queue.dtype != torch.int64 or queue_ptr.dtype != torch.int64 or not id_queue.is_contiguous():
        raise ValueError("id_queue / queue_ptr must be contiguous int64 buffers")
    D, K = v_queue.shape
    N = v_keys.shape[0]
    if K % N != 0:
        raise AssertionError("K %% batch_size != 0 (head.py:101)")
    vk, tk = _f32c(v_keys), _f32c(t_keys)
    ids = ids.detach().reshape(-1).to(torch.int64).contiguous()
    _lib.check(_lib.load().trb_enqueue(_lib.ptr(v_queue), _lib.ptr(t_queue), _lib.ptr(id_queue), _lib.ptr(queue_ptr),
                                       _lib.ptr(vk), _lib.ptr(tk), _lib.ptr(ids), N, D, K,
                                       _lib.stream_ptr(v_queue.device)), "trb_enqueue")


class MomentumUpdater:
    """_momentum_update_key_encoder (head.py:73-94) as ONE launch over a device-resident chunk table
    instead of ~3 elementwise launches per parameter tensor."""

    CHUNK = 1 << 15

    def __init__(self, m: float):
        self.m = float(m)
        self.one_minus_m = 1.0 - float(m)     # formed in double like the reference, rounded once to fp32
        self._key = None
        self._table = None
        self._nchunks = 0

    def _build(self, params_k: Sequence[torch.Tensor], params_q: Sequence[torch.Tensor]):
        rows = []
        for pk, pq in zip(params_k, params_q):
            if pk.numel() == 0:
                continue
            if pk.dtype != torch.float32 or pq.dtype != torch.float32:
                raise ValueError("momentum update expects float32 parameters")
            if not (pk.is_contiguous() and pq.is_contiguous()) or pk.numel() != pq.numel():
                raise ValueError("momentum update expects contiguous parameter pairs of equal size")
            _lib.require_cuda(pk, pq)
            n = pk.numel()
            for off in range(0, n, self.CHUNK):
                rows.append((pk.data_ptr() + 4 * off, pq.data_ptr() + 4 * off, min(self.CHUNK, n - off)))
        arr = np.asarray(rows, dtype=np.uint64).reshape(-1, 3)
        dev = params_k[0].device
        self._table = torch.from_numpy(arr.view(np.int64)).to(dev)
        self._nchunks = len(rows)

    @torch.no_grad()
    def __call__(self, params_k: Iterable[torch.Tensor], params_q: Iterable[torch.Tensor]) -> None:
        pk = [p.data for p in params_k]
        pq = [p.data for p in params_q]
        if not pk:
            return
        key = tuple((a.data_ptr(), b.data_ptr(), a.numel()) for a, b in zip(pk, pq))
        if key != self._key:
            self._build(pk, pq)
            self._key = key
        _lib.check(_lib.load().trb_ema_update_chunks_f32(_lib.ptr(self._table), self._nchunks, self.CHUNK, self.m,
                                                         self.one_minus_m, _lib.stream_ptr(pk[0].device)),
                   "trb_ema_update_chunks_f32")


@torch.no_grad()
def ema_update_flat(p_k: torch.Tensor, p_q: torch.Tensor, m: float) -> None:
    """Momentum update of one contiguous parameter arena."""
    _lib.require_cuda(p_k, p_q)
    _lib.check(_lib.load().trb_ema_update_f32(_lib.ptr(p_k), _lib.ptr(p_q), p_k.numel(), float(m), 1.0 - float(m),
                                              _lib.stream_ptr(p_k.device)), "trb_ema_update_f32")
