"""Host side of the bf16 tensor-core retrieval path (trb_pack_rows_bf16 + trb_retrieval_stream_tc).

Rows of both operands are gathered in pid order while they are packed, so that the relevant gallery
items of a 128-query tile form one contiguous band of the packed gallery: the thresholds
(similarities of relevant pairs) are then captured by a short banded run of the SAME tcgen05
instruction sequence that the full stream uses, which makes them bit-identical to the streamed values.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .evaluation import RetrievalResult


def pack_rows(x: torch.Tensor, perm: Optional[torch.Tensor] = None, normalize: bool = True, eps: float = 1e-12):
    """[rows, D] fp32/bf16 -> packed pre-swizzled bf16 image (uint8 tensor) in tile-major order."""
    _lib.require_cuda(x)
    lib = _lib.load()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    x = x.contiguous()
    rows, dim = x.shape
    nbytes = lib.trb_packed_bytes(rows, dim)
    if nbytes == 0 and rows > 0:
        raise RuntimeError("tensor-core path needs an embedding size that is a multiple of 64 (got %d)" % dim)
    packed = torch.empty(int(nbytes), dtype=torch.uint8, device=x.device)
    if perm is not None:
        perm = perm.to(torch.int64).contiguous()
    _lib.check(lib.trb_pack_rows_bf16(_lib.ptr(x), int(x.dtype == torch.bfloat16), _lib.ptr(perm), int(normalize), eps,
                                      _lib.ptr(packed), rows, dim, _lib.stream_ptr(x.device)), "trb_pack_rows_bf16")
    return packed


def unpack_rows(packed: torch.Tensor, rows: torch.Tensor, dim: int) -> torch.Tensor:
    """Read rows back out of a packed image (the TRB-P layout of csrc/tc_common.cuh): [len(rows), dim] bf16, exactly the
    operand values the tensor cores consume.  Index arithmetic in torch; for tests, parity checks and debugging."""
    kchunks = dim // 64
    r = rows.to(torch.int64).reshape(-1, 1)                       # packed row numbers
    c16 = torch.arange(dim // 8, device=packed.device).reshape(1, -1)
    rb, rr, kc, c = r >> 7, r & 127, c16 >> 3, c16 & 7
    byte = (rb * kchunks + kc) * 16384 + (rr >> 3) * 1024 + (rr & 7) * 128 + ((c ^ (rr & 7)) << 4)
    elem = (byte >> 1).unsqueeze(-1) + torch.arange(8, device=packed.device)
    return packed.view(torch.bfloat16)[elem.reshape(r.shape[0], -1)]


def choose_nsplit_tc(num_qtiles: int, num_gtiles: int, sms: int) -> int:
    """Gallery split factor for the query tiles of the LAST (partial) wave of the persistent grid.

    The kernel gives whole waves (multiples of the SM count) of query tiles one unit each -- a unit then streams the
    entire gallery, which keeps the per-unit start-up (query tile load, top-10 warm-up) negligible -- and cuts only the
    remaining ``num_qtiles % sms`` tiles into ``nsplit`` gallery pieces so that the last wave is short instead of
    leaving most SMs idle.  With fewer query tiles than SMs every tile is split."""
    rem = num_qtiles % sms
    if rem == 0:
        return 1
    return int(max(1, min(sms // rem, max(1, num_gtiles // 4), 32)))


def retrieve_tc(text_embed, image_embed, q_pids, g_pids, topk=(1, 5, 10), get_mAP=True, normalized=False,
                nsplit: Optional[int] = None) -> RetrievalResult:
    """Single-GPU tensor-core evaluation = the sharded protocol with one shard and no collectives."""
    from .sharded import retrieve_sharded_local
    return retrieve_sharded_local(text_embed, [image_embed], q_pids, [g_pids], topk, get_mAP, "bf16", None, nsplit, normalized)
