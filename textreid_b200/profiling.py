"""Kernel-level timing helpers for bench.py (CUDA events on torch's current stream, which is the stream the
library launches on)."""
from __future__ import annotations

import torch

from .sharded import CudaBackend, ShardWorker


def time_stream_kernel(text, image, q_pid, g_pid, precision, iters=5, flush=None, g_base=0, g_pids_all=None):
    """Median duration (ms) of the gallery stream kernel alone, thresholds already captured."""
    q_pids = q_pid.reshape(-1).to(torch.int64).contiguous()
    g_all = (g_pid if g_pids_all is None else g_pids_all).reshape(-1).to(torch.int64).contiguous()
    w = ShardWorker(text, image, q_pids, g_all, g_base, True, precision, CudaBackend())
    thr = w.local_thresholds()
    w.stream(thr)                       # warm-up
    w.record_events = True
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1)
        w.stream(thr)
        torch.cuda.synchronize()
        a, b = w.stream_events
        times.append(a.elapsed_time(b))
    times.sort()
    return times[len(times) // 2]
