"""Kernel-level timing helpers for bench.py (CUDA events on torch's current stream, which is the stream the
library launches on)."""
from __future__ import annotations

import torch

from .sharded import CudaBackend, ShardWorker


def time_stream_kernel(text, image, q_pid, g_pid, precision, iters=5, flush=None):
    """Median duration (ms) of the gallery stream kernel alone (single shard), thresholds already captured."""
    q_pids = q_pid.reshape(-1).to(torch.int64).contiguous()
    g_pids = g_pid.reshape(-1).to(torch.int64).contiguous()
    w = ShardWorker(text, image, q_pids, g_pids, 0, True, precision, CudaBackend())
    w.set_layout(w.local_counts().unsqueeze(0), 0)
    w.set_thresholds(*w.local_thresholds())
    w.stream()                          # warm-up
    w.record_events = True
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1)
        w.stream()
        torch.cuda.synchronize()
        a, b = w.stream_events
        times.append(a.elapsed_time(b))
    times.sort()
    return times[len(times) // 2]
