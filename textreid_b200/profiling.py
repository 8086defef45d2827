"""Kernel-level timing helpers for bench.py (CUDA events on torch's current stream, which is the stream the
library launches on)."""
from __future__ import annotations

import torch

from .sharded import CudaBackend, ShardWorker, _as_pid, _local_plans


def time_stream_kernel(text, image, q_pid, g_pid, precision, iters=5, flush=None):
    """Median duration (ms) of the gallery stream kernel alone (single shard), thresholds already captured."""
    plan = _local_plans(_as_pid(q_pid), [_as_pid(g_pid)], [0], True, precision)[0]
    w = ShardWorker(text, image, plan, CudaBackend())
    w.set_thresholds(w.local_thresholds())
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
