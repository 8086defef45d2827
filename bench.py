#!/usr/bin/env python
"""Benchmark of the TextReID hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu] [--workload ...]

Headline (BASELINE.json metric): retrieval queries/s -- similarity + top-10 + R@k + mAP -- on the scaled
gallery configuration (configs[3]: 100k text queries x 1M gallery images, D=256, bf16 storage), the
configuration the 1/2/4/8-GPU numbers are quoted on.  The gallery is sharded over the N ranks (strong
scaling: total work fixed).  The same JSON line carries the second half of the metric -- MoCo loss
steps/s at batch 128 / queue 2048 (configs[1]) -- and the CUHK-PEDES-sized evaluation (configs[0]) under
"secondary", each with its own roofline fraction.

One "step" = one full evaluation of the workload: pid bookkeeping, normalise+pack, threshold capture,
gallery stream (the GEMM), merge, metrics.  `value` has inputs resident in HBM; `e2e` goes through the
public Python API from pinned HOST buffers with the H2D copies and the D2H read of R@k/mAP timed.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/, kind "port": the
reference is pure Python/PyTorch and cannot travel to the GPU box) on a bounded sample of the workload: a few hundred
queries against the FULL gallery, so the G log G sort and the memory footprint are measured, not extrapolated; only the
(embarrassingly parallel) query count is scaled.  `--impl reference-gpu` runs the same restated ATen call sequence
(normalise, matmul, argsort, per-column AP loop) on the B200 under torch 2.11 for context.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Q, G, D, ids, images per id)
    "retrieval_1m": dict(Q=100_000, G=1_000_000, D=256, n_ids=250_000, desc="configs[3]: 100k text queries x 1M gallery, D=256, bf16 storage"),
    "retrieval_cuhk": dict(Q=6156, G=3074, D=256, n_ids=1000, desc="configs[0]: CUHK-PEDES-shaped 6156 x 3074, D=256"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:       # one sampler per job (rank 0): N concurrent nvidia-smi loops perturb the timed region
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median of the upper half = clocks under load (idle samples before/after drag a plain median down)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_eval_data(*a, **k):
    from textreid_b200.synthetic import eval_data
    return eval_data(*a, **k)


def shard_bounds(G, world, rank):
    per = -(-G // world)
    per = -(-per // 256) * 256
    lo = min(G, rank * per)
    return lo, min(G, lo + per)


# ---------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------------
def reference_retrieval_sample(cfg, steps, warmup, device="cpu", Qs=None, Gs=None, budget_s=25.0):
    """The reference's algorithm (oracle.retrieve: normalise + fp32 matmul + full stable argsort + the per-column AP loop of
    evaluation.py:33) on Qs queries x the workload's WHOLE gallery.  Queries are independent, so queries/s = Qs / time is a
    measurement at the stated gallery size, not an extrapolation over G.  The first run sizes the sample: if one step takes
    longer than `budget_s` the number of timed steps is cut so that the whole arm stays within a few minutes."""
    from oracle import textreid_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    G = cfg["G"] if Gs is None else min(Gs, cfg["G"])
    if Qs is None:
        Qs = 128 if G > 200_000 else min(cfg["Q"], 2048)
    Qs = min(Qs, cfg["Q"])
    text, q_pid, image, g_pid = make_eval_data(Qs, G, cfg["D"], cfg["n_ids"] if G == cfg["G"] else max(G // 4, 1), 0, G, device,
                                               torch.float32, seed=1)
    sync = (lambda: torch.cuda.synchronize()) if str(device).startswith("cuda") else (lambda: None)
    times = []
    planned = warmup + steps
    i = 0
    while i < planned:
        sync()
        t0 = time.perf_counter()
        O.retrieve(text, image, q_pid, g_pid, (1, 5, 10), get_mAP=True, per_column_loop=True)
        sync()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        elif dt > budget_s:              # a slow box: the warm-up run becomes the measurement
            times.append(dt)
            break
        if len(times) >= 1 and sum(times) > 4 * budget_s:
            break
        i += 1
    mean = sum(times) / len(times)
    where = "CPU, %d threads" % cores if device == "cpu" else "B200, torch %s ATen kernels" % torch.__version__
    return dict(value=Qs / mean, unit="queries/s", cores=cores, kind="port",
                sample="oracle.retrieve (normalise + fp32 matmul + full stable argsort + reference per-column AP loop) on %d queries x "
                       "the full %d-row gallery (%s), %d timed run(s), %.2f s each; queries are independent, so queries/s = %d / time "
                       "at the stated gallery size" % (Qs, G, where, len(times), mean, Qs),
                sample_seconds=mean, ms_per_step=mean * 1e3, steps_timed=len(times))


def cpu_retrieval_sample(cfg, steps, warmup):
    return reference_retrieval_sample(cfg, steps, warmup, "cpu")


def cpu_loss_sample(steps=5, warmup=2, N=128, D=256, K=2048, C=11003):
    from oracle import textreid_oracle as O
    from textreid_b200.synthetic import loss_inputs as synth_loss_inputs
    torch.set_num_threads(os.cpu_count() or 1)
    inp = synth_loss_inputs(N, D, K, C, seed=0)
    args = [inp[k] for k in ("v_embed", "t_embed", "v_key", "t_key", "labels", "v_queue", "t_queue", "id_queue", "projection")]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.moco_loss_dict_with_grads(*args, epsilon=0.1)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return dict(value=1.0 / mean, unit="steps/s", cores=os.cpu_count() or 1, kind="port",
                sample="oracle loss dict fwd+bwd (autograd) N=%d K=%d D=%d C=%d, %d runs" % (N, K, D, C, len(times)))


# ---------------------------------------------------------------------------------------------------
# secondary measurements (rank 0, N=1): MoCo loss steps/s, EMA GB/s, CUHK-sized eval
# ---------------------------------------------------------------------------------------------------
def time_cuda(fn, iters, warmup, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


LOSS_SHAPES = {
    # algorithmic bytes (each operand once) and FLOPs of the loss dict fwd+bwd
    "n128_k2048": dict(N=128, D=256, K=2048, C=11003, bytes=27.8e6, flops=4.89e9, desc="configs[1]: bs128, queue 2048"),
    "n256_k4096": dict(N=256, D=256, K=4096, C=11003, bytes=33.0e6, flops=10.9e9, desc="configs[2]: bs256, queue 4096"),
}


def loss_step_line(trb, pk, device, flush, shape_key, prec, mode, iters=40):
    """One trainer-style loss step (trainer.py:81-90 on the head's output): loss dict -> sum -> backward (+ enqueue), timed with
    CUDA events, L2 flushed between steps.  The labels of the NEXT step are drawn during the (untimed) flush, like a data
    loader would: new ids every step keep the queue from degenerating into "every slot masked"."""
    from textreid_b200.synthetic import loss_inputs as synth_loss_inputs
    sh = LOSS_SHAPES[shape_key]
    N, D, K, C = sh["N"], sh["D"], sh["K"], sh["C"]
    inp = {k: v.to(device) for k, v in synth_loss_inputs(N, D, K, C, seed=0).items()}
    ve, te, pr = inp["v_embed"].requires_grad_(True), inp["t_embed"].requires_grad_(True), inp["projection"].requires_grad_(True)
    ptr = torch.zeros(1, dtype=torch.int64, device=device)
    labels = inp["labels"]

    def loss_step(inner_graph):
        d = trb.moco_loss_dict(ve, te, inp["v_key"], inp["t_key"], labels, inp["v_queue"], inp["t_queue"], inp["id_queue"],
                               ptr, pr, epsilon=0.1, enqueue=True, precision=prec, cuda_graph=inner_graph)
        (d["instance_loss"] + d["infonce_loss"] + d["global_align_loss"]).backward()

    def flush_and_relabel():
        labels.add_(97).remainder_(C)
        flush()

    if mode == "stepgraph":
        # the whole step captured once with torch.cuda.graph and replayed: the way a production loop removes the host from
        # the path; gradients land in static .grad tensors
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                ve.grad = te.grad = pr.grad = None
                loss_step(False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ve.grad = te.grad = pr.grad = None
        whole = torch.cuda.CUDAGraph()
        with torch.cuda.graph(whole):
            loss_step(False)
        fn = whole.replay
    else:
        def fn(inner=(mode == "graph")):
            ve.grad = te.grad = pr.grad = None
            loss_step(inner)
    med, best = time_cuda(fn, iters, 8, flush_and_relabel)
    t_s = med * 1e-3
    hbm_frac = sh["bytes"] / t_s / 1e9 / pk["hbm"]
    tensor_frac = sh["flops"] / t_s / 1e12 / pk["tf_burst"]
    roof_s = max(sh["bytes"] / (pk["hbm"] * 1e9), sh["flops"] / (pk["tf_burst"] * 1e12))
    shape = trb._lib.MocoShape(N, D, K, C)
    launches = int(trb._lib.load().trb_moco_step_launches(__import__("ctypes").byref(shape), 1 if prec == "bf16" else 0))
    return {
        "metric": "MoCo loss steps/s (loss dict fwd+bwd + enqueue, %s, D=%d, C=%d)" % (sh["desc"], D, C), "value": 1e3 / med,
        "unit": "steps/s", "ms_per_step": med, "ms_best": best, "dtype": "bf16" if prec == "bf16" else "f32",
        "mode": {"stepgraph": "whole step captured in one CUDA graph (torch.cuda.graph around loss dict + backward)",
                 "graph": "library call replayed from its own CUDA graph, autograd glue eager",
                 "eager": "every launch issued from Python"}[mode],
        "l2_flushed": True, "library_launches": launches,
        "roofline": {"bound": "hbm" if sh["bytes"] / (pk["hbm"] * 1e9) >= sh["flops"] / (pk["tf_burst"] * 1e12) else "tensor",
                     "achieved": sh["bytes"] / t_s / 1e9, "peak": pk["hbm"], "unit": "GB/s", "frac": hbm_frac,
                     "tensor_frac": tensor_frac, "roofline_time_us": roof_s * 1e6, "frac_of_roofline_time": roof_s / t_s,
                     "traffic": None,
                     "note": "whole step (loss dict + gradients + enqueue in the library call, loss sum, backward combine); algorithmic "
                             "%.1f MB / %.2f GFLOP (BASELINE.md section 3); tensor peak = measured burst (a kernel timed alone)"
                             % (sh["bytes"] / 1e6, sh["flops"] / 1e9)}}


def loss_e2e_line(trb, device, shape_key="n128_k2048", prec="bf16", iters=30):
    """The loss step end to end from HOST buffers: pinned embeddings / keys / labels -> device, loss dict + backward + enqueue,
    losses and the three gradients back to pinned host memory; every copy inside the timed region."""
    from textreid_b200.synthetic import loss_inputs as synth_loss_inputs
    sh = LOSS_SHAPES[shape_key]
    N, D, K, C = sh["N"], sh["D"], sh["K"], sh["C"]
    inp = synth_loss_inputs(N, D, K, C, seed=0)
    host = {k: inp[k].pin_memory() for k in ("v_embed", "t_embed", "v_key", "t_key", "labels")}
    dev = {k: inp[k].to(device) for k in ("v_queue", "t_queue", "id_queue")}
    pr = inp["projection"].to(device).requires_grad_(True)
    ptr = torch.zeros(1, dtype=torch.int64, device=device)
    out_host = {"losses": torch.empty(3).pin_memory(), "gv": torch.empty(N, D).pin_memory(), "gt": torch.empty(N, D).pin_memory(),
                "gp": torch.empty(D, C).pin_memory()}

    def step():
        ve = host["v_embed"].to(device, non_blocking=True).requires_grad_(True)
        te = host["t_embed"].to(device, non_blocking=True).requires_grad_(True)
        vk, tk = host["v_key"].to(device, non_blocking=True), host["t_key"].to(device, non_blocking=True)
        lab = host["labels"].to(device, non_blocking=True)
        pr.grad = None
        d = trb.moco_loss_dict(ve, te, vk, tk, lab, dev["v_queue"], dev["t_queue"], dev["id_queue"], ptr, pr, epsilon=0.1,
                               enqueue=True, precision=prec)
        (d["instance_loss"] + d["infonce_loss"] + d["global_align_loss"]).backward()
        out_host["losses"].copy_(torch.stack([d[k].detach() for k in ("instance_loss", "infonce_loss", "global_align_loss")]), non_blocking=True)
        out_host["gv"].copy_(ve.grad, non_blocking=True)
        out_host["gt"].copy_(te.grad, non_blocking=True)
        out_host["gp"].copy_(pr.grad, non_blocking=True)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())
    return {"metric": "MoCo loss steps/s end to end from pinned host buffers (%s)" % sh["desc"], "value": 1e3 / ms, "unit": "steps/s",
            "ms_per_step": ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "dtype": prec,
            "note": "back-to-back eager steps (2 library launches + autograd glue); the 11.3 MB projection gradient dominates the D2H"}


def reference_gpu_loss_line(device, shape_key="n128_k2048", iters=20):
    """The reference's ATen call sequence for the loss dict (oracle restatement: queue mask with nonzero/unique syncs, gathers,
    CPU one-hot replaced by a device scatter) forward + autograd backward on the B200, for context beside the fused step."""
    from oracle import textreid_oracle as O
    from textreid_b200.synthetic import loss_inputs as synth_loss_inputs
    sh = LOSS_SHAPES[shape_key]
    inp = {k: v.to(device) for k, v in synth_loss_inputs(sh["N"], sh["D"], sh["K"], sh["C"], seed=0).items()}
    args = [inp[k] for k in ("v_embed", "t_embed", "v_key", "t_key", "labels", "v_queue", "t_queue", "id_queue", "projection")]
    for _ in range(3):
        O.moco_loss_dict_with_grads(*args, epsilon=0.1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        O.moco_loss_dict_with_grads(*args, epsilon=0.1)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / iters * 1e3
    return {"metric": "reference ATen loss dict fwd + autograd bwd on the B200 (oracle restatement, torch %s, eager, wall clock)" % torch.__version__,
            "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms}


def train_step_lines(trb, device, N=128, iters=6, warmup=3):
    """BASELINE configs[4]: one whole train step -- CLIP RN50 image encoder + bi-GRU text encoder (PyTorch / cuDNN, out of the
    accelerated scope) + MoCo head, batch 128, 384 x 128 images, captions of up to 100 tokens, queue 2048, 11003 classes,
    Adam -- once with the reference's head as an ATen call sequence (oracle/reference_head.py) and once with FusedMoCoHead
    (bf16 fused loss step, one-launch momentum update).  Same encoders, same data, same optimizer in both arms."""
    from types import SimpleNamespace
    from oracle.reference_head import ReferenceStyleHead
    from textreid_b200.encoders import BiGRUTextEncoder, ClipResNetEncoder, synthetic_vocab_table
    from textreid_b200.synthetic import train_batch
    cfg = SimpleNamespace(MODEL=SimpleNamespace(EMBEDDING=SimpleNamespace(FEATURE_SIZE=256, EPSILON=0.1),
                                                MOCO=SimpleNamespace(K=2048, M=0.999, FC=False), NUM_CLASSES=11003))
    table = synthetic_vocab_table()
    batches = [train_batch(N, 11003, table.shape[0], seed=s, device=device) for s in range(2)]
    out = {}
    for arm in ("reference_head", "fused_head"):
        torch.manual_seed(0)
        vis, txt = ClipResNetEncoder(), BiGRUTextEncoder(table)
        head = (ReferenceStyleHead(cfg, vis, txt) if arm == "reference_head" else trb.FusedMoCoHead(cfg, vis, txt, precision="bf16"))
        head = head.to(device).train()
        opt = torch.optim.Adam([p for p in head.parameters() if p.requires_grad], lr=1e-4)

        def step(i):
            images, caps, _ = batches[i % 2]
            opt.zero_grad(set_to_none=True)
            losses = head(images, caps)
            sum(losses.values()).backward()
            opt.step()
            return losses

        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        syncs0 = getattr(head, "host_syncs", 0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for i in range(iters):
            losses = step(i)
        b.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / iters * 1e3
        line = {"ms_per_step": a.elapsed_time(b) / iters, "wall_ms_per_step": wall, "steps_per_s": 1e3 / (a.elapsed_time(b) / iters),
                "losses": {k: float(v) for k, v in losses.items()}}
        if arm == "reference_head":
            line["host_syncs_per_step_in_head"] = (head.host_syncs - syncs0) // iters
        try:        # kernel launches of ONE step (encoders included), counted by the profiler
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step(0)
                torch.cuda.synchronize()
            line["cuda_kernel_launches_per_step"] = sum(e.count for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA)
        except Exception as e:      # pragma: no cover
            line["cuda_kernel_launches_per_step"] = None
        out[arm] = line
        del head, opt, vis, txt
        torch.cuda.empty_cache()
    r, f = out["reference_head"], out["fused_head"]
    out["metric"] = "end-to-end train step, RN50 + bi-GRU + MoCo head, bs%d, 384x128, <=100 tokens, K=2048, C=11003, Adam (BASELINE configs[4])" % N
    out["speedup"] = r["ms_per_step"] / f["ms_per_step"]
    if r.get("cuda_kernel_launches_per_step") and f.get("cuda_kernel_launches_per_step"):
        out["launches_removed_per_step"] = r["cuda_kernel_launches_per_step"] - f["cuda_kernel_launches_per_step"]
    out["note"] = ("encoders run in fp32 (cuDNN TF32 convolutions, torch defaults) in both arms and dominate the step; the arms differ in "
                   "the head only: per-parameter momentum loop + host-synchronising loss sequence vs one-launch momentum update + "
                   "two-launch fused loss step")
    return out


def secondary_measurements(pk, device):
    import textreid_b200 as trb
    out = {}
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)      # > 126 MB L2

    def flush():
        flush_buf.fill_(1)

    # ---- MoCo loss step, configs[1] (N=128, K=2048) in three host modes and both precisions; configs[2] (N=256, K=4096) ----
    for prec in ("bf16", "fp32"):
        for mode in ("stepgraph", "graph", "eager"):
            out["moco_loss_%s_%s" % (prec, mode)] = loss_step_line(trb, pk, device, flush, "n128_k2048", prec, mode)
    for prec in ("bf16", "fp32"):
        out["moco_loss_%s_stepgraph_n256_k4096" % prec] = loss_step_line(trb, pk, device, flush, "n256_k4096", prec, "stepgraph", iters=20)
    out["moco_loss_kernel_launches"] = {p_: out["moco_loss_%s_eager" % p_]["library_launches"] for p_ in ("bf16", "fp32")}
    out["moco_loss_bf16_e2e"] = loss_e2e_line(trb, device)
    try:
        out["moco_loss_reference_gpu"] = reference_gpu_loss_line(device)
    except Exception as e:      # pragma: no cover
        out["moco_loss_reference_gpu"] = {"error": str(e)[:200]}
    # ---- EMA over an RN50+GRU-sized arena: 41,755,488 fp32 parameters ----
    P = 41_755_488
    pk_, pq_ = torch.randn(P, device=device), torch.randn(P, device=device)
    med, best = time_cuda(lambda: trb.ema_update_flat(pk_, pq_, 0.999), 20, 3, flush)
    out["ema_rn50_gru"] = {"metric": "momentum update of 41,755,488 fp32 parameters", "ms": med,
                           "roofline": {"bound": "hbm", "achieved": 12.0 * P / (med * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                        "frac": 12.0 * P / (med * 1e-3) / 1e9 / pk["hbm"], "traffic": None}}
    del pk_, pq_
    # ---- CUHK-PEDES-sized evaluation, configs[0]: repeated evaluations of one split (plan cached, like trainer.py:124) ----
    cfg = WORKLOADS["retrieval_cuhk"]
    text, q_pid, image, g_pid = make_eval_data(cfg["Q"], cfg["G"], cfg["D"], cfg["n_ids"], 0, cfg["G"], device, torch.float32, seed=2)
    for prec in ("fp32", "bf16"):
        med, best = time_cuda(lambda: trb.retrieve(text, image, q_pid, g_pid, (1, 5, 10), True, prec), 10, 3, flush)
        out["retrieval_cuhk_" + prec] = {"metric": "retrieval queries/s, 6156 x 3074, D=256 (R@k + mAP; pid plan cached across evaluations)",
                                         "value": cfg["Q"] / (med * 1e-3), "unit": "queries/s", "ms_per_step": med, "dtype": prec}
        trb.clear_plan_cache()
        t0 = time.perf_counter()
        trb.retrieve(text, image, q_pid, g_pid, (1, 5, 10), True, prec)
        torch.cuda.synchronize()
        out["retrieval_cuhk_" + prec]["first_evaluation_ms"] = (time.perf_counter() - t0) * 1e3
    # the reference's entry points on the same split: evaluation() fast path (trainer.py:124) and with re-ranking (test_net.py:107)
    n = cfg["Q"]

    class _Split:
        def __init__(self, image_ids, pids): self.image_ids, self.pids = image_ids, pids
        def __len__(self): return len(self.pids)
        def get_id_info(self, idx): return self.image_ids[idx], self.pids[idx]

    gen = torch.Generator().manual_seed(3)
    img_of_caption = torch.randint(0, cfg["G"], (n,), generator=gen)
    img_of_caption[:cfg["G"]] = torch.arange(cfg["G"])                        # every image has at least one caption
    ds = _Split(img_of_caption.tolist(), g_pid.cpu()[img_of_caption].tolist())
    image_all = image[img_of_caption.to(device)].contiguous()
    idx = list(range(n))
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        for name, kw in (("evaluation_entry_cuhk_fast", dict(save_data=False, rerank=False)),
                         ("evaluation_entry_cuhk_rerank", dict(save_data=False, rerank=True))):
            fn = lambda: trb.evaluate_embeddings(ds, idx, image_all, text, tmp, [1, 5, 10], **kw)
            med, best = time_cuda(fn, 5, 2, flush)
            out[name] = {"metric": "evaluation() core on [n, D] device tensors, 6156 captions x 3074 images, %s" % (
                             "t2i + i2t top-k (trainer.py:124)" if "fast" in name else "4 rankings incl. k-reciprocal re-rank (test_net.py:107)"),
                         "ms_per_call": med, "value": 1e3 / med, "unit": "evaluations/s"}
    del text, image
    torch.cuda.empty_cache()
    # ---- configs[4]: the whole train step, reference-style head vs fused head around the same encoders ----
    try:
        out["train_step_config5"] = train_step_lines(trb, device)
    except Exception as e:      # pragma: no cover
        out["train_step_config5"] = {"error": str(e)[:300]}
    torch.cuda.empty_cache()
    # ---- configs[3] on the fp32 FFMA parity path and at D = 512 on the tensor-core path (one step each) ----
    big = WORKLOADS["retrieval_1m"]
    for name, D, prec, Qn in (("retrieval_1m_fp32", 256, "fp32", 20_000), ("retrieval_1m_d512_bf16", 512, "bf16", 50_000)):
        try:
            dt = torch.float32 if prec == "fp32" else torch.bfloat16
            text, q_pid, image, g_pid = make_eval_data(Qn, big["G"], D, big["n_ids"], 0, big["G"], device, dt)
            trb.retrieve(text, image, q_pid, g_pid, (1, 5, 10), True, prec)
            med, best = time_cuda(lambda: trb.retrieve(text, image, q_pid, g_pid, (1, 5, 10), True, prec), 2, 0)
            fl = 2.0 * Qn * big["G"] * D
            out[name] = {"metric": "retrieval queries/s, %d x %d, D=%d, %s" % (Qn, big["G"], D, prec), "value": Qn / (med * 1e-3),
                         "unit": "queries/s", "ms_per_step": med, "tflops": fl / (med * 1e-3) / 1e12,
                         "frac_of_bf16_sustained": (fl / (med * 1e-3) / 1e12 / pk["tf_sustained"]) if prec == "bf16" else None}
            del text, image
            torch.cuda.empty_cache()
        except Exception as e:      # pragma: no cover
            out[name] = {"error": str(e)[:200]}
    return out


def stored_traffic(workload, precision, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the stream kernel, extracted from the committed ncu capture by
    tools/ncu_traffic.py into profiles/stream_traffic.json (never a literal here)."""
    path = os.path.join(ROOT, "profiles", "stream_traffic.json")
    if world != 1 or not os.path.exists(path):
        return None
    try:
        t = json.load(open(path))
        e = t.get("%s/%s" % (workload, precision))
        return None if e is None else {"bytes": e["dram_bytes"], "source": e["source"]}
    except Exception:      # pragma: no cover
        return None


# R@1 / R@5 / R@10 / mAP of the seeded synthetic workloads (single-GPU bf16 run of round 1, BENCH_r01.json): exact integer
# artefacts, so every later run -- any number of GPUs, any kernel revision -- must reproduce them bit for bit.
STORED_RESULTS = {
    "retrieval_1m/bf16": {"R@1": 39.49300003051758, "R@5": 60.90899658203125, "R@10": 69.00799560546875, "mAP": 21.91847038269043},
}


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="retrieval_1m", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "%s (%s)" % (args.workload, cfg["desc"]), "Q": cfg["Q"], "G": cfg["G"], "D": cfg["D"],
              "precision": args.precision, "sharding": "gallery rows over %d rank(s)" % world,
              "cache": "inputs larger than L2 (gallery %.0f MB per rank)" % (cfg["G"] / world * cfg["D"] * 2 / 1e6)
              if cfg["G"] * cfg["D"] * 2 / world > 130e6 else "L2 flushed between timed iterations"}

    if args.impl in ("reference", "reference-gpu"):
        if rank != 0:
            return 0
        on_gpu = args.impl == "reference-gpu"
        if on_gpu and not torch.cuda.is_available():
            print(json.dumps({"impl": "reference-gpu", "unavailable": "no CUDA device"}))
            return 0
        r = reference_retrieval_sample(cfg, max(1, args.steps), max(0, min(args.warmup, 1)), "cuda" if on_gpu else "cpu",
                                       Qs=1024 if on_gpu else None)
        line = {"impl": args.impl, "metric": "retrieval queries/s (sim+top-k+R@k/mAP)", "value": r["value"], "unit": "queries/s",
                "n_gpus": args.gpus, "steps": r["steps_timed"], "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch.distributed as dist
    import textreid_b200 as trb
    from textreid_b200 import _lib
    from textreid_b200.sharded import retrieve_sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    pk = peaks()
    Q, G, D = cfg["Q"], cfg["G"], cfg["D"]
    dtype = torch.bfloat16 if args.precision == "bf16" else torch.float32
    g_lo, g_hi = shard_bounds(G, world, rank)
    text, q_pid, image, g_pid = make_eval_data(Q, G, D, cfg["n_ids"], g_lo, g_hi, device, dtype)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device) if "flushed" in config["cache"] else None
    sizes = [shard_bounds(G, world, r)[1] - shard_bounds(G, world, r)[0] for r in range(world)]

    def step(text_d, image_d, q_pid_d, g_pid_d):
        if world > 1:
            return retrieve_sharded(text_d, image_d, q_pid_d, g_pid_d, (1, 5, 10), True, args.precision, shard_sizes=sizes)
        return trb.retrieve(text_d, image_d, q_pid_d, g_pid_d, (1, 5, 10), True, args.precision)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: repeated evaluations of one (dataset, sharding), like the trainer's periodic evaluation
    #      (trainer.py:124): the pid bookkeeping of the split is planned once, during the warm-up ----
    for _ in range(args.warmup):
        res = step(text, image, q_pid, g_pid)
    barrier()
    _lib.reset_launch_count()
    with ClockSampler(local_rank, enabled=(rank == 0)) as clocks:
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # TRB_BENCH_CUDA_PROFILER=1: the timed steps, and nothing else, sit between cudaProfilerStart / Stop, so that
        # `ncu --profile-from-start off` lists exactly the launches of the timed region (tools/final_run.sh)
        prof_window = bool(os.environ.get("TRB_BENCH_CUDA_PROFILER"))
        if prof_window:
            torch.cuda.cudart().cudaProfilerStart()
        ev0.record()
        for _ in range(args.steps):
            if flush_buf is not None:
                flush_buf.fill_(1)
            res = step(text, image, q_pid, g_pid)
        ev1.record()
        barrier()
        if prof_window:
            torch.cuda.cudart().cudaProfilerStop()
    launches = _lib.launch_count()
    if os.environ.get("TRB_PROFILE_PHASES") and world > 1:     # every rank takes part in the collectives
        from textreid_b200.sharded import PhaseTimer
        PhaseTimer.marks = []
        step(text, image, q_pid, g_pid)
        rep = PhaseTimer.report()
        if rank == 0:
            sys.stderr.write("PHASES " + rep + "\n")
    ms = ev0.elapsed_time(ev1) / args.steps
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- dominant kernel alone, CUDA events on the launching stream (torch's current stream) ----
    kern_ms = None
    try:
        from textreid_b200 import profiling
        kern_ms = profiling.time_stream_kernel(text, image, q_pid, g_pid, args.precision, iters=max(3, args.steps), flush=flush_buf)
    except Exception as e:          # pragma: no cover
        config["kernel_timing_error"] = str(e)[:200]

    # ---- sampled parity check of the result at this very size (N = 1: the whole gallery is on this device) ----
    parity = None
    if world == 1 and args.precision == "bf16":
        try:
            from textreid_b200 import verify
            rep = verify.sampled_check(text, image, q_pid, g_pid, res, n_sample=64, seed=0, margin=1e-6)
            parity = {"status": rep["status"], "queries": rep["n_queries"], "top10_match": "%d/%d" % (rep["top10_match"], rep["top10_decided"]),
                      "hit_ranks_exact": "%d/%d decided, %d outside the margin interval" % (rep["slots_decided_exact"], rep["slots_decided"],
                                                                                          rep["slots_outside_interval"]),
                      "max_similarity_error": rep["max_sim_err"],
                      "how": "float64 re-evaluation of 64 random queries x the full gallery from the kernel's own bf16 operands "
                             "(textreid_b200/verify.py; cross-check against the CPU restatement in tests/test_gpu_retrieval.py)"}
        except Exception as e:      # pragma: no cover
            parity = {"status": "error", "error": str(e)[:200]}

    # ---- end to end through the public API from pinned host buffers ----
    h_text, h_image = text.cpu().pin_memory(), image.cpu().pin_memory()
    h_qpid, h_gpid = q_pid.cpu().pin_memory(), g_pid.cpu().pin_memory()

    # N > 1: every rank needs ALL queries on its device, but nothing says each must pull them over its own PCIe link: rank r
    # uploads rows [r*Qc, (r+1)*Qc) and the ranks all-gather the block over NVLink (the way inference() hands the embeddings
    # over: textreid_b200.evaluation.gather_embeddings).  The gallery shard and the pid vectors are uploaded whole.
    Qc = -(-Q // world)
    h_text_part = h_text[rank * Qc: min(Q, (rank + 1) * Qc)]
    h2d_rank = sum(x.numel() * x.element_size() for x in ((h_text_part if world > 1 else h_text), h_image, h_qpid, h_gpid))
    t = torch.tensor([h2d_rank], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    h2d = int(t.item())                      # whole job, all ranks

    def upload_queries():
        if world == 1:
            return h_text.to(device, non_blocking=True)
        part = torch.zeros(Qc, D, dtype=h_text.dtype, device=device)
        part[:h_text_part.shape[0]].copy_(h_text_part, non_blocking=True)
        full = torch.empty(Qc * world, D, dtype=h_text.dtype, device=device)
        dist.all_gather_into_tensor(full, part)
        return full[:Q]

    def e2e_step():
        # everything is uploaded every step, pid vectors included: new device tensors each time, which the plan cache
        # recognises by content (one device-side comparison) instead of re-planning the split
        r = step(upload_queries(), h_image.to(device, non_blocking=True),
                 h_qpid.to(device, non_blocking=True), h_gpid.to(device, non_blocking=True))
        return torch.cat([r.cmc, r.mAP.reshape(1)]).cpu()

    for _ in range(2):
        out_host = e2e_step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out_host = e2e_step()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1) / args.steps
    t = torch.tensor([e2e_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    if rank == 0:
        flops = 2.0 * Q * (g_hi - g_lo) * D          # per rank, the stream kernel's algorithmic work
        roof = None
        if kern_ms:
            ach = flops / (kern_ms * 1e-3) / 1e12
            tr = stored_traffic(args.workload, args.precision, world)
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "traffic": tr["bytes"] if tr else None, "traffic_source": tr["source"] if tr else None,
                    "kernel": "retrieval_tc_kernel<0>" if args.precision == "bf16" else "stream_f32_kernel",
                    "kernel_ms": kern_ms, "peak_source": pk["source"] + " bf16 sustained (kernel timed back to back inside a long step)",
                    "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": (Q + (g_hi - g_lo)) * D * 2.0}
            if args.precision == "fp32":
                roof["note"] = "fp32 FFMA parity path; the tensor-core denominator does not apply"
        result = {"R@1": float(res.cmc[0]), "R@5": float(res.cmc[1]), "R@10": float(res.cmc[2]), "mAP": float(res.mAP)}
        stored = STORED_RESULTS.get("%s/%s" % (args.workload, args.precision))
        if stored is not None:
            result["matches_stored_single_gpu_result"] = all(result[k] == stored[k] for k in stored)
        if parity is not None:
            result["parity_sample"] = parity["status"]
            result["parity_sample_detail"] = parity
        line = {"metric": "retrieval queries/s (sim+top-k+R@k/mAP)", "value": Q / (ms * 1e-3), "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": config,
                "clocks": clocks.summary(), "gpu_launches": launches,
                "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": int(out_host.numel() * 4) * world, "ms_per_step": e2e_ms,
                        "note": "bytes are whole-job totals over all ranks; for N > 1 each rank uploads 1/N of the query block (all-gathered "
                                "over NVLink), its gallery shard and the pid vectors"},
                "roofline": roof, "result": result}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_retrieval_sample(cfg, 1, 0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if world == 1 and not args.no_secondary:
            del text, image, h_text, h_image
            torch.cuda.empty_cache()
            try:
                line["secondary"] = secondary_measurements(pk, device)
                line["secondary"]["moco_loss_cpu_baseline"] = cpu_loss_sample()
                best = line["secondary"].get("moco_loss_bf16_stepgraph")
                if best:     # the other half of BASELINE's metric, lifted to the top level for readers of the one line
                    base = line["secondary"].get("moco_loss_cpu_baseline") or {}
                    e2e_l = line["secondary"].get("moco_loss_bf16_e2e") or {}
                    line["loss_step"] = {"metric": "MoCo loss steps/s at bs128 (fwd+bwd+enqueue, bf16 fused kernel, whole step in one CUDA graph)",
                                         "value": best["value"], "unit": "steps/s", "us_per_step": best["ms_per_step"] * 1e3,
                                         "hbm_roofline_frac": (best.get("roofline") or {}).get("frac"),
                                         "library_launches": best.get("library_launches"),
                                         "e2e": {"value": e2e_l.get("value"), "unit": "steps/s", "h2d_bytes_per_step": e2e_l.get("h2d_bytes_per_step"),
                                                 "d2h_bytes_per_step": e2e_l.get("d2h_bytes_per_step")},
                                         "cpu_baseline_steps_per_s": base.get("value")}
            except Exception as e:   # pragma: no cover
                line["secondary"] = {"error": str(e)[:300]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
