#!/usr/bin/env python
"""Benchmark of the TextReID hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

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
reference is pure Python/PyTorch and cannot travel to the GPU box) on a bounded sample of the workload.
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
def cpu_retrieval_sample(cfg, steps, warmup, Qs=1024, Gs=32768):
    from oracle import textreid_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Qs, Gs = min(Qs, cfg["Q"]), min(Gs, cfg["G"])
    text, q_pid, image, g_pid = make_eval_data(Qs, Gs, cfg["D"], max(Gs // 4, 1), 0, Gs, "cpu", torch.float32, seed=1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.retrieve(text, image, q_pid, g_pid, (1, 5, 10), get_mAP=True, per_column_loop=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    pair_rate = Qs * Gs / mean
    return dict(value=pair_rate / cfg["G"], unit="queries/s", cores=cores, kind="port",
                sample="oracle.retrieve (normalise + fp32 matmul + full argsort + reference per-column AP loop) on %d queries x %d "
                       "gallery, %d timed runs, %.2f s each; pair rate scaled to the workload's G=%d" % (Qs, Gs, len(times), mean, cfg["G"]),
                sample_seconds=mean, ms_per_step=mean * 1e3)


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


def secondary_measurements(pk, device):
    import textreid_b200 as trb
    from textreid_b200.synthetic import loss_inputs as synth_loss_inputs
    out = {}
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)      # > 126 MB L2

    def flush():
        flush_buf.fill_(1)

    # ---- MoCo loss dict fwd+bwd (+ enqueue), configs[1]: N=128, K=2048, D=256, C=11003 ----
    # one step = what trainer.py:81-90 does with the head's output: loss dict -> sum -> backward.  Labels change every step
    # (device-side) so that the queue never degenerates into "every slot masked".
    N, D, K, C = 128, 256, 2048, 11003
    bytes_alg = 27.8e6            # queues + projection read + dProjection write + embeddings
    launches = {}
    for prec in ("bf16", "fp32"):
        for mode in ("stepgraph", "graph", "eager"):
            inp = {k: v.to(device) for k, v in synth_loss_inputs(N, D, K, C, seed=0).items()}
            ve, te, pr = inp["v_embed"].requires_grad_(True), inp["t_embed"].requires_grad_(True), inp["projection"].requires_grad_(True)
            ptr = torch.zeros(1, dtype=torch.int64, device=device)
            labels = inp["labels"]

            def loss_step(inner_graph):
                labels.add_(97).remainder_(C)
                d = trb.moco_loss_dict(ve, te, inp["v_key"], inp["t_key"], labels, inp["v_queue"], inp["t_queue"], inp["id_queue"],
                                       ptr, pr, epsilon=0.1, enqueue=True, precision=prec, cuda_graph=inner_graph)
                (d["instance_loss"] + d["infonce_loss"] + d["global_align_loss"]).backward()

            if mode == "stepgraph":
                # the whole trainer-style step (loss dict -> sum -> backward -> enqueue) captured once with torch.cuda.graph and
                # replayed: the way a production loop removes the host from the path; gradients land in static .grad tensors
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

            med, best = time_cuda(fn, 40, 8, flush)
            key = "moco_loss_%s_%s" % (prec, mode)
            out[key] = {
                "metric": "MoCo loss steps/s (loss dict fwd+bwd + enqueue, bs128, queue 2048, D=256, C=11003)", "value": 1e3 / med,
                "unit": "steps/s", "ms_per_step": med, "ms_best": best, "dtype": "bf16" if prec == "bf16" else "f32",
                "mode": {"stepgraph": "whole step captured in one CUDA graph (torch.cuda.graph around loss dict + backward)",
                         "graph": "library call replayed from its own CUDA graph, autograd glue eager",
                         "eager": "every launch issued from Python"}[mode],
                "l2_flushed": True,
                "roofline": {"bound": "hbm", "achieved": bytes_alg / (med * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": bytes_alg / (med * 1e-3) / 1e9 / pk["hbm"], "traffic": None,
                             "note": "whole step (two label kernels, loss dict + gradients, loss sum, backward combine, enqueue); "
                                     "algorithmic 27.8 MB"}}
            del fn
        shape = trb._lib.MocoShape(N, D, K, C)
        launches[prec] = int(trb._lib.load().trb_moco_loss_launches(__import__("ctypes").byref(shape), 1 if prec == "bf16" else 0))
    out["moco_loss_kernel_launches"] = launches
    # ---- EMA over an RN50+GRU-sized arena: 41,755,488 fp32 parameters ----
    P = 41_755_488
    pk_, pq_ = torch.randn(P, device=device), torch.randn(P, device=device)
    med, best = time_cuda(lambda: trb.ema_update_flat(pk_, pq_, 0.999), 20, 3, flush)
    out["ema_rn50_gru"] = {"metric": "momentum update of 41,755,488 fp32 parameters", "ms": med,
                           "roofline": {"bound": "hbm", "achieved": 12.0 * P / (med * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                        "frac": 12.0 * P / (med * 1e-3) / 1e9 / pk["hbm"], "traffic": None}}
    del pk_, pq_
    # ---- CUHK-PEDES-sized evaluation, configs[0] ----
    cfg = WORKLOADS["retrieval_cuhk"]
    text, q_pid, image, g_pid = make_eval_data(cfg["Q"], cfg["G"], cfg["D"], cfg["n_ids"], 0, cfg["G"], device, torch.float32, seed=2)
    for prec in ("fp32", "bf16"):
        med, best = time_cuda(lambda: trb.retrieve(text, image, q_pid, g_pid, (1, 5, 10), True, prec), 10, 3, flush)
        out["retrieval_cuhk_" + prec] = {"metric": "retrieval queries/s, 6156 x 3074, D=256 (full step incl. pid bookkeeping)",
                                         "value": cfg["Q"] / (med * 1e-3), "unit": "queries/s", "ms_per_step": med, "dtype": prec}
    return out


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
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

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_retrieval_sample(cfg, max(1, args.steps), max(0, min(args.warmup, 1)))
        line = {"impl": "reference", "metric": "retrieval queries/s (sim+top-k+R@k/mAP)", "value": r["value"], "unit": "queries/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
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
    stream_ms = []

    def step(text_d, image_d, q_pid_d, g_pid_d):
        if world > 1:
            sizes = [shard_bounds(G, world, r)[1] - shard_bounds(G, world, r)[0] for r in range(world)]
            return retrieve_sharded(text_d, image_d, q_pid_d, g_pid_d, (1, 5, 10), True, args.precision, shard_sizes=sizes)
        return trb.retrieve(text_d, image_d, q_pid_d, g_pid_d, (1, 5, 10), True, args.precision)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        res = step(text, image, q_pid, g_pid)
    barrier()
    _lib.reset_launch_count()
    with ClockSampler(local_rank, enabled=(rank == 0)) as clocks:
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            if flush_buf is not None:
                flush_buf.fill_(1)
            res = step(text, image, q_pid, g_pid)
        ev1.record()
        barrier()
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

    # ---- end to end through the public API from pinned host buffers ----
    h_text, h_image = text.cpu().pin_memory(), image.cpu().pin_memory()
    h_qpid, h_gpid = q_pid.cpu().pin_memory(), g_pid.cpu().pin_memory()
    h2d = sum(x.numel() * x.element_size() for x in (h_text, h_image, h_qpid, h_gpid))

    def e2e_step():
        r = step(h_text.to(device, non_blocking=True), h_image.to(device, non_blocking=True),
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
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
            # (profiles/r01_ncu_retrieval_tc_final.md: 100k x 1M x 256 on one GPU); algorithmic bytes are 563 MB
            traffic = 4.24e9 if (world == 1 and args.workload == "retrieval_1m" and args.precision == "bf16") else None
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "traffic": traffic, "kernel": "retrieval_tc_kernel<0>" if args.precision == "bf16" else "stream_f32_kernel",
                    "kernel_ms": kern_ms, "peak_source": pk["source"] + " bf16 sustained (kernel timed back to back inside a long step)",
                    "algorithmic_flops_per_launch": flops}
            if args.precision == "fp32":
                roof["note"] = "fp32 FFMA parity path; the tensor-core denominator does not apply"
        line = {"metric": "retrieval queries/s (sim+top-k+R@k/mAP)", "value": Q / (ms * 1e-3), "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": config,
                "clocks": clocks.summary(), "gpu_launches": launches,
                "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": e2e_ms},
                "roofline": roof,
                "result": {"R@1": float(res.cmc[0]), "R@5": float(res.cmc[1]), "R@10": float(res.cmc[2]), "mAP": float(res.mAP)}}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_retrieval_sample(cfg, 2, 1)
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
                    line["loss_step"] = {"metric": "MoCo loss steps/s at bs128 (fwd+bwd+enqueue, bf16 fused kernel, whole step in one CUDA graph)",
                                         "value": best["value"], "unit": "steps/s", "us_per_step": best["ms_per_step"] * 1e3,
                                         "hbm_roofline_frac": (best.get("roofline") or {}).get("frac"),
                                         "library_launches": (line["secondary"].get("moco_loss_kernel_launches") or {}).get("bf16"),
                                         "cpu_baseline_steps_per_s": base.get("value")}
            except Exception as e:   # pragma: no cover
                line["secondary"] = {"error": str(e)[:300]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
