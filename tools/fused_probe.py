"""Fused loss kernel probe (GPU box): compares every output of trb_moco_loss(precision=1, fused roles R) with the fp32 path,
output by output, and times the step.  Each configuration runs in its own subprocess under a timeout so that a trap or a
hang in one of them cannot take the others down.

    python tools/fused_probe.py [--full]         # driver: fused roles 7 (and, with --full, 1/2/4/0), then roles 7 with phase stamps
    python tools/fused_probe.py child R 0        # one configuration (R = bit mask of fused branches: 1 instance, 2 InfoNCE, 4 align)
"""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [(128, 256, 2048, 11003, "some"), (32, 64, 128, 1000, "empty"), (24, 128, 384, 700, "some"), (100, 192, 256, 257, "some")]


def call(lib, _lib, inp, shape, hp, precision, grads=True):
    import torch
    N, D, K, Cn = shape
    dev = "cuda"
    out = dict(losses=torch.full((3,), float("nan"), device=dev), vkn=torch.zeros(N, D, device=dev), tkn=torch.zeros(N, D, device=dev))
    if grads:
        for k in ("d_inst", "d_nce", "d_ga"):
            out[k] = torch.full((2, N, D), float("nan"), device=dev)
        out["d_proj"] = torch.full((D, Cn), float("nan"), device=dev)
    sh = _lib.MocoShape(N, D, K, Cn)
    nbytes = lib.trb_moco_loss_workspace_bytes(C.byref(sh), precision)
    ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=dev)
    p = _lib.ptr
    rc = lib.trb_moco_loss(p(inp["v_embed"]), p(inp["t_embed"]), p(inp["v_embed"]), p(inp["t_embed"]), p(inp["v_key"]), p(inp["t_key"]), 0,
                           p(out["vkn"]), p(out["tkn"]), p(inp["labels"]), p(inp["v_queue"]), p(inp["t_queue"]), p(inp["id_queue"]),
                           p(inp["projection"]), C.byref(sh), C.byref(hp), precision, p(out["losses"]), p(out.get("d_inst")),
                           p(out.get("d_nce")), p(out.get("d_ga")), p(out.get("d_proj")), p(ws), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "trb_moco_loss")
    torch.cuda.synchronize()
    return out, ws



def print_stamps(lib, roles, K, Cn, tag, ws=None, sh=None):
    import numpy as np
    from textreid_b200 import _lib
    buf = np.zeros(160 * 16, dtype=np.uint64)
    _lib.check(lib.trb_moco_loss_debug_stamps(_lib.ptr(ws), C.byref(sh), buf.ctypes.data_as(C.c_void_p)), "trb_moco_loss_debug_stamps")
    st = buf.reshape(160, 16).astype(np.int64)
    T_inst, T_k = (Cn + 127) // 128, (K + 127) // 128
    n_inst = T_inst if roles & 1 else 0
    n_nce = 2 * T_k if roles & 2 else 0
    n_tiles = n_inst + n_nce + (1 if roles & 4 else 0)
    groups = [("inst", 0, n_inst), ("nce", n_inst, n_inst + n_nce), ("align", n_inst + n_nce, n_tiles), ("spare", n_tiles, 148)]
    if sh.N > 128:      # an InfoNCE CTA per tile index (both modalities in turn), one align CTA per 128-row window; the per-window
        n_nce = T_k if roles & 2 else 0      # stamps show the LAST window
        n_ga = (sh.N + 127) // 128 if roles & 4 else 0
        n_tiles = n_inst + n_nce + n_ga
        groups = [("inst", 0, n_inst), ("nce", n_inst, n_inst + n_nce), ("align", n_inst + n_nce, n_tiles), ("spare", n_tiles, 148)]
    t0 = st[:n_tiles, 0]
    t0 = int(t0[t0 > 0].min())
    for name, lo, hi in groups:
        if hi <= lo:
            continue
        seg = st[lo:hi]
        print("STAMPS[%s] %s: " % (tag, name) + "  ".join(
            "%d:%.1f..%.1f" % (k, (seg[:, k][seg[:, k] > 0].min() - t0) / 1e3, (seg[:, k].max() - t0) / 1e3)
            for k in range(16) if (seg[:, k] > 0).any()), flush=True)

def child(roles, variant):
    os.environ["TRB_FUSED_ROLES"] = str(roles)
    import torch
    from textreid_b200 import _lib
    from textreid_b200.synthetic import loss_inputs
    lib = _lib.load()
    hp = _lib.MocoHParams(0.07, 0.1, 0.6, 0.4, 10.0, 40.0)
    ok_all = True
    for (N, D, K, Cn, masked) in SHAPES:
        inp = {k: v.cuda().contiguous() for k, v in loss_inputs(N, D, K, Cn, seed=N + K, masked=masked).items()}
        inp["id_queue"] = inp["id_queue"].reshape(-1).contiguous()
        ref, _ = call(lib, _lib, inp, (N, D, K, Cn), hp, 0)
        got, _ = call(lib, _lib, inp, (N, D, K, Cn), hp, 1)
        line = ["N%d D%d K%d C%d roles=%d var=%d" % (N, D, K, Cn, roles, variant)]
        ok = True
        for i, nm in enumerate(("inst", "nce", "ga")):
            e = abs(float(got["losses"][i]) - float(ref["losses"][i])) / max(abs(float(ref["losses"][i])), 1e-6)
            line.append("L.%s=%.1e" % (nm, e))
            ok &= e < 2e-3
        for nm in ("d_inst", "d_nce", "d_ga", "d_proj"):
            g, r = got[nm].double(), ref[nm].double()
            e = float((g - r).abs().max() / r.abs().max().clamp_min(1e-30))
            bad = int((~torch.isfinite(g)).sum())
            line.append("%s=%.1e%s" % (nm, e, ("(nonfinite %d)" % bad) if bad else ""))
            ok &= (e < 3e-2) and bad == 0
            if e >= 3e-2 and bad == 0:         # localise: per modality / per column tile
                if nm == "d_proj":
                    tiles = [(float((g[:, c:c + 128] - r[:, c:c + 128]).abs().max() / r.abs().max())) for c in range(0, Cn, 128)]
                    line.append("   d_proj tile errs first=%.1e last=%.1e max=%.1e rows0-127=%.1e rows128+=%.1e" % (
                        tiles[0], tiles[-1], max(tiles), float((g[:128] - r[:128]).abs().max() / r.abs().max()),
                        float((g[128:] - r[128:]).abs().max() / r.abs().max()) if D > 128 else 0.0))
                else:
                    line.append("   %s v=%.1e t=%.1e cols0-127=%.1e cols128+=%.1e" % (
                        nm, float((g[0] - r[0]).abs().max() / r.abs().max()), float((g[1] - r[1]).abs().max() / r.abs().max()),
                        float((g[..., :128] - r[..., :128]).abs().max() / r.abs().max()),
                        float((g[..., 128:] - r[..., 128:]).abs().max() / r.abs().max()) if D > 128 else 0.0))
        print(("PASS " if ok else "FAIL ") + " ".join(line), flush=True)
        ok_all &= ok
    # timing at the headline shape (or TRB_PROBE_TIMED="N,D,K,C"): eager library call and CUDA-graph replay of it
    N, D, K, Cn, masked = SHAPES[0]
    if os.environ.get("TRB_PROBE_TIMED"):
        N, D, K, Cn = (int(x) for x in os.environ["TRB_PROBE_TIMED"].split(","))
    inp = {k: v.cuda().contiguous() for k, v in loss_inputs(N, D, K, Cn, seed=1, masked=masked).items()}
    inp["id_queue"] = inp["id_queue"].reshape(-1).contiguous()
    sh = _lib.MocoShape(N, D, K, Cn)
    out = dict(losses=torch.zeros(3, device="cuda"), vkn=torch.zeros(N, D, device="cuda"), tkn=torch.zeros(N, D, device="cuda"),
               d_inst=torch.zeros(2, N, D, device="cuda"), d_nce=torch.zeros(2, N, D, device="cuda"), d_ga=torch.zeros(2, N, D, device="cuda"),
               d_proj=torch.zeros(D, Cn, device="cuda"))
    ws = torch.zeros(int(lib.trb_moco_loss_workspace_bytes(C.byref(sh), 1)), dtype=torch.uint8, device="cuda")
    p = _lib.ptr
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    qptr = torch.zeros(1, dtype=torch.int64, device="cuda")
    with_enqueue = os.environ.get("TRB_PROBE_ENQUEUE", "1") != "0"      # the product call: loss + gradients + enqueue

    def step():
        if with_enqueue:
            _lib.check(lib.trb_moco_step(p(inp["v_embed"]), p(inp["t_embed"]), p(inp["v_embed"]), p(inp["t_embed"]), p(inp["v_key"]), p(inp["t_key"]), 0,
                                         p(out["vkn"]), p(out["tkn"]), p(inp["labels"]), p(inp["v_queue"]), p(inp["t_queue"]), p(inp["id_queue"]), p(qptr),
                                         p(inp["projection"]), C.byref(sh), C.byref(hp), 1, p(out["losses"]), p(out["d_inst"]), p(out["d_nce"]),
                                         p(out["d_ga"]), p(out["d_proj"]), p(ws), ws.numel(), _lib.stream_ptr("cuda")), "trb_moco_step")
            return
        _lib.check(lib.trb_moco_loss(p(inp["v_embed"]), p(inp["t_embed"]), p(inp["v_embed"]), p(inp["t_embed"]), p(inp["v_key"]), p(inp["t_key"]), 0,
                                     p(out["vkn"]), p(out["tkn"]), p(inp["labels"]), p(inp["v_queue"]), p(inp["t_queue"]), p(inp["id_queue"]),
                                     p(inp["projection"]), C.byref(sh), C.byref(hp), 1, p(out["losses"]), p(out["d_inst"]), p(out["d_nce"]),
                                     p(out["d_ga"]), p(out["d_proj"]), p(ws), ws.numel(), _lib.stream_ptr("cuda")), "trb_moco_loss")

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        for name, fl in (("hot L2", False), ("L2 flushed", True)):
            ts = []
            for _ in range(20):
                if fl:
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(s)
                step()
                b.record(s)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            print("TIME roles=%d var=%d eager %s: median %.1f us  min %.1f us" % (roles, variant, name, ts[len(ts) // 2], ts[0]), flush=True)
            if os.environ.get("TRB_FUSED_DEBUG") and roles:
                print_stamps(lib, roles, K, Cn, name, ws, sh)
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                step()
            for name, fl in (("hot L2", False), ("L2 flushed", True)):
                ts = []
                for _ in range(20):
                    if fl:
                        flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(s)
                    g.replay()
                    b.record(s)
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b) * 1e3)
                ts.sort()
                print("TIME roles=%d graph %s: median %.1f us  min %.1f us" % (roles, name, ts[len(ts) // 2], ts[0]), flush=True)
        except Exception as e:  # noqa: BLE001
            print("GRAPH capture failed: %r" % (e,), flush=True)
    return 0 if ok_all else 1


def driver():
    results = {}
    full = "--full" in sys.argv
    plan = [(r, v, False) for r in ((1, 2, 4, 7, 0) if full else (7,)) for v in (0,)] + [(7, 0, True)]
    for roles, variant, debug in plan:
        env = dict(os.environ)
        if debug:
            env["TRB_FUSED_DEBUG"] = "1"
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(roles), str(variant)], timeout=150,
                               capture_output=True, text=True, env=env)
            rc, txt = r.returncode, r.stdout + r.stderr[-3000:]
        except subprocess.TimeoutExpired as e:
            rc, txt = 124, "TIMEOUT\n" + ((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""))
        results[(roles, variant, debug)] = rc
        print("=== roles=%d variant=%d debug=%d rc=%d\n%s" % (roles, variant, debug, rc, txt), flush=True)
    print("SUMMARY", results)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        sys.exit(child(int(sys.argv[2]), int(sys.argv[3])))
    driver()
