"""GPU parity tests of the MoCo loss step (loss dict + gradients, EMA, enqueue) against the golden
fixtures of the unmodified reference and against the CPU oracle.  `pytest -m gpu` on a B200."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

import textreid_b200 as trb
from oracle import textreid_oracle as O

DEV = "cuda"
KEYS = ("instance_loss", "infonce_loss", "global_align_loss")


def load(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name + ".npz")).items()}


def T(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def run_fused(g_or_inputs, eps, weights=(1.0, 1.0, 1.0), precision="fp32", enqueue=False):
    a = {k: (T(v) if isinstance(v, np.ndarray) else v.to(DEV)) for k, v in g_or_inputs.items()}
    ve = a["v_embed"].clone().requires_grad_(True)
    te = a["t_embed"].clone().requires_grad_(True)
    pr = a["projection"].clone().requires_grad_(True)
    K = a["v_queue"].shape[1]
    ptr = torch.zeros(1, dtype=torch.int64, device=DEV)
    d = trb.moco_loss_dict(ve, te, a["v_key"], a["t_key"], a["labels"], a["v_queue"].clone(), a["t_queue"].clone(),
                           a["id_queue"].reshape(1, K).clone(), ptr, pr, epsilon=eps, enqueue=enqueue, precision=precision)
    total = sum(w * d[k] for w, k in zip(weights, KEYS))
    total.backward()
    return d, ve.grad, te.grad, pr.grad


@pytest.mark.parametrize("name", ["loss_fn_small", "loss_fn_nomask", "loss_fn_allmask", "loss_fn_eps0", "loss_fn_eps02"])
def test_loss_dict_matches_reference_golden(golden_dir, name):
    g = load(golden_dir, name)
    eps = float(g["eps"])
    inputs = {k: g[k] for k in ("v_embed", "t_embed", "v_key", "t_key", "labels", "v_queue", "t_queue", "id_queue", "projection")}
    for i, (key, short) in enumerate(zip(KEYS, ("instance", "infonce", "global_align"))):
        w = [0.0, 0.0, 0.0]
        w[i] = 1.0
        d, gv, gt, gp = run_fused(inputs, eps, w)
        torch.testing.assert_close(d[key].cpu(), torch.from_numpy(g[key]), rtol=1e-5, atol=1e-6)      # north star: 1e-5 fp32
        torch.testing.assert_close(gv.cpu(), torch.from_numpy(g[f"grad.{short}.v"]), rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(gt.cpu(), torch.from_numpy(g[f"grad.{short}.t"]), rtol=1e-4, atol=2e-6)
        if f"grad.{short}.p" in g:
            torch.testing.assert_close(gp.cpu(), torch.from_numpy(g[f"grad.{short}.p"]), rtol=1e-4, atol=2e-6)
        else:
            assert float(gp.abs().max()) == 0.0


from textreid_b200.synthetic import loss_inputs as synth_loss_inputs


@pytest.mark.parametrize("N,D,K,C,masked", [(128, 256, 2048, 11003, "some"), (256, 256, 4096, 11003, "some"),
                                              (32, 64, 128, 1000, "empty"), (20, 48, 60, 77, "some")])
def test_loss_dict_vs_oracle_full_size(N, D, K, C, masked):
    inp = synth_loss_inputs(N, D, K, C, seed=N + K, masked=masked)
    eps = 0.1
    d, gv, gt, gp = run_fused(inp, eps)
    args = [inp[k].double() if inp[k].dtype.is_floating_point else inp[k]
            for k in ("v_embed", "t_embed", "v_key", "t_key", "labels", "v_queue", "t_queue", "id_queue", "projection")]
    losses, rv, rt, rp = O.moco_loss_dict_with_grads(*args, epsilon=eps)          # fp64 oracle
    for k in KEYS:
        torch.testing.assert_close(d[k].cpu().double(), losses[k], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(gv.cpu().double(), rv, rtol=2e-4, atol=1e-6)
    torch.testing.assert_close(gt.cpu().double(), rt, rtol=2e-4, atol=1e-6)
    torch.testing.assert_close(gp.cpu().double(), rp, rtol=2e-4, atol=1e-7)


def test_upstream_gradient_weights_are_applied_on_device():
    inp = synth_loss_inputs(16, 32, 64, 101, seed=4)
    w = (0.5, 2.0, -1.5)
    d, gv, gt, gp = run_fused(inp, 0.1, w)
    args = [inp[k] for k in ("v_embed", "t_embed", "v_key", "t_key", "labels", "v_queue", "t_queue", "id_queue", "projection")]
    _, rv, rt, rp = O.moco_loss_dict_with_grads(*args, weights=w, epsilon=0.1)
    torch.testing.assert_close(gv.cpu(), rv, rtol=2e-4, atol=2e-6)
    torch.testing.assert_close(gt.cpu(), rt, rtol=2e-4, atol=2e-6)
    torch.testing.assert_close(gp.cpu(), rp, rtol=2e-4, atol=2e-6)


class StubEncoder(nn.Module):
    def __init__(self, in_dim, out_channels, take_captions=False):
        super().__init__()
        self.out_channels = out_channels
        self.take_captions = take_captions
        self.lin = nn.Linear(in_dim, out_channels)

    def forward(self, x):
        if self.take_captions:
            x = torch.stack([c.feat for c in x])
        return self.lin(x)


class StubCaption:
    def __init__(self, feat, pid):
        self.feat, self._id = feat, pid

    def get_field(self, name):
        return self._id


@pytest.mark.parametrize("name", ["moco_head_small", "moco_head_fc"])
def test_fused_moco_head_replays_reference_steps(golden_dir, name):
    """FusedMoCoHead loaded with the reference's state dict reproduces, step by step, the reference
    MoCoHead's losses, parameter gradients, EMA'd key encoders (bit-exact) and queue contents."""
    g = load(golden_dir, name)
    N, F, D, K, C, steps, fc = [int(x) for x in g["meta"]]
    cfg = SimpleNamespace(MODEL=SimpleNamespace(EMBEDDING=SimpleNamespace(FEATURE_SIZE=D, EPSILON=float(g["eps"])),
                                                MOCO=SimpleNamespace(K=K, M=0.999, FC=bool(fc)), NUM_CLASSES=C))
    head = trb.FusedMoCoHead(cfg, StubEncoder(F, F), StubEncoder(F, F, take_captions=True), precision="fp32")   # 1e-5 parity path
    state = {k[len("state0."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("state0.")}
    missing, unexpected = head.load_state_dict(state, strict=True)      # same names/shapes as the reference
    head.to(DEV).train()
    for s in range(steps):
        images, cfeat, labels = T(g[f"s{s}.images"]), T(g[f"s{s}.cfeat"]), T(g[f"s{s}.labels"])
        caps = [StubCaption(cfeat[i], labels[i]) for i in range(N)]
        head.zero_grad()
        losses = head(images, caps)
        assert set(losses) == set(KEYS)
        sum(losses.values()).backward()
        for k in KEYS:
            torch.testing.assert_close(losses[k].detach().cpu(), torch.from_numpy(g[f"s{s}.loss.{k}"]), rtol=1e-5, atol=2e-6)
        for k, p in head.named_parameters():
            gk = f"s{s}.grad.{k}"
            if gk in g:
                torch.testing.assert_close(p.grad.cpu(), torch.from_numpy(g[gk]), rtol=3e-4, atol=3e-6)
        sd = head.state_dict()
        for k in sd:
            gk = f"s{s}.state.{k}"
            if gk not in g:
                continue
            ref = torch.from_numpy(g[gk])
            if "encoder_k" in k or "fc_k" in k or not ref.dtype.is_floating_point:
                assert torch.equal(sd[k].cpu(), ref), k            # EMA / ids / pointer: bit-exact
            else:
                torch.testing.assert_close(sd[k].cpu(), ref, rtol=1e-5, atol=1e-6)


def test_fused_head_eval_branch_and_checkpoint_layout():
    cfg = SimpleNamespace(MODEL=SimpleNamespace(EMBEDDING=SimpleNamespace(FEATURE_SIZE=32, EPSILON=0.1),
                                                MOCO=SimpleNamespace(K=64, M=0.999, FC=False), NUM_CLASSES=50))
    head = trb.FusedMoCoHead(cfg, StubEncoder(8, 16), StubEncoder(8, 16, take_captions=True)).to(DEV).eval()
    sd = head.state_dict()
    assert sd["t_queue"].shape == (32, 64) and sd["v_queue"].dtype == torch.float32
    assert sd["id_queue"].shape == (1, 64) and sd["id_queue"].dtype == torch.int64 and int(sd["id_queue"].max()) == -1
    assert sd["queue_ptr"].shape == (1,) and sd["queue_ptr"].dtype == torch.int64
    assert sd["loss_evaluator.projection"].shape == (32, 50)
    x = torch.randn(4, 8, device=DEV)
    caps = [StubCaption(x[i], torch.tensor(i)) for i in range(4)]
    out = head(x, caps)
    assert isinstance(out, list) and len(out) == 2 and out[0].shape == (4, 32)      # un-normalised [v_embed, t_embed]


def test_enqueue_wraps_and_rejects_indivisible_batch():
    D, K, N = 16, 24, 8
    vq, tq = torch.zeros(D, K, device=DEV), torch.zeros(D, K, device=DEV)
    idq = -torch.ones(1, K, dtype=torch.int64, device=DEV)
    ptr = torch.tensor([16], dtype=torch.int64, device=DEV)
    rv, rt, rid, rptr = vq.cpu().clone(), tq.cpu().clone(), idq.cpu().clone(), ptr.cpu().clone()
    for step in range(4):           # 16 -> 0 -> 8 -> 16 -> 0 : wraps twice
        vk, tk = torch.randn(N, D), torch.randn(N, D)
        ids = torch.arange(N) + 100 * step
        trb.dequeue_and_enqueue(vq, tq, idq, ptr, vk.to(DEV), tk.to(DEV), ids.to(DEV))
        O.enqueue(rv, rt, rid, rptr, vk, tk, ids)
        assert torch.equal(vq.cpu(), rv) and torch.equal(tq.cpu(), rt)
        assert torch.equal(idq.cpu(), rid) and torch.equal(ptr.cpu(), rptr)
    with pytest.raises(AssertionError):
        trb.dequeue_and_enqueue(vq, tq, idq, ptr, torch.randn(5, D, device=DEV), torch.randn(5, D, device=DEV),
                                torch.arange(5, device=DEV))


@pytest.mark.parametrize("n", [1, 3, 4099, 1 << 20])
def test_ema_bit_exact(golden_dir, n):
    g = torch.Generator().manual_seed(n)
    k, q = torch.randn(n, generator=g), torch.randn(n, generator=g)
    want = k * 0.999 + q * (1.0 - 0.999)          # the reference expression, head.py:81
    kd = k.to(DEV)
    trb.ema_update_flat(kd, q.to(DEV), 0.999)
    assert torch.equal(kd.cpu(), want)
    # multi-tensor table, including a mis-aligned view
    big_k, big_q = torch.randn(n + 3, generator=g).to(DEV), torch.randn(n + 3, generator=g).to(DEV)
    pk, pq = [kd.clone(), big_k[1:n + 1]], [q.to(DEV), big_q[1:n + 1]]
    want2 = [p.cpu() * 0.999 + r.cpu() * (1.0 - 0.999) for p, r in zip(pk, pq)]
    upd = trb.MomentumUpdater(0.999)
    upd(pk, pq)
    for a, b in zip(pk, want2):
        assert torch.equal(a.cpu(), b)
    g2 = load(golden_dir, "ema")
    kd = T(g2["k"]).clone()
    trb.ema_update_flat(kd, T(g2["q"]), float(g2["m"]))
    assert torch.equal(kd.cpu(), torch.from_numpy(g2["k1"]))


@pytest.mark.parametrize("N,D,K,C,masked", [(128, 256, 2048, 11003, "some"), (32, 64, 128, 1000, "empty"), (20, 48, 60, 77, "some"),
                                              (256, 256, 4096, 11003, "some")])
def test_loss_dict_bf16_tensor_core_path(N, D, K, C, masked):
    """precision="bf16": every contraction runs on tcgen05 with bf16 operands / fp32 accumulation (row-wise softmax math
    stays fp32).  Losses within 1e-3 of the fp64 oracle (north star: 1e-3 on the bf16 path); gradients carry the bf16
    operand rounding (2^-9 relative per element) and are compared at that level."""
    inp = synth_loss_inputs(N, D, K, C, seed=N + K, masked=masked)
    d, gv, gt, gp = run_fused(inp, 0.1, precision="bf16")
    args = [inp[k].double() if inp[k].dtype.is_floating_point else inp[k]
            for k in ("v_embed", "t_embed", "v_key", "t_key", "labels", "v_queue", "t_queue", "id_queue", "projection")]
    losses, rv, rt, rp = O.moco_loss_dict_with_grads(*args, epsilon=0.1)
    for k in KEYS:
        torch.testing.assert_close(d[k].cpu().double(), losses[k], rtol=1e-3, atol=1e-4)
    for got, ref in ((gv, rv), (gt, rt), (gp, rp)):
        got = got.cpu().double()
        err = (got - ref).abs().max() / ref.abs().max()
        # measured on B200: 7.5e-3 (embedding gradients, fused kernel: bf16 operands + bf16 partial tiles), 4e-3 (projection)
        assert float(err) < (1.2e-2 if N >= 32 else 2.5e-2), float(err)
        cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0)
        assert float(cos) > 0.9995, float(cos)
    # and the two precisions agree with each other far inside the bf16 budget
    d32, gv32, _, gp32 = run_fused(inp, 0.1, precision="fp32")
    for k in KEYS:
        torch.testing.assert_close(d[k], d32[k], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cuda_graph_replay_matches_eager(precision):
    """cuda_graph=True replays loss + gradients + enqueue as one graph: same numbers, same queue evolution."""
    inp = synth_loss_inputs(32, 64, 128, 257, seed=7)
    def fresh():
        a = {k: v.clone().to(DEV) for k, v in inp.items()}
        a["ptr"] = torch.zeros(1, dtype=torch.int64, device=DEV)
        return a
    runs = {}
    for graph in (False, True):
        a = fresh()
        ve, te, pr = a["v_embed"].requires_grad_(True), a["t_embed"].requires_grad_(True), a["projection"].requires_grad_(True)
        hist = []
        for step in range(5):                      # 128 / 32 = 4 slots: the pointer wraps
            d = trb.moco_loss_dict(ve, te, a["v_key"], a["t_key"], a["labels"], a["v_queue"], a["t_queue"], a["id_queue"],
                                   a["ptr"], pr, epsilon=0.1, enqueue=True, precision=precision, cuda_graph=graph)
            ve.grad = te.grad = pr.grad = None
            sum(d.values()).backward()
            hist.append(([float(d[k]) for k in KEYS], ve.grad.clone(), pr.grad.clone(), int(a["ptr"]), a["v_queue"].clone(),
                         a["id_queue"].clone()))
        runs[graph] = hist
    for e, g in zip(runs[False], runs[True]):
        assert e[0] == g[0] and e[3] == g[3]
        assert torch.equal(e[1], g[1]) and torch.equal(e[2], g[2]) and torch.equal(e[4], g[4]) and torch.equal(e[5], g[5])


# ---------------------------------------------------------------------------------------------------
# fused cooperative kernel (csrc/loss_fused.cu): precision="bf16" takes it whenever the shape fits
# ---------------------------------------------------------------------------------------------------
import ctypes as C

from textreid_b200 import _lib

ARGS = ("v_embed", "t_embed", "v_key", "t_key", "labels", "v_queue", "t_queue", "id_queue", "projection")


def _launches(N, D, K, Cn, precision):
    shape = _lib.MocoShape(N, D, K, Cn)
    return int(_lib.load().trb_moco_loss_launches(C.byref(shape), precision))


def test_fused_shape_gate():
    """Two launches (prologue + cooperative kernel) exactly where the fused path applies."""
    assert _launches(128, 256, 2048, 11003, 1) == 2          # BASELINE configs[1]
    assert _launches(32, 64, 128, 257, 1) == 2               # D padded to 128
    assert _launches(100, 192, 200, 257, 1) == 2             # ragged K and C tiles
    # 128 < N <= 256 (BASELINE configs[2]): the 128-row windows are walked inside the same kernel -- still two launches
    assert _launches(256, 256, 4096, 11003, 1) == 2
    assert _launches(200, 128, 512, 700, 1) == 2                       # ragged second window
    # N > 256: instance + InfoNCE fused (one launch), the global-align branch unfused: shared prologue + 11 launches on the
    # tensor-core sequence + fused prologue and ONE cooperative launch + loss reduce
    assert _launches(384, 64, 768, 300, 1) == 1 + 11 + 2 + 1
    assert _launches(20, 48, 60, 77, 1) > 2                  # D not a multiple of 64
    assert _launches(128, 256, 2048, 40000, 1) > 2           # more tiles than SMs
    assert _launches(128, 256, 2048, 11003, 0) > 2           # fp32 path is never fused


@pytest.mark.parametrize("N,D,K,Cn,masked", [(128, 256, 2048, 11003, "some"), (100, 192, 200, 257, "some"), (32, 64, 128, 1000, "empty"),
                                              (8, 128, 384, 130, "some"), (256, 256, 4096, 11003, "some"), (200, 128, 512, 700, "some"),
                                              (384, 64, 256, 300, "empty"), (384, 64, 768, 300, "some"),
                                              (32, 64, 1024, 16640, "some"), (32, 64, 1024, 16768, "some")])
def test_fused_matches_unfused_bf16(monkeypatch, N, D, K, Cn, masked):
    """The fused kernel and the generic bf16 launch sequence round the same operands to bf16: losses agree to fp32 level,
    gradients far inside the bf16 budget.  Both are checked against the fp64 oracle.  N > 128 runs the fused kernel in 128-row
    windows (instance and InfoNCE tiles; projection gradient accumulated over the windows; queue mask from the whole batch --
    the (384, .., "some") case has batch ids beyond the first 256 rows that must mask their queue slots too).  The two shapes with
    130 / 131 class tiles leave one / no SM without a tile on a 148-SM device: every CTA then takes a share of the final
    reductions (no early arrival of the instance tiles), and without a spare CTA tile 0 forms the row losses."""
    assert _launches(N, D, K, Cn, 1) == (2 if N <= 256 else 15)
    inp = synth_loss_inputs(N, D, K, Cn, seed=N + K, masked=masked)
    fused = run_fused(inp, 0.1, precision="bf16")
    monkeypatch.setenv("TRB_FUSED_ROLES", "0")
    plain = run_fused(inp, 0.1, precision="bf16")
    monkeypatch.delenv("TRB_FUSED_ROLES")
    for k in KEYS:
        torch.testing.assert_close(fused[0][k], plain[0][k], rtol=2e-5, atol=1e-5)
    for a, b in zip(fused[1:], plain[1:]):
        err = (a - b).abs().max() / b.abs().max()
        assert float(err) < 1e-2, float(err)
    args = [inp[k].double() if inp[k].dtype.is_floating_point else inp[k] for k in ARGS]
    losses, rv, rt, rp = O.moco_loss_dict_with_grads(*args, epsilon=0.1)
    for k in KEYS:
        torch.testing.assert_close(fused[0][k].cpu().double(), losses[k], rtol=1e-3, atol=1e-4)
    for got, ref in zip(fused[1:], (rv, rt, rp)):
        cos = torch.nn.functional.cosine_similarity(got.cpu().double().flatten(), ref.flatten(), dim=0)
        assert float(cos) > 0.9995, float(cos)


def test_fused_edge_cases_all_masked_and_eps0():
    """Every queue slot masked (K' = 0: only the positive logit is left, infonce = 0) and label smoothing off."""
    N, D, K, Cn = 16, 64, 128, 300
    inp = synth_loss_inputs(N, D, K, Cn, seed=5, masked="some")
    inp["id_queue"][:] = inp["labels"][0]
    d, gv, gt, gp = run_fused(inp, 0.0, precision="bf16")
    args = [inp[k].double() if inp[k].dtype.is_floating_point else inp[k] for k in ARGS]
    losses, rv, rt, rp = O.moco_loss_dict_with_grads(*args, epsilon=0.0)
    assert float(d["infonce_loss"]) == 0.0 and float(losses["infonce_loss"]) == 0.0
    for k in KEYS:
        torch.testing.assert_close(d[k].cpu().double(), losses[k], rtol=1e-3, atol=1e-4)
    for got, ref in ((gv, rv), (gt, rt), (gp, rp)):
        assert float((got.cpu().double() - ref).abs().max() / ref.abs().max()) < 4e-2
    assert torch.isfinite(gv).all() and torch.isfinite(gp).all()


def test_fused_separate_queries_and_key_normalisation():
    """cfg.MODEL.MOCO.FC = True (head.py:118-124): InfoNCE queries come from their own head; keys arrive un-normalised."""
    N, D, K, Cn = 64, 256, 512, 2000
    inp = {k: v.to(DEV) for k, v in synth_loss_inputs(N, D, K, Cn, seed=11).items()}
    g = torch.Generator().manual_seed(3)
    vq_raw, tq_raw = torch.randn(N, D, generator=g).to(DEV), torch.randn(N, D, generator=g).to(DEV)
    vk_raw, tk_raw = 3.0 * inp["v_key"], 0.5 * inp["t_key"]
    leaves = [inp["v_embed"].clone().requires_grad_(True), inp["t_embed"].clone().requires_grad_(True),
              vq_raw.clone().requires_grad_(True), tq_raw.clone().requires_grad_(True), inp["projection"].clone().requires_grad_(True)]
    ptr = torch.zeros(1, dtype=torch.int64, device=DEV)
    d = trb.moco_loss_dict(leaves[0], leaves[1], vk_raw, tk_raw, inp["labels"], inp["v_queue"].clone(), inp["t_queue"].clone(),
                           inp["id_queue"].clone(), ptr, leaves[4], epsilon=0.1, enqueue=False, v_embed_q=leaves[2], t_embed_q=leaves[3],
                           normalize_keys=True, precision="bf16")
    sum(d.values()).backward()
    ref_leaves = [x.detach().cpu().double().requires_grad_(True) for x in leaves]
    c = {k: (v.cpu().double() if v.dtype.is_floating_point else v.cpu()) for k, v in inp.items()}
    rd = O.moco_loss_dict(ref_leaves[0], ref_leaves[1], c["v_key"], c["t_key"], c["labels"], c["v_queue"], c["t_queue"], c["id_queue"],
                          ref_leaves[4], epsilon=0.1, v_embed_q=ref_leaves[2], t_embed_q=ref_leaves[3])
    sum(rd.values()).backward()
    for k in KEYS:
        torch.testing.assert_close(d[k].detach().cpu().double(), rd[k].detach(), rtol=1e-3, atol=1e-4)
    for got, ref in zip(leaves, ref_leaves):
        err = (got.grad.cpu().double() - ref.grad).abs().max() / ref.grad.abs().max()
        assert float(err) < 2e-2, float(err)


@pytest.mark.parametrize("shape", [(128, 256, 2048, 11003), (256, 256, 4096, 11003), (200, 128, 512, 700)])
def test_fused_forward_only_and_repeatability(shape):
    """No tensor requires grad -> the library skips the backward half; repeated calls are bit-identical (fixed-order reductions,
    no floating-point atomics; the work-stealing reduction of the partial tiles hands units to whichever CTA is free, which must
    not change any sum).  Also at the windowed shapes (two 128-row windows inside the kernel, ragged second window)."""
    inp = {k: v.to(DEV) for k, v in synth_loss_inputs(*shape, seed=9).items()}
    ptr = torch.zeros(1, dtype=torch.int64, device=DEV)

    def call(grad):
        ve, te, pr = (inp[k].clone().requires_grad_(grad) for k in ("v_embed", "t_embed", "projection"))
        d = trb.moco_loss_dict(ve, te, inp["v_key"], inp["t_key"], inp["labels"], inp["v_queue"], inp["t_queue"], inp["id_queue"],
                               ptr, pr, epsilon=0.1, enqueue=False, precision="bf16")
        if grad:
            sum(d.values()).backward()
        return [d[k].detach().clone() for k in KEYS], ([ve.grad.clone(), te.grad.clone(), pr.grad.clone()] if grad else None)

    l0, _ = call(False)
    l1, g1 = call(True)
    l2, g2 = call(True)
    for a, b, c in zip(l0, l1, l2):
        assert torch.equal(a, b) and torch.equal(b, c)
    for a, b in zip(g1, g2):
        assert torch.equal(a, b)


def test_grad_combine_single_launch_backward():
    """trb_moco_grad_combine = g_inst * d_inst + g_nce * d_nce + g_ga * d_ga (+ projection), upstream scalars read on the device."""
    lib = _lib.load()
    N, D, Cn = 24, 64, 131
    g = torch.Generator().manual_seed(0)
    d_inst, d_nce, d_ga = (torch.randn(2, N, D, generator=g).to(DEV) for _ in range(3))
    d_proj = torch.randn(D, Cn, generator=g).to(DEV)
    gs = [torch.tensor(x, device=DEV) for x in (0.5, -2.0, 3.0)]
    for separate in (0, 1):
        for g_inst in (gs[0], torch.tensor(1.0, device=DEV), None):
            ov, ot, ovq, otq = (torch.full((N, D), float("nan"), device=DEV) for _ in range(4))
            op = torch.full((D, Cn), float("nan"), device=DEV)
            _lib.check(lib.trb_moco_grad_combine(_lib.ptr(d_inst), _lib.ptr(d_nce), _lib.ptr(d_ga), _lib.ptr(d_proj), _lib.ptr(g_inst),
                                                 _lib.ptr(gs[1]), _lib.ptr(gs[2]), separate, N * D, D * Cn, _lib.ptr(ov), _lib.ptr(ot),
                                                 _lib.ptr(ovq) if separate else None, _lib.ptr(otq) if separate else None, _lib.ptr(op),
                                                 _lib.stream_ptr(DEV)), "trb_moco_grad_combine")
            gi = 0.0 if g_inst is None else float(g_inst)
            for m, (o, oq) in enumerate(((ov, ovq), (ot, otq))):
                want = gi * d_inst[m] + 3.0 * d_ga[m] + (0 if separate else -2.0 * d_nce[m])
                torch.testing.assert_close(o, want, rtol=1e-6, atol=1e-6)
                if separate:
                    torch.testing.assert_close(oq, -2.0 * d_nce[m], rtol=0, atol=0)
            torch.testing.assert_close(op, gi * d_proj, rtol=0, atol=0)


def test_fused_workspace_reuse_has_no_stale_operands():
    """The operand images of the prologue live in the cached workspace.  A second call with different data must not see the
    first call's images, and the barrier words must be ready for the next call."""
    from textreid_b200 import losses as L
    L._workspaces.clear()
    shape = (64, 256, 1024, 3000)
    a = synth_loss_inputs(*shape, seed=21)
    b = synth_loss_inputs(*shape, seed=22)
    run_fused(a, 0.1, precision="bf16")
    second = run_fused(b, 0.1, precision="bf16")              # reuses the workspace the first call left behind
    L._workspaces.clear()
    fresh = run_fused(b, 0.1, precision="bf16")               # brand-new zero-filled workspace
    for k in KEYS:
        assert torch.equal(second[0][k], fresh[0][k])
    for x, y in zip(second[1:], fresh[1:]):
        assert torch.equal(x, y)


def test_fused_head_bf16_tracks_fp32_head():
    """FusedMoCoHead(precision="bf16") -- the fused cooperative kernel behind the reference's module surface -- follows the fp32
    head step by step (losses, parameter gradients, queue pointer, ids bit-exact; queue contents to fp32 rounding)."""
    N, F, D, K, Cn = 32, 24, 64, 128, 500
    cfg = SimpleNamespace(MODEL=SimpleNamespace(EMBEDDING=SimpleNamespace(FEATURE_SIZE=D, EPSILON=0.1),
                                                MOCO=SimpleNamespace(K=K, M=0.999, FC=False), NUM_CLASSES=Cn))
    torch.manual_seed(0)
    ref = trb.FusedMoCoHead(cfg, StubEncoder(F, F), StubEncoder(F, F, take_captions=True), precision="fp32").to(DEV).train()
    torch.manual_seed(0)
    fused = trb.FusedMoCoHead(cfg, StubEncoder(F, F), StubEncoder(F, F, take_captions=True), precision="bf16").to(DEV).train()
    fused.load_state_dict(ref.state_dict())
    assert _launches(N, D, K, Cn, 1) == 2
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        images = torch.randn(N, F, generator=g).to(DEV)
        cfeat = torch.randn(N, F, generator=g).to(DEV)
        labels = torch.randint(0, Cn, (N // 4,), generator=g).repeat_interleave(4).to(DEV)
        caps = [StubCaption(cfeat[i], labels[i]) for i in range(N)]
        outs = []
        for head in (ref, fused):
            head.zero_grad()
            losses = head(images, caps)
            sum(losses.values()).backward()
            outs.append(losses)
        for k in KEYS:
            torch.testing.assert_close(outs[1][k], outs[0][k], rtol=2e-3, atol=2e-4)
        for (name, p), (_, q) in zip(ref.named_parameters(), fused.named_parameters()):
            if p.grad is None:
                continue
            cos = torch.nn.functional.cosine_similarity(p.grad.flatten(), q.grad.flatten(), dim=0)
            assert float(cos) > 0.999, (name, float(cos))
        assert torch.equal(ref.queue_ptr, fused.queue_ptr) and torch.equal(ref.id_queue, fused.id_queue)
        torch.testing.assert_close(fused.v_queue, ref.v_queue, rtol=1e-6, atol=1e-7)


# ---------------------------------------------------------------------------------------------------
# round 2: step call with the enqueue inside the kernel, robustness guards, logits-level parity of the fused path
# ---------------------------------------------------------------------------------------------------
def test_cuda_graph_cache_hits_with_fresh_inputs_every_step():
    """A real training loop hands freshly allocated embeddings / keys / labels to every step: the library-level graph must be
    captured once (static input buffers inside the cache entry), not once per step."""
    from textreid_b200 import losses as L
    L._GRAPHS.clear()
    inp = synth_loss_inputs(32, 64, 128, 257, seed=9)
    vq, tq = inp["v_queue"].to(DEV), inp["t_queue"].to(DEV)
    idq, ptr = inp["id_queue"].to(DEV), torch.zeros(1, dtype=torch.int64, device=DEV)
    pr = inp["projection"].to(DEV).requires_grad_(True)
    rvq, rtq, ridq, rptr = vq.clone(), tq.clone(), idq.clone(), ptr.clone()
    keep = []                                          # hold every input alive: the allocator cannot hand an address back
    for step in range(6):
        g = torch.Generator().manual_seed(100 + step)
        ve = (0.05 * torch.randn(32, 64, generator=g)).to(DEV).requires_grad_(True)
        te = (0.05 * torch.randn(32, 64, generator=g)).to(DEV).requires_grad_(True)
        vk, tk = torch.randn(32, 64, generator=g).to(DEV), torch.randn(32, 64, generator=g).to(DEV)
        lab = torch.randint(0, 257, (32,), generator=g).to(DEV)
        keep += [ve, te, vk, tk, lab]
        d = trb.moco_loss_dict(ve, te, vk, tk, lab, vq, tq, idq, ptr, pr, epsilon=0.1, normalize_keys=True, precision="fp32",
                               cuda_graph=True)
        pr.grad = None
        sum(d.values()).backward()
        ve2, te2 = ve.detach().clone().requires_grad_(True), te.detach().clone().requires_grad_(True)
        pr2 = pr.detach().clone().requires_grad_(True)
        e = trb.moco_loss_dict(ve2, te2, vk, tk, lab, rvq, rtq, ridq, rptr, pr2, epsilon=0.1, normalize_keys=True, precision="fp32")
        sum(e.values()).backward()
        for k in KEYS:
            assert float(d[k]) == float(e[k]), (step, k)
        assert torch.equal(ve.grad, ve2.grad) and torch.equal(te.grad, te2.grad) and torch.equal(pr.grad, pr2.grad)
        assert torch.equal(vq, rvq) and torch.equal(idq, ridq) and int(ptr) == int(rptr)
        assert len(L._GRAPHS) == 1, "step %d re-captured the graph" % step


def test_enqueue_with_a_foreign_pointer_wraps_instead_of_overrunning():
    """queue_ptr from a checkpoint written with another batch size (here 20 with N = 8, K = 24): the reference's slice
    assignment raises (head.py:104); the kernels wrap modulo K -- no write past a queue row or past id_queue."""
    D, K, N = 16, 24, 8
    guard = 64
    store = torch.zeros(D * K + guard, device=DEV)
    vq, tq = store[:D * K].view(D, K), torch.zeros(D, K, device=DEV)
    ids_store = torch.full((K + guard,), -7, dtype=torch.int64, device=DEV)
    idq = ids_store[:K].view(1, K)
    ptr = torch.tensor([20], dtype=torch.int64, device=DEV)
    vk, tk = torch.randn(N, D, device=DEV), torch.randn(N, D, device=DEV)
    ids = torch.arange(N, device=DEV) + 500
    trb.dequeue_and_enqueue(vq, tq, idq, ptr, vk, tk, ids)
    cols = (20 + torch.arange(N)) % K
    want = torch.zeros(D, K)
    want[:, cols] = vk.cpu().t()
    assert torch.equal(vq.cpu(), want)
    assert torch.equal(idq.cpu()[0, cols], ids.cpu()) and int(ptr) == (20 + N) % K
    assert float(store[D * K:].abs().max()) == 0.0 and bool((ids_store[K:] == -7).all())      # guard zones untouched


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_out_of_range_label_gives_nan_instance_loss_not_a_wild_read(precision):
    inp = synth_loss_inputs(32, 64, 128, 257, seed=3)
    inp["labels"] = inp["labels"].clone()
    inp["labels"][5] = 257              # == NUM_CLASSES: the reference raises in scatter_ (losses.py:33)
    d, gv, gt, gp = run_fused(inp, 0.1, precision=precision)
    assert torch.isnan(d["instance_loss"]) and torch.isfinite(d["infonce_loss"]) and torch.isfinite(d["global_align_loss"])


@pytest.mark.parametrize("N,D,K,Cn,tile", [(128, 256, 2048, 11003, 0), (128, 256, 2048, 11003, 85), (100, 192, 200, 257, 1)])
def test_fused_logits_tile_within_1e3_of_the_bf16_operand_reference(monkeypatch, N, D, K, Cn, tile):
    """North star: logits within 1e-3 (relative) on the bf16 path.  On the product path the logits never leave the SM; the debug
    read-back dumps one instance tile.  Reference = the same bf16-rounded operands accumulated in fp64, column norm from the
    bf16-rounded W like the kernel."""
    from textreid_b200.losses import fused_debug_logits
    monkeypatch.setenv("TRB_FUSED_DEBUG_LOGITS", str(tile))
    inp = synth_loss_inputs(N, D, K, Cn, seed=N + tile)
    assert _launches(N, D, K, Cn, 1) == 2
    run_fused(inp, 0.1, precision="bf16")
    got = fused_debug_logits((N, D, K, Cn)).double()
    c0, c1 = tile * 128, min(Cn, tile * 128 + 128)
    W = inp["projection"][:, c0:c1]
    Wb = W.to(torch.bfloat16).double()
    scale = 1.0 / W.double().norm(dim=0).clamp_min(1e-12)          # losses.py:51 (the kernel takes the norm of the fp32 column)
    for mod, emb in enumerate((inp["v_embed"], inp["t_embed"])):
        ref = (emb.to(torch.bfloat16).double() @ Wb) * scale
        z = got[mod * 128: mod * 128 + N, : c1 - c0]
        err = (z - ref).abs().max() / ref.abs().max()
        assert float(err) < 1e-3, (mod, float(err))
        # and against the un-rounded fp32 reference logits: operand rounding only (2^-8 relative per operand)
        ref32 = (emb.double() @ W.double()) * scale
        assert float((z - ref32).abs().max() / ref32.abs().max()) < 1.5e-2


def test_fused_step_enqueues_inside_the_kernel_like_the_separate_call():
    """trb_moco_step on the fused path (2 launches, the enqueue rides in the cooperative kernel) leaves the queues, ids and
    pointer exactly as loss + trb_enqueue does, over several steps including the wrap-around."""
    N, D, K, Cn = 32, 64, 128, 1000
    inp = synth_loss_inputs(N, D, K, Cn, seed=21)
    shape = _lib.MocoShape(N, D, K, Cn)
    assert _lib.load().trb_moco_step_launches(C.byref(shape), 1) == 2
    state = [{k: inp[k].clone().to(DEV) for k in ("v_queue", "t_queue", "id_queue")} for _ in range(2)]
    ptrs = [torch.zeros(1, dtype=torch.int64, device=DEV) for _ in range(2)]
    pr = inp["projection"].to(DEV)
    for step in range(6):
        g = torch.Generator().manual_seed(step)
        ve, te = (0.05 * torch.randn(N, D, generator=g)).to(DEV), (0.05 * torch.randn(N, D, generator=g)).to(DEV)
        import torch.nn.functional as F
        vk = F.normalize(torch.randn(N, D, generator=g), dim=1).to(DEV)      # pre-normalised keys: both paths enqueue these bits
        tk = F.normalize(torch.randn(N, D, generator=g), dim=1).to(DEV)
        lab = torch.randint(0, Cn, (N,), generator=g).to(DEV)
        a = trb.moco_loss_dict(ve, te, vk, tk, lab, state[0]["v_queue"], state[0]["t_queue"], state[0]["id_queue"], ptrs[0], pr,
                               epsilon=0.1, normalize_keys=False, precision="bf16", enqueue=True)
        b = trb.moco_loss_dict(ve, te, vk, tk, lab, state[1]["v_queue"], state[1]["t_queue"], state[1]["id_queue"], ptrs[1], pr,
                               epsilon=0.1, normalize_keys=False, precision="bf16", enqueue=False)
        trb.dequeue_and_enqueue(state[1]["v_queue"], state[1]["t_queue"], state[1]["id_queue"], ptrs[1], vk, tk, lab)
        for k in KEYS:
            assert float(a[k]) == float(b[k]), (step, k)
        for k in state[0]:
            assert torch.equal(state[0][k], state[1][k]), (step, k)
        assert int(ptrs[0]) == int(ptrs[1]) == ((step + 1) * N) % K


# ---------------------------------------------------------------------------------------------------
# one whole train step with REAL encoders (no stubs) -- FusedMoCoHead vs the reference's step
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol_loss,tol_grad", [("fp32", 2e-5, 2e-3), ("bf16", 2e-3, 5e-2)])
def test_train_step_with_real_encoders_matches_the_reference_step(precision, tol_loss, tol_grad):
    """CLIP-style ResNet + bi-GRU (textreid_b200/encoders.py, pinned to the reference's encoder modules by tests/test_encoders.py)
    under FusedMoCoHead, against oracle/reference_head.py (the reference's MoCoHead step, pinned to reference-recorded goldens):
    same initial state, two steps; losses, parameter gradients, momentum encoders, queues, pointer."""
    from oracle.reference_head import ReferenceStyleHead
    from textreid_b200.encoders import BiGRUTextEncoder, ClipResNetEncoder, synthetic_vocab_table
    from textreid_b200.synthetic import train_batch
    torch.manual_seed(0)
    D, K, C, N = 64, 64, 97, 16
    cfg = SimpleNamespace(MODEL=SimpleNamespace(EMBEDDING=SimpleNamespace(FEATURE_SIZE=D, EPSILON=0.1),
                                                MOCO=SimpleNamespace(K=K, M=0.999, FC=False), NUM_CLASSES=C))
    table = synthetic_vocab_table(300, 32, seed=1)

    def encoders():
        return (ClipResNetEncoder(layers=[1, 1, 1, 1], output_dim=48, heads=4, input_resolution=(64, 32), width=8),
                BiGRUTextEncoder(table, hidden_dim=24, embed_size=32))

    fused = trb.FusedMoCoHead(cfg, *encoders(), precision=precision)
    ref = ReferenceStyleHead(cfg, *encoders())
    ref.load_state_dict(fused.state_dict(), strict=True)
    fused.to(DEV).train()
    ref.to(DEV).train()
    for step in range(2):
        images, caps, ids = train_batch(N, C, 300, max_len=12, min_tokens=3, max_tokens=12, height=64, width=32, seed=step, device=DEV)
        fused.zero_grad()
        ref.zero_grad()
        lf, lr = fused(images, caps), ref(images, caps)
        sum(lf.values()).backward()
        sum(lr.values()).backward()
        for k in KEYS:
            assert abs(float(lf[k]) - float(lr[k])) <= tol_loss * max(1.0, abs(float(lr[k]))), (step, k, float(lf[k]), float(lr[k]))
        pr = dict(ref.named_parameters())
        for k, p in fused.named_parameters():
            if p.grad is None:
                assert pr[k].grad is None or float(pr[k].grad.abs().max()) == 0.0, k
                continue
            scale = float(pr[k].grad.abs().max())
            if scale < 1e-4 and k.endswith("k_proj.bias"):   # analytically zero (softmax is shift-invariant): rounding noise only
                assert float(p.grad.abs().max()) < 1e-3, (step, k)
                continue
            assert float((p.grad - pr[k].grad).abs().max()) / scale <= tol_grad, (step, k, scale)
        sf, sr = fused.state_dict(), ref.state_dict()
        for k in sf:
            if "encoder_k" in k or k in ("id_queue", "queue_ptr"):
                assert torch.equal(sf[k], sr[k]), (step, k)                       # momentum update / ids / pointer: bit-exact
            elif k in ("v_queue", "t_queue"):
                torch.testing.assert_close(sf[k], sr[k], rtol=1e-5, atol=1e-6)
        # keep the two models on the same trajectory for the second step
        with torch.no_grad():
            for k, p in fused.named_parameters():
                if p.grad is not None:
                    p -= 0.01 * p.grad
                    pr[k] -= 0.01 * p.grad
