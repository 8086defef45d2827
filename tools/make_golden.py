#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference) on seeded synthetic inputs.  Runs only in the build container;
the fixtures it writes are committed so that the tests never need the reference.

Shims (none touches reference arithmetic):
  * a 9-line ``yacs.config`` stand-in on sys.path (yacs is not installed), so
    that ``lib.data.metrics.evaluation`` and ``lib.engine.inference`` import;
  * ``torch.Tensor.cuda`` -> identity, because ``.cuda()`` is hard-coded at
    head.py:154 and losses.py:36,215 and this container has no GPU;
  * for the tie-heavy ranking fixtures only, ``torch.argsort`` is called with
    ``stable=True`` (the north star pins ties to gallery-index order; the
    reference leaves them undefined).

Usage: python tools/make_golden.py [--out tests/golden]
"""
import argparse
import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"


def _install_shims():
    stub_dir = tempfile.mkdtemp(prefix="yacs_stub_")
    os.makedirs(os.path.join(stub_dir, "yacs"))
    with open(os.path.join(stub_dir, "yacs", "__init__.py"), "w") as f:
        f.write("")
    with open(os.path.join(stub_dir, "yacs", "config.py"), "w") as f:
        f.write(
            "class CfgNode(dict):\n"
            "    def __getattr__(self, k):\n"
            "        try: return self[k]\n"
            "        except KeyError: raise AttributeError(k)\n"
            "    def __setattr__(self, k, v): self[k] = v\n"
            "    def freeze(self): pass\n"
            "    def merge_from_file(self, *a): pass\n"
            "    def merge_from_list(self, *a): pass\n"
            "    def clone(self): return self\n"
        )
    sys.path.insert(0, stub_dir)
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self  # CPU container


class StubEncoder(nn.Module):
    """Stand-in for the CLIP-ResNet / bi-GRU encoders (out of scope): one Linear so
    the momentum update has parameters to move."""

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
        self.feat = feat
        self._id = pid

    def get_field(self, name):
        assert name == "id"
        return self._id


def make_cfg(D, K, C, eps, fc):
    return SimpleNamespace(MODEL=SimpleNamespace(
        EMBEDDING=SimpleNamespace(FEATURE_SIZE=D, EPSILON=eps),
        MOCO=SimpleNamespace(K=K, M=0.999, FC=fc),
        NUM_CLASSES=C))


def tnp(t):
    return t.detach().cpu().numpy().copy()  # copy: buffers are mutated in place by later steps


def golden_moco_head(out, name, *, N, F, D, K, C, eps, fc, steps, seed, id_pool, store_inputs=True):
    """Run reference MoCoHead.forward + backward for ``steps`` steps."""
    from lib.models.embeddings.moco_head.head import MoCoHead

    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    cfg = make_cfg(D, K, C, eps, fc)
    head = MoCoHead(cfg, StubEncoder(F, F), StubEncoder(F, F, take_captions=True))
    head.train()
    # make q and k encoders differ so the EMA moves something
    with torch.no_grad():
        for p in list(head.v_encoder_k.parameters()) + list(head.t_encoder_k.parameters()):
            p.add_(0.01 * torch.randn(p.shape, generator=g))
    # warm queues: column-normalised gaussians, some slots carrying ids of the pool
    with torch.no_grad():
        head.v_queue.copy_(nn.functional.normalize(torch.randn(D, K, generator=g), dim=0))
        head.t_queue.copy_(nn.functional.normalize(torch.randn(D, K, generator=g), dim=0))
        ids = torch.randint(0, id_pool, (K,), generator=g)
        ids[torch.rand(K, generator=g) < 0.25] = -1
        head.id_queue.copy_(ids.reshape(1, K))
        head.queue_ptr[0] = (K // N // 2) * N

    rec = {"meta": np.array([N, F, D, K, C, steps, int(fc)], dtype=np.int64),
           "eps": np.array(eps, dtype=np.float64)}
    state0 = {k: v.clone() for k, v in head.state_dict().items()}
    for k, v in state0.items():
        rec["state0." + k] = tnp(v)
    for s in range(steps):
        images = torch.randn(N, F, generator=g)
        cfeat = torch.randn(N, F, generator=g)
        labels = torch.randint(0, id_pool, (N // 4,), generator=g).repeat_interleave(4)[:N]
        caps = [StubCaption(cfeat[i], labels[i]) for i in range(N)]
        head.zero_grad()
        losses = head(images, caps)
        total = sum(losses.values())
        total.backward()
        rec[f"s{s}.images"] = tnp(images)
        rec[f"s{s}.cfeat"] = tnp(cfeat)
        rec[f"s{s}.labels"] = tnp(labels)
        for k, v in losses.items():
            rec[f"s{s}.loss.{k}"] = tnp(v)
        for k, p in head.named_parameters():
            if p.grad is not None:
                rec[f"s{s}.grad.{k}"] = tnp(p.grad)
        for k, v in head.state_dict().items():
            if "queue" in k or "encoder_k" in k or "fc_k" in k:
                rec[f"s{s}.state.{k}"] = tnp(v)
    np.savez_compressed(os.path.join(out, name + ".npz"), **rec)
    print("wrote", name, {k: float(v) for k, v in losses.items()})


def golden_loss_functions(out, name, *, N, D, K, C, eps, seed, mask_mode="some"):
    """Function-level: reference losses.* on embedding-level inputs; neg_idx built with the
    reference's own eq/nonzero/unique/counts recipe (head.py:148-157, inline there)."""
    import lib.models.losses as L

    g = torch.Generator().manual_seed(seed)
    v_embed = (0.05 * torch.randn(N, D, generator=g)).requires_grad_(True)
    t_embed = (0.05 * torch.randn(N, D, generator=g)).requires_grad_(True)
    v_key = nn.functional.normalize(torch.randn(N, D, generator=g), dim=1)
    t_key = nn.functional.normalize(torch.randn(N, D, generator=g), dim=1)
    v_queue = nn.functional.normalize(torch.randn(D, K, generator=g), dim=0)
    t_queue = nn.functional.normalize(torch.randn(D, K, generator=g), dim=0)
    labels = torch.randint(0, C, (max(N // 4, 1),), generator=g).repeat_interleave(4)[:N]
    if mask_mode == "none":
        id_queue = -torch.ones(1, K, dtype=torch.long)
    elif mask_mode == "all":
        id_queue = labels[torch.randint(0, N, (K,), generator=g)].reshape(1, K)
    else:
        id_queue = torch.randint(0, C, (1, K), generator=g)
        sel = torch.randperm(K, generator=g)[: max(K // 8, 1)]
        id_queue[0, sel] = labels[torch.randint(0, N, (sel.numel(),), generator=g)]
    bound = (6.0 / (D + C)) ** 0.5
    projection = ((torch.rand(D, C, generator=g) * 2 - 1) * bound).requires_grad_(True)

    v_q = nn.functional.normalize(v_embed, dim=1)
    t_q = nn.functional.normalize(t_embed, dim=1)
    pos_idx = id_queue.expand(N, K).eq(labels.unsqueeze(-1)).nonzero(as_tuple=False)[:, 1]
    unique, counts = torch.unique(torch.cat([torch.arange(K).long(), pos_idx]), return_counts=True)
    neg_idx = unique[counts == 1]
    v_pos = torch.einsum("nc,nc->n", [v_q, t_key]).unsqueeze(-1)
    v_neg = torch.einsum("nc,ck->nk", [v_q, t_queue.clone().detach()[:, neg_idx]])
    t_pos = torch.einsum("nc,nc->n", [t_q, v_key]).unsqueeze(-1)
    t_neg = torch.einsum("nc,ck->nk", [t_q, v_queue.clone().detach()[:, neg_idx]])
    li = L.instance_loss(projection, v_embed, t_embed, labels, epsilon=eps)
    ln = L.infonce_loss(v_pos, v_neg, t_pos, t_neg, 0.07)
    lg = L.global_align_loss(v_embed, t_embed, labels)
    rec = dict(v_embed=tnp(v_embed), t_embed=tnp(t_embed), v_key=tnp(v_key), t_key=tnp(t_key),
               v_queue=tnp(v_queue), t_queue=tnp(t_queue), labels=tnp(labels), id_queue=tnp(id_queue),
               projection=tnp(projection), eps=np.array(eps), neg_idx=tnp(neg_idx),
               instance_loss=tnp(li), infonce_loss=tnp(ln), global_align_loss=tnp(lg))
    for nm, l in (("instance", li), ("infonce", ln), ("global_align", lg)):
        gv, gt, gp = torch.autograd.grad(l, [v_embed, t_embed, projection], retain_graph=True,
                                         allow_unused=True)
        rec[f"grad.{nm}.v"] = tnp(gv if gv is not None else torch.zeros_like(v_embed))
        rec[f"grad.{nm}.t"] = tnp(gt if gt is not None else torch.zeros_like(t_embed))
        if gp is not None:
            rec[f"grad.{nm}.p"] = tnp(gp)
    np.savez_compressed(os.path.join(out, name + ".npz"), **rec)
    print("wrote", name, float(li), float(ln), float(lg), "K'=", neg_idx.numel())


def _load_eval():
    import importlib
    importlib.import_module("lib.data.metrics")  # the package re-exports the function under the same name
    return sys.modules["lib.data.metrics.evaluation"]


def eval_inputs(Q, G, D, n_ids, seed, exact=False, max_per_id=None):
    g = torch.Generator().manual_seed(seed)
    img_pid = torch.cat([torch.arange(n_ids), torch.randint(0, n_ids, (G - n_ids,), generator=g)])
    if max_per_id is not None:
        # at most `max_per_id` images per id, so per-query AP is summation-order free
        img_pid = torch.arange(G) % n_ids if G <= max_per_id * n_ids else img_pid
    img_pid = img_pid[torch.randperm(G, generator=g)]
    src = torch.randint(0, G, (Q,), generator=g)
    txt_pid = img_pid[src]
    if exact:
        image = (torch.randint(0, 2, (G, D), generator=g).float() * 2 - 1) / 16.0
        text = (torch.randint(0, 2, (Q, D), generator=g).float() * 2 - 1) / 16.0
        if D != 256:
            raise ValueError("exact fixture is defined for D=256")
    else:
        image = torch.randn(G, D, generator=g)
        text = 0.5 * image[src] + torch.randn(Q, D, generator=g)
    return text, image, txt_pid, img_pid


def golden_rank(out, name, *, Q, G, D, n_ids, seed, exact, max_per_id=None):
    E = _load_eval()
    text, image, txt_pid, img_pid = eval_inputs(Q, G, D, n_ids, seed, exact, max_per_id)
    tn = nn.functional.normalize(text, p=2, dim=1)
    im = nn.functional.normalize(image, p=2, dim=1)
    sim = torch.matmul(tn, im.t())
    topk = torch.tensor([1, 5, 10])
    real_argsort = torch.argsort
    if exact:
        torch.argsort = lambda *a, **k: real_argsort(*a, **{**k, "stable": True})
    try:
        t2i_cmc, t2i_map, t2i_idx = E.rank(sim, txt_pid, img_pid, topk, get_mAP=True)
        i2t_cmc, i2t_map, i2t_idx = E.rank(sim.t(), img_pid, txt_pid, topk, get_mAP=True)
    finally:
        torch.argsort = real_argsort
    rec = dict(text=tnp(text), image=tnp(image), text_pid=tnp(txt_pid), image_pid=tnp(img_pid),
               similarity=tnp(sim), t2i_cmc=tnp(t2i_cmc), t2i_mAP=tnp(t2i_map),
               t2i_top10=tnp(t2i_idx[:, :10]), i2t_cmc=tnp(i2t_cmc), i2t_mAP=tnp(i2t_map),
               i2t_top10=tnp(i2t_idx[:, :10]), exact=np.array(int(exact)))
    if not exact:
        # top-k path of the reference (torch.topk); tie-free input so order is defined
        c2, idx2 = E.rank(sim, txt_pid, img_pid, topk, get_mAP=False)
        rec["t2i_cmc_topk"] = tnp(c2)
        rec["t2i_idx_topk"] = tnp(idx2)
        # guard: fixture must be tie-free in every row for the reference order to be defined
        srt = sim.sort(dim=1, descending=True)[0]
        assert (srt[:, 1:] < srt[:, :-1]).all(), "gauss fixture has an exact tie; change the seed"
    np.savez_compressed(os.path.join(out, name + ".npz"), **rec)
    print("wrote", name, tnp(t2i_cmc), float(t2i_map), tnp(i2t_cmc), float(i2t_map))


class _DS:
    def __init__(self, image_ids, pids):
        self.image_ids, self.pids = image_ids, pids

    def __len__(self):
        return len(self.pids)

    def get_id_info(self, idx):
        return self.image_ids[idx], self.pids[idx]


def golden_evaluation(out, name, *, n_caps, n_imgs, D, n_ids, seed):
    """Full reference evaluation() (dedup by first image_id, normalise, similarity, rank,
    k-reciprocal re-rank) on a duck-typed dataset; records R@1 returns and the npz it saves."""
    E = _load_eval()
    import logging
    g = torch.Generator().manual_seed(seed)
    img_pid_u = torch.cat([torch.arange(n_ids), torch.randint(0, n_ids, (n_imgs - n_ids,), generator=g)])
    cap_img = torch.cat([torch.arange(n_imgs), torch.randint(0, n_imgs, (n_caps - n_imgs,), generator=g)])
    cap_img = cap_img[torch.randperm(n_caps, generator=g)]
    img_feat = torch.randn(n_imgs, D, generator=g)
    v = img_feat[cap_img] * 1.0
    t = 0.6 * img_feat[cap_img] + torch.randn(n_caps, D, generator=g)
    image_ids = [int(x) for x in cap_img]
    pids = [int(img_pid_u[i]) for i in cap_img]
    ds = _DS(image_ids, pids)
    preds = {i: [v[i], t[i]] for i in range(n_caps)}
    rec = dict(v=tnp(v), t=tnp(t), image_ids=np.array(image_ids), pids=np.array(pids))
    for rr in (False, True):
        with tempfile.TemporaryDirectory() as td:
            r1 = E.evaluation(ds, preds, td, [1, 5, 10], save_data=True, rerank=rr)
            data = np.load(os.path.join(td, "inference_data.npz"))
            tag = "rerank" if rr else "plain"
            rec[f"{tag}.r1"] = tnp(r1)
            for k in data.files:
                rec[f"{tag}.npz.{k}"] = data[k]
            # every ranking the reference logs, recomputed with its own rank()
            sim = torch.tensor(data["similarity"])
            ip, tp = torch.tensor(data["image_pid"]), torch.tensor(data["text_pid"])
            topk = torch.tensor([1, 5, 10])
            if rr:
                c, m, _ = E.rank(sim, tp, ip, topk, get_mAP=True)
                rec["rerank.t2i_cmc"], rec["rerank.t2i_mAP"] = tnp(c), tnp(m)
                c, m, _ = E.rank(sim.t(), ip, tp, topk, get_mAP=True)
                rec["rerank.i2t_cmc"], rec["rerank.i2t_mAP"] = tnp(c), tnp(m)
                rvn, rtn = torch.tensor(data["rvn_mat"]), torch.tensor(data["rtn_mat"])
                c, m, _ = E.rank(rvn + sim, tp, ip, topk, get_mAP=True)
                rec["rerank.re_t2i_cmc"], rec["rerank.re_t2i_mAP"] = tnp(c), tnp(m)
                c, m, _ = E.rank(rtn + sim.t(), ip, tp, topk, get_mAP=True)
                rec["rerank.re_i2t_cmc"], rec["rerank.re_i2t_mAP"] = tnp(c), tnp(m)
            else:
                c, _ = E.rank(sim, tp, ip, topk, get_mAP=False)
                rec["plain.t2i_cmc"] = tnp(c)
                c, _ = E.rank(sim.t(), ip, tp, topk, get_mAP=False)
                rec["plain.i2t_cmc"] = tnp(c)
    np.savez_compressed(os.path.join(out, name + ".npz"), **rec)
    print("wrote", name, float(rec["plain.r1"]), float(rec["rerank.r1"]))


def golden_encoders(out, name, seed):
    """The reference's own encoders (m_resnet.py ModifiedResNet, gru.py GRU), small instances, one forward each: the fixture
    that pins textreid_b200/encoders.py (same state dict -> same outputs).  The vocabulary table is written to a scratch
    <root>/datasets/cuhkpedes/clip_vocab_vit.npy because GRU.__init__ loads it from disk (directory.py:20-23)."""
    from lib.models.backbones.gru import GRU
    from lib.models.backbones.m_resnet import ModifiedResNet
    from lib.utils.caption import Caption
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    vis = ModifiedResNet(layers=[1, 2, 1, 1], output_dim=32, heads=4, last_stride=1, input_resolution=(64, 32), width=8).eval()
    with torch.no_grad():           # non-trivial running statistics so that the BatchNorm layers matter
        for m in vis.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
    images = torch.randn(3, 3, 64, 32, generator=g)
    with torch.no_grad():
        v_out = vis(images)
    root = tempfile.mkdtemp(prefix="trb_vocab_")
    os.makedirs(os.path.join(root, "datasets", "cuhkpedes"))
    V, E, H, L = 50, 24, 16, 9
    table = torch.randn(V, E, generator=g)
    np.save(os.path.join(root, "datasets", "cuhkpedes", "clip_vocab_vit.npy"), table.numpy())
    rec = {}
    for tag, embed in (("same", E), ("proj", 12)):        # vocab_size == embed_size (no Linear) and the projected variant
        txt = GRU(H, E, embed, 1, 0.0, True, "clip_vit", root).eval()
        lengths = [9, 4, 7, 1, 9]
        caps = []
        for n_tok in lengths:
            toks = torch.randint(1, V, (n_tok,), generator=g)
            caps.append(Caption([toks], max_length=L, padded=False))
        with torch.no_grad():
            t_out = txt(caps)
        rec.update({"%s.tokens" % tag: tnp(torch.stack([c.text for c in caps], 1).view(-1, L)),
                    "%s.lengths" % tag: np.asarray(lengths, dtype=np.int64), "%s.out" % tag: tnp(t_out)})
        rec.update({"%s.state.%s" % (tag, k): tnp(v) for k, v in txt.state_dict().items()})
    rec.update({"vis.images": tnp(images), "vis.out": tnp(v_out), "table": tnp(table)})
    rec.update({"vis.state." + k: tnp(v) for k, v in vis.state_dict().items()})
    np.savez_compressed(os.path.join(out, name + ".npz"), **rec)
    print("wrote", name, v_out.shape, t_out.shape)


def golden_ema(out, name, seed):
    """Reference momentum update arithmetic (head.py:78-85) on one tensor."""
    g = torch.Generator().manual_seed(seed)
    k = torch.randn(4099, generator=g)
    q = torch.randn(4099, generator=g)
    m = 0.999
    k1 = k * m + q * (1.0 - m)
    np.savez_compressed(os.path.join(out, name + ".npz"), k=tnp(k), q=tnp(q), k1=tnp(k1), m=np.array(m))
    print("wrote", name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    args = ap.parse_args()
    out = os.path.abspath(args.out)
    os.makedirs(out, exist_ok=True)
    _install_shims()
    torch.set_num_threads(1)  # summation order independent of the machine's core count

    golden_loss_functions(out, "loss_fn_small", N=16, D=32, K=64, C=101, eps=0.1, seed=1)
    golden_loss_functions(out, "loss_fn_nomask", N=8, D=32, K=32, C=37, eps=0.1, seed=2, mask_mode="none")
    golden_loss_functions(out, "loss_fn_allmask", N=8, D=32, K=32, C=37, eps=0.1, seed=3, mask_mode="all")
    golden_loss_functions(out, "loss_fn_eps0", N=8, D=64, K=32, C=50, eps=0.0, seed=4)
    # EPSILON other than 0.1: the reference only tests `epsilon > 0` and smooths with the class default 0.1 (losses.py:56-57,18)
    golden_loss_functions(out, "loss_fn_eps02", N=8, D=32, K=32, C=41, eps=0.2, seed=7)
    golden_moco_head(out, "moco_head_small", N=16, F=24, D=32, K=64, C=101, eps=0.1, fc=False,
                     steps=3, seed=5, id_pool=40)
    golden_moco_head(out, "moco_head_fc", N=8, F=16, D=32, K=32, C=53, eps=0.1, fc=True,
                     steps=2, seed=6, id_pool=20)
    golden_rank(out, "rank_gauss", Q=96, G=61, D=32, n_ids=20, seed=7, exact=False)
    golden_rank(out, "rank_exact", Q=80, G=70, D=256, n_ids=25, seed=8, exact=True)
    golden_rank(out, "rank_exact_le2", Q=64, G=60, D=256, n_ids=30, seed=9, exact=True, max_per_id=2)
    golden_evaluation(out, "evaluation_small", n_caps=60, n_imgs=31, D=32, n_ids=12, seed=10)
    golden_encoders(out, "encoders_small", seed=13)
    golden_ema(out, "ema", seed=11)


if __name__ == "__main__":
    main()
