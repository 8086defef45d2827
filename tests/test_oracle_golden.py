"""Pin the CPU oracle (oracle/textreid_oracle.py) against fixtures produced by the
unmodified reference (tools/make_golden.py).  No GPU, no /root/reference needed."""
import os

import numpy as np
import pytest
import torch

from oracle import textreid_oracle as O


def load(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name + ".npz")).items()}


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(autouse=True)
def _single_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", ["loss_fn_small", "loss_fn_nomask", "loss_fn_allmask", "loss_fn_eps0", "loss_fn_eps02"])
def test_loss_functions_match_reference(golden_dir, name):
    g = load(golden_dir, name)
    eps = float(g["eps"])
    args = [T(g[k]) for k in ("v_embed", "t_embed", "v_key", "t_key", "labels", "v_queue", "t_queue",
                              "id_queue", "projection")]
    mask = O.queue_positive_mask(args[7], args[4])
    keep = (~mask).nonzero().reshape(-1)
    assert torch.equal(keep, T(g["neg_idx"]))  # head.py:148-157 column set
    for i, (key, short) in enumerate((("instance_loss", "instance"), ("infonce_loss", "infonce"),
                                      ("global_align_loss", "global_align"))):
        w = [0.0, 0.0, 0.0]
        w[i] = 1.0
        losses, gv, gt, gp = O.moco_loss_dict_with_grads(*args, weights=w, epsilon=eps)
        torch.testing.assert_close(losses[key], T(g[key]), rtol=2e-6, atol=1e-6)
        torch.testing.assert_close(gv, T(g[f"grad.{short}.v"]), rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(gt, T(g[f"grad.{short}.t"]), rtol=1e-4, atol=2e-6)
        if f"grad.{short}.p" in g:
            torch.testing.assert_close(gp, T(g[f"grad.{short}.p"]), rtol=1e-4, atol=2e-6)
        else:
            assert gp is None or float(gp.abs().max()) == 0.0


def _linear(x, w, b):
    return x @ w.t() + b


@pytest.mark.parametrize("name", ["moco_head_small", "moco_head_fc"])
def test_moco_head_steps_match_reference(golden_dir, name):
    """Replays MoCoHead.forward/backward (head.py:111-176) step by step with oracle pieces:
    EMA -> key embeddings -> loss dict + grads -> enqueue."""
    g = load(golden_dir, name)
    N, F, D, K, C, steps, fc = [int(x) for x in g["meta"]]
    eps = float(g["eps"])
    st = {k[len("state0."):]: T(v).clone() for k, v in g.items() if k.startswith("state0.")}
    m = 0.999

    def enc(prefix, x):
        return _linear(x, st[prefix + ".lin.weight"], st[prefix + ".lin.bias"])

    def fc_head(prefix, x):
        h = _linear(x, st[prefix + ".0.weight"], st[prefix + ".0.bias"]).relu()
        return _linear(h, st[prefix + ".2.weight"], st[prefix + ".2.bias"])

    for s in range(steps):
        images, cfeat, labels = T(g[f"s{s}.images"]), T(g[f"s{s}.cfeat"]), T(g[f"s{s}.labels"])
        leaf_names = [k for k in st if ("encoder_q" in k or "embed_layer" in k or "fc_q" in k
                                         or k.endswith("projection"))]
        for k in leaf_names:
            st[k] = st[k].detach().clone().requires_grad_(True)
        vf, tf = enc("v_encoder_q", images), enc("t_encoder_q", cfeat)
        v_embed = _linear(vf, st["v_embed_layer.weight"], st["v_embed_layer.bias"])
        t_embed = _linear(tf, st["t_embed_layer.weight"], st["t_embed_layer.bias"])
        vq_raw = fc_head("v_fc_q", vf) if fc else None
        tq_raw = fc_head("t_fc_q", tf) if fc else None
        with torch.no_grad():
            pairs = [("v_encoder_q", "v_encoder_k"), ("t_encoder_q", "t_encoder_k")]
            if fc:
                pairs += [("v_fc_q", "v_fc_k"), ("t_fc_q", "t_fc_k")]
            for q, k in pairs:
                names = [n for n in st if n.startswith(q + ".")]
                O.ema_update([st[n.replace(q, k, 1)] for n in names], [st[n] for n in names], m)
            vk, tk = enc("v_encoder_k", images), enc("t_encoder_k", cfeat)
            if fc:
                vk, tk = fc_head("v_fc_k", vk), fc_head("t_fc_k", tk)
            else:
                vk = _linear(vk, st["v_embed_layer.weight"], st["v_embed_layer.bias"])
                tk = _linear(tk, st["t_embed_layer.weight"], st["t_embed_layer.bias"])
            vk, tk = O.normalize_rows(vk), O.normalize_rows(tk)
        d = O.moco_loss_dict(v_embed, t_embed, vk, tk, labels, st["v_queue"], st["t_queue"],
                             st["id_queue"], st["loss_evaluator.projection"], epsilon=eps,
                             v_embed_q=vq_raw, t_embed_q=tq_raw)
        sum(d.values()).backward()
        for k, v in d.items():
            torch.testing.assert_close(v.detach(), T(g[f"s{s}.loss.{k}"]), rtol=3e-6, atol=1e-6)
        for k in leaf_names:
            gk = f"s{s}.grad.{k}"
            if gk in g:
                torch.testing.assert_close(st[k].grad, T(g[gk]), rtol=2e-4, atol=2e-6)
        with torch.no_grad():
            O.enqueue(st["v_queue"], st["t_queue"], st["id_queue"], st["queue_ptr"], vk, tk, labels)
        for k in ("v_queue", "t_queue", "id_queue", "queue_ptr"):
            ref = T(g[f"s{s}.state.{k}"])
            if ref.dtype.is_floating_point:
                torch.testing.assert_close(st[k], ref, rtol=1e-5, atol=1e-6)
            else:
                assert torch.equal(st[k], ref)
        for k in [n for n in st if "encoder_k" in n or "fc_k" in n]:
            assert torch.equal(st[k], T(g[f"s{s}.state.{k}"])), k  # EMA is bit-exact
        for k in leaf_names:
            st[k] = st[k].detach()


def test_ema_bit_exact(golden_dir):
    g = load(golden_dir, "ema")
    k, q = T(g["k"]).clone(), T(g["q"])
    O.ema_update([k], [q], float(g["m"]))
    assert torch.equal(k, T(g["k1"]))


@pytest.mark.parametrize("name", ["rank_gauss", "rank_exact", "rank_exact_le2"])
@pytest.mark.parametrize("loop", [True, False])
def test_rank_matches_reference(golden_dir, name, loop):
    g = load(golden_dir, name)
    sim, tp, ip = T(g["similarity"]), T(g["text_pid"]), T(g["image_pid"])
    cmc, mAP, idx = O.rank(sim, tp, ip, (1, 5, 10), get_mAP=True, per_column_loop=loop)
    assert torch.equal(cmc, T(g["t2i_cmc"]))
    assert torch.equal(mAP, T(g["t2i_mAP"]))
    assert torch.equal(idx[:, :10], T(g["t2i_top10"]))
    cmc, mAP, idx = O.rank(sim.t(), ip, tp, (1, 5, 10), get_mAP=True, per_column_loop=loop)
    assert torch.equal(cmc, T(g["i2t_cmc"]))
    ref_map = T(g["i2t_mAP"])
    assert torch.equal(mAP, ref_map) or (torch.isnan(mAP) and torch.isnan(ref_map))  # num_rel = 0 -> NaN
    assert torch.equal(idx[:, :10], T(g["i2t_top10"]))
    if "t2i_cmc_topk" in g:
        cmc, idx = O.rank(sim, tp, ip, (1, 5, 10), get_mAP=False)
        assert torch.equal(cmc, T(g["t2i_cmc_topk"]))
        assert torch.equal(idx, T(g["t2i_idx_topk"]))


def test_similarity_and_exact_fixture(golden_dir):
    g = load(golden_dir, "rank_gauss")
    sim = O.similarity_matrix(T(g["text"]), T(g["image"]))
    torch.testing.assert_close(sim, T(g["similarity"]), rtol=1e-5, atol=1e-6)
    g = load(golden_dir, "rank_exact")
    sim = O.similarity_matrix(T(g["text"]), T(g["image"]))
    assert torch.equal(sim, T(g["similarity"]))  # +-1/16 fixture: every dot product is exact


def test_evaluation_pipeline(golden_dir):
    """evaluation.py:101-120 (dedup by first image id, normalise, similarity) and the rankings."""
    g = load(golden_dir, "evaluation_small")
    v, t = T(g["v"]), T(g["t"])
    keep = O.first_occurrence(list(g["image_ids"]))
    pids = T(g["pids"])
    image, image_pid, text_pid = v[keep], pids[keep], pids
    assert torch.equal(image_pid, T(g["plain.npz.image_pid"]))
    sim = O.similarity_matrix(t, image)
    torch.testing.assert_close(sim, T(g["plain.npz.similarity"]), rtol=1e-5, atol=1e-6)
    sim = T(g["plain.npz.similarity"])
    cmc, _ = O.rank(sim, text_pid, image_pid, (1, 5, 10), get_mAP=False)
    assert torch.equal(cmc, T(g["plain.t2i_cmc"]))
    assert torch.equal(cmc[0], T(g["plain.r1"]))
    cmc, _ = O.rank(sim.t(), image_pid, text_pid, (1, 5, 10), get_mAP=False)
    assert torch.equal(cmc, T(g["plain.i2t_cmc"]))
    # re-rank matrices (evaluation.py:40-65,122-124); float64 like the reference
    tn, im = O.normalize_rows(t), O.normalize_rows(image)
    rvn = O.jaccard_rerank_matrix(tn, im)
    rtn = O.jaccard_rerank_matrix(im, tn)
    assert rvn.dtype == torch.float64
    torch.testing.assert_close(rvn, T(g["rerank.npz.rvn_mat"]), rtol=0, atol=1e-12)
    torch.testing.assert_close(rtn, T(g["rerank.npz.rtn_mat"]), rtol=0, atol=1e-12)
    cmc, mAP, _ = O.rank(rvn + sim, text_pid, image_pid, (1, 5, 10), get_mAP=True)
    torch.testing.assert_close(cmc, T(g["rerank.re_t2i_cmc"]).to(cmc.dtype))
    torch.testing.assert_close(mAP, T(g["rerank.re_t2i_mAP"]).to(mAP.dtype))


def test_rank_zero_relevant_is_nan():
    sim = torch.tensor([[0.3, 0.2, 0.1]])
    cmc, mAP, _ = O.rank(sim, torch.tensor([7]), torch.tensor([1, 2, 3]), (1,), get_mAP=True)
    assert torch.isnan(mAP) and float(cmc[0]) == 0.0


def test_hit_ranks_consistent_with_rank(golden_dir):
    g = load(golden_dir, "rank_exact")
    sim, tp, ip = T(g["similarity"]), T(g["text_pid"]), T(g["image_pid"])
    ranks = O.hit_ranks(sim, tp, ip)
    ap = torch.tensor([sum((j + 1) / (int(r) + 1) for j, r in enumerate(rs)) / len(rs) for rs in ranks])
    _, mAP, _ = O.rank(sim, tp, ip, (1, 5, 10), get_mAP=True)
    assert abs(float(ap.mean() * 100) - float(mAP)) < 1e-4


# ---------------------------------------------------------------------------------------------------
# oracle/reference_head.py: the reference's step as an ATen call sequence (bench baseline arm) -- pinned to the same fixtures
# ---------------------------------------------------------------------------------------------------
class _StubEncoder(torch.nn.Module):
    def __init__(self, in_dim, out_channels, take_captions=False):
        super().__init__()
        self.lin = torch.nn.Linear(in_dim, out_channels)
        self.out_channels = out_channels
        self.take_captions = take_captions

    def forward(self, x):
        if self.take_captions:
            x = torch.stack([c.feat for c in x])
        return self.lin(x)


class _StubCaption:
    def __init__(self, feat, pid):
        self.feat, self._id = feat, pid

    def get_field(self, name):
        return self._id


@pytest.mark.parametrize("name", ["moco_head_small", "moco_head_fc"])
def test_reference_style_head_replays_reference_steps(golden_dir, name):
    from types import SimpleNamespace
    from oracle.reference_head import ReferenceStyleHead
    g = load(golden_dir, name)
    N, F, D, K, C, steps, fc = [int(x) for x in g["meta"]]
    cfg = SimpleNamespace(MODEL=SimpleNamespace(EMBEDDING=SimpleNamespace(FEATURE_SIZE=D, EPSILON=float(g["eps"])),
                                                MOCO=SimpleNamespace(K=K, M=0.999, FC=bool(fc)), NUM_CLASSES=C))
    head = ReferenceStyleHead(cfg, _StubEncoder(F, F), _StubEncoder(F, F, take_captions=True))
    head.load_state_dict({k[len("state0."):]: T(v) for k, v in g.items() if k.startswith("state0.")}, strict=True)
    head.train()
    for s in range(steps):
        images, cfeat, labels = T(g[f"s{s}.images"]), T(g[f"s{s}.cfeat"]), T(g[f"s{s}.labels"])
        head.zero_grad()
        losses = head(images, [_StubCaption(cfeat[i], labels[i]) for i in range(N)])
        sum(losses.values()).backward()
        for k in losses:
            torch.testing.assert_close(losses[k].detach(), T(g[f"s{s}.loss.{k}"]), rtol=2e-6, atol=1e-6)
        for k, p in head.named_parameters():
            if f"s{s}.grad.{k}" in g:
                torch.testing.assert_close(p.grad, T(g[f"s{s}.grad.{k}"]), rtol=1e-4, atol=2e-6)
        sd = head.state_dict()
        for k in sd:
            if f"s{s}.state.{k}" in g:
                ref = T(g[f"s{s}.state.{k}"])
                if ref.dtype.is_floating_point and "encoder_k" not in k and "fc_k" not in k:
                    torch.testing.assert_close(sd[k], ref, rtol=1e-5, atol=1e-6)
                else:
                    assert torch.equal(sd[k], ref), k
