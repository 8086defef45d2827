"""Seeded synthetic inputs of the shapes BASELINE.json names.  Used by bench.py,
__graft_entry__.smoke() and the tests; no dataset or checkpoint is needed (there is no network)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def loss_inputs(N, D, K, C, seed, masked="some"):
    """MoCo loss step inputs: TripletSampler-shaped labels (ids x 4), warm column-normalised queues,
    Xavier-uniform projection, post-Linear embeddings ~ N(0,1)*0.05.  CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, C, (max(N // 4, 1),), generator=g).repeat_interleave(4)[:N]
    id_queue = torch.randint(0, C, (max(K // 4, 1),), generator=g).repeat_interleave(4)[:K]
    if masked == "some":
        m = max(K // 16, 1)
        id_queue[:m] = labels[torch.randint(0, N, (m,), generator=g)]
    elif masked == "empty":
        id_queue[:] = -1
    bound = (6.0 / (D + C)) ** 0.5
    return dict(
        v_embed=0.05 * torch.randn(N, D, generator=g), t_embed=0.05 * torch.randn(N, D, generator=g),
        v_key=F.normalize(torch.randn(N, D, generator=g), dim=1), t_key=F.normalize(torch.randn(N, D, generator=g), dim=1),
        labels=labels, v_queue=F.normalize(torch.randn(D, K, generator=g), dim=0),
        t_queue=F.normalize(torch.randn(D, K, generator=g), dim=0), id_queue=id_queue.reshape(1, K),
        projection=(torch.rand(D, C, generator=g) * 2 - 1) * bound)


def eval_data(Q, G, D, n_ids, g_lo, g_hi, device, dtype, seed=0, signal=0.55):
    """Retrieval inputs with identity structure: embedding = signal * centre[pid] + N(0,1), gallery pid = g mod n_ids
    (G / n_ids images per identity), query pid uniform.  Queries are identical for every caller; the gallery slice
    [g_lo, g_hi) is generated in fixed 65536-row blocks so that any sharding yields the same global gallery."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    centres = torch.randn(n_ids, D, generator=gen, device=dev, dtype=torch.float32)
    q_pid = torch.randint(0, n_ids, (Q,), generator=gen, device=dev)
    text = (signal * centres[q_pid] + torch.randn(Q, D, generator=gen, device=dev)).to(dtype)
    BLK = 65536
    imgs, pids = [], []
    for b0 in range((g_lo // BLK) * BLK, g_hi, BLK):
        gen.manual_seed(seed * 1000003 + 17 + b0 // BLK)
        n = min(BLK, G - b0)
        pid = torch.arange(b0, b0 + n, device=dev) % n_ids
        img = (signal * centres[pid] + torch.randn(n, D, generator=gen, device=dev)).to(dtype)
        lo, hi = max(g_lo, b0) - b0, min(g_hi, b0 + n) - b0
        imgs.append(img[lo:hi])
        pids.append(pid[lo:hi])
    if not imgs:
        return text, q_pid, torch.zeros(0, D, device=dev, dtype=dtype), torch.zeros(0, dtype=torch.int64, device=dev)
    return text, q_pid, torch.cat(imgs), torch.cat(pids)


class SyntheticCaption:
    """Duck-typed stand-in for lib/utils/caption.py's Caption as the encoders and the head use it: ``text`` [1, L] int64 token
    ids (zero padded), ``length`` [1] int64, ``get_field("id")`` the person id, ``to(device)``."""

    def __init__(self, text, length, pid):
        self.text, self.length, self._id = text, length, pid

    def to(self, device):
        return SyntheticCaption(self.text.to(device), self.length.to(device), self._id.to(device))

    def get_field(self, name):
        if name != "id":
            raise KeyError(name)
        return self._id


def train_batch(N, n_classes, vocab, max_len=105, min_tokens=20, max_tokens=100, height=384, width=128, seed=0, device="cpu"):
    """One synthetic training batch of BASELINE configs[4]: ``N`` 384 x 128 images, captions of up to 100 tokens padded to 105
    (lib/data/build.py:26), TripletSampler-shaped ids (N/4 identities x 4)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(N, 3, height, width, generator=g)
    ids = torch.randint(0, n_classes, (max(N // 4, 1),), generator=g).repeat_interleave(4)[:N]
    caps = []
    for i in range(N):
        n_tok = int(torch.randint(min_tokens, max_tokens + 1, (1,), generator=g))
        text = torch.zeros(1, max_len, dtype=torch.int64)
        text[0, :n_tok] = torch.randint(1, vocab, (n_tok,), generator=g)
        caps.append(SyntheticCaption(text.to(device), torch.tensor([n_tok], dtype=torch.int64, device=device), ids[i].to(device)))
    return images.to(device), caps, ids.to(device)
