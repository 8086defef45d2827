"""TEST / BASELINE INFRASTRUCTURE -- restatement of the reference's MoCoHead train step as the reference executes it
(lib/models/embeddings/moco_head/head.py:73-176, moco_head/loss.py:21-39, lib/models/losses.py:6-62,102-128,206-217): the same
ATen call sequence, including the per-parameter momentum loop, the nonzero / unique queue mask, the gathered negatives, the
host-built one-hot targets and the ``int(queue_ptr)`` read.  Device-agnostic (runs on CPU here and on the B200 as the
"reference ATen sequence" arm of bench.py).  Nothing under textreid_b200/ imports this file.

Pinned like the rest of the oracle: ``ReferenceStyleHead`` driven with the fixtures of tests/golden/moco_head_*.npz reproduces the
unmodified reference step by step (tests/test_oracle_golden.py::test_reference_style_head_replays_reference_steps).
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import textreid_oracle as O


class ReferenceStyleHead(nn.Module):
    """Same modules / buffers / names as MoCoHead (head.py:9-62), so state dicts are interchangeable with the reference and
    with textreid_b200.FusedMoCoHead."""

    def __init__(self, cfg, visual_model, textual_model):
        super().__init__()
        self.embed_size = cfg.MODEL.EMBEDDING.FEATURE_SIZE
        self.K, self.m, self.fc = cfg.MODEL.MOCO.K, cfg.MODEL.MOCO.M, cfg.MODEL.MOCO.FC
        self.epsilon = cfg.MODEL.EMBEDDING.EPSILON
        self.v_encoder_q, self.t_encoder_q = visual_model, textual_model
        self.v_encoder_k, self.t_encoder_k = copy.deepcopy(visual_model), copy.deepcopy(textual_model)
        frozen = list(self.v_encoder_k.parameters()) + list(self.t_encoder_k.parameters())
        if self.fc:
            def mlp(i):
                return nn.Sequential(nn.Linear(i, self.embed_size), nn.ReLU(), nn.Linear(self.embed_size, self.embed_size))
            self.v_fc_q, self.t_fc_q = mlp(visual_model.out_channels), mlp(textual_model.out_channels)
            self.v_fc_k, self.t_fc_k = copy.deepcopy(self.v_fc_q), copy.deepcopy(self.t_fc_q)
            frozen += list(self.v_fc_k.parameters()) + list(self.t_fc_k.parameters())
        for p in frozen:
            p.requires_grad = False
        self.v_embed_layer = nn.Linear(visual_model.out_channels, self.embed_size)
        self.t_embed_layer = nn.Linear(textual_model.out_channels, self.embed_size)
        self.register_buffer("t_queue", F.normalize(torch.rand(self.embed_size, self.K), dim=0))
        self.register_buffer("v_queue", F.normalize(torch.rand(self.embed_size, self.K), dim=0))
        self.register_buffer("id_queue", -torch.ones((1, self.K), dtype=torch.long))
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))
        self.loss_evaluator = nn.Module()
        self.loss_evaluator.projection = nn.Parameter(torch.randn(self.embed_size, cfg.MODEL.NUM_CLASSES))
        nn.init.xavier_uniform_(self.loss_evaluator.projection.data, gain=1)
        self.T = 0.07
        self.host_syncs = 0          # counted for the bench line: reads that stall the host on the device

    @torch.no_grad()
    def _momentum(self):
        """head.py:73-94: one (mul, mul, add) triple of element-wise launches per parameter tensor."""
        pairs = [(self.v_encoder_q, self.v_encoder_k), (self.t_encoder_q, self.t_encoder_k)]
        if self.fc:
            pairs += [(self.v_fc_q, self.v_fc_k), (self.t_fc_q, self.t_fc_k)]
        for q, k in pairs:
            for pq, pk in zip(q.parameters(), k.parameters()):
                pk.data = pk.data * self.m + pq.data * (1.0 - self.m)

    def _smoothed_ce_reference_style(self, logits, labels):
        """losses.py:26-39: the one-hot target is built on the HOST (scatter_ on a CPU tensor) and uploaded."""
        log_probs = torch.log_softmax(logits, dim=1)
        onehot = torch.zeros(log_probs.size()).scatter_(1, labels.unsqueeze(1).data.cpu(), 1)
        self.host_syncs += 1
        target = onehot.to(logits.device)
        target = (1 - O.REFERENCE_SMOOTHING) * target + O.REFERENCE_SMOOTHING / logits.shape[1]
        return (-target * log_probs).mean(0).sum()

    def forward(self, images, captions):
        n = images.shape[0]
        v_feat, t_feat = self.v_encoder_q(images), self.t_encoder_q(captions)
        v_embed, t_embed = self.v_embed_layer(v_feat), self.t_embed_layer(t_feat)
        if not self.training:
            return [v_embed, t_embed]
        v_q = F.normalize(self.v_fc_q(v_feat) if self.fc else v_embed, dim=1)
        t_q = F.normalize(self.t_fc_q(t_feat) if self.fc else t_embed, dim=1)
        id_q = torch.stack([c.get_field("id") for c in captions]).long().to(v_embed.device)
        with torch.no_grad():
            self._momentum()
            v_k, t_k = self.v_encoder_k(images), self.t_encoder_k(captions)
            v_k = F.normalize(self.v_fc_k(v_k) if self.fc else self.v_embed_layer(v_k), dim=1)
            t_k = F.normalize(self.t_fc_k(t_k) if self.fc else self.t_embed_layer(t_k), dim=1)
        # head.py:148-157 -- nonzero and unique both return data-dependent shapes: two host syncs
        pos_idx = self.id_queue.expand(n, self.K).eq(id_q.unsqueeze(-1)).nonzero(as_tuple=False)[:, 1]
        uniq, counts = torch.unique(torch.cat([torch.arange(self.K, device=pos_idx.device), pos_idx]), return_counts=True)
        neg_idx = uniq[counts == 1]
        self.host_syncs += 3
        v_pos = (v_q * t_k).sum(1, keepdim=True)
        v_neg = v_q @ self.t_queue.clone().detach()[:, neg_idx]
        t_pos = (t_q * v_k).sum(1, keepdim=True)
        t_neg = t_q @ self.v_queue.clone().detach()[:, neg_idx]
        proj = self.loss_evaluator.projection
        w_hat = proj / proj.norm(dim=0, keepdim=True).clamp_min(1e-12)
        zv, zt = v_embed @ w_hat, t_embed @ w_hat
        if self.epsilon > 0:
            inst = self._smoothed_ce_reference_style(zv, id_q) + self._smoothed_ce_reference_style(zt, id_q)
        else:
            inst = F.cross_entropy(zv, id_q) + F.cross_entropy(zt, id_q)
        losses = {"instance_loss": inst, "infonce_loss": O.infonce_loss(v_pos, v_neg, t_pos, t_neg, self.T),
                  "global_align_loss": O.global_align_loss(v_embed, t_embed, id_q)}      # boolean-mask indexing: 2 more syncs
        self.host_syncs += 2
        with torch.no_grad():        # head.py:96-109
            ptr = int(self.queue_ptr)
            self.host_syncs += 1
            assert self.K % n == 0
            self.v_queue[:, ptr:ptr + n] = v_k.T
            self.t_queue[:, ptr:ptr + n] = t_k.T
            self.id_queue[:, ptr:ptr + n] = id_q.unsqueeze(0)
            self.queue_ptr[0] = (ptr + n) % self.K
        return losses
