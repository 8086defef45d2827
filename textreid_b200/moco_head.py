"""FusedMoCoHead: drop-in for lib/models/embeddings/moco_head/head.py:9-187 (MoCoHead) and
moco_head/loss.py:8-43 (LossComputation).

Same constructor ``(cfg, visual_model, textual_model)``, same parameters and buffers (names,
shapes, dtypes -- so reference checkpoints load: ``t_queue``/``v_queue`` [D,K] fp32, ``id_queue``
[1,K] int64, ``queue_ptr`` [1] int64, ``v_embed_layer``, ``t_embed_layer``,
``loss_evaluator.projection``), same ``forward(images, captions)`` returning the 3-key loss dict in
training and ``[v_embed, t_embed]`` in eval.  Encoders stay PyTorch (out of scope); everything after
them runs in libtextreid_b200: normalisation, queue mask, logits, the three losses with their
gradients, the momentum update and the enqueue -- with no host synchronisation.
"""
from __future__ import annotations

import copy
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from .losses import MomentumUpdater, moco_loss_dict


class LossComputation(nn.Module):
    """Holds ``projection`` [D, num_classes] (Xavier-uniform), ``T`` = 0.07 and ``epsilon``
    (moco_head/loss.py:8-19).  The arithmetic lives in ``moco_loss_dict``."""

    def __init__(self, cfg):
        super().__init__()
        self.projection = Parameter(torch.randn(cfg.MODEL.EMBEDDING.FEATURE_SIZE, cfg.MODEL.NUM_CLASSES),
                                    requires_grad=True)
        self.epsilon = cfg.MODEL.EMBEDDING.EPSILON
        self.T = 0.07
        nn.init.xavier_uniform_(self.projection.data, gain=1)


def make_loss_evaluator(cfg):
    return LossComputation(cfg)


class FusedMoCoHead(nn.Module):
    def __init__(self, cfg, visual_model, textual_model, precision: str = "bf16", cuda_graph: bool = False,
                 cross_rank_queue: bool = False):
        super().__init__()
        self.embed_size = cfg.MODEL.EMBEDDING.FEATURE_SIZE
        self.K = cfg.MODEL.MOCO.K
        self.m = cfg.MODEL.MOCO.M
        self.fc = cfg.MODEL.MOCO.FC
        self.precision = precision
        self.cuda_graph = cuda_graph
        # opt-in: one global queue under data parallelism (keys of all ranks all-gathered before the enqueue) instead of the
        # reference's per-rank queues (train_net.py:50-56 wraps the model in DDP with broadcast_buffers=False)
        self.cross_rank_queue = cross_rank_queue

        self.v_encoder_q = visual_model
        self.t_encoder_q = textual_model
        self.v_encoder_k = copy.deepcopy(visual_model)
        self.t_encoder_k = copy.deepcopy(textual_model)
        for p in list(self.v_encoder_k.parameters()) + list(self.t_encoder_k.parameters()):
            p.requires_grad = False

        if self.fc:
            def mlp(in_dim):
                return nn.Sequential(nn.Linear(in_dim, self.embed_size), nn.ReLU(),
                                     nn.Linear(self.embed_size, self.embed_size))
            self.v_fc_q = mlp(visual_model.out_channels)
            self.t_fc_q = mlp(textual_model.out_channels)
            self.v_fc_k = copy.deepcopy(self.v_fc_q)
            self.t_fc_k = copy.deepcopy(self.t_fc_q)
            for p in list(self.v_fc_k.parameters()) + list(self.t_fc_k.parameters()):
                p.requires_grad = False

        self.v_embed_layer = nn.Linear(visual_model.out_channels, self.embed_size)
        self.t_embed_layer = nn.Linear(textual_model.out_channels, self.embed_size)

        self.register_buffer("t_queue", F.normalize(torch.rand(self.embed_size, self.K), dim=0))
        self.register_buffer("v_queue", F.normalize(torch.rand(self.embed_size, self.K), dim=0))
        self.register_buffer("id_queue", -torch.ones((1, self.K), dtype=torch.long))   # -1 = empty slot
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))

        self.loss_evaluator = make_loss_evaluator(cfg)
        self._momentum = MomentumUpdater(self.m)
        self._init_weight()
        self.register_load_state_dict_post_hook(FusedMoCoHead._check_queue_ptr)

    @staticmethod
    def _check_queue_ptr(module, incompatible_keys):
        """A checkpoint's ``queue_ptr`` is consumed on the device without a host read (head.py:100 does ``int(self.queue_ptr)``
        every step).  Validate it once, at load time: the reference's slice assignment (head.py:104) raises when the pointer
        does not leave room for a whole batch; the kernels wrap modulo K instead, so say so here."""
        ptr = int(module.queue_ptr.reshape(-1)[0])
        if not 0 <= ptr < module.K:
            raise ValueError("checkpoint queue_ptr=%d lies outside the queue [0, %d)" % (ptr, module.K))

    def _init_weight(self):
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.kaiming_normal_(mod.weight, a=0, mode="fan_out")
                nn.init.constant_(mod.bias, 0)
            elif isinstance(mod, nn.BatchNorm1d):
                nn.init.constant_(mod.weight, 1)
                nn.init.constant_(mod.bias, 0)

    def _ema_pairs(self):
        pairs = [(self.v_encoder_q, self.v_encoder_k), (self.t_encoder_q, self.t_encoder_k)]
        if self.fc:
            pairs += [(self.v_fc_q, self.v_fc_k), (self.t_fc_q, self.t_fc_k)]
        pq, pk = [], []
        for q, k in pairs:
            pq += list(q.parameters())
            pk += list(k.parameters())
        return pk, pq

    @torch.no_grad()
    def _momentum_update_key_encoder(self):
        pk, pq = self._ema_pairs()
        self._momentum(pk, pq)

    def forward(self, images, captions):
        v_feat = self.v_encoder_q(images)
        t_feat = self.t_encoder_q(captions)
        v_embed = self.v_embed_layer(v_feat)
        t_embed = self.t_embed_layer(t_feat)
        if not self.training:
            return [v_embed, t_embed]

        v_q = self.v_fc_q(v_feat) if self.fc else None
        t_q = self.t_fc_q(t_feat) if self.fc else None
        id_q = torch.stack([c.get_field("id") for c in captions]).long().to(v_embed.device)
        with torch.no_grad():
            self._momentum_update_key_encoder()
            v_k = self.v_encoder_k(images)
            t_k = self.t_encoder_k(captions)
            if self.fc:
                v_k, t_k = self.v_fc_k(v_k), self.t_fc_k(t_k)
            else:
                v_k, t_k = self.v_embed_layer(v_k), self.t_embed_layer(t_k)
        ev = self.loss_evaluator
        return moco_loss_dict(v_embed, t_embed, v_k, t_k, id_q, self.v_queue, self.t_queue, self.id_queue,
                              self.queue_ptr, ev.projection, T=ev.T, epsilon=ev.epsilon, enqueue=True,
                              v_embed_q=v_q, t_embed_q=t_q, normalize_keys=True, precision=self.precision,
                              cuda_graph=self.cuda_graph,
                              gather_group=True if (self.cross_rank_queue and torch.distributed.is_available()
                                                    and torch.distributed.is_initialized()) else None)


def build_moco_head(cfg, visual_model, textual_model):
    """Same factory signature as the reference (moco_head/head.py:185-187), so the import swap of INTEGRATION.md section 2 is
    all a maintainer does.  Default arithmetic = the product path: bf16 operands / fp32 accumulation, which at the reference's
    shapes (batch <= 128, FEATURE_SIZE <= 256) is the fused tcgen05 step of csrc/loss_fused.cu -- two launches, enqueue
    included.  ``MODEL.MOCO.PRECISION = "fp32"`` (a key the reference's config does not have) or TRB_LOSS_PRECISION=fp32 selects
    the 1e-5 FFMA parity path.  ``MODEL.MOCO.CUDA_GRAPH`` / TRB_LOSS_GRAPH=1 replays the loss step from a library-level CUDA
    graph; it is off by default because the fused step is already two launches (a graph would add an input copy and save
    nothing) -- it pays for the unfused sequences (fp32, or bf16 outside the fused gate such as batch 256)."""
    moco = cfg.MODEL.MOCO
    precision = getattr(moco, "PRECISION", None) or os.environ.get("TRB_LOSS_PRECISION", "bf16")
    graph = getattr(moco, "CUDA_GRAPH", None)
    if graph is None:
        graph = os.environ.get("TRB_LOSS_GRAPH", "0") not in ("0", "", "false", "False")
    cross = getattr(moco, "CROSS_RANK_QUEUE", None)
    if cross is None:
        cross = os.environ.get("TRB_CROSS_RANK_QUEUE", "0") not in ("0", "", "false", "False")
    return FusedMoCoHead(cfg, visual_model, textual_model, precision=precision, cuda_graph=bool(graph), cross_rank_queue=bool(cross))
