"""The two encoders of the reference's MoCo model, for the end-to-end train step (BASELINE configs[4]): a CLIP-style
"modified" ResNet for 384 x 128 person crops and a bidirectional GRU over CLIP token embeddings.

The encoders are NOT part of the accelerated hot path (out of scope, they stay PyTorch / cuDNN); they
exist here so that the fused loss head can be exercised and timed inside a real step without the reference's pretrained
files.  Parameter names and shapes follow the reference modules (lib/models/backbones/m_resnet.py:11-217, gru.py:7-82), so a
reference checkpoint's ``v_encoder_q.*`` / ``t_encoder_q.*`` entries load unchanged; the forward passes are written for the
B200 step: fused scaled-dot-product attention for the pooling head, no host read of the caption lengths (masking instead of
pack_padded_sequence), channels-last friendly.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv_bn(cin, cout, k, stride=1):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=False), nn.BatchNorm2d(cout)


class AntiAliasedBottleneck(nn.Module):
    """1x1 -> 3x3 -> (avg-pool when striding) -> 1x1, residual; the shortcut strides with an avg-pool too (m_resnet.py:11-67)."""
    expansion = 4

    def __init__(self, cin, planes, stride=1):
        super().__init__()
        self.conv1, self.bn1 = _conv_bn(cin, planes, 1)
        self.conv2, self.bn2 = _conv_bn(planes, planes, 3)
        self.avgpool = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
        self.conv3, self.bn3 = _conv_bn(planes, planes * self.expansion, 1)
        self.downsample = None
        if stride > 1 or cin != planes * self.expansion:
            self.downsample = nn.Sequential(OrderedDict([
                ("-1", nn.AvgPool2d(stride)), ("0", nn.Conv2d(cin, planes * self.expansion, 1, bias=False)),
                ("1", nn.BatchNorm2d(planes * self.expansion))]))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)), inplace=True)
        y = F.relu(self.bn2(self.conv2(y)), inplace=True)
        y = self.bn3(self.conv3(self.avgpool(y)))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)), inplace=True)


class AttentionPool(nn.Module):
    """QKV attention pooling of the final feature map: the query is the mean token (m_resnet.py:70-139).  Only the pooled
    token's output is needed, so the attention runs with ONE query row per image instead of HW + 1."""

    def __init__(self, spatial, embed_dim, heads, out_dim):
        super().__init__()
        self.positional_embedding = nn.Parameter(torch.randn(spatial[0] * spatial[1] + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj, self.q_proj, self.v_proj = nn.Linear(embed_dim, embed_dim), nn.Linear(embed_dim, embed_dim), nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, out_dim)
        self.num_heads = heads

    def forward(self, x):
        n, c = x.shape[0], x.shape[1]
        tok = x.flatten(2).transpose(1, 2)                                   # [N, HW, C]
        tok = torch.cat([tok.mean(dim=1, keepdim=True), tok], dim=1) + self.positional_embedding.to(tok.dtype)
        h, d = self.num_heads, c // self.num_heads
        q = self.q_proj(tok[:, :1]).view(n, 1, h, d).transpose(1, 2)
        k = self.k_proj(tok).view(n, -1, h, d).transpose(1, 2)
        v = self.v_proj(tok).view(n, -1, h, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)                          # [N, h, 1, d]
        return self.c_proj(o.transpose(1, 2).reshape(n, c))


class ClipResNetEncoder(nn.Module):
    """CLIP's modified ResNet (3-conv stem, anti-aliased strides, attention pooling), m_resnet.py:142-217.
    ``layers=[3, 4, 6, 3], output_dim=1024, heads=32`` is RN50; ``[3, 4, 23, 3], 512, 32`` is RN101."""

    def __init__(self, layers: Sequence[int] = (3, 4, 6, 3), output_dim: int = 1024, heads: int = 32, last_stride: int = 1,
                 input_resolution=(384, 128), width: int = 64):
        super().__init__()
        self.out_channels = self.output_dim = output_dim
        self.input_resolution = tuple(input_resolution)
        self.conv1, self.bn1 = nn.Conv2d(3, width // 2, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(width // 2)
        self.conv2, self.bn2 = _conv_bn(width // 2, width // 2, 3)
        self.conv3, self.bn3 = _conv_bn(width // 2, width, 3)
        self.avgpool = nn.AvgPool2d(2)
        cin = width
        for i, (mult, blocks, stride) in enumerate(zip((1, 2, 4, 8), layers, (1, 2, 2, last_stride)), start=1):
            stage = [AntiAliasedBottleneck(cin, width * mult, stride)]
            cin = width * mult * AntiAliasedBottleneck.expansion
            stage += [AntiAliasedBottleneck(cin, width * mult) for _ in range(1, blocks)]
            setattr(self, "layer%d" % i, nn.Sequential(*stage))
        down = 16 if last_stride == 1 else 32
        self.attnpool = AttentionPool((input_resolution[0] // down, input_resolution[1] // down), width * 32, heads, output_dim)

    def forward(self, x):
        x = x.to(self.conv1.weight.dtype)
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)):
            x = F.relu(bn(conv(x)), inplace=True)
        x = self.avgpool(x)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.attnpool(x)


class _BiGRUWeights(nn.Module):
    """The four weight matrices of a one-layer bidirectional bias-free nn.GRU under nn.GRU's parameter names (so the
    reference's ``gru.*`` state-dict entries load) and with nn.GRU's initialisation -- but as plain parameters: nn.GRU
    flattens its weights into one cuDNN buffer laid out for the BIDIRECTIONAL network, and handing views of that buffer to
    a unidirectional cuDNN call silently mis-reads them (measured on B200, torch 2.11: 0.36 absolute error)."""

    def __init__(self, embed_size, hidden_dim):
        super().__init__()
        bound = 1.0 / hidden_dim ** 0.5
        for name, cols in (("weight_ih_l0", embed_size), ("weight_hh_l0", hidden_dim), ("weight_ih_l0_reverse", embed_size),
                           ("weight_hh_l0_reverse", hidden_dim)):
            setattr(self, name, nn.Parameter(torch.empty(3 * hidden_dim, cols).uniform_(-bound, bound)))


class BiGRUTextEncoder(nn.Module):
    """gru.py:7-82 with ONEHOT = "clip_vit": frozen token table [V, vocab_size] -> (Linear when vocab_size != embed_size) ->
    bidirectional GRU without biases -> max over time.  ``captions`` is the reference's list of Caption objects (``.text``
    [1, L] int64, ``.length`` [1]) or a pair of tensors ``(tokens [N, L], lengths [N])``.

    The reference packs the padded batch (one host read of the lengths per call).  Here both directions run on the padded
    batch: the forward direction is causal, so its outputs at t < length are those of the packed run; the backward direction
    runs on the per-sequence REVERSED tokens (a gather), which is exactly what packing does for it.  Positions >= length give
    0 like pad_packed_sequence, and the maximum is taken over the batch's longest length like the reference's padded output."""

    def __init__(self, vocab_table: torch.Tensor, hidden_dim: int = 512, embed_size: int = 512, num_layers: int = 1,
                 dropout: float = 0.0, bidirectional: bool = True):
        super().__init__()
        vocab_size = vocab_table.shape[1]
        self.embed = None if vocab_size == embed_size else nn.Linear(vocab_size, embed_size)
        self.register_buffer("vocab_dict", vocab_table.float(), persistent=False)      # a plain attribute in the reference
        if num_layers != 1 or not bidirectional or dropout != 0.0:
            raise NotImplementedError("the reference's configs use one bidirectional layer without dropout (cfg.MODEL.GRU.*)")
        self.gru = _BiGRUWeights(embed_size, hidden_dim)
        self.hidden_dim = hidden_dim
        self.out_channels = hidden_dim * 2

    @staticmethod
    def _unpack(captions):
        if isinstance(captions, (tuple, list)) and len(captions) == 2 and torch.is_tensor(captions[0]) and captions[0].dim() == 2:
            return captions[0], captions[1].reshape(-1)
        text = torch.stack([c.text for c in captions], dim=1)
        length = torch.stack([c.length for c in captions], dim=1)
        return text.view(-1, text.size(-1)), length.view(-1)

    def _run(self, x, w_ih, w_hh):
        # one direction of the (bias-free) GRU over a padded batch-first input, zero initial state: cuDNN through torch's op
        h0 = x.new_zeros(1, x.shape[0], self.hidden_dim)
        out, _ = torch._VF.gru(x, h0, [w_ih, w_hh], False, 1, 0.0, self.training, False, True)
        return out

    def forward(self, captions):
        tokens, length = self._unpack(captions)
        n, L = tokens.shape
        x = self.vocab_dict[tokens.reshape(-1)].reshape(n, L, -1)
        if self.embed is not None:
            x = self.embed(x)
        pos = torch.arange(L, device=tokens.device).unsqueeze(0)
        valid = pos < length.unsqueeze(1)                                             # [N, L]
        rev = torch.where(valid, length.unsqueeze(1) - 1 - pos, pos)                  # reverse inside the valid prefix
        g = self.gru
        fwd = self._run(x, g.weight_ih_l0, g.weight_hh_l0)
        bwd = self._run(torch.gather(x, 1, rev.unsqueeze(-1).expand_as(x)), g.weight_ih_l0_reverse, g.weight_hh_l0_reverse)
        bwd = torch.gather(bwd, 1, rev.unsqueeze(-1).expand_as(bwd))
        out = torch.cat([fwd, bwd], dim=2) * valid.unsqueeze(-1).to(fwd.dtype)
        # the reference's padded output is as long as the batch's longest caption: zeros enter the maximum of every shorter one
        longest = (pos < length.max()).unsqueeze(-1)
        return out.masked_fill(~longest, float("-inf")).amax(dim=1)


def synthetic_vocab_table(vocab: int = 49408, dim: int = 512, seed: int = 0) -> torch.Tensor:
    """Stand-in for datasets/cuhkpedes/clip_vocab_vit.npy (CLIP's token embedding table, asserted [*, 512] at gru.py:32-34)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(vocab, dim, generator=g) * 0.02
