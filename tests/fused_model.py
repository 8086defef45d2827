"""Numerical model (torch, CPU) of the fused bf16 loss kernel, csrc/loss_fused.cu -- test infrastructure only.

It restates the kernel's ALGORITHM with its rounding points, not its code: operands rounded once to bf16, fp32 accumulation,
128-column tiles with per-tile base-2 softmax statistics combined after the fact, logit gradients rounded to bf16 before the
backward contractions, dE partial tiles rounded to bf16 before the fp32 reduction, <dWs, What> taken as sum_rows dz' * z.
tests/test_fused_algorithm_cpu.py checks this model against the fp64 oracle at the tolerances the GPU tests use, so the precision
contract of the design is pinned on the CPU as well (the CUDA kernel itself is checked on the GPU against the same oracle).
"""
import math

import torch

TILE = 128
LOG2E, LN2 = 1.4426950408889634, 0.6931471805599453


def bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def _tile_pass(E, B, scale, valid, lse_init, target_col, uni, onehot, N, want_dw):
    """One family of tiles (instance: B = W [D, C]; InfoNCE: B = queue [D, K]).
    E [rows, D] fp32 (already bf16-rounded), B fp32, scale [cols] (0 for excluded columns), valid [cols] bool,
    lse_init = (M2, S) running start per row, target_col [rows] (-1 = none).
    Returns lse2 [rows], sz [rows], zy [rows], dE [rows, D], dW [D, cols] or None."""
    rows, D = E.shape
    cols = B.shape[1]
    tiles = (cols + TILE - 1) // TILE
    Bb = bf16(B)
    sc2 = (scale * LOG2E).float()
    M = lse_init[0].clone()
    S = lse_init[1].clone()
    sz = torch.zeros(rows)
    zy = torch.zeros(rows)
    z2_tiles = []
    for t in range(tiles):
        c0, c1 = t * TILE, min(cols, (t + 1) * TILE)
        acc = E @ Bb[:, c0:c1]                                   # fp32 accumulation of bf16 products
        z2 = acc * sc2[c0:c1]
        z2_tiles.append(z2)
        v = valid[c0:c1]
        sz += (z2 * v).sum(1) * LN2
        in_tile = (target_col >= c0) & (target_col < c1)
        if in_tile.any():
            idx = (target_col[in_tile] - c0).long()
            zy[in_tile] = z2[in_tile, idx] * LN2
        zm = torch.where(v, z2, torch.full_like(z2, -1e30))
        m_t = zm.max(1).values
        s_t = torch.exp2(zm - m_t[:, None]).sum(1)
        has = v.any()
        if has:
            newM = torch.maximum(M, m_t)
            S = S * torch.exp2(M - newM) + s_t * torch.exp2(m_t - newM)
            M = newM
    lse2 = M + torch.log2(S)
    dE = torch.zeros(rows, D)
    dWs = torch.zeros(D, cols) if want_dw else None
    dot = torch.zeros(cols)
    for t in range(tiles):
        c0, c1 = t * TILE, min(cols, (t + 1) * TILE)
        z2 = z2_tiles[t]
        g = torch.exp2(z2 - lse2[:, None]) - uni
        hit = (target_col[:, None] == torch.arange(c0, c1)[None, :])
        g = g - hit.float() * onehot
        dz = g * (sc2[c0:c1] * (LN2 / N))                        # = (softmax - target) / N * scale; 0 where scale == 0
        dzb = bf16(dz)
        dE += bf16(dzb @ Bb[:, c0:c1].t())                       # partial tile rounded to bf16, summed in fp32
        if want_dw:
            dWs[:, c0:c1] = E.t() @ dzb
            dot[c0:c1] = (dz * z2).sum(0) * LN2
    dW = None
    if want_dw:
        dW = dWs - Bb * (scale * dot)[None, :]
    return lse2, sz, zy, dE, dW


def fused_loss_model(v_embed, t_embed, v_key, t_key, labels, v_queue, t_queue, id_queue, projection, *, T=0.07, epsilon=0.1,
                     alpha=0.6, beta=0.4, scale_pos=10.0, scale_neg=40.0):
    """Returns (losses dict, d_v_embed, d_t_embed, d_projection) for upstream gradients (1, 1, 1)."""
    N, D = v_embed.shape
    C = projection.shape[1]
    K = v_queue.shape[1]
    f = torch.float32
    ve, te = v_embed.to(f), t_embed.to(f)
    E = torch.cat([ve, te], 0)
    norms = E.norm(dim=1).clamp_min(1e-12)
    en = E / norms[:, None]
    y = labels.long()
    y2 = torch.cat([y, y])

    # ---- instance loss (losses.py:42-62): z = e @ W/||W||_col, label smoothing
    scale = 1.0 / projection.to(f).norm(dim=0).clamp_min(1e-12)
    lse2, sz, zy, dE_inst, dW = _tile_pass(bf16(E), projection.to(f), scale, torch.ones(C, dtype=torch.bool),
                                           (torch.full((2 * N,), -1e30), torch.zeros(2 * N)), y2, epsilon / C, 1.0 - epsilon, N, True)
    inst = (lse2 * LN2 - (1.0 - epsilon) * zy - (epsilon / C) * sz).sum() / N

    # ---- InfoNCE (head.py:148-170, losses.py:206-217): q = normalised embeds, masked queue slots, positive logit separate
    mask = (id_queue.reshape(-1)[:, None] == y[None, :]).any(1)
    keys = [t_key.to(f), v_key.to(f)]
    queues = [t_queue.to(f), v_queue.to(f)]
    nce = 0.0
    d_nce = []
    for m in range(2):
        q = en[m * N:(m + 1) * N]
        pos = (q * keys[m]).sum(1)
        z02 = pos / T * LOG2E
        sc = torch.where(mask, torch.zeros(K), torch.full((K,), 1.0 / T))
        l2, _, _, dq, _ = _tile_pass(bf16(q), queues[m], sc, ~mask, (z02.clone(), torch.ones(N)), torch.full((N,), -1), 0.0, 0.0, N, False)
        nce = nce + ((l2 - z02) * LN2).sum() / N
        dpos = (torch.exp2(z02 - l2) - 1.0) / (N * T)
        gq = dq + dpos[:, None] * keys[m]
        d_nce.append((gq - (gq * q).sum(1, keepdim=True) * q) / norms[m * N:(m + 1) * N, None])

    # ---- global align (losses.py:102-128) on the bf16-rounded normalised embeds
    eb = bf16(en)
    S = eb[:N] @ eb[N:].t()
    same = y[:, None] == y[None, :]
    x = torch.where(same, -scale_pos * (S - alpha), scale_neg * (S - beta))
    e = torch.exp(x)
    ga = torch.log1p(e).sum() * 2.0 / N
    dS = bf16(torch.where(same, torch.full_like(S, -scale_pos), torch.full_like(S, scale_neg)) * (e / (1 + e)) * (2.0 / N))
    gv, gt = dS @ eb[N:], dS.t() @ eb[:N]
    d_ga_v = (gv - (gv * eb[:N]).sum(1, keepdim=True) * eb[:N]) / norms[:N, None]
    d_ga_t = (gt - (gt * eb[N:]).sum(1, keepdim=True) * eb[N:]) / norms[N:, None]

    losses = {"instance_loss": inst, "infonce_loss": nce, "global_align_loss": ga}
    dv = dE_inst[:N] + d_nce[0] + d_ga_v
    dt = dE_inst[N:] + d_nce[1] + d_ga_t
    return losses, dv, dt, dW
