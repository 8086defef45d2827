"""textreid_b200/encoders.py against outputs of the reference's own encoder modules (fixture written by tools/make_golden.py
from the unmodified lib/models/backbones/m_resnet.py and gru.py): same state dict, same inputs, same outputs."""
import os

import numpy as np
import torch

from textreid_b200.encoders import BiGRUTextEncoder, ClipResNetEncoder


def load(golden_dir):
    return {k: v for k, v in np.load(os.path.join(golden_dir, "encoders_small.npz")).items()}


def test_clip_resnet_encoder_matches_reference(golden_dir):
    g = load(golden_dir)
    enc = ClipResNetEncoder(layers=[1, 2, 1, 1], output_dim=32, heads=4, last_stride=1, input_resolution=(64, 32), width=8).eval()
    state = {k[len("vis.state."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("vis.state.")}
    enc.load_state_dict(state, strict=True)                      # same parameter / buffer names as the reference module
    with torch.no_grad():
        out = enc(torch.from_numpy(g["vis.images"]))
    torch.testing.assert_close(out, torch.from_numpy(g["vis.out"]), rtol=1e-4, atol=1e-5)


def test_bigru_text_encoder_matches_reference_without_packing(golden_dir):
    g = load(golden_dir)
    table = torch.from_numpy(g["table"])
    for tag, embed in (("same", table.shape[1]), ("proj", 12)):
        enc = BiGRUTextEncoder(table, hidden_dim=16, embed_size=embed).eval()
        state = {k[len(tag) + 7:]: torch.from_numpy(v) for k, v in g.items() if k.startswith(tag + ".state.")}
        enc.load_state_dict(state, strict=True)
        tokens, lengths = torch.from_numpy(g[tag + ".tokens"]), torch.from_numpy(g[tag + ".lengths"])
        with torch.no_grad():
            out = enc((tokens, lengths))
        torch.testing.assert_close(out, torch.from_numpy(g[tag + ".out"]), rtol=1e-4, atol=1e-5)
