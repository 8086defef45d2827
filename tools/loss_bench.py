"""Loss-step bench lines alone (GPU box): the `secondary.moco_loss_*` entries of bench.py without the retrieval workload.

    python tools/loss_bench.py [mode ...]        # modes: stepgraph graph eager (default: stepgraph)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import textreid_b200 as trb


def main():
    modes = sys.argv[1:] or ["stepgraph"]
    dev = torch.device("cuda", 0)
    pk = bench.peaks()
    buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for shape in ("n128_k2048", "n256_k4096"):
        for mode in modes:
            r = bench.loss_step_line(trb, pk, dev, buf.zero_, shape, "bf16", mode)
            print(shape, mode, "%.1f us (best %.1f)  hbm frac %.3f  launches %d" % (
                r["ms_per_step"] * 1e3, r["ms_best"] * 1e3, r["roofline"]["frac"], r["library_launches"]), flush=True)


if __name__ == "__main__":
    main()
