"""CPU-side tests: the C-ABI library loads and exports every symbol the header declares, and the
host-side index logic (relevance CSR, first-occurrence dedup, work splitting, packed layout) is right.
No compute call is made here (no GPU in the build container)."""
import os
import re

import numpy as np
import pytest
import torch

import textreid_b200
from textreid_b200 import _lib
from textreid_b200.evaluation import build_relevance, first_occurrence_index, _choose_nsplit
from textreid_b200.retrieval_tc import choose_nsplit_tc
from oracle import textreid_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from textreid_b200.build import build_library
    build_library()
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "textreid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree: %s" % (
        set(names) ^ set(_lib.SIGNATURES))


def test_version_and_argument_validation_without_gpu(lib):
    assert lib.trb_version() == 100
    # rejected before any CUDA call: null pointers / bad shapes
    rc = lib.trb_l2_normalize_rows_f32(None, None, None, 4, 4, 1e-12, None)
    assert rc == -1 and b"null" in lib.trb_last_error_string()
    rc = lib.trb_enqueue(None, None, None, None, None, None, None, 3, 4, 8, None)
    assert rc == -1
    assert lib.trb_packed_rows(1) == 256 and lib.trb_packed_rows(257) == 512 and lib.trb_packed_rows(0) == 0
    assert lib.trb_packed_bytes(100, 256) == 256 * 256 * 2
    assert lib.trb_packed_bytes(100, 100) == 0     # D must be a multiple of 64 on the tensor-core path


def test_product_refuses_cpu_tensors(lib):
    x = torch.randn(4, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        textreid_b200.l2_normalize_rows(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        textreid_b200.rank(torch.randn(3, 5), torch.arange(3), torch.arange(5))
    with pytest.raises(RuntimeError, match="CUDA"):
        textreid_b200.retrieve(x, x, torch.arange(4), torch.arange(4))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="not built"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "textreid_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), "%s mentions the oracle" % f


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_build_relevance_matches_bruteforce(seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randint(0, 9, (23,), generator=g)
    gp = torch.randint(0, 12, (31,), generator=g)
    rel = build_relevance(q, gp)
    for i in range(23):
        want = (gp == q[i]).nonzero().reshape(-1)
        got = rel.rel_col[rel.rel_ptr[i]:rel.rel_ptr[i + 1]]
        assert torch.equal(got, want)
    assert rel.total == int(rel.rel_ptr[-1])
    empty = build_relevance(q, torch.zeros(0, dtype=torch.long))
    assert empty.total == 0 and int(empty.rel_ptr.sum()) == 0


def test_first_occurrence_matches_reference_semantics():
    ids = [5, 3, 5, 9, 3, 3, 7, 9, 1]
    assert torch.equal(first_occurrence_index(ids, "cpu"), O.first_occurrence(ids))


def test_work_split_choices():
    assert choose_nsplit_tc(782, 3907, 148) == 3        # 782 = 5*148 + 42 -> the 42 remainder tiles are cut in 3
    assert choose_nsplit_tc(296, 3907, 148) == 1        # whole waves only
    assert choose_nsplit_tc(1, 1, 148) == 1
    assert choose_nsplit_tc(49, 13, 148) == 3
    assert _choose_nsplit(6156, 3074, 128, 128, 148) >= 1
    assert _choose_nsplit(10, 5, 128, 128, 148) == 1


def packed_offset(row, chunk, kchunks):
    rb, r, kc, c = row >> 7, row & 127, chunk >> 3, chunk & 7
    return (rb * kchunks + kc) * 16384 + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)


def test_packed_layout_is_a_bijection_and_block_contiguous():
    rows, dim = 256, 256
    kch = dim // 64
    offs = np.array([[packed_offset(r, c, kch) for c in range(dim // 8)] for r in range(rows)])
    assert sorted(offs.reshape(-1).tolist()) == list(range(0, rows * dim * 2, 16))
    # every (128-row, 64-col) block is one contiguous 16 KiB span, 8-row groups 1024 B apart
    blk = offs[:128, :8]
    assert blk.min() == 0 and blk.max() == 16384 - 16
    assert offs[8, 0] - offs[0, 0] == 1024
    # Swizzle<3,4,3>: 16-byte chunk index XOR (row & 7)
    assert offs[3, 5] == 3 * 128 + ((5 ^ 3) << 4)


def test_fused_loss_gate_and_argument_validation_without_gpu(lib):
    """Pure host queries of the loss entry points: the shape gate of the fused kernel (148 SMs assumed when no device is
    present), workspace sizes, and argument validation that happens before any CUDA call."""
    import ctypes as C

    def launches(N, D, K, Cn, precision):
        shape = _lib.MocoShape(N, D, K, Cn)
        return int(lib.trb_moco_loss_launches(C.byref(shape), precision))

    assert launches(128, 256, 2048, 11003, 1) == 2         # prologue + one cooperative kernel
    assert launches(128, 256, 2048, 11003, 0) > 2          # fp32 parity path: launch sequence
    assert launches(256, 256, 4096, 11003, 1) == 2         # BASELINE configs[2]: 128-row windows walked inside the same kernel
    assert launches(384, 256, 4096, 11003, 1) == 15        # N > 256: the global-align branch stays on the unfused sequence
    assert launches(128, 320, 2048, 11003, 1) > 2          # D > 256
    assert launches(128, 256, 2048, 20000, 1) > 2          # 157 + 32 + 1 tiles > 148 SMs
    assert launches(0, 256, 2048, 11003, 1) == -1          # rejected shape
    fits, big = _lib.MocoShape(128, 256, 2048, 11003), _lib.MocoShape(256, 256, 4096, 11003)
    w_fused = lib.trb_moco_loss_workspace_bytes(C.byref(fits), 1)
    w_f32 = lib.trb_moco_loss_workspace_bytes(C.byref(fits), 0)
    assert w_fused > w_f32 > 0 and lib.trb_moco_loss_workspace_bytes(C.byref(big), 1) > 0
    assert lib.trb_moco_loss_workspace_bytes(C.byref(fits), 7) == -1
    rc = lib.trb_moco_grad_combine(None, None, None, None, None, None, None, 0, 16, 16, None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.trb_last_error_string()


def test_build_moco_head_precision_selection(monkeypatch):
    """build_moco_head keeps the reference's factory signature; the arithmetic path comes from optional config keys or the
    environment and defaults to the product path (bf16 operands; the fused tcgen05 step at the reference's shapes)."""
    from types import SimpleNamespace
    import torch.nn as nn

    class Enc(nn.Module):
        out_channels = 8

        def __init__(self):
            super().__init__()
            self.lin = nn.Linear(4, 8)

    def cfg(**moco):
        return SimpleNamespace(MODEL=SimpleNamespace(EMBEDDING=SimpleNamespace(FEATURE_SIZE=16, EPSILON=0.1),
                                                    MOCO=SimpleNamespace(K=32, M=0.999, FC=False, **moco), NUM_CLASSES=10))

    monkeypatch.delenv("TRB_LOSS_PRECISION", raising=False)
    monkeypatch.delenv("TRB_LOSS_GRAPH", raising=False)
    head = textreid_b200.build_moco_head(cfg(), Enc(), Enc())
    assert head.precision == "bf16" and head.cuda_graph is False
    head = textreid_b200.build_moco_head(cfg(PRECISION="fp32", CUDA_GRAPH=True), Enc(), Enc())
    assert head.precision == "fp32" and head.cuda_graph is True
    monkeypatch.setenv("TRB_LOSS_PRECISION", "fp32")
    assert textreid_b200.build_moco_head(cfg(), Enc(), Enc()).precision == "fp32"
    # a checkpoint whose queue pointer lies outside the queue is rejected at load time (the kernels never read it on the host)
    sd = head.state_dict()
    sd["queue_ptr"] = sd["queue_ptr"].clone().fill_(32)
    import pytest as _pytest
    with _pytest.raises(ValueError):
        head.load_state_dict(sd)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores; the only place besides tests / smoke where oracle code runs)
    must print ONE JSON line with the keys the driver reads, on the same metric / unit / direction as the CUDA arm."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "retrieval_cuhk",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "retrieval queries/s (sim+top-k+R@k/mAP)" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
