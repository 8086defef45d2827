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
