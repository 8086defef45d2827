"""CPU check of the rank-counting PROTOCOL of csrc/retrieval_tc.cu (tests/stream_model.py: nextbelow / strict compare values,
count-first with corrections, padded tail tile, column groups, gallery shards) against the definition of a hit rank under the
pinned order (similarity desc, gallery index asc) -- the order evaluation.py:23-36 leaves undefined on ties.  The kernel itself is
checked on the GPU (tests/test_gpu_retrieval.py); this pins the bookkeeping it implements."""
import numpy as np
import pytest

from tests.stream_model import next_below, oracle_ranks, stream_counts


def tie_heavy_row(G, rng, levels=9):
    """similarities on a coarse grid (multiples of 1/8 in [-0.5, 0.5]): many exact ties, positive, zero and negative values"""
    return (rng.integers(0, levels, size=G).astype(np.float32) - (levels // 2)) / np.float32(8)


@pytest.mark.parametrize("G,shards", [(700, 1), (1000, 3), (257, 2), (512, 1), (31, 1), (1031, 4)])
def test_count_first_protocol_reproduces_the_hit_ranks(G, shards):
    rng = np.random.default_rng(G * 10 + shards)
    cuts = sorted(rng.choice(np.arange(1, G), size=shards - 1, replace=False).tolist()) if shards > 1 else []
    bounds = list(zip([0] + cuts, cuts + [G]))
    for trial in range(6):
        sim = tie_heavy_row(G, rng)
        rel = np.sort(rng.choice(G, size=int(rng.integers(1, 7)), replace=False)).tolist()
        got = stream_counts(sim, rel, bounds)
        assert np.array_equal(got, oracle_ranks(sim, rel)), (trial, rel)


def test_items_in_the_padded_tail_tile_and_negative_thresholds():
    """Relevant items inside the last (padded) tile, with negative similarity: the zero padding rows would out-rank them if they were
    not taken back; an item in the very last valid row; ties on both sides of it."""
    G = 256 + 77
    rng = np.random.default_rng(3)
    sim = tie_heavy_row(G, rng)
    sim[300:] = np.float32(-0.25)                  # one long run of ties through the end of the gallery
    rel = [5, 299, 310, G - 1]
    assert np.array_equal(stream_counts(sim, rel, [(0, G)]), oracle_ranks(sim, rel))
    assert np.array_equal(stream_counts(sim, rel, [(0, 200), (200, G)]), oracle_ranks(sim, rel))


def test_all_equal_row_is_ranked_by_index_alone():
    G = 600
    sim = np.full(G, np.float32(0.125))
    rel = [0, 31, 32, 255, 256, 599]
    assert np.array_equal(stream_counts(sim, rel, [(0, G)]), np.asarray(rel))
    assert np.array_equal(stream_counts(sim, rel, [(0, 100), (100, 400), (400, G)]), np.asarray(rel))


def test_next_below_is_the_largest_smaller_float():
    for x in (1.0, 0.125, -0.125, 1e-30, -1e-30):
        y = next_below(np.float32(x))
        assert y < np.float32(x) and np.nextafter(y, np.float32(np.inf), dtype=np.float32) == np.float32(x)
